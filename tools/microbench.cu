// microbench.cu -- measured SM pipe peaks on the box (roofline denominators for compute-bound kernels):
// FFMA, FFMA2 (fma.rn.f32x2), MUFU.EX2, broadcast LDS.128, SHFL.  nvcc -arch=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, float seed) {
    __shared__ __align__(16) float sm[1024];
    sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    float a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    float2 b0 = {a0, a1}, b1 = {a2, a3}, b2 = {a4, a5}, b3 = {a6, a7}, b4 = {a1, a0}, b5 = {a3, a2}, b6 = {a5, a4}, b7 = {a7, a6};
    float2 m = {seed * 0.5f, seed * 0.25f}, c = {0.001f, 0.002f};
    int lane = threadIdx.x & 31;
    for (int i = 0; i < ITERS; ++i) {
        if (MODE == 0) {
            a0 = fmaf(a0, m.x, c.x); a1 = fmaf(a1, m.x, c.x); a2 = fmaf(a2, m.x, c.x); a3 = fmaf(a3, m.x, c.x);
            a4 = fmaf(a4, m.x, c.x); a5 = fmaf(a5, m.x, c.x); a6 = fmaf(a6, m.x, c.x); a7 = fmaf(a7, m.x, c.x);
        } else if (MODE == 1) {
            b0 = __ffma2_rn(b0, m, c); b1 = __ffma2_rn(b1, m, c); b2 = __ffma2_rn(b2, m, c); b3 = __ffma2_rn(b3, m, c);
            b4 = __ffma2_rn(b4, m, c); b5 = __ffma2_rn(b5, m, c); b6 = __ffma2_rn(b6, m, c); b7 = __ffma2_rn(b7, m, c);
        } else if (MODE == 2) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
        } else if (MODE == 3) {            // broadcast LDS.128: all lanes same address
            const float4* p = reinterpret_cast<const float4*>(sm) + ((i & 7) * 8);
            float4 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3], v4 = p[4], v5 = p[5], v6 = p[6], v7 = p[7];
            a0 += v0.x + v0.y + v0.z + v0.w; a1 += v1.x + v1.y + v1.z + v1.w; a2 += v2.x + v2.y + v2.z + v2.w; a3 += v3.x + v3.y + v3.z + v3.w;
            a4 += v4.x + v4.y + v4.z + v4.w; a5 += v5.x + v5.y + v5.z + v5.w; a6 += v6.x + v6.y + v6.z + v6.w; a7 += v7.x + v7.y + v7.z + v7.w;
        } else if (MODE == 4) {
            a0 = __shfl_xor_sync(0xffffffffu, a0, 1); a1 = __shfl_xor_sync(0xffffffffu, a1, 2); a2 = __shfl_xor_sync(0xffffffffu, a2, 4);
            a3 = __shfl_xor_sync(0xffffffffu, a3, 8); a4 = __shfl_xor_sync(0xffffffffu, a4, 16); a5 = __shfl_xor_sync(0xffffffffu, a5, 1);
            a6 = __shfl_xor_sync(0xffffffffu, a6, 2); a7 = __shfl_xor_sync(0xffffffffu, a7, 4);
        } else if (MODE == 5) {            // LDS.128 with 6 distinct rows per warp (the fan_lse access pattern)
            const float4* p = reinterpret_cast<const float4*>(sm) + (lane / 5) * 5 + ((i & 3) * 40);
            float4 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3], v4 = p[4];
            a0 += v0.x + v0.y + v0.z + v0.w; a1 += v1.x + v1.y + v1.z + v1.w; a2 += v2.x + v2.y + v2.z + v2.w; a3 += v3.x + v3.y + v3.z + v3.w;
            a4 += v4.x + v4.y + v4.z + v4.w;
        } else if (MODE == 6) {            // FFMA with 3 distinct register operands
            a0 = fmaf(a1, a2, a0); a3 = fmaf(a4, a5, a3); a6 = fmaf(a7, a1, a6); a2 = fmaf(a5, a4, a2);
            a0 = fmaf(a3, a6, a0); a3 = fmaf(a2, a7, a3); a6 = fmaf(a0, a1, a6); a2 = fmaf(a3, a4, a2);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + b0.x + b0.y + b1.x + b1.y + b2.x + b2.y + b3.x + b3.y +
                                                 b4.x + b4.y + b5.x + b5.y + b6.x + b6.y + b7.x + b7.y;
}
template <int MODE> void run(const char* name, double per_iter_ops, float* out, int sms, int mhz) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = sms * 2, block = 1024;
    k<MODE><<<grid, block>>>(out, 1.0f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<MODE><<<grid, block>>>(out, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double ops = per_iter_ops * ITERS * (double)grid * block;
    double rate = ops / (best * 1e-3);
    printf("%-28s %8.3f ms  %10.3f Gop/s  = %7.2f lane-ops/clk/SM at %d MHz max clock\n", name, best, rate / 1e9, rate / sms / (mhz * 1e6), mhz);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int mhz = p.clockRate / 1000;
    printf("device %s  SMs %d  clock %d MHz\n", p.name, p.multiProcessorCount, mhz);
    float* out; cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 2 * 1024);
    run<0>("FFMA (imm-like, 8 chains)", 8, out, p.multiProcessorCount, mhz);
    run<6>("FFMA (3 reg operands)", 8, out, p.multiProcessorCount, mhz);
    run<1>("FFMA2 (as scalar FMAs)", 16, out, p.multiProcessorCount, mhz);
    run<2>("MUFU.EX2", 8, out, p.multiProcessorCount, mhz);
    run<3>("LDS.128 broadcast (floats)", 32, out, p.multiProcessorCount, mhz);
    run<5>("LDS.128 6 rows/warp (floats)", 20, out, p.multiProcessorCount, mhz);
    run<4>("SHFL", 8, out, p.multiProcessorCount, mhz);
    return 0;
}
