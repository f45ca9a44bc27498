// fused.cuh -- tuned kernels the planner selects for recognised factor patterns.
//
// normal_fan: the dominant factor of hierarchical Gaussian models (MovieLens-shaped cfg-2/5,
// radon): log N(value; loc, scale) summed over the event dim, where value/loc carry the "row"
// axes (plates and their K axes) and scale carries one extra "fan" K axis of its own:
//
//     out[row, f] = - sum_d (v[row,d] - l[row,d])^2 * w[f,d] - c[f],
//     w[f,d] = 1 / (2 scale[f,d]^2),  c[f] = sum_d log scale[f,d] + D * log sqrt(2 pi)
//
// i.e. a skinny GEMM  T[row, d] x W[d, f]  with the squared residual T built on the fly in
// registers (never stored) and W^T staged once per CTA in shared memory.  Each thread owns two
// rows (2*D residuals in registers) and walks the fan axis four columns at a time: one
// broadcast LDS.128 of W feeds eight FFMAs.  reference call site: TorchDimDist.log_prob
// (src/alan/TorchDimDist.py:127-162) on the `[M,Kz,d,Kmu,Kpsi]` broadcast of SURVEY.md §2.4 K1.
#pragma once
#include "kernels.cuh"

template <typename T>
struct FanParams {
    Dims rd;                              // row dims (n_a = nd)
    i64 vstride[AB_MAXD], lstride[AB_MAXD], ostride[AB_MAXD];
    i64 v_ev, l_ev;                       // event strides of value / loc (0 = broadcast)
    const T* v; const T* l; const T* s;
    i64 s_f, s_ev;                        // scale strides along fan axis / event
    int F;                                // fan extent (1 if scale has no axis of its own)
    i64 o_f;                              // out stride along the fan axis
    T* out;
    i64 n_rows;
};

template <typename T> struct Vec4 { T x, y, z, w; };

template <typename T, int D>
__global__ void __launch_bounds__(256) normal_fan_kernel(const __grid_constant__ FanParams<T> p) {
    extern __shared__ __align__(16) unsigned char fan_smem[];
    const int FP = (p.F + 3) & ~3;
    T* Wt = (T*)fan_smem;                 // [D][FP]
    T* cc = Wt + D * FP;                  // [FP]
    for (int i = threadIdx.x; i < D * FP; i += blockDim.x) {
        int d = i / FP, f = i - d * FP;
        T w = T(0);
        if (f < p.F) { T sc = p.s[f * p.s_f + d * p.s_ev]; w = T(1) / (T(2) * (sc * sc)); }
        Wt[i] = w;
    }
    for (int f = threadIdx.x; f < FP; f += blockDim.x) {
        T c = T(0);
        if (f < p.F) {
            for (int d = 0; d < D; ++d) c += ab_log(p.s[f * p.s_f + d * p.s_ev]);
            c += T(D) * T(HALF_LOG_2PI);
        }
        cc[f] = c;
    }
    __syncthreads();

    const i64 chunk = 2 * (i64)blockDim.x;
    for (i64 base = (i64)blockIdx.x * chunk; base < p.n_rows; base += (i64)gridDim.x * chunk) {
        T tt[2][D];
        i64 ooff[2];
        bool live[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            i64 row = base + threadIdx.x + (i64)r * blockDim.x;
            live[r] = row < p.n_rows;
            i64 voff = 0, loff = 0, oo = 0;
            if (live[r]) {
                i64 lin = row;
#pragma unroll 1
                for (int k = p.rd.nd - 1; k >= 0; --k) {
                    int sz = p.rd.size[k];
                    i64 q = lin / sz;
                    int ix = (int)(lin - q * sz);
                    lin = q;
                    voff += ix * p.vstride[k]; loff += ix * p.lstride[k]; oo += ix * p.ostride[k];
                }
            }
            ooff[r] = oo;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                T df = live[r] ? p.v[voff + d * p.v_ev] - p.l[loff + d * p.l_ev] : T(0);
                tt[r][d] = df * df;
            }
        }
        for (int f0 = 0; f0 < FP; f0 += 4) {
            T acc[2][4];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[r][j] = T(0);
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const Vec4<T> w = *reinterpret_cast<const Vec4<T>*>(&Wt[d * FP + f0]);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    acc[r][0] += tt[r][d] * w.x; acc[r][1] += tt[r][d] * w.y;
                    acc[r][2] += tt[r][d] * w.z; acc[r][3] += tt[r][d] * w.w;
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (!live[r]) continue;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (f0 + j < p.F) p.out[ooff[r] + (i64)(f0 + j) * p.o_f] = -acc[r][j] - cc[f0 + j];
            }
        }
    }
}

template <typename T, int D>
static void launch_fan_D(const FanParams<T>& p, cudaStream_t stream, int sm_count) {
    const int FP = (p.F + 3) & ~3;
    size_t smem = (size_t)(D * FP + FP) * sizeof(T);
    i64 blocks = (p.n_rows + 511) / 512;
    i64 cap = (i64)sm_count * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    normal_fan_kernel<T, D><<<(int)blocks, 256, smem, stream>>>(p);
}

static bool fan_supported_D(int D) {
    switch (D) { case 1: case 2: case 3: case 4: case 6: case 8: case 12: case 16: case 18: case 24: case 32: return true; }
    return false;
}

template <typename T>
static int launch_fan(const FanParams<T>& p, int D, cudaStream_t stream, int sm_count) {
    switch (D) {
        case 1: launch_fan_D<T, 1>(p, stream, sm_count); break;
        case 2: launch_fan_D<T, 2>(p, stream, sm_count); break;
        case 3: launch_fan_D<T, 3>(p, stream, sm_count); break;
        case 4: launch_fan_D<T, 4>(p, stream, sm_count); break;
        case 6: launch_fan_D<T, 6>(p, stream, sm_count); break;
        case 8: launch_fan_D<T, 8>(p, stream, sm_count); break;
        case 12: launch_fan_D<T, 12>(p, stream, sm_count); break;
        case 16: launch_fan_D<T, 16>(p, stream, sm_count); break;
        case 18: launch_fan_D<T, 18>(p, stream, sm_count); break;
        case 24: launch_fan_D<T, 24>(p, stream, sm_count); break;
        case 32: launch_fan_D<T, 32>(p, stream, sm_count); break;
        default: return 1;
    }
    return 0;
}


// dot: out[o] = sum_e a[o,e] * b[o,e] over the trailing event dim, both operands broadcast
// through strides.  Covers `lambda z, x: z @ x` (movielens.py:40) without the factor VM.
template <typename T>
struct DotParams {
    Dims d;                 // n_a = output dims, one reduced dim
    Opnd a, b;
    T* out;
    i64 n_out, n_red;
};

template <typename T>
__global__ void __launch_bounds__(256) dot_kernel(const __grid_constant__ DotParams<T> p) {
    int idx[AB_MAXD];
    const T* A = (const T*)p.a.ptr;
    const T* B = (const T*)p.b.ptr;
    const i64 sa = p.a.stride[p.d.n_a], sb = p.b.stride[p.d.n_a];
    for (i64 o = (i64)blockIdx.x * blockDim.x + threadIdx.x; o < p.n_out; o += (i64)gridDim.x * blockDim.x) {
        unravel(o, p.d, 0, p.d.n_a, idx);
        i64 ba = dot_stride(p.a, idx, 0, p.d.n_a), bb = dot_stride(p.b, idx, 0, p.d.n_a);
        T acc = T(0);
        for (i64 e = 0; e < p.n_red; ++e) acc += A[ba + e * sa] * B[bb + e * sb];
        p.out[o] = acc;
    }
}


// ------------------------------------------------------------------------------------------
// fan_lse: factor kernel fused with the log-semiring contraction that consumes it
// (north_star (1)+(2): densities straight into registers, online max/rescale LSE, nothing
// materialised).  For every rho (the row axes except the contracted axis kappa):
//
//     out[rho, f] = log( sum_kappa exp( A[rho,kappa,f] + B[rho,kappa] - max ) + eps ) + max + cadd
//     A = - sum_d (v[rho,kappa,d] - l[rho,kappa,d])^2 w[f,d] - c[f]        (normal_fan above)
//     B = sum_i coeff_i * b_i[rho,kappa]                                   (the other, small factors)
//
// One warp owns one rho at a time.  It first builds the squared-residual tile T[kappa][d] of
// that rho in its private shared-memory slice (coalesced reads of v), then lane f keeps
// w[f, 0..D) in registers and walks kappa: each T row is one broadcast LDS.128 stream shared by
// all lanes, and the LSE over kappa is an online max/rescale in registers -- no cross-lane
// traffic at all in the forward pass.  reference: logsumexp_sum (src/alan/reduce_Ks.py:249-251)
// over the factor of TorchDimDist.log_prob (TorchDimDist.py:127-162).
//
// The adjoint kernel recomputes A the same way and emits gS[rho,kappa] = sum_f gout[rho,f] *
// exp(A + B + cadd - out[rho,f]), the adjoint of the small-factor sum (what RWS needs: the
// gradient w.r.t. log Q); the sum over f is a fixed-order warp butterfly.
// ------------------------------------------------------------------------------------------
template <typename T>
struct FanLseParams {
    Dims rd;                              // rho dims (n_a = nd)
    i64 vstride[AB_MAXD], lstride[AB_MAXD], ostride[AB_MAXD];
    i64 v_k, l_k, v_ev, l_ev;
    const T* v; const T* l; const T* s;
    i64 s_f, s_ev;
    int F; i64 o_f;
    int Kk;
    int nb;
    const T* b[AB_MAXL];
    i64 bstride[AB_MAXL][AB_MAXD];
    i64 b_k[AB_MAXL];
    T bcoeff[AB_MAXL];
    T cadd;
    T* out;                               // fwd: result; bwd: unused
    const T* lse; const T* gout;          // bwd
    T* gS;                                // bwd: [rho, kappa] contiguous
    i64 n_rho;
};

#define FANLSE_WARPS 8

// base-2 exponent / logarithm on the SFU for the fp32 instantiation (one MUFU each); the fp64
// instantiation keeps exp()/log().  Inputs are pre-scaled by log2(e) once per tile, so the inner
// loop has no multiply in front of the exponential.
template <typename T> struct FastExp;
template <> struct FastExp<float> {
    static __device__ __forceinline__ float scale() { return 1.4426950408889634f; }     // log2(e)
    static __device__ __forceinline__ float unscale() { return 0.6931471805599453f; }   // ln(2)
    static __device__ __forceinline__ float ex(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
};
template <> struct FastExp<double> {
    static __device__ __forceinline__ double scale() { return 1.0; }
    static __device__ __forceinline__ double unscale() { return 1.0; }
    static __device__ __forceinline__ double ex(double x) { return exp(x); }
};

template <typename T, int D, bool BWD>
__global__ void __launch_bounds__(FANLSE_WARPS * 32) fan_lse_kernel(const __grid_constant__ FanLseParams<T> p) {
    extern __shared__ __align__(16) unsigned char fan_smem[];
    constexpr int DP = (D + 3) & ~3;
    const int FP = (p.F + 3) & ~3;
    const int K4 = (p.Kk + 3) & ~3;         // kappa padded to a multiple of four rows
    T* Wt = (T*)fan_smem;                   // [D][FP]   w * log2e
    T* cc = Wt + D * FP;                    // [FP]      c * log2e
    T* warp_base = cc + FP;
    const int warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = K4 * DP + K4;
    T* Tt = warp_base + warp_in_cta * per_warp;         // [K4][DP]  squared residuals
    T* Bs = Tt + K4 * DP;                               // [K4]      small-factor sum * log2e (-inf on padding)
    const T LS = FastExp<T>::scale();

    for (int i = threadIdx.x; i < D * FP; i += blockDim.x) {
        int d = i / FP, f = i - d * FP;
        T w = T(0);
        if (f < p.F) { T sc = p.s[f * p.s_f + d * p.s_ev]; w = LS / (T(2) * (sc * sc)); }
        Wt[i] = w;
    }
    for (int f = threadIdx.x; f < FP; f += blockDim.x) {
        T c = T(0);
        if (f < p.F) {
            for (int d = 0; d < D; ++d) c += ab_log(p.s[f * p.s_f + d * p.s_ev]);
            c = (c + T(D) * T(HALF_LOG_2PI)) * LS;
        }
        cc[f] = c;
    }
    // padding rows / columns of the per-warp tile never change
    for (int e = lane; e < K4 * DP; e += 32) Tt[e] = T(0);
    for (int k = lane; k < K4; k += 32) Bs[k] = neg_inf<T>();
    __syncthreads();

    const unsigned n_rho = (unsigned)p.n_rho;
    const unsigned warp = blockIdx.x * FANLSE_WARPS + warp_in_cta;
    const unsigned nwarps = gridDim.x * FANLSE_WARPS;
    const bool contig = (p.v_ev == 1 && p.v_k == D);    // value rows of one rho are one contiguous run
    for (unsigned rho = warp; rho < n_rho; rho += nwarps) {
        i64 voff = 0, loff = 0, ooff = 0;
        i64 boff[AB_MAXL];
#pragma unroll
        for (int i = 0; i < AB_MAXL; ++i) boff[i] = 0;
        {
            unsigned lin = rho;
#pragma unroll 1
            for (int k = p.rd.nd - 1; k >= 0; --k) {
                unsigned sz = (unsigned)p.rd.size[k];
                unsigned q = lin / sz;
                unsigned ix = lin - q * sz;
                lin = q;
                voff += ix * p.vstride[k]; loff += ix * p.lstride[k]; ooff += ix * p.ostride[k];
#pragma unroll
                for (int i = 0; i < AB_MAXL; ++i) if (i < p.nb) boff[i] += ix * p.bstride[i][k];
            }
        }
        // squared residual tile of this rho: T[k][d] = (v - l)^2
        {
            const T* vp = p.v + voff;
            const T* lp = p.l + loff;
            int k = lane / D, d = lane - k * D;             // (k, d) of element e = lane, stepped by 32
            constexpr int SK = 32 / D, SD = 32 - SK * D;
            for (int e = lane; e < p.Kk * D; e += 32) {
                T vv = contig ? vp[e] : vp[k * p.v_k + d * p.v_ev];
                T df = vv - lp[k * p.l_k + d * p.l_ev];
                Tt[k * DP + d] = df * df;
                k += SK; d += SD;
                if (d >= D) { d -= D; k += 1; }
            }
        }
        for (int k = lane; k < p.Kk; k += 32) {
            T b = T(0);
#pragma unroll
            for (int i = 0; i < AB_MAXL; ++i) if (i < p.nb) b += p.bcoeff[i] * p.b[i][boff[i] + k * p.b_k[i]];
            Bs[k] = b * LS;
        }
        __syncwarp();

        T keep[4] = {T(0), T(0), T(0), T(0)};       // bwd: gS for kappa = lane + 32 q
        for (int f0 = 0; f0 < p.F; f0 += 32) {
            const int f = f0 + lane;
            const bool active = f < p.F;
            T wr[DP];
#pragma unroll
            for (int d = 0; d < DP; ++d) wr[d] = (active && d < D) ? Wt[d * FP + f] : T(0);
            const T cf = active ? cc[f] : T(0);
            if (!BWD) {
                T m = neg_inf<T>(), sum = T(0);
                for (int k0 = 0; k0 < K4; k0 += 4) {
                    T sv[4];
                    const Vec4<T> b4 = *reinterpret_cast<const Vec4<T>*>(&Bs[k0]);
                    const T bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const T* row = Tt + (k0 + j) * DP;
                        T a0 = T(0), a1 = T(0);
#pragma unroll
                        for (int d = 0; d < DP; d += 4) {
                            const Vec4<T> t4 = *reinterpret_cast<const Vec4<T>*>(row + d);
                            a0 += t4.x * wr[d];
                            if (d + 1 < D) a1 += t4.y * wr[d + 1];
                            if (d + 2 < D) a0 += t4.z * wr[d + 2];
                            if (d + 3 < D) a1 += t4.w * wr[d + 3];
                        }
                        sv[j] = bb[j] - (a0 + a1) - cf;
                    }
                    T mn = ab_max(ab_max(ab_max(sv[0], sv[1]), ab_max(sv[2], sv[3])), m);
                    sum = sum * FastExp<T>::ex(m - mn);
#pragma unroll
                    for (int j = 0; j < 4; ++j) sum += FastExp<T>::ex(sv[j] - mn);
                    m = mn;
                }
                if (active)
                    p.out[ooff + (i64)f * p.o_f] = ab_log(sum + Eps<T>::v()) + m * FastExp<T>::unscale() + p.cadd;
            } else {
                const T lz = active ? (p.lse[ooff + (i64)f * p.o_f] - p.cadd) * LS : -neg_inf<T>();   // idle lanes: weight 0
                const T gz = active ? p.gout[ooff + (i64)f * p.o_f] : T(0);
                for (int kb = 0; kb < K4; kb += 32) {
                    T wv[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        wv[j] = T(0);
                        if (kb + j < K4) {                                  // warp-uniform
                            const T* row = Tt + (kb + j) * DP;
                            T a0 = T(0), a1 = T(0);
#pragma unroll
                            for (int d = 0; d < DP; d += 4) {
                                const Vec4<T> t4 = *reinterpret_cast<const Vec4<T>*>(row + d);
                                a0 += t4.x * wr[d];
                                if (d + 1 < D) a1 += t4.y * wr[d + 1];
                                if (d + 2 < D) a0 += t4.z * wr[d + 2];
                                if (d + 3 < D) a1 += t4.w * wr[d + 3];
                            }
                            wv[j] = gz * FastExp<T>::ex(Bs[kb + j] - (a0 + a1) - cf - lz);   // 0 on padding rows
                        }
                    }
                    // fixed-order butterfly reduce-scatter over the lanes (f): lane j ends with the sum for
                    // kappa = kb + j in wv[0]; 31 shuffles for 32 kappas instead of 5 per kappa.
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
                        for (int i = 0; i < off; ++i) {
                            const bool up = (lane & off) != 0;
                            T send = up ? wv[i] : wv[i + off];
                            T mine = up ? wv[i + off] : wv[i];
                            wv[i] = mine + __shfl_xor_sync(0xffffffffu, send, off);
                        }
                    }
                    keep[(kb >> 5) & 3] += wv[0];
                }
            }
        }
        if (BWD) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int k = lane + 32 * q;
                if (k < p.Kk) p.gS[(i64)rho * p.Kk + k] = keep[q];
            }
        }
        __syncwarp();
    }
}

template <typename T, int D>
static int launch_fan_lse_D(const FanLseParams<T>& p, bool bwd, cudaStream_t stream, int sm_count) {
    constexpr int DP = (D + 3) & ~3;
    const int FP = (p.F + 3) & ~3;
    if (bwd && p.Kk > 128) return 2;
    const int K4 = (p.Kk + 3) & ~3;
    size_t smem = (size_t)(D * FP + FP + FANLSE_WARPS * (K4 * DP + K4)) * sizeof(T);
    if (smem > 200 * 1024) return 2;
    i64 blocks = (p.n_rho + FANLSE_WARPS - 1) / FANLSE_WARPS;
    i64 cap = (i64)sm_count * 6;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (bwd) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(fan_lse_kernel<T, D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        fan_lse_kernel<T, D, true><<<(int)blocks, FANLSE_WARPS * 32, smem, stream>>>(p);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(fan_lse_kernel<T, D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        fan_lse_kernel<T, D, false><<<(int)blocks, FANLSE_WARPS * 32, smem, stream>>>(p);
    }
    return 0;
}

template <typename T>
static int launch_fan_lse(const FanLseParams<T>& p, int D, bool bwd, cudaStream_t stream, int sm_count) {
    switch (D) {
        case 1: return launch_fan_lse_D<T, 1>(p, bwd, stream, sm_count);
        case 2: return launch_fan_lse_D<T, 2>(p, bwd, stream, sm_count);
        case 3: return launch_fan_lse_D<T, 3>(p, bwd, stream, sm_count);
        case 4: return launch_fan_lse_D<T, 4>(p, bwd, stream, sm_count);
        case 6: return launch_fan_lse_D<T, 6>(p, bwd, stream, sm_count);
        case 8: return launch_fan_lse_D<T, 8>(p, bwd, stream, sm_count);
        case 12: return launch_fan_lse_D<T, 12>(p, bwd, stream, sm_count);
        case 16: return launch_fan_lse_D<T, 16>(p, bwd, stream, sm_count);
        case 18: return launch_fan_lse_D<T, 18>(p, bwd, stream, sm_count);
        case 24: return launch_fan_lse_D<T, 24>(p, bwd, stream, sm_count);
        case 32: return launch_fan_lse_D<T, 32>(p, bwd, stream, sm_count);
    }
    return 1;
}
