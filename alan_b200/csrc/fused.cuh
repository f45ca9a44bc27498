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
