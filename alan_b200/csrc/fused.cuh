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

template <typename T, int D, bool BWD>
__global__ void __launch_bounds__(FANLSE_WARPS * 32) fan_lse_kernel(const __grid_constant__ FanLseParams<T> p) {
    extern __shared__ __align__(16) unsigned char fan_smem[];
    constexpr int DP = (D + 3) & ~3;
    const int FP = (p.F + 3) & ~3;
    T* Wt = (T*)fan_smem;                 // [D][FP]
    T* cc = Wt + D * FP;                  // [FP]
    T* warp_base = cc + FP;
    const int warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = p.Kk * DP + ((p.Kk + 3) & ~3);
    T* Tt = warp_base + (i64)warp_in_cta * per_warp;    // [Kk][DP]
    T* Bs = Tt + p.Kk * DP;                             // [Kk]

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

    const i64 warp = (i64)blockIdx.x * FANLSE_WARPS + warp_in_cta;
    const i64 nwarps = (i64)gridDim.x * FANLSE_WARPS;
    for (i64 rho = warp; rho < p.n_rho; rho += nwarps) {
        i64 voff = 0, loff = 0, ooff = 0;
        i64 boff[AB_MAXL];
        for (int i = 0; i < p.nb; ++i) boff[i] = 0;
        {
            i64 lin = rho;
#pragma unroll 1
            for (int k = p.rd.nd - 1; k >= 0; --k) {
                int sz = p.rd.size[k];
                i64 q = lin / sz;
                int ix = (int)(lin - q * sz);
                lin = q;
                voff += ix * p.vstride[k]; loff += ix * p.lstride[k]; ooff += ix * p.ostride[k];
                for (int i = 0; i < p.nb; ++i) boff[i] += ix * p.bstride[i][k];
            }
        }
        // squared residual tile of this rho
        for (int e = lane; e < p.Kk * D; e += 32) {
            int k = e / D, d = e - k * D;
            T df = p.v[voff + k * p.v_k + d * p.v_ev] - p.l[loff + k * p.l_k + d * p.l_ev];
            Tt[k * DP + d] = df * df;
        }
        if (DP != D)
            for (int e = lane; e < p.Kk * (DP - D); e += 32) {
                int k = e / (DP - D), d = D + (e - k * (DP - D));
                Tt[k * DP + d] = T(0);
            }
        for (int k = lane; k < p.Kk; k += 32) {
            T b = T(0);
            for (int i = 0; i < p.nb; ++i) b += p.bcoeff[i] * p.b[i][boff[i] + k * p.b_k[i]];
            Bs[k] = b;
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
                for (int k0 = 0; k0 < p.Kk; k0 += 4) {
                    T sv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = k0 + j;
                        if (k < p.Kk) {
                            T acc = T(0);
#pragma unroll
                            for (int d = 0; d < DP; d += 4) {
                                const Vec4<T> t4 = *reinterpret_cast<const Vec4<T>*>(&Tt[k * DP + d]);
                                acc += t4.x * wr[d]; acc += t4.y * wr[d + 1];
                                acc += t4.z * wr[d + 2]; acc += t4.w * wr[d + 3];
                            }
                            sv[j] = Bs[k] - acc - cf;
                        } else sv[j] = neg_inf<T>();
                    }
                    T mx = ab_max(ab_max(sv[0], sv[1]), ab_max(sv[2], sv[3]));
                    T mn = ab_max(m, mx);
                    sum = sum * ab_exp(m - mn);
#pragma unroll
                    for (int j = 0; j < 4; ++j) sum += ab_exp(sv[j] - mn);
                    m = mn;
                }
                if (active) p.out[ooff + (i64)f * p.o_f] = ab_log(sum + Eps<T>::v()) + m + p.cadd;
            } else {
                const T lz = active ? p.lse[ooff + (i64)f * p.o_f] : T(0);
                const T gz = active ? p.gout[ooff + (i64)f * p.o_f] : T(0);
                for (int k = 0; k < p.Kk; ++k) {
                    T acc = T(0);
#pragma unroll
                    for (int d = 0; d < DP; d += 4) {
                        const Vec4<T> t4 = *reinterpret_cast<const Vec4<T>*>(&Tt[k * DP + d]);
                        acc += t4.x * wr[d]; acc += t4.y * wr[d + 1];
                        acc += t4.z * wr[d + 2]; acc += t4.w * wr[d + 3];
                    }
                    T wv = active ? gz * ab_exp(Bs[k] - acc - cf + p.cadd - lz) : T(0);
                    T tot = warp_sum(wv);
                    if ((k & 31) == lane) keep[(k >> 5) & 3] += tot;
                }
            }
        }
        if (BWD) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int k = lane + 32 * q;
                if (k < p.Kk) p.gS[rho * p.Kk + k] = keep[q];
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
    size_t smem = (size_t)(D * FP + FP + FANLSE_WARPS * (p.Kk * DP + ((p.Kk + 3) & ~3))) * sizeof(T);
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
