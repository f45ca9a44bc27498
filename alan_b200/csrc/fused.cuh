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


// ------------------------------------------------------------------------------------------
// Adjoint of normal_fan (the VI / reparameterised path: gradients w.r.t. the sample values, the
// location and the scale).  G[row, f] is the adjoint of out[row, f]; with T = (v - l)^2,
//     R[row, d]  = 2 (v - l) sum_f G[row, f] w[f, d]          -> d/dv = -R, d/dl = +R (summed by reduce ops)
//     V[f, d]    = sum_row G[row, f] T[row, d],  Wsum[f] = sum_row G[row, f]
//                                                          -> d/dscale[f, d] = V / scale^3 - Wsum / scale
// Two kernels: fan_bwd_rows (one thread per row, w broadcast from shared memory) writes R;
// fan_bwd_scale (lanes = f, warps stride over rows) writes per-CTA partials [cta][f][D] and [cta][f] that a
// fixed-order reduce op sums.  reference: autograd through TorchDimDist.log_prob (TorchDimDist.py:127-162).
// ------------------------------------------------------------------------------------------
template <typename T>
struct FanBwdParams {
    FanParams<T> f;                       // geometry of the forward factor; f.out = G (adjoint of out)
    T* R;                                 // [n_rows, D] contiguous
    T* partial;                           // [n_cta, F, D]   per-CTA partials of V
    T* partial_w;                         // [n_cta, F]      per-CTA partials of Wsum
    int n_cta;
};

template <typename T, int D>
__global__ void __launch_bounds__(256) fan_bwd_rows_kernel(const __grid_constant__ FanBwdParams<T> q) {
    const FanParams<T>& p = q.f;
    extern __shared__ __align__(16) unsigned char fan_smem[];
    T* W = (T*)fan_smem;                  // [F][D]   1 / (2 scale^2)
    for (int i = threadIdx.x; i < p.F * D; i += blockDim.x) {
        int f = i / D, d = i - f * D;
        T sc = p.s[f * p.s_f + d * p.s_ev];
        W[i] = T(1) / (T(2) * (sc * sc));
    }
    __syncthreads();
    for (i64 row = (i64)blockIdx.x * blockDim.x + threadIdx.x; row < p.n_rows; row += (i64)gridDim.x * blockDim.x) {
        i64 voff = 0, loff = 0, goff = 0, lin = row;
#pragma unroll 1
        for (int k = p.rd.nd - 1; k >= 0; --k) {
            int sz = p.rd.size[k];
            i64 qq = lin / sz;
            int ix = (int)(lin - qq * sz);
            lin = qq;
            voff += ix * p.vstride[k]; loff += ix * p.lstride[k]; goff += ix * p.ostride[k];
        }
        T df[D], U[D];
#pragma unroll
        for (int d = 0; d < D; ++d) { df[d] = p.v[voff + d * p.v_ev] - p.l[loff + d * p.l_ev]; U[d] = T(0); }
        for (int f = 0; f < p.F; ++f) {
            const T g = p.out[goff + (i64)f * p.o_f];
            const T* w = W + f * D;
#pragma unroll
            for (int d = 0; d < D; ++d) U[d] += g * w[d];
        }
        // out = -sum T w - c  =>  d out / d v = -2 (v - l) w
#pragma unroll
        for (int d = 0; d < D; ++d) q.R[row * D + d] = T(2) * df[d] * U[d];
    }
}

template <typename T, int D>
__global__ void __launch_bounds__(256) fan_bwd_scale_kernel(const __grid_constant__ FanBwdParams<T> q) {
    const FanParams<T>& p = q.f;
    __shared__ T red[8][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const i64 rows_per_cta = (p.n_rows + gridDim.x - 1) / gridDim.x;
    const i64 r0 = (i64)blockIdx.x * rows_per_cta, r1 = r0 + rows_per_cta < p.n_rows ? r0 + rows_per_cta : p.n_rows;
    for (int f0 = 0; f0 < p.F; f0 += 32) {
        const int f = f0 + lane;
        T acc[D + 1];
#pragma unroll
        for (int d = 0; d <= D; ++d) acc[d] = T(0);
        for (i64 row = r0 + wid; row < r1; row += nw) {
            i64 voff = 0, loff = 0, goff = 0, lin = row;
#pragma unroll 1
            for (int k = p.rd.nd - 1; k >= 0; --k) {
                int sz = p.rd.size[k];
                i64 qq = lin / sz;
                int ix = (int)(lin - qq * sz);
                lin = qq;
                voff += ix * p.vstride[k]; loff += ix * p.lstride[k]; goff += ix * p.ostride[k];
            }
            const T g = f < p.F ? p.out[goff + (i64)f * p.o_f] : T(0);
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const T df = p.v[voff + d * p.v_ev] - p.l[loff + d * p.l_ev];      // same address on every lane: broadcast
                acc[d] += g * (df * df);
            }
            acc[D] += g;
        }
        // fixed-order sum over the warps of the CTA
#pragma unroll 1
        for (int d = 0; d <= D; ++d) {
            red[wid][lane] = acc[d];
            __syncthreads();
            if (wid == 0 && f < p.F) {
                T a = T(0);
                for (int w = 0; w < nw; ++w) a += red[w][lane];
                if (d < D) q.partial[((i64)blockIdx.x * p.F + f) * D + d] = a;
                else q.partial_w[(i64)blockIdx.x * p.F + f] = a;
            }
            __syncthreads();
        }
    }
}

template <typename T, int D>
static void launch_fan_bwd_D(const FanBwdParams<T>& q, int which, cudaStream_t stream, int sm_count) {
    if (which == 0) {
        size_t smem = (size_t)q.f.F * D * sizeof(T);
        i64 blocks = (q.f.n_rows + 255) / 256, cap = (i64)sm_count * 8;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        if (smem > 48 * 1024) cudaFuncSetAttribute(fan_bwd_rows_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        fan_bwd_rows_kernel<T, D><<<(int)blocks, 256, smem, stream>>>(q);
    } else {
        fan_bwd_scale_kernel<T, D><<<q.n_cta, 256, 0, stream>>>(q);
    }
}

template <typename T>
static int launch_fan_bwd(const FanBwdParams<T>& q, int D, int which, cudaStream_t stream, int sm_count) {
    switch (D) {
        case 1: launch_fan_bwd_D<T, 1>(q, which, stream, sm_count); break;
        case 2: launch_fan_bwd_D<T, 2>(q, which, stream, sm_count); break;
        case 3: launch_fan_bwd_D<T, 3>(q, which, stream, sm_count); break;
        case 4: launch_fan_bwd_D<T, 4>(q, which, stream, sm_count); break;
        case 6: launch_fan_bwd_D<T, 6>(q, which, stream, sm_count); break;
        case 8: launch_fan_bwd_D<T, 8>(q, which, stream, sm_count); break;
        case 12: launch_fan_bwd_D<T, 12>(q, which, stream, sm_count); break;
        case 16: launch_fan_bwd_D<T, 16>(q, which, stream, sm_count); break;
        case 18: launch_fan_bwd_D<T, 18>(q, which, stream, sm_count); break;
        case 24: launch_fan_bwd_D<T, 24>(q, which, stream, sm_count); break;
        case 32: launch_fan_bwd_D<T, 32>(q, which, stream, sm_count); break;
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
    i64 gstride[AB_MAXD]; i64 g_f;        // bwd: strides of gout over the rho dims / fan axis (0 = broadcast)
    T* gS;                                // bwd: [rho, kappa] contiguous
    int gs_compact;                       // bwd, dense tcgen05 kernel only: > 0 = gS is [users, gs_compact fan groups, kappa]
    T* psum; int psum_rows;               // fwd, dense tcgen05 kernel only: per-(CTA, team) sums of out over the users,
    i64 ps_lam, ps_f, ps_row;             //      psum[row * ps_row + lam * ps_lam + f * ps_f]
    // dense tcgen05 kernel only: one more small factor evaluated IN the kernel from the value rows its builders
    // already hold,  q_coeff * sum_d log N(v[rho,kappa,d]; q_l[rho,d], q_s[rho,d])  -- the mean-field Gaussian Q
    // factor of the same latent (north_star (1): densities and the P-minus-Q difference straight into registers)
    int qn;
    const T* q_l; const T* q_s;
    i64 q_lstride[AB_MAXD], q_sstride[AB_MAXD], q_lev, q_sev;
    T q_coeff;
    i64 n_rho;
};

// base-2 exponent on the SFU for the fp32 instantiation (one MUFU each); the fp64 instantiation keeps
// exp().  Inputs are pre-scaled by log2(e) once per tile, so the inner loop has no multiply in
// front of the exponential.
template <typename T> struct FastExp;
template <> struct FastExp<float> {
    static __device__ __forceinline__ float scale() { return 1.4426950408889634f; }     // log2(e)
    static __device__ __forceinline__ float unscale() { return 0.6931471805599453f; }   // ln(2)
    static __device__ __forceinline__ float ex(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
    static __device__ __forceinline__ float lowest() { return -3.0e38f; }
};
template <> struct FastExp<double> {
    static __device__ __forceinline__ double scale() { return 1.0; }
    static __device__ __forceinline__ double unscale() { return 1.0; }
    static __device__ __forceinline__ double ex(double x) { return exp(x); }
    static __device__ __forceinline__ double lowest() { return -1.0e300; }
};

// Packed pairs: fp32 pairs go through FFMA2 (fma.rn.f32x2, sm_100): one issue slot per two FMAs,
// which leaves issue bandwidth for the LDS / MUFU traffic of the same warp.
template <typename T> struct Pair2;
template <> struct Pair2<float>  { typedef float2 type; };
template <> struct Pair2<double> { typedef double2 type; };
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ double2 fma2(double2 a, double2 b, double2 c) { return make_double2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 mk2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ double2 mk2(double a, double b) { return make_double2(a, b); }

// ------------------------------------------------------------------------------------------
// bern_dot_sum: data factor of logistic-regression-like likelihoods, fused with the plate sum that
// consumes it (MovieLens `obs ~ Bernoulli(logits = z @ x)` inside plate_2):
//
//     out[o] = cadd + sum_n  log Bernoulli(y[o, n]; logits = sum_d a[o, d] * b[o, n, d])
//
// o runs over the kept axes (plates and K axes), n over the summed plate, d over the event dim.  One
// thread per o keeps its a-row (D values) in registers; the b rows are the same for all threads that
// differ only in axes b does not carry (the K axis of z), so their loads are L1 broadcasts.  Neither
// the logits nor the per-(o, n) log-likelihoods are ever written: HBM traffic is a + b + y + out.
// reference: the lambda `z @ x` (movielens.py:40) -> TorchDimDist.log_prob (TorchDimDist.py:127-162)
// -> lp.sum(plate) (logpq.py:149).
// ------------------------------------------------------------------------------------------
template <typename T>
struct BernDotParams {
    Dims d;                 // n_a = kept dims, the rest are the summed plate dims
    Opnd a, b, y;           // strides over d (a has stride 0 on the summed dims)
    i64 a_ev, b_ev;         // event strides
    T cadd;
    T* out;
    i64 n_out, n_red;
    int vec2;               // a / b rows are contiguous and 2-element aligned
    // side output (plan.py Planner.fuse_side_factors): qout[o] = sum_d log N(a[o, d]; ql[o, d], qs[o, d]) -- a Gaussian
    // factor of the same rows (logQ(z) of a mean-field Q), evaluated from the registers that hold a[o, :] anyway
    int side;
    Opnd ql, qs;
    i64 ql_ev, qs_ev;
    T* qout;
};

// log Bernoulli(y; logits = x) = y x - softplus(x) = -(1 - y) x + min(x, 0) - log1p(exp(-|x|)).  fp32: the
// log1p(exp(.)) tail on the SFU (ex2 / lg2, two MUFU instead of ~45 libdevice instructions); its absolute error is
// <= 1e-7 on a term of magnitude |x| that is then summed over the plate.  fp64 keeps the libdevice path.
__device__ __forceinline__ float bern_logits_fast(float y, float x) {
    float t, l;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-fabsf(x) * 1.4426950408889634f));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + t));
    return -((1.0f - y) * x) + fminf(x, 0.0f) - l * 0.6931471805599453f;
}
__device__ __forceinline__ double bern_logits_fast(double y, double x) { return bern_logits_lp(y, x); }

template <typename T, int D>
__global__ void __launch_bounds__(256) bern_dot_sum_kernel(const __grid_constant__ BernDotParams<T> p) {
    typedef typename Pair2<T>::type P2;
    int idx[AB_MAXD];
    const T* A = (const T*)p.a.ptr;
    const T* B = (const T*)p.b.ptr;
    const T* Y = (const T*)p.y.ptr;
    const int nred = p.d.nd - p.d.n_a;
    for (i64 o = (i64)blockIdx.x * blockDim.x + threadIdx.x; o < p.n_out; o += (i64)gridDim.x * blockDim.x) {
        unravel(o, p.d, 0, p.d.n_a, idx);
        const i64 ab = dot_stride(p.a, idx, 0, p.d.n_a), bb = dot_stride(p.b, idx, 0, p.d.n_a), yb = dot_stride(p.y, idx, 0, p.d.n_a);
        T ar[D];
        if (p.vec2) {
#pragma unroll
            for (int q = 0; q < D / 2; ++q) { const P2 v = *reinterpret_cast<const P2*>(A + ab + 2 * q); ar[2 * q] = v.x; ar[2 * q + 1] = v.y; }
        } else {
#pragma unroll
            for (int dd = 0; dd < D; ++dd) ar[dd] = A[ab + dd * p.a_ev];
        }
        if (p.side) {
            const T* L = (const T*)p.ql.ptr;
            const T* S = (const T*)p.qs.ptr;
            const i64 lb = dot_stride(p.ql, idx, 0, p.d.n_a), sb = dot_stride(p.qs, idx, 0, p.d.n_a);
            T q = T(0);
#pragma unroll
            for (int dd = 0; dd < D; ++dd) q += normal_lp(ar[dd], L[lb + dd * p.ql_ev], S[sb + dd * p.qs_ev]);
            p.qout[o] = q;
        }
        T acc = T(0);
        for (i64 n = 0; n < p.n_red; ++n) {
            i64 bo = bb, yo = yb;
            if (nred == 1) { bo += n * p.b.stride[p.d.n_a]; yo += n * p.y.stride[p.d.n_a]; }
            else {
                unravel(n, p.d, p.d.n_a, p.d.nd, idx);
                bo += dot_stride(p.b, idx, p.d.n_a, p.d.nd); yo += dot_stride(p.y, idx, p.d.n_a, p.d.nd);
            }
            T l0 = T(0), l1 = T(0);
            if (p.vec2) {
#pragma unroll
                for (int q = 0; q < D / 2; ++q) {
                    const P2 v = *reinterpret_cast<const P2*>(B + bo + 2 * q);
                    l0 += ar[2 * q] * v.x; l1 += ar[2 * q + 1] * v.y;
                }
            } else {
#pragma unroll
                for (int dd = 0; dd < D; ++dd) l0 += ar[dd] * B[bo + dd * p.b_ev];
            }
            acc += bern_logits_fast(Y[yo], l0 + l1);
        }
        p.out[o] = acc + p.cadd;
    }
}

// Shared-memory staged variant for the common layout: one summed plate dim whose b rows ([n, d] block of one "user")
// are contiguous, and an innermost kept dim (the K axis of the sample) that neither b nor y carries.  A CTA takes UPC
// users at a time: their b blocks are read from HBM ONCE with coalesced loads (the thread-per-output kernel above
// re-reads every 8-byte piece through the LSU from all K threads of a user: LSU-bound at ~12 % of the HBM roofline on
// B200), padded to 16-byte rows in shared memory, and every thread walks them with broadcast LDS.128.  A thread owns
// KPT consecutive k of one user.  KPT = 2 halves the shared-memory reads per cell but was measured SLOWER at cfg-5
// (62 us against 45: three resident CTAs instead of five; ncu shows the kernel issue- and latency-bound, 58 % issue
// utilisation with 7 of 16 warp slots filled, not LDS-bound), so it is opt-in (ALAN_B200_BDS_KPT=2).
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" :: "r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}

#define BDS_THREADS 256
#define BDS_SMEM_BYTES (40 * 1024)
#define BDS_SMEM_BYTES2 (72 * 1024)
#define BDS_MAX_UPC 32
template <typename T, int D, int KPT>
__global__ void __launch_bounds__(BDS_THREADS) bern_dot_sum_smem_kernel(const __grid_constant__ BernDotParams<T> p, int Kin, int UPC, int NCH) {
    constexpr int DP = (D + 3) & ~3;                                   // row pitch (elements): 16-byte rows for float
    extern __shared__ __align__(16) unsigned char bds_smem[];
    T* xs = reinterpret_cast<T*>(bds_smem);                            // [UPC][NCH][DP]
    T* ys = xs + (size_t)UPC * NCH * DP;                               // [UPC][NCH]
    T* ql_s = ys + (size_t)UPC * NCH;                                  // [UPC][DP]   side factor: loc
    T* qi_s = ql_s + (size_t)UPC * DP;                                 // [UPC][DP]   1 / scale
    T* qc_s = qi_s + (size_t)UPC * DP;                                 // [UPC]       -(sum log scale) - D/2 log 2 pi
    const T* A = (const T*)p.a.ptr;
    const T* B = (const T*)p.b.ptr;
    const T* Y = (const T*)p.y.ptr;
    const i64 n_users = p.n_out / Kin;
    const int N = (int)p.n_red;
    const i64 bn = p.b.stride[p.d.n_a], yn = p.y.stride[p.d.n_a];
    const int TPU = (Kin + KPT - 1) / KPT;                             // threads per user
    const int tu = threadIdx.x / TPU, tk = (threadIdx.x - tu * TPU) * KPT;   // (user slot, first k) of this thread
    const bool worker = tu < UPC;
    int idx[AB_MAXD];
    __shared__ i64 s_bb[BDS_MAX_UPC], s_yb[BDS_MAX_UPC], s_lb[BDS_MAX_UPC], s_sb[BDS_MAX_UPC];   // per user slot: base offsets
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (i64 u0 = (i64)blockIdx.x * UPC; u0 < n_users; u0 += (i64)gridDim.x * UPC) {
        const i64 u = u0 + tu;
        const bool live = worker && u < n_users;
        T ar[KPT][D];
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            if (live && tk + j < Kin) {
                unravel(u * Kin + tk + j, p.d, 0, p.d.n_a, idx);
                const i64 ab = dot_stride(p.a, idx, 0, p.d.n_a);
#pragma unroll
                for (int dd = 0; dd < D; ++dd) ar[j][dd] = A[ab + dd * p.a_ev];
            } else {
#pragma unroll
                for (int dd = 0; dd < D; ++dd) ar[j][dd] = T(0);
            }
        }
        __syncthreads();                                               // the previous users have been consumed
        if (threadIdx.x < UPC) {
            const i64 uu = u0 + threadIdx.x;
            i64 bb = -1, yb = -1, lb = 0, sb = 0;
            if (uu < n_users) {
                unravel(uu * Kin, p.d, 0, p.d.n_a, idx);
                bb = dot_stride(p.b, idx, 0, p.d.n_a);
                yb = dot_stride(p.y, idx, 0, p.d.n_a);
                if (p.side) { lb = dot_stride(p.ql, idx, 0, p.d.n_a); sb = dot_stride(p.qs, idx, 0, p.d.n_a); }
            }
            s_bb[threadIdx.x] = bb; s_yb[threadIdx.x] = yb; s_lb[threadIdx.x] = lb; s_sb[threadIdx.x] = sb;
        }
        T acc0[KPT], acc1[KPT];
#pragma unroll
        for (int j = 0; j < KPT; ++j) acc0[j] = acc1[j] = T(0);
        for (int n0 = 0; n0 < N; n0 += NCH) {
            const int nc = min(NCH, N - n0);
            __syncthreads();                                           // offsets visible; the previous chunk has been consumed
            // stage: one warp per user slot, consecutive lanes read consecutive (pairs of) elements of the [nc, D] block
            for (int us = warp; us < UPC; us += BDS_THREADS / 32) {
                const i64 bb = s_bb[us], yb = s_yb[us];
                if (bb < 0) continue;
                const T* src = B + bb + (i64)n0 * bn;
                T* dst = xs + (size_t)us * NCH * DP;
                // cp.async (LDGSTS): every copy of the lane is in flight at once, nothing passes through registers
                if (p.vec2) {
                    constexpr int H = D / 2 > 0 ? D / 2 : 1;
                    for (int i = lane; i < nc * H; i += 32) {
                        const int n = i / H, q = i - n * H;
                        cp_async<2 * sizeof(T)>(dst + (size_t)n * DP + 2 * q, src + 2 * i);
                    }
                } else {
                    for (int i = lane; i < nc * D; i += 32) {
                        const int n = i / D, dd = i - n * D;
                        cp_async<sizeof(T)>(dst + (size_t)n * DP + dd, src + i);
                    }
                }
                for (int n = lane; n < nc; n += 32) ys[us * NCH + n] = Y[yb + (i64)(n0 + n) * yn];
                if (p.side && n0 == 0) {
                    // the user's loc, 1 / scale and log-normaliser, once per user (not once per (user, k, d))
                    const T* L = (const T*)p.ql.ptr;
                    const T* S = (const T*)p.qs.ptr;
                    const i64 lb = s_lb[us], sb = s_sb[us];
                    T ls = T(0);
                    for (int dd = lane; dd < D; dd += 32) {
                        const T sv = S[sb + dd * p.qs_ev];
                        ql_s[us * DP + dd] = L[lb + dd * p.ql_ev];
                        qi_s[us * DP + dd] = T(1) / sv;
                        ls += ab_log(sv);
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
                    if (lane == 0) qc_s[us] = -ls - T(D) * T(HALF_LOG_2PI);
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
            if (live) {
                const T* xr = xs + (size_t)tu * NCH * DP;
                const T* yr = ys + tu * NCH;
#pragma unroll 2
                for (int n = 0; n < nc; ++n) {
                    T row[DP];
                    if (sizeof(T) == 4) {
#pragma unroll
                        for (int q = 0; q < DP / 4; ++q)
                            *reinterpret_cast<float4*>(reinterpret_cast<float*>(row) + 4 * q) =
                                *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(xr) + (size_t)n * DP + 4 * q);
                    } else {
#pragma unroll
                        for (int dd = 0; dd < D; ++dd) row[dd] = xr[(size_t)n * DP + dd];
                    }
                    const T yv = yr[n];
                    // packed pairs (FFMA2 in fp32): half the FMA issue slots, two independent chains per k
                    typedef typename Pair2<T>::type P2;
#pragma unroll
                    for (int j = 0; j < KPT; ++j) {
                        P2 l2 = mk2(T(0), T(0)), l3 = mk2(T(0), T(0));
#pragma unroll
                        for (int dd = 0; dd + 1 < D; dd += 2) {
                            if ((dd >> 1) & 1) l3 = fma2(mk2(ar[j][dd], ar[j][dd + 1]), mk2(row[dd], row[dd + 1]), l3);
                            else l2 = fma2(mk2(ar[j][dd], ar[j][dd + 1]), mk2(row[dd], row[dd + 1]), l2);
                        }
                        T l0 = (l2.x + l3.x), l1 = (l2.y + l3.y);
                        if (D & 1) l0 += ar[j][D - 1] * row[D - 1];
                        const T lp = bern_logits_fast(yv, l0 + l1);
                        if ((n0 + n) & 1) acc1[j] += lp; else acc0[j] += lp;      // by the row's own parity: the order does not depend on NCH
                    }
                }
            }
        }
        if (live) {
#pragma unroll
            for (int j = 0; j < KPT; ++j) {
                if (tk + j >= Kin) continue;
                p.out[u * Kin + tk + j] = (acc0[j] + acc1[j]) + p.cadd;
                if (p.side) {
                    T q0 = T(0), q1 = T(0);
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) {
                        const T tt = (ar[j][dd] - ql_s[tu * DP + dd]) * qi_s[tu * DP + dd];
                        if (dd & 1) q1 = fma(tt, tt, q1); else q0 = fma(tt, tt, q0);
                    }
                    p.qout[u * Kin + tk + j] = qc_s[tu] - T(0.5) * (q0 + q1);
                }
            }
        }
    }
}

template <typename T, int D, int KPT>
static void launch_bern_dot_smem(const BernDotParams<T>& p, int Kin, int UPC, int NCH, size_t smem, int blocks, cudaStream_t stream) {
    static const cudaError_t attr = cudaFuncSetAttribute(bern_dot_sum_smem_kernel<T, D, KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         BDS_SMEM_BYTES2 + 8 * 1024);       // once per process and instantiation
    (void)attr;
    bern_dot_sum_smem_kernel<T, D, KPT><<<blocks, BDS_THREADS, smem, stream>>>(p, Kin, UPC, NCH);
}

template <typename T, int D>
static void launch_bern_dot_D(BernDotParams<T> p, cudaStream_t stream, int sm_count) {
    bool v2 = (D % 2 == 0) && p.a_ev == 1 && p.b_ev == 1 && ((uintptr_t)p.a.ptr % (2 * sizeof(T)) == 0) &&
              ((uintptr_t)p.b.ptr % (2 * sizeof(T)) == 0);
    for (int k = 0; k < p.d.nd && v2; ++k) v2 = (p.a.stride[k] % 2 == 0) && (p.b.stride[k] % 2 == 0);
    p.vec2 = v2 ? 1 : 0;
    // staged variant: one summed dim with contiguous b rows, innermost kept dim absent from b and y (and from the side
    // factor's loc / scale), out contiguous
    const int nred = p.d.nd - p.d.n_a, ka = p.d.n_a - 1;
    if (nred == 1 && p.d.n_a >= 1 && p.b_ev == 1 && p.b.stride[p.d.n_a] == D && p.b.stride[ka] == 0 && p.y.stride[ka] == 0 &&
        (!p.side || (p.ql.stride[ka] == 0 && p.qs.stride[ka] == 0)) &&
        p.d.size[ka] <= BDS_THREADS && p.d.size[ka] >= 8 && p.n_red >= 2 && p.n_out / p.d.size[ka] >= 64) {
        constexpr int DP = (D + 3) & ~3;
        const int Kin = p.d.size[ka];
        const char* e_kpt = getenv("ALAN_B200_BDS_KPT");                 // tuning aids
        const char* e_cap = getenv("ALAN_B200_BDS_CAP");
        const char* e_smem = getenv("ALAN_B200_BDS_SMEM_KB");
        const int KPT = (Kin >= 16 && sizeof(T) == 4 && e_kpt && atoi(e_kpt) == 2) ? 2 : 1;
        const int TPU = (Kin + KPT - 1) / KPT;
        const int UPC = BDS_THREADS / TPU;
        // 64-bit staging loads: even D, b block bases and the step between summed rows 2-element aligned
        bool sv2 = (D % 2 == 0) && ((uintptr_t)p.b.ptr % (2 * sizeof(T)) == 0);
        for (int k = 0; k < p.d.nd && sv2; ++k) sv2 = (p.b.stride[k] % 2 == 0);
        p.vec2 = sv2 ? 1 : 0;
        const size_t side_bytes = (size_t)UPC * (2 * DP + 1) * sizeof(T);
        const size_t budget = e_smem ? (size_t)atoi(e_smem) * 1024 : (KPT == 2 ? BDS_SMEM_BYTES2 : BDS_SMEM_BYTES);
        int NCH = (int)(budget / ((size_t)UPC * (DP + 1) * sizeof(T)));
        if (NCH > p.n_red) NCH = (int)p.n_red;
        if (NCH >= 2 && UPC <= BDS_MAX_UPC) {
            const size_t smem = (size_t)UPC * NCH * (DP + 1) * sizeof(T) + side_bytes;
            const i64 n_users = p.n_out / Kin;
            i64 blocks = (n_users + UPC - 1) / UPC, cap = (i64)sm_count * (e_cap ? atoi(e_cap) : (KPT == 2 ? 3 : 5));
            if (blocks > cap) blocks = cap;             // (equal trip counts per CTA measured slower: 50 us against 45 at cfg-5)
            if (KPT == 2) launch_bern_dot_smem<T, D, 2>(p, Kin, UPC, NCH, smem, (int)blocks, stream);
            else launch_bern_dot_smem<T, D, 1>(p, Kin, UPC, NCH, smem, (int)blocks, stream);
            return;
        }
    }
    i64 blocks = (p.n_out + 255) / 256, cap = (i64)sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    bern_dot_sum_kernel<T, D><<<(int)blocks, 256, 0, stream>>>(p);
}

template <typename T>
static int launch_bern_dot(const BernDotParams<T>& p, int D, cudaStream_t stream, int sm_count) {
    switch (D) {
        case 1: launch_bern_dot_D<T, 1>(p, stream, sm_count); break;
        case 2: launch_bern_dot_D<T, 2>(p, stream, sm_count); break;
        case 3: launch_bern_dot_D<T, 3>(p, stream, sm_count); break;
        case 4: launch_bern_dot_D<T, 4>(p, stream, sm_count); break;
        case 6: launch_bern_dot_D<T, 6>(p, stream, sm_count); break;
        case 8: launch_bern_dot_D<T, 8>(p, stream, sm_count); break;
        case 12: launch_bern_dot_D<T, 12>(p, stream, sm_count); break;
        case 16: launch_bern_dot_D<T, 16>(p, stream, sm_count); break;
        case 18: launch_bern_dot_D<T, 18>(p, stream, sm_count); break;
        case 24: launch_bern_dot_D<T, 24>(p, stream, sm_count); break;
        case 32: launch_bern_dot_D<T, 32>(p, stream, sm_count); break;
        default: return 1;
    }
    return 0;
}


// one tile row -> NPU register pairs with 16-byte shared-memory loads (rows are 16-byte aligned)
template <int NPU>
__device__ __forceinline__ void load_row_pairs(const float* row, float2* t2) {
#pragma unroll
    for (int q4 = 0; q4 < (NPU + 1) / 2; ++q4) {
        const float4 t4 = *reinterpret_cast<const float4*>(row + 4 * q4);
        t2[2 * q4] = make_float2(t4.x, t4.y);
        if (2 * q4 + 1 < NPU) t2[2 * q4 + 1] = make_float2(t4.z, t4.w);
    }
}
template <int NPU>
__device__ __forceinline__ void load_row_pairs(const double* row, double2* t2) {
#pragma unroll
    for (int q = 0; q < NPU; ++q) t2[q] = *reinterpret_cast<const double2*>(row + 2 * q);
}

#define FL2_WARPS 4
// fan columns (f) owned by one lane: the lane keeps w[f, 0..D) for FPL columns in registers
// (FPL * D values), so every squared residual read from shared memory feeds FPL FMAs.  Measured on
// B200 (tools/microbench.cu): broadcast LDS.128 delivers ~59 floats/clk/SM against 128 FMA/clk/SM,
// so FPL >= 3 is needed to be FMA- rather than LDS-bound; 6 leaves headroom.
template <typename T, int D> struct FanTile {
    // row layout: [0, D) squared residuals, [D] the row's small-factor sum (bias), [D+1] the constant 1;
    // the matching w "columns" hold 1 and -c[f], so the whole S = b - c - sum_d T w comes out of the
    // FFMA2 chain with a zero initial accumulator.
    static constexpr int DP = (D + 2 + 3) & ~3;             // shared-memory row pitch (16-byte rows)
    static constexpr int NPU = (D + 2 + 1) / 2;             // register pairs per row actually used
    static constexpr int W32 = NPU * 2 * (int)(sizeof(T) / 4);
    static constexpr int FPL_ = 120 / W32;
    static constexpr int FPL = FPL_ >= 6 ? 6 : (FPL_ >= 4 ? 4 : (FPL_ >= 3 ? 3 : (FPL_ >= 2 ? 2 : 1)));
};

struct FanLse2Cfg {
    int FG;          // lanes per rho (each owns FPL fan columns)
    int RPW;         // rho processed concurrently by one warp (RPW * FG <= 32)
    int KP;          // kappa padded to a multiple of RB
    int FC;          // fan columns per chunk = FG * FPL
    int n_chunks;
    int vec2;        // value / loc rows are 2-element aligned and contiguous: 64-bit loads
    int tile_pitch;  // elements between the tiles of one warp (== 4 mod 32: distinct banks per tile)
    int warp_elems;  // shared-memory elements (T) per warp
    int off_words;   // i64 words per warp for the per-rho offsets
};

// One warp owns RPW rho at a time.  It builds their tiles T[kappa][d] = (v - l)^2 (plus the bias and
// constant slots) in its shared-memory slice with coalesced 64-bit reads of v, then lane (rs, fg) walks
// the kappa rows of rho rs for its FPL fan columns: each row is DP/4 LDS.128 feeding FPL * NPU FFMA2,
// and the LSE over kappa is a blocked online max/rescale in registers (RB rows per block) -- no
// cross-lane traffic at all in the forward pass.  The adjoint recomputes S the same way and reduces
// the softmax weights over the fan axis through a per-warp shared-memory transpose in a fixed order.
template <typename T, int D, int RB, bool BWD>
__global__ void __launch_bounds__(FL2_WARPS * 32, 2) fan_lse2_kernel(const __grid_constant__ FanLseParams<T> p,
                                                                    const __grid_constant__ FanLse2Cfg cfg) {
    typedef FanTile<T, D> FT;
    typedef typename Pair2<T>::type P2;
    constexpr int DP = FT::DP, NPU = FT::NPU, FPL = FT::FPL;
    extern __shared__ __align__(16) unsigned char fan_smem[];
    const int FG = cfg.FG, RPW = cfg.RPW, KP = cfg.KP, FC = cfg.FC, TP = cfg.tile_pitch;
    const int KPo = KP | 1;
    T* Wt = (T*)fan_smem;                       // [FC][DP]  -w * log2e of the current fan chunk, then (1, -c * log2e)
    const int warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* Tt = Wt + FC * DP + (size_t)warp_in_cta * cfg.warp_elems;     // [RPW] tiles of [KP][DP], pitch TP
    T* Pp = Tt + RPW * TP;                      // bwd: [32][KPo] per-lane partial weights
    i64* offs = (i64*)(fan_smem + (((size_t)(FC * DP + FL2_WARPS * cfg.warp_elems) * sizeof(T) + 15) & ~(size_t)15))
                + (size_t)warp_in_cta * cfg.off_words;          // [RPW][3 + nb]
    const int NO = 4 + p.nb;
    const T LS = FastExp<T>::scale();
    const int rs = lane / FG, fg = lane - rs * FG;
    const bool lane_on = rs < RPW;
    const int Kk = p.Kk;

    // padding rows / columns of the per-warp tiles never change: T = 0, bias = -inf, constant slot = 1
    for (int e = lane; e < RPW * TP; e += 32) Tt[e] = T(0);
    __syncwarp();
    for (int e = lane; e < RPW * KP; e += 32) {
        int t = e / KP, k = e - t * KP;
        Tt[t * TP + k * DP + D] = neg_inf<T>();
        Tt[t * TP + k * DP + D + 1] = T(1);
    }
    __syncwarp();

    const unsigned n_rho = (unsigned)p.n_rho;
    const unsigned n_groups = (n_rho + RPW - 1) / RPW;
    const unsigned warp = blockIdx.x * FL2_WARPS + warp_in_cta;
    const unsigned nwarps = gridDim.x * FL2_WARPS;
    // 32-bit strides of the tile walk (the host checks that they fit)
    const int vk = (int)p.v_k, lk = (int)p.l_k, vev = (int)p.v_ev, lev = (int)p.l_ev;

    for (int chunk = 0; chunk < cfg.n_chunks; ++chunk) {
        const int f_base = chunk * FC;
        __syncthreads();
        for (int i = threadIdx.x; i < FC * DP; i += blockDim.x) {
            int fl = i / DP, d = i - fl * DP, f = f_base + fl;
            T w = T(0);
            if (f < p.F) {
                if (d < D) { T sc = p.s[f * p.s_f + d * p.s_ev]; w = -LS / (T(2) * (sc * sc)); }
                else if (d == D) w = T(1);
                else if (d == D + 1) {
                    T c = T(0);
                    for (int dd = 0; dd < D; ++dd) c += ab_log(p.s[f * p.s_f + dd * p.s_ev]);
                    w = -(c + T(D) * T(HALF_LOG_2PI)) * LS;
                }
            }
            Wt[i] = w;
        }
        __syncthreads();
        // this lane's fan columns stay in registers for the whole chunk
        P2 W2[FPL][NPU];
#pragma unroll
        for (int j = 0; j < FPL; ++j) {
            const T* wrow = Wt + (lane_on ? fg * FPL + j : 0) * DP;
#pragma unroll
            for (int q = 0; q < NPU; ++q) W2[j][q] = mk2(wrow[2 * q], wrow[2 * q + 1]);
        }

        for (unsigned g = warp; g < n_groups; g += nwarps) {
            const unsigned rho0 = g * RPW;
            // per-rho offsets: lane t decodes rho0 + t
            if (lane < RPW) {
                i64 voff = 0, loff = 0, ooff = 0, goff = 0;
                i64 boff[AB_MAXL];
#pragma unroll
                for (int i = 0; i < AB_MAXL; ++i) boff[i] = 0;
                unsigned lin = rho0 + lane < n_rho ? rho0 + lane : n_rho - 1;
#pragma unroll 1
                for (int k = p.rd.nd - 1; k >= 0; --k) {
                    unsigned sz = (unsigned)p.rd.size[k];
                    unsigned q = lin / sz;
                    unsigned ix = lin - q * sz;
                    lin = q;
                    voff += ix * p.vstride[k]; loff += ix * p.lstride[k]; ooff += ix * p.ostride[k];
                    if (BWD) goff += ix * p.gstride[k];
#pragma unroll
                    for (int i = 0; i < AB_MAXL; ++i) if (i < p.nb) boff[i] += ix * p.bstride[i][k];
                }
                i64* o = offs + lane * NO;
                o[0] = voff; o[1] = loff; o[2] = ooff; o[3] = goff;
#pragma unroll
                for (int i = 0; i < AB_MAXL; ++i) if (i < p.nb) o[4 + i] = boff[i];
            }
            __syncwarp();
            // squared residual tiles: T[t][k][d] = (v - l)^2; (k, d) advance incrementally, 32-bit
            for (int t = 0; t < RPW; ++t) {
                const T* vp = p.v + offs[t * NO];
                const T* lp = p.l + offs[t * NO + 1];
                T* tile = Tt + t * TP;
                if (cfg.vec2) {
                    constexpr int DV = D / 2 > 0 ? D / 2 : 1;
                    constexpr int SK = 32 / DV, SD = 32 % DV;
                    int k = lane / DV, dv = lane - k * DV;
                    int vo = k * vk + 2 * dv, lo = k * lk + 2 * dv, to = k * DP + 2 * dv;
                    const int n_pairs = Kk * DV;
                    for (int e = lane; e < n_pairs; e += 32) {
                        const P2 vv = *reinterpret_cast<const P2*>(vp + vo);
                        const P2 ll = *reinterpret_cast<const P2*>(lp + lo);
                        const T d0 = vv.x - ll.x, d1 = vv.y - ll.y;
                        *reinterpret_cast<P2*>(tile + to) = mk2(d0 * d0, d1 * d1);
                        dv += SD; vo += SK * vk + 2 * SD; lo += SK * lk + 2 * SD; to += SK * DP + 2 * SD;
                        if (dv >= DV) { dv -= DV; vo += vk - 2 * DV; lo += lk - 2 * DV; to += DP - 2 * DV; }
                    }
                } else {
                    constexpr int SK = 32 / D, SD = 32 % D;
                    int k = lane / D, d = lane - k * D;
                    int vo = k * vk + d * vev, lo = k * lk + d * lev, to = k * DP + d;
                    const int n_el = Kk * D;
                    for (int e = lane; e < n_el; e += 32) {
                        const T df = vp[vo] - lp[lo];
                        tile[to] = df * df;
                        d += SD; vo += SK * vk + SD * vev; lo += SK * lk + SD * lev; to += SK * DP + SD;
                        if (d >= D) { d -= D; vo += vk - D * vev; lo += lk - D * lev; to += DP - D; }
                    }
                }
            }
            // bias slot: sum of the small factors of (rho, kappa), pre-scaled by log2e
            for (int t = 0; t < RPW; ++t) {
                for (int k = lane; k < Kk; k += 32) {
                    T b = T(0);
#pragma unroll
                    for (int i = 0; i < AB_MAXL; ++i) if (i < p.nb) b += p.bcoeff[i] * p.b[i][offs[t * NO + 4 + i] + k * p.b_k[i]];
                    Tt[t * TP + k * DP + D] = b * LS;
                }
            }
            __syncwarp();

            const bool live = lane_on && rho0 + rs < n_rho;
            const int rsc = lane_on ? rs : 0;
            const T* tile = Tt + rsc * TP;
            const i64 ooff = offs[rsc * NO + 2];
            if (!BWD) {
                T m[FPL], sum[FPL];
#pragma unroll
                for (int j = 0; j < FPL; ++j) { m[j] = FastExp<T>::lowest(); sum[j] = T(0); }
#pragma unroll 1
                for (int k0 = 0; k0 < KP; k0 += RB) {
                    T s[RB][FPL];
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        P2 t2[NPU];
                        load_row_pairs<NPU>(tile + (k0 + r) * DP, t2);
#pragma unroll
                        for (int j = 0; j < FPL; ++j) {
                            P2 acc = mk2(T(0), T(0));
#pragma unroll
                            for (int q = 0; q < NPU; ++q) acc = fma2(t2[q], W2[j][q], acc);
                            s[r][j] = acc.x + acc.y;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < FPL; ++j) {
                        T mn = m[j];
#pragma unroll
                        for (int r = 0; r < RB; ++r) mn = ab_max(mn, s[r][j]);
                        T a = sum[j] * FastExp<T>::ex(m[j] - mn);
#pragma unroll
                        for (int r = 0; r < RB; ++r) a += FastExp<T>::ex(s[r][j] - mn);
                        sum[j] = a; m[j] = mn;
                    }
                }
                if (live) {
#pragma unroll
                    for (int j = 0; j < FPL; ++j) {
                        const int f = f_base + fg * FPL + j;
                        if (f < p.F) p.out[ooff + (i64)f * p.o_f] = ab_log(sum[j] + Eps<T>::v()) + m[j] * FastExp<T>::unscale() + p.cadd;
                    }
                }
            } else {
                // fold -lse into the constant column: exponent = S - lse comes straight out of the chain
                T gz[FPL];
                P2 wl[FPL];
#pragma unroll
                for (int j = 0; j < FPL; ++j) {
                    const int f = f_base + fg * FPL + j;
                    const bool on = live && f < p.F;
                    const T lz = on ? (p.lse[ooff + (i64)f * p.o_f] - p.cadd) * LS : -neg_inf<T>();   // idle columns: weight 0
                    gz[j] = on ? p.gout[offs[rsc * NO + 3] + (i64)f * p.g_f] : T(0);
                    wl[j] = W2[j][NPU - 1];
                    if (((D + 1) & 1) == 0) wl[j].x -= lz; else wl[j].y -= lz;           // slot D + 1 holds -c
                }
                T* mine = Pp + lane * KPo;
#pragma unroll 1
                for (int k0 = 0; k0 < KP; k0 += RB) {
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        P2 t2[NPU];
                        load_row_pairs<NPU>(tile + (k0 + r) * DP, t2);
                        T w = T(0);
#pragma unroll
                        for (int j = 0; j < FPL; ++j) {
                            P2 acc = mk2(T(0), T(0));
#pragma unroll
                            for (int q = 0; q < NPU - 1; ++q) acc = fma2(t2[q], W2[j][q], acc);
                            acc = fma2(t2[NPU - 1], wl[j], acc);
                            w += gz[j] * FastExp<T>::ex(acc.x + acc.y);          // 0 on padding rows (bias = -inf)
                        }
                        mine[k0 + r] = w;
                    }
                }
                __syncwarp();
                // fixed-order sum over the FG lanes of each rho; one (rho, kappa) per thread, coalesced store
                for (int t = 0; t < RPW; ++t) {
                    if (rho0 + t < n_rho) {
                        for (int k = lane; k < Kk; k += 32) {
                            T a = T(0);
                            for (int q = 0; q < FG; ++q) a += Pp[(t * FG + q) * KPo + k];
                            const i64 gi = (i64)(rho0 + t) * Kk + k;
                            if (chunk > 0) a += p.gS[gi];
                            p.gS[gi] = a;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
}

template <typename T, int D, int RB>
static int launch_fan_lse_DR(const FanLseParams<T>& p, bool bwd, cudaStream_t stream, int sm_count) {
    typedef FanTile<T, D> FT;
    FanLse2Cfg c;
    const int fgn = (p.F + FT::FPL - 1) / FT::FPL;
    c.FG = fgn < 32 ? fgn : 32;
    c.FC = c.FG * FT::FPL;
    c.n_chunks = (p.F + c.FC - 1) / c.FC;
    c.KP = (p.Kk + RB - 1) / RB * RB;
    c.tile_pitch = c.KP * FT::DP;
    while (c.tile_pitch % 32 != 4) c.tile_pitch += 4;
    const int KPo = c.KP | 1;
    const size_t tile_bytes = (size_t)c.tile_pitch * sizeof(T);
    int rpw = 32 / c.FG;
    const size_t budget = 18 * 1024;                     // per-warp tile budget: 4 warps x 2 CTAs per SM
    while (rpw > 1 && rpw * tile_bytes > budget) --rpw;
    if (tile_bytes > 44 * 1024) return 2;
    c.RPW = rpw;
    const i64 lim = (i64)1 << 30;
    if (p.v_k >= lim || p.l_k >= lim || p.v_ev >= lim || p.l_ev >= lim || (i64)p.Kk * (p.v_k > p.l_k ? p.v_k : p.l_k) >= lim) return 3;
    bool ev2 = (D % 2 == 0) && p.v_ev == 1 && p.l_ev == 1 && p.v_k % 2 == 0 && p.l_k % 2 == 0 &&
               ((uintptr_t)p.v % (2 * sizeof(T)) == 0) && ((uintptr_t)p.l % (2 * sizeof(T)) == 0);
    for (int k = 0; k < p.rd.nd && ev2; ++k) ev2 = (p.vstride[k] % 2 == 0) && (p.lstride[k] % 2 == 0);
    c.vec2 = ev2 ? 1 : 0;
    c.warp_elems = c.RPW * c.tile_pitch + (bwd ? ((32 * KPo + 3) & ~3) : 0);
    c.off_words = c.RPW * (4 + p.nb);
    size_t smem = (((size_t)(c.FC * FT::DP + FL2_WARPS * c.warp_elems) * sizeof(T) + 15) & ~(size_t)15)
                  + (size_t)FL2_WARPS * c.off_words * sizeof(i64);
    if (smem > 200 * 1024) return 2;
    const i64 n_groups = (p.n_rho + c.RPW - 1) / c.RPW;
    i64 blocks = (n_groups + FL2_WARPS - 1) / FL2_WARPS;
    int per_sm = (int)(220 * 1024 / (smem + 1024));
    if (per_sm > 2) per_sm = 2;
    if (per_sm < 1) per_sm = 1;
    i64 cap = (i64)sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (bwd) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(fan_lse2_kernel<T, D, RB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        fan_lse2_kernel<T, D, RB, true><<<(int)blocks, FL2_WARPS * 32, smem, stream>>>(p, c);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(fan_lse2_kernel<T, D, RB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        fan_lse2_kernel<T, D, RB, false><<<(int)blocks, FL2_WARPS * 32, smem, stream>>>(p, c);
    }
    return 0;
}

template <typename T, int D>
static int launch_fan_lse_D(const FanLseParams<T>& p, bool bwd, cudaStream_t stream, int sm_count) {
    // rows per online-LSE block: 5 when that pads kappa less than 4 does (K = 30 -> no padding)
    const int pad4 = (p.Kk + 3) / 4 * 4, pad5 = (p.Kk + 4) / 5 * 5;
    if (pad5 < pad4) return launch_fan_lse_DR<T, D, 5>(p, bwd, stream, sm_count);
    return launch_fan_lse_DR<T, D, 4>(p, bwd, stream, sm_count);
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
