// chain.cuh -- the Timeseries chain (reference utils.py:478-510 chain_logmmexp / chain_reduce / logmmexp, call site
// logpq.py:131-143) as a FEW launches instead of one per tree level.
//
// The reference reduces ms[T, K, K] by a binary tree of log-space matrix products, the odd tail of a level carried
// unreduced to the end of the next one; eps and the tree shape are observable at 1e-5, so the tree is reproduced
// exactly.  A segment of S = 2^L consecutive matrices that starts at a multiple of S pairs up identically in the
// global tree for L levels (only the global tail can be odd, and it is the tail of the LAST segment), so one CTA
// takes a segment through L levels with the matrices resident in shared memory -- each product costs a few hundred
// cycles instead of a launch plus global round trips -- and writes every level to the `levels` buffer on the way
// (the adjoint reads them).  T = 1000, K = 16: 125 CTAs x 3 levels, 16 CTAs x 3 levels, then 1 CTA x 4 levels and the
// final row-wise logsumexp in the same launch: 3 launches against 11.  The adjoint walks the same phases in reverse.
//   C[i,k] = log( sum_j exp(A[i,j]-a_i) exp(B[j,k]-b_k) + eps ) + a_i + b_k,  a_i = max_j A[i,j], b_k = max_j B[j,k]
#pragma once
#include "kernels.cuh"

#define CH_MAXL 6           // levels per phase (S <= 64)
#define CH_MAXG 32          // product groups per CTA

template <typename T>
struct ChainPhase {
    const T* X0;            // input level [outer][n0][K K]
    i64 n0;
    int L, S, K, G, GT;       // G groups of GT threads each work on one product at a time (GT = 32: a warp, no barriers)
    T* Y[CH_MAXL];          // output of local level l: [outer][nY[l]][K K]
    i64 nY[CH_MAXL];
    T* final_out;           // last phase only: out[outer][K] = logsumexp over the rows of the single remaining matrix
    // adjoint
    const T* gfinal;        // d/d out[outer][K] (last phase only)
    const T* out;           // forward out (last phase only)
    T* gX0;                 // adjoint of the input level
    T* gY[CH_MAXL];         // adjoint of local level l's output (scratch)
};

__device__ __forceinline__ void group_sync(int g, int G, int GT) {
    if (GT == 32) __syncwarp();
    else if (G == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" :: "r"(1 + g), "r"(GT) : "memory");
}

// one product by one group of CH_GT threads; A, B anywhere, C_s (shared, may be null) and C_g (global, may be null)
template <typename T>
__device__ void chain_product(const T* A, const T* B, T* C_s, T* C_g, T* EA, T* EB, T* am, T* bm, int K, int tid, int g, int G, int CH_GT) {
    const int KK = K * K;
    for (int e = tid; e < KK; e += CH_GT) { EA[e] = A[e]; EB[e] = B[e]; }
    group_sync(g, G, CH_GT);
    for (int i = tid; i < K; i += CH_GT) {
        T m = neg_inf<T>(), n = neg_inf<T>();
        for (int j = 0; j < K; ++j) { m = ab_max(m, EA[i * K + j]); n = ab_max(n, EB[j * K + i]); }
        am[i] = m; bm[i] = n;
    }
    group_sync(g, G, CH_GT);
    for (int e = tid; e < KK; e += CH_GT) {
        EA[e] = ab_exp(EA[e] - am[e / K]);
        EB[e] = ab_exp(EB[e] - bm[e % K]);
    }
    group_sync(g, G, CH_GT);
    for (int e = tid; e < KK; e += CH_GT) {
        const int i = e / K, k = e % K;
        T acc = T(0);
        for (int j = 0; j < K; ++j) acc += EA[i * K + j] * EB[j * K + k];
        const T v = ab_log(acc + Eps<T>::v()) + am[i] + bm[k];
        if (C_s) C_s[e] = v;
        if (C_g) C_g[e] = v;
    }
    group_sync(g, G, CH_GT);                                              // scratch free for the group's next product
}

template <typename T>
__global__ void __launch_bounds__(1024) chain_multi_fwd_kernel(const __grid_constant__ ChainPhase<T> p) {
    extern __shared__ __align__(16) unsigned char ch_smem[];
    const int K = p.K, KK = K * K, S = p.S, G = p.G, CH_GT = p.GT;
    T* buf0 = reinterpret_cast<T*>(ch_smem);                        // [S][KK]
    T* buf1 = buf0 + (size_t)S * KK;                                // [S/2][KK]
    T* scratch = buf1 + (size_t)(S / 2 > 0 ? S / 2 : 1) * KK;       // per group: EA, EB, am, bm
    const int g = threadIdx.x / CH_GT, tid = threadIdx.x % CH_GT;
    T* EA = scratch + (size_t)g * (2 * KK + 2 * K);
    T* EB = EA + KK;
    T* am = EB + KK;
    T* bm = am + K;
    const i64 c = blockIdx.x, outer = blockIdx.y;
    i64 m = p.n0 - c * S;
    if (m > S) m = S;
    const T* src_g = p.X0 + ((i64)outer * p.n0 + c * S) * KK;
    for (i64 e = threadIdx.x; e < m * KK; e += blockDim.x) buf0[e] = src_g[e];
    __syncthreads();
    T* src = buf0;
    T* dst = buf1;
    for (int l = 0; l < p.L; ++l) {
        const i64 np = m / 2, mo = np + (m & 1);
        T* Yl = p.Y[l] + ((i64)outer * p.nY[l] + c * (S >> (l + 1))) * KK;
        for (i64 q = g; q < np; q += G)
            chain_product<T>(src + 2 * q * KK, src + (2 * q + 1) * KK, dst + q * KK, Yl + q * KK, EA, EB, am, bm, K, tid, g, G, CH_GT);
        __syncthreads();
        if (m & 1) {                                               // carried tail
            for (int e = threadIdx.x; e < KK; e += blockDim.x) {
                const T v = src[(m - 1) * KK + e];
                dst[np * KK + e] = v;
                Yl[np * KK + e] = v;
            }
            __syncthreads();
        }
        T* t = src; src = dst; dst = t;
        m = mo;
    }
    if (p.final_out) {                                             // no eps: torch.logsumexp (logpq.py:139)
        for (int i = threadIdx.x; i < K; i += blockDim.x) {
            const T* x = src + i * K;
            T mx = neg_inf<T>();
            for (int k = 0; k < K; ++k) mx = ab_max(mx, x[k]);
            const T mm = (mx == neg_inf<T>()) ? T(0) : mx;
            T a = T(0);
            for (int k = 0; k < K; ++k) a += ab_exp(x[k] - mm);
            p.final_out[(i64)outer * K + i] = ab_log(a) + mm;
        }
    }
}

// adjoint of one product by one group: A, B, g (= d/dC) global -> gA, gB global.  Includes the path through the amax
// shifts (weight eps / (P + eps), split evenly among ties as torch.amax does).
template <typename T>
__device__ void chain_product_bwd(const T* A, const T* B, const T* gC, T* gA, T* gB, T* EA, T* EB, T* D, T* am, T* bm, T* ga, T* gb,
                                  int K, int tid, int g, int G, int CH_GT) {
    const int KK = K * K;
    for (int e = tid; e < KK; e += CH_GT) { EA[e] = A[e]; EB[e] = B[e]; }
    group_sync(g, G, CH_GT);
    for (int i = tid; i < K; i += CH_GT) {
        T m = neg_inf<T>(), n = neg_inf<T>();
        for (int j = 0; j < K; ++j) { m = ab_max(m, EA[i * K + j]); n = ab_max(n, EB[j * K + i]); }
        am[i] = m; bm[i] = n;
    }
    group_sync(g, G, CH_GT);
    for (int e = tid; e < KK; e += CH_GT) {
        EA[e] = ab_exp(EA[e] - am[e / K]);
        EB[e] = ab_exp(EB[e] - bm[e % K]);
    }
    group_sync(g, G, CH_GT);
    for (int e = tid; e < KK; e += CH_GT) {
        const int i = e / K, k = e % K;
        T acc = T(0);
        for (int j = 0; j < K; ++j) acc += EA[i * K + j] * EB[j * K + k];
        D[e] = gC[e] / (acc + Eps<T>::v());
    }
    group_sync(g, G, CH_GT);
    for (int i = tid; i < K; i += CH_GT) {
        T sa = T(0), sb = T(0);
        for (int k = 0; k < K; ++k) { sa += D[i * K + k]; sb += D[k * K + i]; }
        ga[i] = sa * Eps<T>::v();
        gb[i] = sb * Eps<T>::v();
    }
    group_sync(g, G, CH_GT);
    for (int e = tid; e < KK; e += CH_GT) {
        const int i = e / K, j = e % K;
        T s = T(0);
        for (int k = 0; k < K; ++k) s += D[i * K + k] * EB[j * K + k];
        T v = s * EA[e];
        if (A[e] == am[i]) { int ties = 0; for (int jj = 0; jj < K; ++jj) ties += (A[i * K + jj] == am[i]); v += ga[i] / T(ties); }
        gA[e] = v;
        const int jr = i, k = j;                                   // gB[i,j] viewed as B[j' = i, k = j]
        T s2 = T(0);
        for (int ii = 0; ii < K; ++ii) s2 += EA[ii * K + jr] * D[ii * K + k];
        T v2 = s2 * EB[e];
        if (B[e] == bm[k]) { int ties = 0; for (int jj = 0; jj < K; ++jj) ties += (B[jj * K + k] == bm[k]); v2 += gb[k] / T(ties); }
        gB[e] = v2;
    }
    group_sync(g, G, CH_GT);
}

template <typename T>
__global__ void __launch_bounds__(1024) chain_multi_bwd_kernel(const __grid_constant__ ChainPhase<T> p) {
    extern __shared__ __align__(16) unsigned char ch_smem[];
    const int K = p.K, KK = K * K, S = p.S, G = p.G, CH_GT = p.GT;
    const int g = threadIdx.x / CH_GT, tid = threadIdx.x % CH_GT;
    T* EA = reinterpret_cast<T*>(ch_smem) + (size_t)g * (3 * KK + 4 * K);
    T* EB = EA + KK;
    T* D = EB + KK;
    T* am = D + KK;
    T* bm = am + K;
    T* ga = bm + K;
    T* gb = ga + K;
    const i64 c = blockIdx.x, outer = blockIdx.y;
    i64 m0 = p.n0 - c * S;
    if (m0 > S) m0 = S;
    i64 ml[CH_MAXL + 1];                                           // local element count entering level l
    ml[0] = m0;
    for (int l = 0; l < p.L; ++l) ml[l + 1] = ml[l] / 2 + (ml[l] & 1);
    if (p.gfinal) {
        // adjoint of the final row-wise logsumexp of the single remaining matrix (top of the last phase)
        const T* Xl = p.L ? p.Y[p.L - 1] + (i64)outer * p.nY[p.L - 1] * KK : p.X0 + (i64)outer * p.n0 * KK;
        T* gXl = p.L ? p.gY[p.L - 1] + (i64)outer * p.nY[p.L - 1] * KK : p.gX0 + (i64)outer * p.n0 * KK;
        for (int e = threadIdx.x; e < KK; e += blockDim.x) {
            const int r = e / K;
            gXl[e] = p.gfinal[(i64)outer * K + r] * ab_exp(Xl[e] - p.out[(i64)outer * K + r]);
        }
        __threadfence_block();
        __syncthreads();
    }
    for (int l = p.L - 1; l >= 0; --l) {
        const i64 m = ml[l], np = m / 2;
        const T* Xin = l ? p.Y[l - 1] + ((i64)outer * p.nY[l - 1] + c * (S >> l)) * KK : p.X0 + ((i64)outer * p.n0 + c * S) * KK;
        T* gXin = l ? p.gY[l - 1] + ((i64)outer * p.nY[l - 1] + c * (S >> l)) * KK : p.gX0 + ((i64)outer * p.n0 + c * S) * KK;
        const T* gYl = p.gY[l] + ((i64)outer * p.nY[l] + c * (S >> (l + 1))) * KK;
        for (i64 q = g; q < np; q += G)
            chain_product_bwd<T>(Xin + 2 * q * KK, Xin + (2 * q + 1) * KK, gYl + q * KK, gXin + 2 * q * KK, gXin + (2 * q + 1) * KK,
                                 EA, EB, D, am, bm, ga, gb, K, tid, g, G, CH_GT);
        if (m & 1)
            for (int e = threadIdx.x; e < KK; e += blockDim.x) gXin[(m - 1) * KK + e] = gYl[np * KK + e];
        __threadfence_block();
        __syncthreads();
    }
}

// phases of the tree: {first level, element count entering it, levels in the phase, S}
struct ChainPlan { int n_phases; int l0[32]; i64 n0[32]; int L[32]; int S[32]; int G, GT; int Smax; size_t smem_f, smem_b; };

template <typename T>
static bool chain_plan(i64 Tn, i64 K, ChainPlan& cp) {
    const size_t KK = (size_t)K * K, budget = 180 * 1024;
    // A product is bound by instruction throughput inside its group (2 K^2 exp, K^2 log, K^3 FMA), so groups are wide
    // (256 threads; measured: one warp per product is 1.6x slower) and segments short: S = 2 G lets every product of
    // a segment's first level run at once, and many CTAs share the level.  The last phase (one CTA) takes up to 2 S.
    const int GT = KK <= 64 ? 64 : (KK <= 128 ? 128 : 256);
    cp.GT = GT;
    int S = 8;
    for (; S >= 2; S /= 2) {
        int G = S / 2;
        if (G > 1024 / GT) G = 1024 / GT;
        if (G < 1) G = 1;
        const size_t f = ((size_t)2 * S + S) * KK * sizeof(T) + (size_t)G * (2 * KK + 2 * K) * sizeof(T);   // sized for the last phase
        const size_t b = (size_t)G * (3 * KK + 4 * K) * sizeof(T);
        if (f <= budget && b <= budget) { cp.G = G; cp.smem_f = f; cp.smem_b = b; break; }
    }
    if (S < 2) return false;
    cp.Smax = S;
    cp.n_phases = 0;
    i64 n = Tn;
    int l = 0;
    const int Lmax = [&] { int x = 0; while ((1 << x) < S) ++x; return x; }();
    do {
        int L, Sp = S;
        if (n <= 2 * S) { L = 0; for (i64 q = n; q > 1; q = q / 2 + (q & 1)) ++L; Sp = 2 * S; }
        else L = Lmax;
        if (cp.n_phases >= 32) return false;
        cp.l0[cp.n_phases] = l; cp.n0[cp.n_phases] = n; cp.L[cp.n_phases] = L; cp.S[cp.n_phases] = Sp;
        ++cp.n_phases;
        for (int j = 0; j < L; ++j) n = n / 2 + (n & 1);
        l += L;
    } while (n > 1);
    return true;
}

// level offsets in the `levels` buffer (same layout as the per-level kernels: level l's output follows level l-1's)
static void chain_level_table(i64 outer, i64 Tn, i64 K, std::vector<i64>& n_in, std::vector<i64>& off) {
    i64 n = Tn, o = 0;
    while (n > 1) { const i64 no = n / 2 + (n & 1); n_in.push_back(n); off.push_back(o); o += outer * no * K * K; n = no; }
}

template <typename T>
static int launch_chain_fwd(const T* ms, T* levels, T* out, i64 outer, i64 Tn, i64 K, cudaStream_t st) {
    ChainPlan cp;
    if (outer > 65535 || !chain_plan<T>(Tn, K, cp)) return 1;
    std::vector<i64> n_in, off;
    chain_level_table(outer, Tn, K, n_in, off);
    static bool attr_done = false;
    if (!attr_done) { cudaFuncSetAttribute(chain_multi_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_done = true; }
    for (int ph = 0; ph < cp.n_phases; ++ph) {
        ChainPhase<T> p;
        memset(&p, 0, sizeof(p));
        const int l0 = cp.l0[ph];
        p.X0 = l0 ? levels + off[l0 - 1] : ms;
        p.n0 = cp.n0[ph]; p.L = cp.L[ph]; p.S = cp.S[ph]; p.K = (int)K; p.G = cp.G; p.GT = cp.GT;
        for (int j = 0; j < p.L; ++j) { p.Y[j] = levels + off[l0 + j]; p.nY[j] = n_in[l0 + j] / 2 + (n_in[l0 + j] & 1); }
        p.final_out = ph == cp.n_phases - 1 ? out : nullptr;
        const i64 nc = (p.n0 + p.S - 1) / p.S;
        chain_multi_fwd_kernel<T><<<dim3((unsigned)nc, (unsigned)outer), cp.GT * cp.G, cp.smem_f, st>>>(p);
    }
    return 0;
}

template <typename T>
static int launch_chain_bwd(const T* ms, const T* levels, const T* out, const T* gout, T* glevels, T* gms,
                            i64 outer, i64 Tn, i64 K, cudaStream_t st) {
    ChainPlan cp;
    if (outer > 65535 || !chain_plan<T>(Tn, K, cp)) return 1;
    std::vector<i64> n_in, off;
    chain_level_table(outer, Tn, K, n_in, off);
    static bool attr_done = false;
    if (!attr_done) { cudaFuncSetAttribute(chain_multi_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_done = true; }
    for (int ph = cp.n_phases - 1; ph >= 0; --ph) {
        ChainPhase<T> p;
        memset(&p, 0, sizeof(p));
        const int l0 = cp.l0[ph];
        p.X0 = l0 ? levels + off[l0 - 1] : ms;
        p.gX0 = l0 ? glevels + off[l0 - 1] : gms;
        p.n0 = cp.n0[ph]; p.L = cp.L[ph]; p.S = cp.S[ph]; p.K = (int)K; p.G = cp.G; p.GT = cp.GT;
        for (int j = 0; j < p.L; ++j) {
            p.Y[j] = const_cast<T*>(levels) + off[l0 + j]; p.gY[j] = glevels + off[l0 + j];
            p.nY[j] = n_in[l0 + j] / 2 + (n_in[l0 + j] & 1);
        }
        if (ph == cp.n_phases - 1) { p.gfinal = gout; p.out = out; }
        const i64 nc = (p.n0 + p.S - 1) / p.S;
        chain_multi_bwd_kernel<T><<<dim3((unsigned)nc, (unsigned)outer), cp.GT * cp.G, cp.smem_b, st>>>(p);
    }
    return 0;
}
