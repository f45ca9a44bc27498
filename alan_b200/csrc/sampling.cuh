// sampling.cuh -- kernels of the step BEFORE the logPQ path: ancestral sampling of Q with permuted / resampled parent
// particles (reference src/alan/Plate.py:93-143, dist.py:23-72, Sampler.py:85-169,
// Timeseries.py:89-123).  The draws themselves are factor-VM expressions over explicit base noise (loc + scale * eps,
// ...: ExprOp); what is specific to sampling lives here:
//   perm_kernel      per plate cell, the permutation of the K parent particles = argsort of K uniforms
//                    (PermutationSampler.perm, Sampler.py:143-148) or K uniform categorical picks (CategoricalSampler)
//   kgather_kernel   parent[cell, perm[cell, k], :] -> out[cell, k, :]   (Sampler.resample_scope, Sampler.py:85-116)
//   ts_sample_kernel the T-step recursion of a Timeseries in ONE launch (the reference loops over T in Python,
//                    Timeseries.py:101-121): thread (k, e) of the CTA that owns a plate cell evaluates the transition draw
//                    through the factor VM with `prev` read from shared memory, stores step t, and the CTA permutes the
//                    particles in shared memory (timeseries_perm[t]) before step t + 1
#pragma once
#include "kernels.cuh"

// u: [rows, K] float64 uniforms; perm: [rows, K] int64.  mode 0: perm[row, r] = index of the r-th smallest u (ties by
// index); mode 1: perm[row, k] = floor(u * K).  One warp per row, row staged in shared memory.
__global__ void __launch_bounds__(256) perm_kernel(const double* __restrict__ u, i64* __restrict__ perm, i64 rows, int K, int mode) {
    extern __shared__ double perm_smem[];                      // [8 warps][K]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* row_u = perm_smem + (size_t)warp * K;
    for (i64 row = (i64)blockIdx.x * 8 + warp; row < rows; row += (i64)gridDim.x * 8) {
        for (int k = lane; k < K; k += 32) row_u[k] = u[row * K + k];
        __syncwarp();
        for (int k = lane; k < K; k += 32) {
            if (mode == 1) {
                int pick = (int)(row_u[k] * K);
                perm[row * K + k] = pick < K ? pick : K - 1;
            } else {
                const double mine = row_u[k];
                int rank = 0;
                for (int j = 0; j < K; ++j) { const double o = row_u[j]; rank += (o < mine || (o == mine && j < k)) ? 1 : 0; }
                perm[row * K + rank] = k;
            }
        }
        __syncwarp();
    }
}

// x: [outer, K, inner], perm: [outer, K] -> out[o, k, i] = x[o, perm[o, k], i]
template <typename E>
__global__ void kgather_kernel(const E* __restrict__ x, const i64* __restrict__ perm, E* __restrict__ out,
                               i64 outer, i64 K, i64 inner) {
    const i64 total = outer * K * inner;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const i64 in = e % inner, ok = e / inner, o = ok / K;
        out[e] = x[(o * K + perm[ok]) * inner + in];
    }
}

// dst[idx . dstride] = src[idx . sstride] over a small index box (prediction: the posterior block pasted into the extended draw)
struct PasteParams { int nd; int size[AB_MAXD]; i64 ss[AB_MAXD], ds[AB_MAXD]; i64 total; };
template <typename E>
__global__ void paste_kernel(const E* __restrict__ src, E* __restrict__ dst, const __grid_constant__ PasteParams p) {
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < p.total; e += (i64)gridDim.x * blockDim.x) {
        i64 r = e, so = 0, d_o = 0;
        for (int k = p.nd - 1; k >= 0; --k) {
            const i64 q = r / p.size[k];
            const i64 i = r - q * p.size[k];
            so += i * p.ss[k]; d_o += i * p.ds[k];
            r = q;
        }
        dst[d_o] = src[so];
    }
}

// Timeseries draw.  The expression (ExprParams, nothing summed) is laid out over dims [outer..., T, K, event...] with
// n_a = nd; leaf `prev_leaf` is the previous state.  out: [outer, T, K, E] contiguous (E = product of the event dims);
// init: the (already resampled) initial state [outer, K, E]; perm: timeseries_perm [outer, T, K] (null: no permutation).
template <typename T>
struct TsSampleParams {
    ExprParams<T> e;
    int prev_leaf;
    int t_dim, k_dim;                 // positions of T and K in e.d
    const T* init;
    const i64* perm;
    i64 n_outer; int Tn, Kn, En;
};

template <typename T>
__global__ void __launch_bounds__(1024) ts_sample_kernel(const __grid_constant__ TsSampleParams<T> p) {
    extern __shared__ __align__(16) unsigned char ts_smem[];
    T* prev = reinterpret_cast<T*>(ts_smem);                   // [K * E]
    T* cur = prev + (size_t)p.Kn * p.En;
    T reg[AB_NREG];
    T lv[AB_MAXL];
    int idx[AB_MAXD];
    const int KE = p.Kn * p.En;
    for (i64 outer = blockIdx.x; outer < p.n_outer; outer += gridDim.x) {
        for (int i = threadIdx.x; i < KE; i += blockDim.x) prev[i] = p.init[outer * KE + i];
        __syncthreads();
        for (int t = 0; t < p.Tn; ++t) {
            for (int i = threadIdx.x; i < KE; i += blockDim.x) {
                // linear output index of (outer, t, k, e) in the op's dims (outer dims, T, K, event dims: row-major)
                const i64 o = (outer * p.Tn + t) * KE + i;
                unravel(o, p.e.d, 0, p.e.d.nd, idx);
                for (int l = 0; l < p.e.n_leaves; ++l)
                    lv[l] = (l == p.prev_leaf) ? prev[i] : load_leaf<T>(p.e.leaf[l], dot_stride(p.e.leaf[l], idx, 0, p.e.d.nd), idx);
                const T v = vm_eval(p.e.prog, lv, reg);
                p.e.out[o] = v;
                cur[i] = v;
            }
            __syncthreads();
            // the state handed to step t + 1 is this step's draw with its particles permuted (Timeseries.py:116-120)
            for (int i = threadIdx.x; i < KE; i += blockDim.x) {
                const int k = i / p.En, e = i - k * p.En;
                const i64 src = p.perm ? p.perm[(outer * p.Tn + t) * p.Kn + k] : k;
                prev[i] = cur[(int)src * p.En + e];
            }
            __syncthreads();
        }
    }
}
