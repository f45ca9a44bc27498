// normal_poly.cuh -- regression-style likelihoods summed over their plate in one pass:
//
//     out[row, k] = cadd + sum_z log N( resid = 0 ; ... )
//                 = cadd - 1/(2 s[row,k]^2) * sum_z r(row, k, z)^2 - n_z (log s[row,k] + log sqrt(2 pi))
//     r(row, k, z) = sum_t c_t * Zpart_t(row, z) * Kpart_t(row, k)            (value minus loc, as a polynomial)
//
// This is `logP(data)` of a Normal whose loc is a polynomial in tensors that depend either on the summed plate
// ("Z leaves": data and covariates, e.g. obs[s,c,z], basement[s,c,z]) or on K axes ("K leaves": latent samples, e.g.
// County_mean[s,c,K], Beta_u[s,c,K']) and whose scale is a K leaf, followed by the plate sum over z
// (reference: Dist.log_prob -> TorchDimDist.log_prob -> sum over the plate, src/alan/logpq.py:186-196,149-153;
// radon: examples/models/radon/radon.py:62-102).  The reference materialises the [row, z, K^4] tensor; the generic
// path here interprets the traced expression per cell (VM) and writes it out before a second pass sums it.  Here the
// Z parts of one row are staged once in shared memory (products of the Z leaves: a few hundred floats), every thread
// owns NP_OPT K-tuples, keeps their Kpart coefficients in registers and walks z: n_terms FMAs + one FMA for the
// square per (cell, z) -- 4 for the radon likelihood against ~100 interpreted instructions -- and nothing of size
// [row, z, K...] ever exists.  Bound: FP32 FMA pipe; HBM traffic = inputs + the [row, K...] output once.
#pragma once
#include "kernels.cuh"

#define NP_MAXLEAF 8
#define NP_MAXTERM 8
#define NP_OPT 4            // outputs (K tuples) per thread: the staged Z parts are read once per NP_OPT outputs
#define NP_THREADS 256
#define NP_ZTILE 512        // z points staged at a time

struct NpTerm { double coeff; int z[2]; int k[2]; };           // leaf ids, -1 = none

template <typename T>
struct NormalPolyParams {
    Dims d;                              // [row dims..., K dims..., z dims...]
    int n_row, n_k, n_z;                 // how many dims of each kind
    i64 rows, ks, zs;                    // their total extents
    Opnd zleaf[NP_MAXLEAF]; int n_zleaf;  // strides over all dims (zero over the K dims)
    Opnd kleaf[NP_MAXLEAF]; int n_kleaf;  // strides over all dims (zero over the z dims)
    int n_zt, n_kt;                      // terms with / without a Z part
    NpTerm zt[NP_MAXTERM], kt[NP_MAXTERM];
    int scale_leaf; double scale_const;  // K leaf id of the scale, or -1: constant
    T* out; i64 ostride[AB_MAXD];        // over row and K dims
    T cadd;
};

template <typename T, int NZT>
__global__ void __launch_bounds__(NP_THREADS) normal_poly_sum_kernel(const __grid_constant__ NormalPolyParams<T> p) {
    extern __shared__ __align__(16) unsigned char np_smem[];
    T* zf = reinterpret_cast<T*>(np_smem);                     // [NZT][NP_ZTILE]
    const i64 kchunks = (p.ks + (i64)NP_THREADS * NP_OPT - 1) / ((i64)NP_THREADS * NP_OPT);
    const int nd_rk = p.n_row + p.n_k;
    for (i64 work = blockIdx.x; work < p.rows * kchunks; work += gridDim.x) {
        const i64 row = work / kchunks, kc = work - row * kchunks;
        int idx[AB_MAXD];
        unravel(row, p.d, 0, p.n_row, idx);
        i64 zbase[NP_MAXLEAF], kbase[NP_MAXLEAF];
        for (int l = 0; l < p.n_zleaf; ++l) zbase[l] = dot_stride(p.zleaf[l], idx, 0, p.n_row);
        for (int l = 0; l < p.n_kleaf; ++l) kbase[l] = dot_stride(p.kleaf[l], idx, 0, p.n_row);
        const i64 obase = [&] { i64 o = 0; for (int k = 0; k < p.n_row; ++k) o += (i64)idx[k] * p.ostride[k]; return o; }();
        // ---- this thread's outputs: Kpart coefficients of every term, the constant part, the scale
        T kcf[NP_OPT][NZT], c0[NP_OPT], sg[NP_OPT], acc[NP_OPT];
        i64 ooff[NP_OPT];
#pragma unroll
        for (int j = 0; j < NP_OPT; ++j) {
            const i64 k = kc * ((i64)NP_THREADS * NP_OPT) + (i64)j * NP_THREADS + threadIdx.x;
            ooff[j] = -1; c0[j] = T(0); sg[j] = T(1); acc[j] = T(0);
#pragma unroll
            for (int t = 0; t < NZT; ++t) kcf[j][t] = T(0);
            if (k >= p.ks) continue;
            unravel(k, p.d, p.n_row, nd_rk, idx);
            T kv[NP_MAXLEAF];
            for (int l = 0; l < p.n_kleaf; ++l)
                kv[l] = ((const T*)p.kleaf[l].ptr)[kbase[l] + dot_stride(p.kleaf[l], idx, p.n_row, nd_rk)];
#pragma unroll
            for (int t = 0; t < NZT; ++t) {
                T c = (T)p.zt[t].coeff;
                if (p.zt[t].k[0] >= 0) c *= kv[p.zt[t].k[0]];
                if (p.zt[t].k[1] >= 0) c *= kv[p.zt[t].k[1]];
                kcf[j][t] = c;
            }
            for (int t = 0; t < p.n_kt; ++t) {
                T c = (T)p.kt[t].coeff;
                if (p.kt[t].k[0] >= 0) c *= kv[p.kt[t].k[0]];
                if (p.kt[t].k[1] >= 0) c *= kv[p.kt[t].k[1]];
                c0[j] += c;
            }
            sg[j] = p.scale_leaf >= 0 ? kv[p.scale_leaf] : (T)p.scale_const;
            i64 o = obase;
            for (int q = p.n_row; q < nd_rk; ++q) o += (i64)idx[q] * p.ostride[q];
            ooff[j] = o;
        }
        // ---- walk z in staged tiles
        for (i64 z0 = 0; z0 < p.zs; z0 += NP_ZTILE) {
            const int zn = (int)((p.zs - z0) < NP_ZTILE ? (p.zs - z0) : NP_ZTILE);
            __syncthreads();                                   // the previous tile (or work item) is done with zf
            for (int i = threadIdx.x; i < NZT * zn; i += NP_THREADS) {
                const int t = i / zn, z = i - t * zn;
                int zi[AB_MAXD];
                unravel(z0 + z, p.d, nd_rk, p.d.nd, zi);
                T v = T(1);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int l = p.zt[t].z[e];
                    if (l >= 0) v *= ((const T*)p.zleaf[l].ptr)[zbase[l] + dot_stride(p.zleaf[l], zi, nd_rk, p.d.nd)];
                }
                zf[t * NP_ZTILE + z] = v;
            }
            __syncthreads();
            for (int z = 0; z < zn; ++z) {
                T f[NZT];
#pragma unroll
                for (int t = 0; t < NZT; ++t) f[t] = zf[t * NP_ZTILE + z];
#pragma unroll
                for (int j = 0; j < NP_OPT; ++j) {
                    T r = c0[j];
#pragma unroll
                    for (int t = 0; t < NZT; ++t) r = fma(kcf[j][t], f[t], r);
                    acc[j] = fma(r, r, acc[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NP_OPT; ++j) {
            if (ooff[j] < 0) continue;
            const T s = sg[j];
            p.out[ooff[j]] = p.cadd - acc[j] / (T(2) * s * s) - (T)p.zs * (ab_log(s) + T(HALF_LOG_2PI));
        }
    }
}

template <typename T>
static int launch_normal_poly_sum(const NormalPolyParams<T>& p, cudaStream_t stream, int sm_count) {
    if (p.n_zt < 1 || p.n_zt > 6 || p.n_kt > NP_MAXTERM || p.n_zleaf > NP_MAXLEAF || p.n_kleaf > NP_MAXLEAF) return 1;
    const i64 kchunks = (p.ks + (i64)NP_THREADS * NP_OPT - 1) / ((i64)NP_THREADS * NP_OPT);
    i64 work = p.rows * kchunks;
    i64 grid = work < (i64)sm_count * 16 ? work : (i64)sm_count * 16;
    if (grid < 1) grid = 1;
    const size_t smem = (size_t)p.n_zt * NP_ZTILE * sizeof(T);
    switch (p.n_zt) {
        case 1: normal_poly_sum_kernel<T, 1><<<(int)grid, NP_THREADS, smem, stream>>>(p); break;
        case 2: normal_poly_sum_kernel<T, 2><<<(int)grid, NP_THREADS, smem, stream>>>(p); break;
        case 3: normal_poly_sum_kernel<T, 3><<<(int)grid, NP_THREADS, smem, stream>>>(p); break;
        case 4: normal_poly_sum_kernel<T, 4><<<(int)grid, NP_THREADS, smem, stream>>>(p); break;
        case 5: normal_poly_sum_kernel<T, 5><<<(int)grid, NP_THREADS, smem, stream>>>(p); break;
        case 6: normal_poly_sum_kernel<T, 6><<<(int)grid, NP_THREADS, smem, stream>>>(p); break;
    }
    return 0;
}
