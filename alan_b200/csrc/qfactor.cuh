// qfactor.cuh -- the mean-field Gaussian Q factor  log N(v; loc, scale)  summed over the event dim, whose value v
// carries (users, kappa) and whose loc / scale are per-user parameters (kappa-independent):
//
//     q[u, kappa] = sum_d log N(v[u,kappa,d]; loc[u,d], scale[u,d])          (MovieLens z: loc = z_loc, scale = exp(z_ls))
//
// normal_q_bwd: the whole adjoint of that factor in ONE pass over v (north_star (4): no autograd re-materialisation):
//     G[u, kappa]  = coeff * sum_s gS[u, s, kappa]         (s: the partial slots / loc samples the fan_lse adjoint wrote)
//     g_loc[u, d]  (+)= sum_kappa G (v - loc) / scale^2
//     g_ls[u, d]   (+)= sum_kappa G ((v - loc)^2 / scale^2 - 1)              (scale = exp(ls): gradient w.r.t. the LOG-scale)
//     g_scale[u,d] (+)= sum_kappa G ((v - loc)^2 / scale^3 - 1 / scale)      (scale is a leaf itself)
// It replaces four launches of the generic path (the sum over s, one gather-style ExprBwd per target leaf, the exp
// adjoint), each of which re-read v or a [u, kappa] tensor.  HBM traffic = v + gS + parameters + gradients, once.
// v tiles ([users of the tile, kappa, d], contiguous) are staged with TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx; UBLKCP in SASS), double-buffered so that the copy of tile i + 1 overlaps the arithmetic on tile i;
// layouts that are not one contiguous 16-byte-aligned block per tile take plain coalesced loads instead.
// Sums over kappa are sequential per thread: bit-reproducible.
// reference: autograd through Dist.log_prob of Q (src/alan/logpq.py:221-235, TorchDimDist.py:127-162).
#pragma once
#include "fan_tc.cuh"

template <typename T>
struct NormalQBwdParams {
    Dims ud;                                   // user dims (n_a = nd)
    i64 vstride[AB_MAXD], lstride[AB_MAXD], sstride[AB_MAXD], gstride[AB_MAXD];   // over the user dims
    i64 v_k, v_ev, l_ev, s_ev;
    const T* v; const T* l; const T* s;
    int scale_is_exp;                          // the scale gradient goes to the log-scale leaf (scale = exp(ls))
    const T* G; i64 g_s, g_k; int S;           // gS[user offset + sI * g_s + kappa * g_k], summed over sI < S
    T coeff;
    T* gl; T* gs;                              // gradient tensors laid out like the loc / scale leaves (null: not wanted)
    int acc_l, acc_s;
    int Kk;
    i64 n_users;
};

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(tc::smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}

template <typename T, int D, int NQB_THREADS>
__global__ void __launch_bounds__(NQB_THREADS) normal_q_bwd_kernel(const __grid_constant__ NormalQBwdParams<T> p, int UPB, int bulk) {
    extern __shared__ __align__(128) unsigned char nqb_smem[];
    const int Kk = p.Kk, row = Kk * D, KP = Kk | 1;
    T* zt0 = reinterpret_cast<T*>(nqb_smem);
    T* zt1 = zt0 + (size_t)UPB * row;
    T* Gs = zt1 + (size_t)UPB * row;                                        // [UPB][KP]
    uint64_t* bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(Gs + (size_t)UPB * KP) + 15) & ~(uintptr_t)15);
    const i64 tiles = (p.n_users + UPB - 1) / UPB;
    if (threadIdx.x == 0) {
        tc::mbar_init(&bar[0], 1); tc::mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const i64 u_lin_stride = p.ud.nd ? p.vstride[p.ud.nd - 1] : 0;          // bulk mode: users are one contiguous run
    auto issue = [&](i64 tile, int s) {
        const i64 u0 = tile * UPB;
        const i64 nu = (p.n_users - u0 < UPB) ? p.n_users - u0 : UPB;
        bulk_load(s ? zt1 : zt0, p.v + u0 * u_lin_stride, (uint32_t)(nu * row * sizeof(T)), &bar[s]);
    };
    if (bulk && threadIdx.x == 0 && (i64)blockIdx.x < tiles) issue(blockIdx.x, 0);
    const int tu = threadIdx.x / D, td = threadIdx.x - tu * D;              // (user slot, event element) of this thread
    int idx[AB_MAXD];
    unsigned it = 0;
    for (i64 tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const i64 u0 = tile * UPB;
        const int nu = (int)((p.n_users - u0 < UPB) ? p.n_users - u0 : UPB);
        T* zt = s ? zt1 : zt0;
        if (bulk) {
            if (threadIdx.x == 0 && tile + gridDim.x < tiles) issue(tile + gridDim.x, s ^ 1);
        } else {
            // plain staging: consecutive threads read consecutive elements of a user's [kappa, d] block
            for (int us = 0; us < nu; ++us) {
                unravel(u0 + us, p.ud, 0, p.ud.nd, idx);
                i64 vo = 0;
                for (int k = 0; k < p.ud.nd; ++k) vo += (i64)idx[k] * p.vstride[k];
                for (int e = threadIdx.x; e < row; e += NQB_THREADS) {
                    const int k = e / D, d = e - k * D;
                    zt[(size_t)us * row + e] = p.v[vo + (i64)k * p.v_k + (i64)d * p.v_ev];
                }
            }
        }
        // G[u, kappa] = coeff * sum over the S partial slots, in slot order
        for (int e = threadIdx.x; e < nu * Kk; e += NQB_THREADS) {
            const int us = e / Kk, k = e - us * Kk;
            unravel(u0 + us, p.ud, 0, p.ud.nd, idx);
            i64 go = (i64)k * p.g_k;
            for (int q = 0; q < p.ud.nd; ++q) go += (i64)idx[q] * p.gstride[q];
            T a = T(0);
            for (int sI = 0; sI < p.S; ++sI) a += p.G[go + (i64)sI * p.g_s];
            Gs[us * KP + k] = p.coeff * a;
        }
        if (bulk) tc::mbar_wait(&bar[s], (it >> 1) & 1);
        __syncthreads();
        if (tu < nu) {
            unravel(u0 + tu, p.ud, 0, p.ud.nd, idx);
            i64 lo = (i64)td * p.l_ev, so = (i64)td * p.s_ev;
            for (int q = 0; q < p.ud.nd; ++q) { lo += (i64)idx[q] * p.lstride[q]; so += (i64)idx[q] * p.sstride[q]; }
            const T loc = p.l[lo], sc = p.s[so];
            const T* zr = zt + (size_t)tu * row + td;
            const T* gr = Gs + tu * KP;
            T a1 = T(0), a2 = T(0), a0 = T(0);
#pragma unroll 6
            for (int k = 0; k < Kk; ++k) {
                const T g = gr[k], df = zr[k * D] - loc;
                a0 += g; a1 += g * df; a2 += g * (df * df);
            }
            const T iv = T(1) / (sc * sc);
            if (p.gl) { const T r = a1 * iv; p.gl[lo] = p.acc_l ? p.gl[lo] + r : r; }
            if (p.gs) {
                const T r = p.scale_is_exp ? (a2 * iv - a0) : (a2 * iv - a0) / sc;
                p.gs[so] = p.acc_s ? p.gs[so] + r : r;
            }
        }
        __syncthreads();                                                     // stage and Gs are free again
    }
}

template <typename T, int D, int NT>
static int launch_normal_q_bwd_DT(const NormalQBwdParams<T>& p, cudaStream_t stream, int sm_count) {
    int UPB = NT / D;
    if (UPB < 1) return 1;
    const size_t row = (size_t)p.Kk * D;
    const size_t per_user = (2 * row + (size_t)(p.Kk | 1)) * sizeof(T);
    if ((size_t)UPB * per_user > 96 * 1024) UPB = (int)(96 * 1024 / per_user);      // two CTAs per SM
    if (UPB < 1) return 2;
    const size_t smem = (size_t)UPB * per_user + 16 + 2 * sizeof(uint64_t);
    // TMA bulk staging: the users of a tile form ONE contiguous block whose size and address are multiples of 16 bytes
    bool bulk = p.v_ev == 1 && p.v_k == D && (row * sizeof(T)) % 16 == 0 && ((uintptr_t)p.v % 16) == 0;
    {
        i64 expect = (i64)row;
        for (int k = p.ud.nd - 1; k >= 0 && bulk; --k) { bulk = p.vstride[k] == expect; expect *= p.ud.size[k]; }
    }
    if (bulk && ((size_t)UPB * row * sizeof(T)) % 16 != 0) bulk = false;              // every tile start stays 16-byte aligned
    const i64 tiles = (p.n_users + UPB - 1) / UPB;
    // resident CTAs per SM: bounded by shared memory (227 KB) and by 2048 threads; every CTA gets the same trip count
    i64 per_sm = (i64)((220 * 1024) / (smem + 1024));
    if (per_sm > 2048 / NT) per_sm = 2048 / NT;
    if (per_sm < 1) per_sm = 1;
    const i64 cap = (i64)sm_count * per_sm;
    i64 blocks = tiles;
    if (blocks > cap) { const i64 trips = (tiles + cap - 1) / cap; blocks = (tiles + trips - 1) / trips; }
    if (blocks < 1) blocks = 1;
    static const cudaError_t attr = cudaFuncSetAttribute(normal_q_bwd_kernel<T, D, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    (void)attr;
    normal_q_bwd_kernel<T, D, NT><<<(int)blocks, NT, smem, stream>>>(p, UPB, bulk ? 1 : 0);
    return 0;
}

template <typename T, int D>
static int launch_normal_q_bwd_D(const NormalQBwdParams<T>& p, cudaStream_t stream, int sm_count) {
    // smaller CTAs = more tiles in flight per SM at different phases (copy / wait / arithmetic / store) when there are
    // enough users to fill them; ALAN_B200_NQB_THREADS overrides (tuning aid)
    const char* e = getenv("ALAN_B200_NQB_THREADS");
    const int nt = e ? atoi(e) : 256;
    if (nt == 128 && D <= 64) return launch_normal_q_bwd_DT<T, D, 128>(p, stream, sm_count);
    return launch_normal_q_bwd_DT<T, D, 256>(p, stream, sm_count);
}

template <typename T>
static int launch_normal_q_bwd(const NormalQBwdParams<T>& p, int D, cudaStream_t stream, int sm_count) {
    switch (D) {
        case 1: return launch_normal_q_bwd_D<T, 1>(p, stream, sm_count);
        case 2: return launch_normal_q_bwd_D<T, 2>(p, stream, sm_count);
        case 3: return launch_normal_q_bwd_D<T, 3>(p, stream, sm_count);
        case 4: return launch_normal_q_bwd_D<T, 4>(p, stream, sm_count);
        case 6: return launch_normal_q_bwd_D<T, 6>(p, stream, sm_count);
        case 8: return launch_normal_q_bwd_D<T, 8>(p, stream, sm_count);
        case 12: return launch_normal_q_bwd_D<T, 12>(p, stream, sm_count);
        case 16: return launch_normal_q_bwd_D<T, 16>(p, stream, sm_count);
        case 18: return launch_normal_q_bwd_D<T, 18>(p, stream, sm_count);
        case 24: return launch_normal_q_bwd_D<T, 24>(p, stream, sm_count);
        case 32: return launch_normal_q_bwd_D<T, 32>(p, stream, sm_count);
    }
    return 1;
}
