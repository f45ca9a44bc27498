// fan_tc2b.cuh -- adjoint of the dense fan_lse (fan_tc2.cuh) with the product TRANSPOSED.
//
//     gS[u, kappa] = sum_{f'} gout[u, f'] exp(S[u, kappa, f'] - lse[u, f'])            f' = (lam, f)
//
// The forward needs kappa on the accumulator COLUMNS (its LSE reduces over kappa inside one thread).  The adjoint
// reduces over the fan f' instead, so here the roles of the operands are swapped:
//
//   A (values, one TMEM stage per block of 4 users, written by the builders with tcgen05.st -- no shared-memory
//      round trip, no proxy fence)        rows (user slot, kappa) = 128 TMEM lanes,  K = 2 D + 2 -> KT, hi and lo
//   B (constants, shared memory, once)    rows f' (128 per tile, up to 3 tiles per CTA = one fan group), hi and lo;
//                                         one more K column carries (cadd - C[f']) log2e against a 1 on the value side
//   D = A B^T (TMEM, 2 stages x 128 columns): lane = (user slot, kappa), columns = f'
//
// so that an epilogue thread owns ONE (user, kappa) and walks the fan along its registers:
//     d = D - lse[u, f'] log2e;   acc += ex2(d) gout[u, f']
// with lse / gout of the block staged in shared memory (cp.async, one block ahead) and read as warp-wide broadcasts
// (every lane of a warp belongs to the same user).  The sum over f' is a plain in-thread accumulation: the 31-shuffle
// butterfly per user and block of the lane = f' layout (as many issue slots as three tiles of ex2) is gone; the four
// teams, which split the columns of every tile, combine one value per thread through shared memory in a fixed order.
//
// Roles as in the forward (16 epilogue warps = 4 teams x 4 lane quadrants, 1 MMA warp, 4 builder warps); the
// builder warp of user slot us is the one with warp % 4 == us (a warp reaches only its own TMEM lane quadrant).
#pragma once
#include <cstdio>
#include "fan_tc2.cuh"

namespace tc {

constexpr int T2B_ASTG = 3;       // value-operand stages in TMEM
constexpr int T2B_LSTG = 3;       // lse / gout stages in shared memory

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

template <int D>
__global__ void __launch_bounds__(T2_WARPS * 32, 1) fan_lse_tc2_adj_kernel(const __grid_constant__ FanLseParams<float> p, const __grid_constant__ Tc2Geom geo) {
    constexpr int KT = (2 * D + 2 + 7) / 8 * 8;                            // K extent (40 at D = 18)
    constexpr int NC = KT / 4, KSTEPS = KT / 8;
    constexpr uint32_t LBO = 128 * 16, SBO = 8 * 16;                       // [chunk][128 rows][16 B]
    constexpr uint32_t OPER = NC * LBO;                                    // bytes of one constant tile part (20 KB at D = 18)
    constexpr uint32_t A_COL = 0, D_COL = T2B_ASTG * 2 * KT;
    static_assert(D_COL + T2_ACC * 128 <= T2_TMEM_COLS, "TMEM budget");
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((128u >> 4) << 24);
    constexpr int FW = T2_TILES * 128;                                     // fan columns of one group

    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char* bc_base = tc_smem;                                      // [tile][hi | lo] constant operand
    unsigned char* tail = bc_base + (size_t)T2_TILES * 2 * OPER;
    float* s_lse = reinterpret_cast<float*>(tail);                         // [T2B_LSTG][4 user slots][FW]
    float* s_g = s_lse + T2B_LSTG * T2_US * FW;                            // [T2B_LSTG][4 user slots][FW]
    int* s_ooff = reinterpret_cast<int*>(s_g + T2B_LSTG * T2_US * FW);     // [FW] out / lse offset of f' (-1: padding)
    int* s_goff = s_ooff + FW;                                             // [FW] gout offset of f'
    float* s_red = reinterpret_cast<float*>(s_goff + FW);                  // [2][team][128]
    float* s_cd = s_red + 2 * T2_EPI * 128;                                // [32] centre per event element
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_cd + 32);
    uint64_t* afull = bars;
    uint64_t* aempty = afull + T2B_ASTG;
    uint64_t* tfull = aempty + T2B_ASTG;
    uint64_t* tempty = tfull + T2_ACC;
    uint64_t* lfull = tempty + T2_ACC;
    uint64_t* lempty = lfull + T2B_LSTG;
    uint32_t* tmem_slot = (uint32_t*)(lempty + T2B_LSTG);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float LS = 1.4426950408889634f;
    const int Kk = p.Kk;
    const unsigned n_u = (unsigned)geo.n_u;
    const unsigned n_blocks = (n_u + T2_US - 1) / T2_US;
    int grp = 0;
#pragma unroll
    for (int g = 1; g < T2_MAXG; ++g) if (g < geo.NG && (int)blockIdx.x >= geo.cta_lo[g]) grp = g;
    const unsigned blk0 = blockIdx.x - geo.cta_lo[grp], blk_step = geo.cta_lo[grp + 1] - geo.cta_lo[grp];
    const int fp_lo = grp * FW;
    const int n_tiles = min(T2_TILES, (geo.FP - fp_lo + 127) / 128);
    long long dbg0 = 0, dbg1 = 0, dbg2 = 0, dbg3 = 0, dbg4 = 0, dbg5 = 0; (void)dbg0; (void)dbg1; (void)dbg2; (void)dbg3; (void)dbg4; (void)dbg5;
    const long long dbg_k0 = clock64(); (void)dbg_k0;

    // ---------------------------------------------------------------- prologue
    {
        // centre per event element = mean over lam of the loc (all loc values in one global round trip, staged in the
        // still unused operand memory; lane d adds its column in lam order: reproducible)
        float* stage_f = reinterpret_cast<float*>(tc_smem);
        for (int i = threadIdx.x; i < geo.L * D; i += blockDim.x) {
            const int j = i / D, dd = i - j * D;
            stage_f[i] = p.l[j * geo.l_lam + dd * (int)p.l_ev];
        }
        __syncthreads();
        if (warp == 0) {
            float c = 0.f;
            if (lane < D) {
                for (int j = 0; j < geo.L; ++j) c += stage_f[j * D + lane];
                c /= (float)geo.L;
            }
            s_cd[lane] = c;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < T2B_ASTG; ++s) { mbar_init(&afull[s], 32 * T2_BW); mbar_init(&aempty[s], 1); }
        for (int a = 0; a < T2_ACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 128 * T2_EPI); }
        for (int s = 0; s < T2B_LSTG; ++s) { mbar_init(&lfull[s], 32 * T2_BW); mbar_init(&lempty[s], 4 * T2_EPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == T2_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(T2_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp < 4 * T2_TILES) {
        // ---------------------------------------------------------------- constant operand, once: thread = one row f' of one tile
        // (12 warps side by side: this prologue is the fixed cost that small problems and small shards see)
        const int row = threadIdx.x & 127;
        {
            const int tl = threadIdx.x >> 7;
            const int fp = fp_lo + 128 * tl + row;
            float a[KT];
#pragma unroll
            for (int k = 0; k < KT; ++k) a[k] = 0.f;
            int ooff = -1, goff = 0;
            if (tl < n_tiles && fp < geo.FP) {
                const int lam = fp / p.F, f = fp - lam * p.F;
                float c = 0.f;
#pragma unroll
                for (int dd = 0; dd < D; ++dd) {
                    const float sc = p.s[f * (int)p.s_f + dd * (int)p.s_ev];
                    const float lc = p.l[lam * geo.l_lam + dd * (int)p.l_ev] - s_cd[dd];
                    const float w = 1.f / (2.f * (sc * sc));
                    a[dd] = -w * LS;
                    a[D + dd] = 2.f * lc * w * LS;
                    c += lc * lc * w + logf(sc);
                }
                a[2 * D] = 1.f;
                a[2 * D + 1] = (p.cadd - (c + float(D) * float(HALF_LOG_2PI))) * LS;
                ooff = lam * geo.o_lam + f * (int)p.o_f;
                goff = lam * geo.g_lam + f * geo.g_f;
            }
            s_ooff[tl * 128 + row] = ooff;
            s_goff[tl * 128 + row] = goff;
            float* bh = reinterpret_cast<float*>(bc_base + (size_t)tl * 2 * OPER);
            float* bl = reinterpret_cast<float*>(bc_base + (size_t)tl * 2 * OPER + OPER);
#pragma unroll
            for (int c4 = 0; c4 < NC; ++c4) {
                float4 hh, l;
                hh.x = __uint_as_float(to_tf32(a[4 * c4])); l.x = a[4 * c4] - hh.x;
                hh.y = __uint_as_float(to_tf32(a[4 * c4 + 1])); l.y = a[4 * c4 + 1] - hh.y;
                hh.z = __uint_as_float(to_tf32(a[4 * c4 + 2])); l.z = a[4 * c4 + 2] - hh.z;
                hh.w = __uint_as_float(to_tf32(a[4 * c4 + 3])); l.w = a[4 * c4 + 3] - hh.w;
                const int off = (c4 * 128 + row) * 4;
                *reinterpret_cast<float4*>(bh + off) = hh;
                *reinterpret_cast<float4*>(bl + off) = l;
            }
        }
    }
    for (int i = threadIdx.x; i < T2B_LSTG * T2_US * FW; i += blockDim.x) { s_lse[i] = INFINITY; s_g[i] = 0.f; }   // padding columns: weight 0
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const long long dbg_t0 = clock64(); (void)dbg_t0;

    if (warp < 4 * T2_EPI) {
        // ---------------------------------------------------------------- epilogue: quadrant q = user slot, lane = kappa;
        // team e takes columns [32 e, 32 e + 32) of every tile
        const int q = warp & 3, team = warp >> 2;
        const uint32_t ld_base = tmem + ((uint32_t)(32 * q) << 16) + D_COL + 32 * team;
        const float2 nls = make_float2(-LS, -LS);
        // The accumulator is read 16 columns at a time into two register buffers: the tcgen05.ld of the next 16 columns
        // (the second half of this tile, or the first half of the NEXT tile when its MMAs have already completed) is in
        // flight while the current 16 go through the MUFU pipe, so the TMEM read latency -- which all four warps of a
        // sub-partition would otherwise expose at the same moment, right after the tile's barrier -- stays hidden.
        uint32_t ra[16], rb[16];
        float2 acc[4];
        const float* lrow = s_lse;
        const float* grow = s_g;
        auto half = [&](const uint32_t (&r)[16], int col) {
#ifdef TC_EXP_NOEPI
            acc[0].x += __uint_as_float(r[0]) + __uint_as_float(r[15]); return;
#endif
#pragma unroll
            for (int k = 0; k < 16; k += 4) {
                const float4 l4 = *reinterpret_cast<const float4*>(lrow + col + k);
                const float4 g4 = *reinterpret_cast<const float4*>(grow + col + k);
                const float2 d0 = __ffma2_rn(make_float2(l4.x, l4.y), nls, make_float2(__uint_as_float(r[k]), __uint_as_float(r[k + 1])));
                const float2 d1 = __ffma2_rn(make_float2(l4.z, l4.w), nls, make_float2(__uint_as_float(r[k + 2]), __uint_as_float(r[k + 3])));
                const float2 e0 = make_float2(FastExp<float>::ex(d0.x), FastExp<float>::ex(d0.y));
                const float2 e1 = make_float2(FastExp<float>::ex(d1.x), FastExp<float>::ex(d1.y));
                acc[(k >> 1) & 3] = __ffma2_rn(e0, make_float2(g4.x, g4.y), acc[(k >> 1) & 3]);
                acc[((k >> 1) + 1) & 3] = __ffma2_rn(e1, make_float2(g4.z, g4.w), acc[((k >> 1) + 1) & 3]);
            }
        };
        unsigned tt = 0, it = 0;
        if (blk0 < n_blocks) {                                                   // first half of the very first tile
            MBAR_WAIT(&tfull[0], 0, dbg1);
            tc_fence_after();
            TC_LD16(ra, ld_base);
        }
        for (unsigned blk = blk0; blk < n_blocks; blk += blk_step, ++it) {
            const int ls = it % T2B_LSTG;
            MBAR_WAIT(&lfull[ls], (it / T2B_LSTG) & 1, dbg0);
            lrow = s_lse + (ls * T2_US + q) * FW + 32 * team;
            grow = s_g + (ls * T2_US + q) * FW + 32 * team;
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c] = make_float2(0.f, 0.f);
            const bool more_blocks = blk + blk_step < n_blocks;
#pragma unroll
            for (int tl = 0; tl < T2_TILES; ++tl) {
                if (tl < n_tiles) {
                    const int a = tt % T2_ACC;
                    ++tt;
                    TC_WAIT_LD16(ra);
                    TC_LD16(rb, ld_base + 128 * a + 16);
                    half(ra, 128 * tl);
                    TC_WAIT_LD16(rb);
                    tc_fence_before();
                    mbar_arrive(&tempty[a]);
                    // first half of the next tile (same block or the next one): at once if its MMAs are done, else
                    // after this tile's second half
                    const bool has_next = tl + 1 < n_tiles || more_blocks;
                    const int na = tt % T2_ACC;
                    const uint32_t npa = (tt / T2_ACC) & 1;
                    bool issued = false;
                    if (has_next) {
                        uint32_t ok = 0;
                        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                                     : "=r"(ok) : "r"(smem_u32(&tfull[na])), "r"(npa) : "memory");
                        issued = __shfl_sync(0xffffffffu, ok, 0) != 0;
                        if (issued) { MBAR_WAIT(&tfull[na], npa, dbg1); tc_fence_after(); TC_LD16(ra, ld_base + 128 * na); }
                    }
                    half(rb, 128 * tl + 16);
                    if (has_next && !issued) { MBAR_WAIT(&tfull[na], npa, dbg1); tc_fence_after(); TC_LD16(ra, ld_base + 128 * na); }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&lempty[ls]);
            const float2 t2 = __fadd2_rn(__fadd2_rn(acc[0], acc[1]), __fadd2_rn(acc[2], acc[3]));
            float* red = s_red + (it & 1) * (T2_EPI * 128);
            red[team * 128 + 32 * q + lane] = t2.x + t2.y;
            asm volatile("bar.sync %0, %1;" :: "r"(1 + q), "n"(32 * T2_EPI) : "memory");
            if (team == 0) {
                const float* rp = red + 32 * q + lane;
                float sum = rp[0];
#pragma unroll
                for (int e = 1; e < T2_EPI; ++e) sum += rp[e * 128];
                const unsigned u = T2_US * blk + q;
                if (u < n_u && lane < Kk) {
                    int idx[T2_ND];
                    t2_decode(u, geo, idx);
                    p.gS[((i64)t2_dot(idx, geo.ss) + (i64)grp * geo.s_lam) * Kk + lane] = sum;
                }
            }
        }
    } else if (warp == T2_MMA_WARP) {
        // ---------------------------------------------------------------- MMA issuer
        unsigned it = 0, tt = 0;
        for (unsigned blk = blk0; blk < n_blocks; blk += blk_step, ++it) {
            const int s = it % T2B_ASTG;
            MBAR_WAIT(&afull[s], (it / T2B_ASTG) & 1, dbg0);
            tc_fence_after();
            const uint32_t ahi = tmem + A_COL + s * 2 * KT, alo = ahi + KT;
            for (int tl = 0; tl < n_tiles; ++tl, ++tt) {
                const int a = tt % T2_ACC;
                const uint32_t pa = (tt / T2_ACC) & 1;
                MBAR_WAIT(&tempty[a], pa ^ 1, dbg1);
                tc_fence_after();
                const uint32_t d = tmem + D_COL + 128 * a;
                const uint32_t bhi = smem_u32(bc_base + (size_t)tl * 2 * OPER), blo = bhi + OPER;
                if (elect_one()) {
#ifdef TC_EXP_NOMMA
                    constexpr int KS_ = 1;
#else
                    constexpr int KS_ = KSTEPS;
#endif
#pragma unroll
                    for (int j = 0; j < KS_; ++j) {
                        const uint64_t dh = smem_desc(bhi + j * 2 * LBO, LBO, SBO), dl = smem_desc(blo + j * 2 * LBO, LBO, SBO);
                        mma_tf32_ts(d, ahi + 8 * j, dh, IDESC, j > 0 ? 1u : 0u);
                        mma_tf32_ts(d, alo + 8 * j, dh, IDESC, 1u);
                        mma_tf32_ts(d, ahi + 8 * j, dl, IDESC, 1u);
                    }
                    if (tl == n_tiles - 1) tc_commit(&aempty[s]);
                    tc_commit(&tfull[a]);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------------------------------------------------------- builders: user slot = warp % 4 (TMEM lane quadrant),
        // lane = kappa.  Raw values of the NEXT block are loaded before the current one is written; lse / gout of the
        // next block travel to shared memory asynchronously (cp.async) while this block is built.
        const int us = warp & 3, kz = lane;
        const int vk = (int)p.v_k, vev = (int)p.v_ev, nb = geo.nb, vec2 = geo.vec2;
        const uint32_t lane_base = tmem + ((uint32_t)(32 * us) << 16);
        auto load_raw = [&](unsigned blk, float (&raw)[D], float (&braw)[TC_NB], float& ql, float& qs) {
            const unsigned u = T2_US * blk + us;
            const bool live = u < n_u && kz < Kk;
#pragma unroll
            for (int dd = 0; dd < D; ++dd) raw[dd] = s_cd[dd];                  // idle rows: v' = 0
#pragma unroll
            for (int i = 0; i < TC_NB; ++i) braw[i] = 0.f;
            ql = 0.f; qs = 1.f;
            if (geo.qn && u < n_u && lane < D) {
                int idx[T2_ND];
                t2_decode(u, geo, idx);
                ql = p.q_l[t2_dot(idx, geo.qls) + lane * geo.q_lev];
                qs = p.q_s[t2_dot(idx, geo.qss) + lane * geo.q_sev];
            }
            if (live) {
                int idx[T2_ND];
                t2_decode(u, geo, idx);
#pragma unroll
                for (int i = 0; i < TC_NB; ++i) if (i < nb) braw[i] = p.b[i][t2_dot(idx, geo.bs[i]) + kz * geo.bk[i]];
                const float* vp = p.v + t2_dot(idx, geo.vs) + kz * vk;
                if (vec2) {
#pragma unroll
                    for (int q2 = 0; q2 < D / 2; ++q2) {
                        const float2 vv = *reinterpret_cast<const float2*>(vp + 2 * q2);
                        raw[2 * q2] = vv.x; raw[2 * q2 + 1] = vv.y;
                    }
                } else {
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) raw[dd] = vp[dd * vev];
                }
            }
        };
        auto stage_lg = [&](unsigned blk, unsigned jt) {
            // lse[u, f'] and gout[u, f'] of this warp's user for the fan columns of the group; padding -> weight 0
            const int ls = jt % T2B_LSTG;
            MBAR_WAIT(&lempty[ls], ((jt / T2B_LSTG) & 1) ^ 1, dbg0);
            const unsigned u = T2_US * blk + us;
            float* ld = s_lse + (ls * T2_US + us) * FW;
            float* gd = s_g + (ls * T2_US + us) * FW;
            int uo = 0, ug = 0;
            if (u < n_u) {
                int idx[T2_ND];
                t2_decode(u, geo, idx);
                uo = t2_dot(idx, geo.os); ug = t2_dot(idx, geo.gs);
            }
            if (geo.lg_vec4 && u < n_u) {
                // contiguous rows: 16 bytes per copy; the padding columns (f' >= FP) were set once in the prologue
                const int nv = min(128 * n_tiles, geo.FP - fp_lo);
                const float* lsrc = p.lse + uo + fp_lo;
                const float* gsrc = p.gout + ug + fp_lo;
                for (int j = 4 * lane; j < nv; j += 128) {
                    cp_async16(ld + j, lsrc + j);
                    cp_async16(gd + j, gsrc + j);
                }
            } else {
                for (int j = lane; j < 128 * n_tiles; j += 32) {
                    const int oo = s_ooff[j];
                    if (u < n_u && oo >= 0) {
                        cp_async4(ld + j, p.lse + uo + oo);
                        cp_async4(gd + j, p.gout + ug + s_goff[j]);
                    } else {
                        ld[j] = INFINITY; gd[j] = 0.f;
                    }
                }
            }
        };
        // one block: `c*` hold its raw loads (issued one step ago), `n*` receive the next block's
        auto step = [&](unsigned blk, unsigned it, float (&c)[D], float (&cb)[TC_NB], float& c_ql, float& c_qs,
                        float (&n)[D], float (&nb_)[TC_NB], float& n_ql, float& n_qs) {
            const int s = it % T2B_ASTG;
            const bool more = blk + blk_step < n_blocks;
#ifdef TC_DEBUG_SPIN
            const long long tb0 = clock64();
#endif
            if (more) load_raw(blk + blk_step, n, nb_, n_ql, n_qs);
#ifdef TC_DEBUG_SPIN
            const long long tb1 = clock64();
            dbg3 += tb1 - tb0;
#endif
            MBAR_WAIT(&aempty[s], ((it / T2B_ASTG) & 1) ^ 1, dbg1);
            tc_fence_after();
#ifdef TC_DEBUG_SPIN
            const long long tb2 = clock64();
#endif
            float qsum = 0.f;
            if (geo.qn) {
                const float iv2 = 0.5f / (c_qs * c_qs);
                float lg = lane < D ? logf(c_qs) : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
#pragma unroll
                for (int dd = 0; dd < D; ++dd) {
                    const float l_d = __shfl_sync(0xffffffffu, c_ql, dd), w_d = __shfl_sync(0xffffffffu, iv2, dd);
                    const float df = c[dd] - l_d;
                    qsum = fmaf(-(df * df), w_d, qsum);
                }
                qsum -= lg + float(D) * float(HALF_LOG_2PI);
            }
            float bsum = geo.qc * qsum;
#pragma unroll
            for (int i = 0; i < TC_NB; ++i) if (i < nb) bsum += geo.bc[i] * cb[i];
            bsum *= LS;
            if (kz >= Kk) bsum = -1.0e30f;                                     // padding kappa: weight 0
#pragma unroll
            for (int g8 = 0; g8 < KT / 8; ++g8) {
                // the tensor core reads only the 19 TF32 bits of a word: the raw fp32 value IS the "hi" part and
                // lo = t - trunc(t) its exact remainder
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int k = 8 * g8 + e;
                    const int dk = k < D ? k : (k < 2 * D ? k - D : 0);
                    const float df = c[dk] - s_cd[dk];
                    const float tv = k < D ? df * df : k < 2 * D ? df : k == 2 * D ? bsum : k == 2 * D + 1 ? 1.f : 0.f;
                    hi[e] = __float_as_uint(tv);
                    lo[e] = __float_as_uint(tv - __uint_as_float(hi[e] & 0xFFFFE000u));
                }
                TC_ST8(lane_base + A_COL + s * 2 * KT + 8 * g8, hi, 0);
                TC_ST8(lane_base + A_COL + s * 2 * KT + KT + 8 * g8, lo, 0);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(&afull[s]);
#ifdef TC_DEBUG_SPIN
            const long long tb3 = clock64();
            dbg2 += tb3 - tb2;
#endif
            // lse / gout of THIS block were issued a whole build ago: complete them, then stage the next block (which
            // waits for the epilogue to free a stage: the value operand of this block is already with the MMA issuer)
            asm volatile("cp.async.wait_all;" ::: "memory");
            mbar_arrive(&lfull[it % T2B_LSTG]);
#ifdef TC_DEBUG_SPIN
            const long long tb4 = clock64();
            dbg4 += tb4 - tb3;
#endif
            if (more) stage_lg(blk + blk_step, it + 1);
#ifdef TC_DEBUG_SPIN
            dbg5 += clock64() - tb4;
#endif
        };
        float bufa[D], bufb[D], ba[TC_NB], bb[TC_NB];
        float a_ql = 0.f, a_qs = 1.f, b_ql = 0.f, b_qs = 1.f;                  // inline Q factor: lane d holds loc[u, d], scale[u, d]
        if (blk0 < n_blocks) { load_raw(blk0, bufa, ba, a_ql, a_qs); stage_lg(blk0, 0); }
        unsigned it = 0;
        for (unsigned blk = blk0; blk < n_blocks;) {
            step(blk, it, bufa, ba, a_ql, a_qs, bufb, bb, b_ql, b_qs);
            blk += blk_step; ++it;
            if (blk >= n_blocks) break;
            step(blk, it, bufb, bb, b_ql, b_qs, bufa, ba, a_ql, a_qs);
            blk += blk_step; ++it;
        }
    }

#ifdef TC_DEBUG_SPIN
    if ((blockIdx.x == 0 || blockIdx.x == 60 || blockIdx.x == 147) && lane == 0 && (warp == 0 || warp == 13 || warp == T2_MMA_WARP || warp == T2_MMA_WARP + 1))
        printf("tc2adj cta %d warp %d prologue %lld total %lld wait0 %lld wait1 %lld build %lld loads %lld waitall %lld stage %lld\n", (int)blockIdx.x, warp, dbg_t0 - dbg_k0, clock64() - dbg_t0, dbg0, dbg1, dbg2, dbg3, dbg4, dbg5);
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == T2_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(T2_TMEM_COLS) : "memory");
    }
}

template <int D>
static void launch_fan_lse_tc2_adj(const FanLseParams<float>& p, const Tc2Geom& geo, int blocks, cudaStream_t stream) {
    constexpr int KT = (2 * D + 2 + 7) / 8 * 8;
    constexpr size_t OPER = (size_t)(KT / 4) * 128 * 16;
    constexpr size_t FW = T2_TILES * 128;
    const size_t smem = T2_TILES * 2 * OPER + 2 * T2B_LSTG * T2_US * FW * 4 + 2 * FW * 4 + 2 * T2_EPI * 128 * 4 + 32 * 4 +
                        (2 * T2B_ASTG + 2 * T2_ACC + 2 * T2B_LSTG) * 8 + 16;
    static const cudaError_t attr = cudaFuncSetAttribute(fan_lse_tc2_adj_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // once per process
    (void)attr;
    fan_lse_tc2_adj_kernel<D><<<blocks, T2_WARPS * 32, smem, stream>>>(p, geo);
}

}  // namespace tc
