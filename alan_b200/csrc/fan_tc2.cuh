// fan_tc2.cuh -- fan_lse on tcgen05 for the case where the Normal's loc does not depend on the value's axes:
// the DENSE formulation (no block-diagonal operand, the builders touch every value once, not once per loc sample).
//
// Contract (fan_lse, fused.cuh):
//     out[rho, f] = LSE_eps_kappa( b[rho,kappa] - c[f] - sum_d (v[rho,kappa,d] - l[rho,d])^2 w[f,d] ) + cadd
// Here one rho dim `lam` is special: l depends on lam ONLY, and v and the small factors b do not depend on it
// (MovieLens-shaped models: v = z[m, K_z, d], l = mu[K_mu, d], w from psi[K_psi, d]: lam = K_mu, f = K_psi).
// With u = the remaining rho dims ("users") and the wide fan f' = (lam, f), expand the square around a per-d centre
// cd[d] = mean_lam l[lam,d]  (v' = v - cd, l' = l - cd: identical differences, small cross terms):
//
//     log2e S[u,kappa,f'] = sum_d v'^2 (-w log2e) + sum_d v' (2 l' w log2e) + b log2e        - C[f'] log2e
//                         =         B[(u,kappa), k]  .  A[f', k]                               (k < 2 D + 1)
//
//   A (constant per CTA)  rows f' (128 per M tile, up to 3 tiles per CTA = one "group" of 384 fan columns),
//                         K = 2 D + 1 -> KT;  A_hi and A_lo in TMEM (tcgen05.st, once)
//   B (one stage per block of 4 users)  rows n = (user slot, kappa) = 128,  { v'^2 | v' | b log2e }, hi and lo, smem
//   D = A B^T (TMEM, 2 stages x 128 columns): lane = f', columns = (user slot, kappa)
//   C[f'] = sum_d l'^2 w + sum_d log s + D/2 log 2pi does not depend on kappa: it leaves the LSE and is added in
//   fp32 afterwards.
// 3xTF32 as in fan_tc.cuh: A_hi B_hi + A_lo B_hi + A_hi B_lo.  15 tcgen05.mma (M128 N128 K8) per (tile, block) at
// D = 18, i.e. 120 per 4 users, against 30 per 16 rho (= 225 per 4 users x 30 lam) in the block-diagonal kernel, and
// the builders handle n_u Kk D values instead of n_u L Kk D.
//
// Roles (21 warps, one persistent CTA per SM; the CTAs are split between the fan groups in proportion to their cost):
//   epilogue  4 teams x 4 warps: warp % 4 = TMEM lane quadrant, team e owns user slot e (32 columns) of EVERY tile:
//             four epilogue warps per SM sub-partition keep the MUFU pipe busy through each other's serial sections.
//             forward: max (FMNMX3 tree) / ex2 / packed sums per (user, f'); adjoint: weights accumulated over the
//             tiles of the group in registers, one fixed-order butterfly per block, quadrant partials combined
//             through shared memory in a fixed order
//   MMA       1 warp (elected lane), owns the TMEM allocation
//   builders  4 warps: warp = user slot, lane = kappa; raw loads of the next block are in flight while the current
//             one is squared, split and stored
// The adjoint writes its partial over fan group g into gS[u, g, kappa] -- the layout the planner commits to when it
// finds the pattern (plan.py dense_fan_geometry; the reduce that follows sums NG partials per user) -- or, for plans
// built without it, into gS[u, lam = g, kappa] of the [rho, kappa] layout (the other lam slots stay zero).
//
// Measured on B200, cfg-5 (10 000 users, L = F = 30, D = 18), cycles per CTA (TC_DEBUG_SPIN): forward 240 k, of which
// the MMA issuer is busy 75 % (87 cycles per MMA) and the epilogue warps 90 %; adjoint 310 k, epilogue-bound.
#pragma once
#include "fan_tc.cuh"
#include <cuda_fp16.h>

namespace tc {

// kind::f16 twin of mma_tf32_ts (A in TMEM as packed fp16 pairs, B in shared memory as fp16): M = 128, K = 16
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// x = hi + lo with hi, lo fp16 (11 significant bits each: 22 bits of x, the dropped lo x lo product is 2^-22 relative);
// two values per call, packed the way the tensor core reads a 32-bit word (the lower K index in the lower half)
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 back = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - back.x, x1 - back.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// tcgen05.wait::ld that names the registers of the load it completes ("+r"): the compiler cannot hoist arithmetic on
// them above the wait (a plain asm volatile only orders against other volatile asm and memory accesses)
#define TC_WAIT_LD32(r)                                                                                        \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                              \
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),         \
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),   \
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), \
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])  \
                 :: "memory")

#define TC_LD16(r, addr)                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                    \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"              \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),       \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])  \
                 : "r"(addr) : "memory")
#define TC_WAIT_LD16(r)                                                                                        \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                              \
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),         \
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])    \
                 :: "memory")

constexpr int T2_ND = 4;          // user dims (rho without lam)
constexpr int T2_TILES = 3;       // M tiles (128 fan columns each) per group: A_hi and A_lo of three tiles fit TMEM next to
                                  // two accumulator stages (2 x 3 x 40 + 2 x 128 = 496 columns at D = 18).  Measured
                                  // alternatives: A_lo or all of A in shared memory (SS MMAs) cost 92 / 105 cycles per MMA
                                  // against 87 with A in TMEM -- the MMAs then compete for shared-memory bandwidth
constexpr int T2_MAXG = 8;        // fan groups
constexpr int T2_BLOCK_COST4 = 7; // per-block cost in quarter tiles (CTA split between groups of unequal tile counts)
constexpr int T2_US = 4;          // user slots per block: N = 32 T2_US
constexpr int T2_N = 32 * T2_US;
constexpr int T2_STAGES = 3;
constexpr int T2_ACC = 2;
constexpr int T2_BW = T2_US;      // builder warps: warp = user slot, lane = kappa
constexpr int T2_EPI = 4;         // epilogue teams of 4 warps (one warp per TMEM lane quadrant); 4 teams = 4 epilogue warps per SM
                                  // sub-partition: warp-level parallelism keeps the MUFU pipe fed through the serial sections
constexpr int T2_UPT = T2_US / T2_EPI;   // user slots (32 accumulator columns each) per team and tile
constexpr int T2_MMA_WARP = 4 * T2_EPI;
constexpr int T2_WARPS = 4 * T2_EPI + 1 + T2_BW;
constexpr int T2_TMEM_COLS = 512;

struct Tc2Geom {
    int sz[T2_ND];                          // user dims, right-aligned (unused: extent 1, stride 0)
    int vs[T2_ND], os[T2_ND], gs[T2_ND], ss[T2_ND];   // strides of v / out / gout / gS over the user dims
    int bs[TC_NB][T2_ND];
    int bk[TC_NB];
    float bc[TC_NB];
    int nb;
    int L, l_lam, o_lam, g_lam, s_lam;      // lam: extent and strides (l, out, gout, gS)
    int g_f;
    int n_u, NG, FP;                        // users, fan groups, wide fan extent L F
    int qn, qls[T2_ND], qss[T2_ND], q_lev, q_sev;   // inline Gaussian Q factor: strides of its loc / scale over the user dims
    float qc;
    int cta_lo[T2_MAXG + 1];                // CTAs [cta_lo[g], cta_lo[g + 1]) work on fan group g (proportional to its tiles)
    int vec2;
    int flat;                               // only the last user dim has an extent > 1
    int lg_vec4;                            // adjoint: lse / gout rows of a fan group are contiguous and 16-byte aligned
};

__device__ __forceinline__ void t2_decode(unsigned u, const Tc2Geom& g, int* idx) {
    if (g.flat) {                               // one user dim (the common case): no divisions
#pragma unroll
        for (int k = 0; k < T2_ND - 1; ++k) idx[k] = 0;
        idx[T2_ND - 1] = (int)u;
        return;
    }
#pragma unroll
    for (int k = T2_ND - 1; k >= 0; --k) {
        const unsigned sz = (unsigned)g.sz[k];
        const unsigned q = u / sz;
        idx[k] = (int)(u - q * sz);
        u = q;
    }
}
__device__ __forceinline__ int t2_dot(const int* idx, const int* st) {
    int o = 0;
#pragma unroll
    for (int k = 0; k < T2_ND; ++k) o += idx[k] * st[k];
    return o;
}

// H16: the operands as fp16 pairs (hi, lo) instead of 3xTF32 -- products hi.hi + lo.hi + hi.lo as kind::f16 MMAs of
// K = 16: 9 MMAs per (tile, block) at D = 18 (K = 48) instead of 15 (K = 40 in steps of 8).  The bias b log2e is carried
// by three K columns (A entries 4096, 1, 1; B entries q0 = fp16(b'/4096), q1 = fp16(b' - 4096 q0), q2 = the rest): 33 bits.
// STAG: the four epilogue teams work as two pairs, pair g on the accumulator stage g (every other tile), each team on two
// of the stage's four user slots.  With all teams on the same tile the warps of a sub-partition run in lockstep -- all
// in their issue-bound phase (TMEM read, max tree, sums, log, store), then all in their MUFU-bound phase (32 ex2 each) --
// and the two phases add up; two pairs half a period apart overlap one pair's ex2 with the other pair's issue phase.
// Measured at cfg-5 on one box (fan_lse forward, us): 3xTF32 140.1, STAG 138.5-128 (no gain: the MMA issuer is then the
// limit), H16 140.5 (no gain either: with all teams in lockstep the epilogue is), H16 + STAG 126.6 (-10 %).  A second group
// of builder warps (two blocks in flight) changed nothing: 139.1.
// fp16 range: plain fp16 operands overflow where 3xTF32 does not (|v - centre| > 255, scale < 0.005), so H16 scales the
// constant operand per CTA (2^-ea: max |A'| < 2^14) and the value rows per (block, user) (2^-eb: max v'^2 < 2^14; the
// usual case eb = 0 is a maximum, a vote and a uniform branch in the builder) and undoes both in the epilogue's first
// FFMA2 (d = r 2^(ea + eb) - m instead of the FADD2 r - m); 2^eb travels through a ring of shared-memory slots read once
// per block and team.  The padding rows' mask rides the spare K columns (A entries 32768 outside the scaling) so that it
// stays below every live column at any scale.  Limits: scale > ~4e-6 (beyond it the unit bias columns 2^-ea leave fp16:
// the kernel writes NaN, loudly).  Cost of the scaling, bisected at cfg-5 on one box (H16 + STAG, us): none 124.6,
// per-CTA only 125.3, both 128.4 with a ring read per tile and user (LDS shares the MIO queue with the saturated MUFU),
// ~127 with one read per block; against 3xTF32 on the same boxes 132.4-137.5: -2.5 .. -4 % instead of -6 .. -10 %.
// H16 + STAG stays opt-in (ALAN_B200_TC_F16=1 ALAN_B200_TC_STAG=1, D = 18; the whole GPU suite passes with both set).
template <int D, bool BWD, bool H16 = false, bool STAG = false>
__global__ void __launch_bounds__(T2_WARPS * 32, 1) fan_lse_tc2_kernel(const __grid_constant__ FanLseParams<float> p, const __grid_constant__ Tc2Geom geo) {
    constexpr int KT = H16 ? (2 * D + 3 + 15) / 16 * 16 : (2 * D + 1 + 7) / 8 * 8;   // K extent (40 at D = 18; 48 as fp16)
    constexpr int NC = H16 ? KT / 8 : KT / 4, KSTEPS = H16 ? KT / 16 : KT / 8;       // 16-byte chunks per row, MMAs per product
    static_assert(!H16 || KT - (2 * D + 3) >= 4, "fp16 operands: four spare K columns carry the mask of the padding rows");
    constexpr int ACOLS = H16 ? KT / 2 : KT;                                // TMEM columns of one A part of one tile
    constexpr uint32_t LBO = T2_N * 16, SBO = 8 * 16;                      // [chunk][128 rows][16 B]: A_lo tiles and B alike
    constexpr uint32_t OPER = NC * LBO;                                    // bytes of one operand part (20 KB at D = 18)
    constexpr uint32_t A_HI = 0, A_LO = T2_TILES * ACOLS, D_COL = 2 * T2_TILES * ACOLS;
    static_assert(2 * T2_TILES * ACOLS + T2_ACC * T2_N <= T2_TMEM_COLS, "TMEM budget");
    constexpr uint32_t IDESC = (1u << 4) | ((H16 ? 0u : 2u) << 7) | ((H16 ? 0u : 2u) << 10) | ((uint32_t)(T2_N >> 3) << 17) | ((128u >> 4) << 24);

    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char* stage_base = tc_smem;                                   // T2_STAGES x (B_hi | B_lo)
    unsigned char* tail = stage_base + T2_STAGES * 2 * OPER;
    float* s_cst = reinterpret_cast<float*>(tail);                         // [T2_TILES][128]  C[f'] (natural log units)
    int* s_ooff = reinterpret_cast<int*>(tail + T2_TILES * 128 * 4);       // [T2_TILES][128]  out offset of f' (-1: padding lane)
    int* s_goff = s_ooff + T2_TILES * 128;                                 // [T2_TILES][128]  gout offset of f' (forward: psum offset)
    float* s_red = reinterpret_cast<float*>(s_goff + T2_TILES * 128);      // [2][team][quadrant][user of team][32]
    float* s_cd = s_red + 2 * 4 * T2_US * 32;                                  // [32] centre per event element
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_cd + 32);
    uint64_t* full = bars;
    uint64_t* empty = bars + T2_STAGES;
    uint64_t* tfull = bars + 2 * T2_STAGES;
    uint64_t* tempty = bars + 2 * T2_STAGES + T2_ACC;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * T2_STAGES + 2 * T2_ACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long dbg_k0 = clock64(); (void)dbg_k0;
    unsigned long long dbg_g0 = 0; (void)dbg_g0;
#ifdef TC_DEBUG_SPIN
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(dbg_g0));
#endif
    const float LS = 1.4426950408889634f;
    const int Kk = p.Kk;
    const unsigned n_u = (unsigned)geo.n_u;
    const unsigned n_blocks = (n_u + T2_US - 1) / T2_US;
    int grp = 0;
#pragma unroll
    for (int g = 1; g < T2_MAXG; ++g) if (g < geo.NG && (int)blockIdx.x >= geo.cta_lo[g]) grp = g;
    const unsigned blk0 = blockIdx.x - geo.cta_lo[grp], blk_step = geo.cta_lo[grp + 1] - geo.cta_lo[grp];
    const int fp_lo = grp * (T2_TILES * 128);
    const int n_tiles = min(T2_TILES, (geo.FP - fp_lo + 127) / 128);

    // ---------------------------------------------------------------- prologue
    for (uint32_t i = threadIdx.x; i < (T2_STAGES * 2) * OPER / 16; i += blockDim.x)
        reinterpret_cast<float4*>(tc_smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        // centre per event element = mean over lam of the loc: the L x D loc values are fetched by all threads at once
        // (one global round trip instead of L dependent ones: the loop below used to cost ~15 us of a ~26 us prologue)
        // into the still unused stage memory, then lane d adds its column in lam order (fixed order: reproducible)
        float* stage_f = reinterpret_cast<float*>(tc_smem);
        __syncthreads();                                                   // the zero fill above is complete
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
        for (int i = threadIdx.x; i < geo.L * D; i += blockDim.x) stage_f[i] = 0.f;     // back to the zero fill
    }
    __syncthreads();
    {   // padding rows (kappa >= Kk) carry a hugely negative bias: no column mask in the epilogue
        const float big = -1.0e30f, bh = __uint_as_float(to_tf32(big)), bl = big - bh;
        const int npad = 32 - Kk, per_stage = T2_US * npad;
        for (int i = threadIdx.x; i < T2_STAGES * per_stage; i += blockDim.x) {
            const int s = i / per_stage, r = i - s * per_stage, us = r / npad, kz = Kk + r - us * npad;
            if (H16) {
                // the first bias piece (its A entry is 4096): -60000 x 4096 = -2.5e8 in log2 units, and the spare K columns
                // (A entries 32768, outside the per-CTA scaling; live rows hold 0 there): -60000 x 32768 each.  With the
                // operands scaled to |A'| < 2^14, v'^2 < 2^14 a live column is above -(D 2^28 + D 2^21 + 2^28) = -5.2e9
                // in the accumulator's units, the padding rows sit at -(KT - 2 D - 3) x 2.0e9 (-1.8e10 at D = 18)
                __half* bhh = reinterpret_cast<__half*>(stage_base + (size_t)s * 2 * OPER);
                bhh[(((2 * D) / 8) * T2_N + 32 * us + kz) * 8 + ((2 * D) & 7)] = __float2half_rn(-60000.f);
                for (int k = 2 * D + 3; k < KT; ++k)
                    bhh[((k / 8) * T2_N + 32 * us + kz) * 8 + (k & 7)] = __float2half_rn(-60000.f);
            } else {
                const int off = (((2 * D) / 4) * T2_N + 32 * us + kz) * 4 + ((2 * D) & 3);
                reinterpret_cast<float*>(stage_base + (size_t)s * 2 * OPER)[off] = bh;
                reinterpret_cast<float*>(stage_base + (size_t)s * 2 * OPER + OPER)[off] = bl;
            }
        }
    }
    if (threadIdx.x == 0) {
        if (H16) reinterpret_cast<int*>(s_red)[33] = 0;                    // max |A| of this CTA (the adjoint's s_red is idle here)
        for (int s = 0; s < T2_STAGES; ++s) { mbar_init(&full[s], 32 * T2_BW); mbar_init(&empty[s], 1); }
        for (int a = 0; a < T2_ACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], STAG ? 128 * (T2_EPI / 2) : 128 * T2_EPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == T2_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(T2_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    long long dbg0 = 0, dbg1 = 0, dbg2 = 0, dbg3 = 0; (void)dbg0; (void)dbg1; (void)dbg2; (void)dbg3;
    const long long dbg_t0 = clock64(); (void)dbg_t0;

    if (warp < 4 * T2_TILES) {
        // ---------------------------------------------------------------- A operand, once: thread = one row of one tile (a warp
        // reaches the TMEM lane quadrant warp % 4; 12 warps side by side: this prologue is the fixed cost that small
        // problems and small shards see)
        const int row = 32 * (warp & 3) + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        {
            const int tl = warp >> 2;
            const int fp = fp_lo + 128 * tl + row;
            float a[KT];
#pragma unroll
            for (int k = 0; k < KT; ++k) a[k] = 0.f;
            float cst = 0.f;
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
                a[2 * D] = H16 ? 4096.f : 1.f;
                if (H16) { a[2 * D + 1] = 1.f; a[2 * D + 2] = 1.f; }
                cst = -(c + float(D) * float(HALF_LOG_2PI));
                ooff = lam * geo.o_lam + f * (int)p.o_f;
                goff = BWD ? lam * geo.g_lam + f * geo.g_f : lam * (int)p.ps_lam + f * (int)p.ps_f;
            }
            s_cst[tl * 128 + row] = cst;
            s_ooff[tl * 128 + row] = ooff;
            s_goff[tl * 128 + row] = goff;
            if (H16) {
                // fp16 range: the constant operand of this CTA is scaled by a power of two so that max |A| < 2^14 (small
                // scales give 1 / (2 s^2) beyond 65504); the epilogue undoes it inside its first FFMA2
                float rmax = 0.f;
#pragma unroll
                for (int k = 0; k < KT; ++k) rmax = fmaxf(rmax, fabsf(a[k]));
                rmax = warp_max(rmax);
                if (lane == 0) atomicMax(reinterpret_cast<int*>(s_red) + 33, __float_as_int(rmax));
                asm volatile("bar.sync 7, %0;" :: "n"(128 * T2_TILES) : "memory");
                const float amax = __int_as_float(reinterpret_cast<int*>(s_red)[33]);
                const int ea = amax > 16384.f ? ilogbf(amax) - 13 : 0;
                const float sa = exp2f((float)-ea);
#pragma unroll
                for (int k = 0; k < KT; ++k) a[k] = k < 2 * D + 3 ? a[k] * sa : 32768.f;   // spare columns: the padding rows' mask
                if (warp == 0 && lane == 0) s_red[32] = ea <= 24 ? exp2f((float)ea) : NAN;    // 2^-ea must stay an fp16 value
            }
#pragma unroll
            for (int g8 = 0; g8 < ACOLS / 8; ++g8) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (H16) {
                        split_f16x2(a[16 * g8 + 2 * e], a[16 * g8 + 2 * e + 1], hi[e], lo[e]);
                    } else {
                        const float x = a[8 * g8 + e];
                        hi[e] = to_tf32(x);
                        lo[e] = __float_as_uint(x - __uint_as_float(hi[e]));
                    }
                }
                TC_ST8(lane_base + A_HI + ACOLS * tl + 8 * g8, hi, 0);
                TC_ST8(lane_base + A_LO + ACOLS * tl + 8 * g8, lo, 0);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp < 4 * T2_EPI) {
        // ---------------------------------------------------------------- epilogue: team = warp / 4, quadrant = warp % 4
        const int q = warp & 3, team = warp >> 2, row = 32 * q + lane;
        const float cadd = p.cadd;
        // adjoint: raw lse / gout of the NEXT block are fetched one block ahead (no arithmetic on them before their tile:
        // an in-order warp stalls on the first use of a pending load)
        static_assert(!STAG || (!BWD && T2_ACC == 2 && T2_EPI == 4 && T2_US == 4), "staggered epilogue: forward, two stages, four teams");
        constexpr int UPT = STAG ? 2 : T2_UPT;                                   // user slots per team and tile
        const int ubase = STAG ? 2 * (team & 1) : T2_UPT * team;                 // first user slot of this team
        const unsigned my_stage = (unsigned)(team >> 1);                         // STAG: this team's accumulator stage
        float n_lse[T2_TILES][UPT], n_g[T2_TILES][UPT];
        int n_uoff[UPT];
        auto fetch = [&](unsigned blk) {
#pragma unroll
            for (int uu = 0; uu < UPT; ++uu) {
                const unsigned u = T2_US * blk + ubase + uu;
                int idx[T2_ND];
                t2_decode(u < n_u ? u : 0u, geo, idx);
                const int uo = t2_dot(idx, geo.os);
                n_uoff[uu] = u < n_u ? uo : -1;
                if (BWD) {
                    const int ug = t2_dot(idx, geo.gs);
#pragma unroll
                    for (int tl = 0; tl < T2_TILES; ++tl) {
                        const int oo = s_ooff[tl * 128 + row];
                        n_lse[tl][uu] = INFINITY; n_g[tl][uu] = 0.f;              // idle rows / users: weight 0
                        if (u < n_u && oo >= 0 && tl < n_tiles) {
                            n_lse[tl][uu] = p.lse[uo + oo];
                            n_g[tl][uu] = p.gout[ug + s_goff[tl * 128 + row]];
                        }
                    }
                }
            }
        };
        if (blk0 < n_blocks) fetch(blk0);
        const float sc_a = H16 ? s_red[32] : 1.f;                                // 2^ea of this CTA's constant operand
        float scu[UPT];                                                          // H16: 2^(ea + eb) of this team's users, current block
        unsigned sc_it = ~0u;
#pragma unroll
        for (int uu = 0; uu < UPT; ++uu) scu[uu] = 1.f;
        unsigned tt = 0;                                                         // accumulator stage counter
        unsigned it = 0;
        float ps[T2_TILES];                                                      // forward: running sum of out over this team's users
#pragma unroll
        for (int tl = 0; tl < T2_TILES; ++tl) ps[tl] = 0.f;
        const uint32_t ld_base = tmem + ((uint32_t)(32 * q) << 16) + D_COL + 32 * ubase;
        for (unsigned blk = blk0; blk < n_blocks; blk += blk_step, ++it) {
            float lz[T2_TILES][UPT], gz[T2_TILES][UPT];
            int uoff[UPT];
#pragma unroll
            for (int uu = 0; uu < UPT; ++uu) {
                uoff[uu] = n_uoff[uu];
#pragma unroll
                for (int tl = 0; tl < T2_TILES; ++tl) { lz[tl][uu] = BWD ? n_lse[tl][uu] : 0.f; gz[tl][uu] = BWD ? n_g[tl][uu] : 0.f; }
            }
            if (blk + blk_step < n_blocks) fetch(blk + blk_step);
            float acc[BWD ? UPT : 1][32];
            if (BWD) {
#pragma unroll
                for (int uu = 0; uu < UPT; ++uu)
#pragma unroll
                    for (int k = 0; k < 32; ++k) acc[uu][k] = 0.f;
            }
#pragma unroll
            for (int tl = 0; tl < T2_TILES; ++tl) {
                if (tl < n_tiles) {
                    const int a = tt % T2_ACC;
                    const uint32_t pa = (tt / T2_ACC) & 1;
                    ++tt;
                    if (STAG && (unsigned)a != my_stage) continue;              // the other pair of teams takes this tile
                    MBAR_WAIT(&tfull[a], pa, dbg0);
                    tc_fence_after();
                    if (!BWD) {
                        // H16: the accumulator carries the powers of two of the operand scalings (A per CTA, B per block and
                        // user), undone inside the subtraction of the maximum (an FFMA2 for the FADD2).  The factors of this
                        // team's users are read once per block, at its first tile, ahead of the TMEM loads (a shared-memory
                        // read per tile and user measured 3 us at cfg-5: LDS shares the MIO queue with the saturated MUFU)
                        if (H16 && sc_it != it) {
                            sc_it = it;
                            if (UPT == 2) {
                                const float2 v2 = *reinterpret_cast<const float2*>(&s_red[(it & 7) * T2_US + ubase]);
                                scu[0] = sc_a * v2.x; scu[UPT - 1] = sc_a * v2.y;
                            } else {
#pragma unroll
                                for (int uu = 0; uu < UPT; ++uu) scu[uu] = sc_a * s_red[(it & 7) * T2_US + ubase + uu];
                            }
                        }
                        // all loads of this team first, then the accumulator stage goes straight back to the MMA issuer
                        uint32_t r[UPT][32];
#pragma unroll
                        for (int uu = 0; uu < UPT; ++uu) TC_LD32(r[uu], ld_base + T2_N * a + 32 * uu);
#pragma unroll
                        for (int uu = 0; uu < UPT; ++uu) TC_WAIT_LD32(r[uu]);
                        tc_fence_before();
                        mbar_arrive(&tempty[a]);
#pragma unroll
                        for (int uu = 0; uu < UPT; ++uu) {
#ifdef TC_EXP_NOEPI
                            if (uoff[uu] >= 0 && s_ooff[tl * 128 + row] >= 0) { p.out[uoff[uu] + s_ooff[tl * 128 + row]] = __uint_as_float(r[uu][0]) + __uint_as_float(r[uu][31]); }
                            continue;
#endif
                            // max as a tree of three-input maxima (FMNMX3), then packed subtract / accumulate around the 32 ex2
                            float m8[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                m8[k] = fmaxf(fmaxf(__uint_as_float(r[uu][4 * k]), __uint_as_float(r[uu][4 * k + 1])),
                                              fmaxf(__uint_as_float(r[uu][4 * k + 2]), __uint_as_float(r[uu][4 * k + 3])));
                            const float sc = scu[uu];
                            const float2 sc2 = make_float2(sc, sc);
                            const float m = sc * fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
                            const float2 nm2 = make_float2(-m, -m);
                            float2 s2[4];
#pragma unroll
                            for (int c = 0; c < 4; ++c) s2[c] = make_float2(0.f, 0.f);
#pragma unroll
                            for (int k = 0; k < 32; k += 2) {
                                const float2 rk = make_float2(__uint_as_float(r[uu][k]), __uint_as_float(r[uu][k + 1]));
                                const float2 d2 = H16 ? __ffma2_rn(rk, sc2, nm2) : __fadd2_rn(rk, nm2);
                                s2[(k >> 1) & 3] = __fadd2_rn(s2[(k >> 1) & 3], make_float2(FastExp<float>::ex(d2.x), FastExp<float>::ex(d2.y)));
                            }
                            const float2 t2 = __fadd2_rn(__fadd2_rn(s2[0], s2[1]), __fadd2_rn(s2[2], s2[3]));
                            const float sum = t2.x + t2.y;
                            const int oo = s_ooff[tl * 128 + row];
                            if (uoff[uu] >= 0 && oo >= 0) {
                                // sum is in [1, 32] (max-shifted): lg2.approx is within 2^-21 absolute there, against values of
                                // magnitude 10-100 whose own fp32 spacing is 1e-6 .. 8e-6; libdevice logf costs ~25 instructions
                                float l2;
                                asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(sum + Eps<float>::v()));
                                const float val = (l2 + m) * 0.6931471805599453f + (cadd + s_cst[tl * 128 + row]);
                                p.out[uoff[uu] + oo] = val;
                                ps[tl] += val;
                            }
                        }
                    } else {
                        // 16 columns at a time: the adjoint also holds 32 accumulators per user and the prefetched lse / gout
#pragma unroll
                        for (int uu = 0; uu < T2_UPT; ++uu) {
                            const float nl = (cadd + s_cst[tl * 128 + row] - lz[tl][uu]) * LS;
                            const float2 nl2 = make_float2(nl, nl), g2 = make_float2(gz[tl][uu], gz[tl][uu]);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                uint32_t r[16];
                                TC_LD16(r, ld_base + T2_N * a + 32 * uu + 16 * h);
                                TC_WAIT_LD16(r);
                                if (uu == T2_UPT - 1 && h == 1) { tc_fence_before(); mbar_arrive(&tempty[a]); }
#pragma unroll
                                for (int k = 0; k < 16; k += 2) {
                                    const float2 d0 = __fadd2_rn(make_float2(__uint_as_float(r[k]), __uint_as_float(r[k + 1])), nl2);
                                    const float2 e0 = make_float2(FastExp<float>::ex(d0.x), FastExp<float>::ex(d0.y));
                                    const float2 a2 = __ffma2_rn(e0, g2, make_float2(acc[uu][16 * h + k], acc[uu][16 * h + k + 1]));
                                    acc[uu][16 * h + k] = a2.x; acc[uu][16 * h + k + 1] = a2.y;
                                }
                            }
                        }
                    }
                }
            }
            if (BWD) {
                // sum over the 32 lanes (f') of this warp: fixed-order butterfly reduce-scatter, lane j ends with kappa = j;
                // then the four quadrant warps of the team combine through shared memory in a fixed order
                float* red = s_red + ((it & 1) * T2_EPI + team) * (4 * T2_UPT * 32);
#pragma unroll
                for (int uu = 0; uu < T2_UPT; ++uu) {
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
                        for (int i = 0; i < off; ++i) {
                            const bool up = (lane & off) != 0;
                            const float send = up ? acc[uu][i] : acc[uu][i + off];
                            const float mine = up ? acc[uu][i + off] : acc[uu][i];
                            acc[uu][i] = mine + __shfl_xor_sync(0xffffffffu, send, off);
                        }
                    }
                    red[(q * T2_UPT + uu) * 32 + lane] = acc[uu][0];
                }
                asm volatile("bar.sync %0, 128;" :: "r"(1 + team) : "memory");
                if (q < T2_UPT) {
                    const int uu = q;
                    const float* rp = red + uu * 32 + lane;
                    const float sum = ((rp[0] + rp[T2_UPT * 32]) + rp[2 * T2_UPT * 32]) + rp[3 * T2_UPT * 32];
                    const unsigned u = T2_US * blk + T2_UPT * team + uu;
                    if (u < n_u && lane < Kk) {
                        int idx[T2_ND];
                        t2_decode(u, geo, idx);
                        p.gS[((i64)t2_dot(idx, geo.ss) + (i64)grp * geo.s_lam) * Kk + lane] = sum;
                    }
                }
            }
        }
        if (!BWD && p.psum_rows > 0) {
            // fused plate sum: the four teams' sums are combined in a fixed order through shared memory (the B stages
            // are idle by now: every MMA that read them has been consumed above), row `CTA` of the partial buffer gets
            // the sums for the fan columns of this CTA's group and zeros for the other groups' columns; the rows no
            // CTA owns are zeroed round-robin.  The reduce that follows adds the rows in order.
            float* scratch = reinterpret_cast<float*>(stage_base);               // [team][tile][128]
#pragma unroll
            for (int tl = 0; tl < T2_TILES; ++tl) scratch[(team * T2_TILES + tl) * 128 + row] = ps[tl];
            asm volatile("bar.sync 5, %0;" :: "n"(128 * T2_EPI) : "memory");
            if (team == 0) {
                float* prow = p.psum + (i64)blockIdx.x * p.ps_row;
#pragma unroll
                for (int tl = 0; tl < T2_TILES; ++tl) {
                    if (tl < n_tiles && s_ooff[tl * 128 + row] >= 0) {
                        float v = scratch[tl * 128 + row];
#pragma unroll
                        for (int e = 1; e < T2_EPI; ++e) v += scratch[(e * T2_TILES + tl) * 128 + row];
                        prow[s_goff[tl * 128 + row]] = v;
                    }
                }
            }
            // zeros: the other groups' columns of this CTA's row, and whole rows >= gridDim.x (round-robin over CTAs)
            const int tid = team * 128 + row;
            for (int r0 = (int)blockIdx.x; r0 < p.psum_rows; r0 += (int)gridDim.x) {
                float* zrow = p.psum + (i64)r0 * p.ps_row;
                const bool own = r0 == (int)blockIdx.x;
                for (int fp = tid; fp < geo.FP; fp += 128 * T2_EPI) {
                    if (own && fp >= fp_lo && fp < fp_lo + 128 * n_tiles) continue;
                    const int lam = fp / p.F, f = fp - lam * p.F;
                    zrow[lam * (int)p.ps_lam + f * (int)p.ps_f] = 0.f;
                }
            }
        }
    } else if (warp == T2_MMA_WARP) {
        // ---------------------------------------------------------------- MMA issuer
        unsigned it = 0, tt = 0;
        for (unsigned blk = blk0; blk < n_blocks; blk += blk_step, ++it) {
            const int s = it % T2_STAGES;
            const uint32_t ps = (it / T2_STAGES) & 1;
            MBAR_WAIT(&full[s], ps, dbg0);
            const uint32_t bhi = smem_u32(stage_base + (size_t)s * 2 * OPER), blo = bhi + OPER;
            for (int tl = 0; tl < n_tiles; ++tl, ++tt) {
                const int a = tt % T2_ACC;
                const uint32_t pa = (tt / T2_ACC) & 1;
                MBAR_WAIT(&tempty[a], pa ^ 1, dbg1);
                tc_fence_after();
                const uint32_t d = tmem + D_COL + T2_N * a;
                const uint32_t ahi = tmem + A_HI + ACOLS * tl, alo = tmem + A_LO + ACOLS * tl;
                if (elect_one()) {
#ifdef TC_EXP_NOMMA
                    constexpr int KS_ = 1;
#else
                    constexpr int KS_ = KSTEPS;
#endif
#pragma unroll
                    for (int j = 0; j < KS_; ++j) {
                        const uint64_t dh = smem_desc(bhi + j * 2 * LBO, LBO, SBO), dl = smem_desc(blo + j * 2 * LBO, LBO, SBO);
                        if (H16) {
                            mma_f16_ts(d, ahi + 8 * j, dh, IDESC, j > 0 ? 1u : 0u);
                            mma_f16_ts(d, alo + 8 * j, dh, IDESC, 1u);
                            mma_f16_ts(d, ahi + 8 * j, dl, IDESC, 1u);
                        } else {
                            mma_tf32_ts(d, ahi + 8 * j, dh, IDESC, j > 0 ? 1u : 0u);
                            mma_tf32_ts(d, alo + 8 * j, dh, IDESC, 1u);
                            mma_tf32_ts(d, ahi + 8 * j, dl, IDESC, 1u);
                        }
                    }
                    if (tl == n_tiles - 1) tc_commit(&empty[s]);
                    tc_commit(&tfull[a]);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------------------------------------------------------- builders: warp = user slot, lane = kappa; the raw
        // values of the NEXT block are loaded before the current one is written, so the global-load latency overlaps the
        // wait for the stage and the shared-memory stores
        const int us = warp - T2_MMA_WARP - 1, kz = lane, n = 32 * us + kz;
        const int vk = (int)p.v_k, vev = (int)p.v_ev, nb = geo.nb, vec2 = geo.vec2;
        // raw loads only (no arithmetic on the loaded values here: an in-order warp would stall on the first use)
        float cur[D], nxt[D], cur_b[TC_NB], nxt_b[TC_NB];
        float cur_ql = 0.f, cur_qs = 1.f, nxt_ql = 0.f, nxt_qs = 1.f;          // inline Q factor: lane d holds loc[u, d], scale[u, d]
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
        if (blk0 < n_blocks) load_raw(blk0, cur, cur_b, cur_ql, cur_qs);
        unsigned it = 0;
        for (unsigned blk = blk0; blk < n_blocks; blk += blk_step, ++it) {
            const int s = it % T2_STAGES;
            const uint32_t ps = (it / T2_STAGES) & 1;
#ifdef TC_DEBUG_SPIN
            const long long tb0 = clock64();
#endif
            if (blk + blk_step < n_blocks) load_raw(blk + blk_step, nxt, nxt_b, nxt_ql, nxt_qs);
            MBAR_WAIT(&empty[s], ps ^ 1, dbg0);
#ifdef TC_DEBUG_SPIN
            const long long tb1 = clock64();
#endif
#ifndef TC_EXP_NOBUILD                   // timing experiment: the builders arrive without building (operands stay zero)
            // inline Gaussian Q factor of this (user, kappa): lane d contributes 1 / (2 s_d^2) and log s_d, every lane
            // walks the D shuffled pairs in d order (fixed order) against its own value row
            float qsum = 0.f;
            if (geo.qn) {
                const float iv2 = 0.5f / (cur_qs * cur_qs);
                float lg = lane < D ? logf(cur_qs) : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
#pragma unroll
                for (int dd = 0; dd < D; ++dd) {
                    const float l_d = __shfl_sync(0xffffffffu, cur_ql, dd), w_d = __shfl_sync(0xffffffffu, iv2, dd);
                    const float df = cur[dd] - l_d;
                    qsum = fmaf(-(df * df), w_d, qsum);
                }
                qsum -= lg + float(D) * float(HALF_LOG_2PI);
            }
            float bsc = 1.f;
            float dfv[D];
            if (H16) {
                // fp16 range: the rows of this user (all kappa) are scaled by a power of two so that max v'^2 < 2^14; the
                // factor goes to the epilogue through a ring of slots indexed by the block counter.  The usual case (every
                // |v'| < 128: no scaling) costs the maximum, a vote and a uniform branch; otherwise exponent arithmetic on
                // the bit patterns (non-negative floats order as unsigned integers: one REDUX for the warp maximum)
                float dmax = 0.f;
#pragma unroll
                for (int dd = 0; dd < D; ++dd) { dfv[dd] = cur[dd] - s_cd[dd]; dmax = fmaxf(dmax, fabsf(dfv[dd])); }
                float unsc = 1.f;
                if (__any_sync(0xffffffffu, kz < Kk && !(dmax < 128.f))) {
                    const float dm = __uint_as_float(__reduce_max_sync(0xffffffffu, kz < Kk ? __float_as_uint(dmax) : 0u));
                    const int eb = max((int)(__float_as_uint(dm * dm) >> 23) - 127 - 13, 0);  // floor(log2 max v'^2) - 13
                    bsc = __uint_as_float((unsigned)(127 - eb) << 23);
                    unsc = __uint_as_float((unsigned)(127 + eb) << 23);
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) dfv[dd] *= bsc;
                }
                if (lane == 0) s_red[(it & 7) * T2_US + us] = unsc;
            }
            if (kz < Kk) {                       // rows of users >= n_u are written as zeros: finite, masked later
                float* bh = reinterpret_cast<float*>(stage_base + (size_t)s * 2 * OPER);
                float* bl = reinterpret_cast<float*>(stage_base + (size_t)s * 2 * OPER + OPER);
                float bsum = geo.qc * qsum;
#pragma unroll
                for (int i = 0; i < TC_NB; ++i) if (i < nb) bsum += geo.bc[i] * cur_b[i];
                bsum *= LS;
                if (H16) {
                    bsum *= bsc;
                    // bias pieces: b' = 4096 q0 + q1 + q2, every piece an fp16 value (their lo parts are exactly zero)
                    const float q0 = __half2float(__float2half_rn(bsum * (1.f / 4096.f)));
                    const float r1 = fmaf(-4096.f, q0, bsum);
                    const float q1 = __half2float(__float2half_rn(r1));
                    const float q2 = r1 - q1;
                    unsigned char* bhb = stage_base + (size_t)s * 2 * OPER;
                    unsigned char* blb = bhb + OPER;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        float tv[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int k = 8 * c + e;
                            const int dd = k < D ? k : (k < 2 * D ? k - D : 0);
                            tv[e] = k < D ? (cur[dd] - s_cd[dd]) * dfv[dd] : k < 2 * D ? dfv[dd] : k == 2 * D ? q0 : k == 2 * D + 1 ? q1 : k == 2 * D + 2 ? q2 : 0.f;
                        }
                        uint4 hh, l;
                        split_f16x2(tv[0], tv[1], hh.x, l.x);
                        split_f16x2(tv[2], tv[3], hh.y, l.y);
                        split_f16x2(tv[4], tv[5], hh.z, l.z);
                        split_f16x2(tv[6], tv[7], hh.w, l.w);
                        const int off = (c * T2_N + n) * 16;
                        *reinterpret_cast<uint4*>(bhb + off) = hh;
                        *reinterpret_cast<uint4*>(blb + off) = l;
                    }
                } else {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    // the tensor core reads only the 19 TF32 bits of a word: the raw fp32 value IS the "hi" part and
                    // lo = t - trunc(t) its exact remainder
                    float tv[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int k = 4 * c + e;
                        const float df = cur[k < D ? k : (k < 2 * D ? k - D : 0)] - s_cd[k < D ? k : (k < 2 * D ? k - D : 0)];
                        tv[e] = k < D ? df * df : k < 2 * D ? df : k == 2 * D ? bsum : 0.f;
                    }
                    float4 hh, l;
                    hh.x = tv[0]; l.x = hh.x - __uint_as_float(__float_as_uint(hh.x) & 0xFFFFE000u);
                    hh.y = tv[1]; l.y = hh.y - __uint_as_float(__float_as_uint(hh.y) & 0xFFFFE000u);
                    hh.z = tv[2]; l.z = hh.z - __uint_as_float(__float_as_uint(hh.z) & 0xFFFFE000u);
                    hh.w = tv[3]; l.w = hh.w - __uint_as_float(__float_as_uint(hh.w) & 0xFFFFE000u);
                    const int off = (c * T2_N + n) * 4;
                    *reinterpret_cast<float4*>(bh + off) = hh;
                    *reinterpret_cast<float4*>(bl + off) = l;
                }
                }
            }
#endif
#ifdef TC_DEBUG_SPIN
            const long long tb2 = clock64();
#endif
            fence_async_smem();
            mbar_arrive(&full[s]);
#pragma unroll
            for (int dd = 0; dd < D; ++dd) cur[dd] = nxt[dd];
#pragma unroll
            for (int i = 0; i < TC_NB; ++i) cur_b[i] = nxt_b[i];
            cur_ql = nxt_ql; cur_qs = nxt_qs;
#ifdef TC_DEBUG_SPIN
            dbg1 += tb2 - tb1;                   // build + store
            dbg2 += clock64() - tb2;             // fence + arrive
            dbg3 += tb1 - tb0;                   // load issue + wait for the stage
#endif
        }
    }

#ifdef TC_DEBUG_SPIN
    if ((blockIdx.x == 0 || blockIdx.x == 60 || blockIdx.x == 120 || blockIdx.x == 147) && lane == 0 && (warp == 0 || warp == T2_MMA_WARP || warp == T2_MMA_WARP + 1)) {
        unsigned long long g1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
        printf("tc2 bwd=%d cta %d warp %d prologue %lld total %lld wait0 %lld wait1 %lld d2 %lld d3 %lld ns %llu\n", (int)BWD, (int)blockIdx.x, warp, dbg_t0 - dbg_k0, clock64() - dbg_t0, dbg0, dbg1, dbg2, dbg3, g1 - dbg_g0);
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == T2_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(T2_TMEM_COLS) : "memory");
    }
}

template <int D>
static void launch_fan_lse_tc2_adj(const FanLseParams<float>& p, const Tc2Geom& geo, int blocks, cudaStream_t stream);   // fan_tc2b.cuh

// index of the lam dim in p.rd, -1 if the loc is a constant vector (L = 1), -2 if the dense formulation does not apply
static int fan_lse_tc2_lam(const FanLseParams<float>& p) {
    if (p.l_k != 0) return -2;
    int lam = -1;
    for (int k = 0; k < p.rd.nd; ++k) {
        if (p.lstride[k] == 0 || p.rd.size[k] == 1) continue;
        if (lam >= 0) return -2;
        lam = k;
    }
    if (lam < 0) return -1;
    if (p.vstride[lam] != 0) return -2;
    for (int i = 0; i < p.nb; ++i) if (p.bstride[i][lam] != 0) return -2;
    return lam;
}

static bool fan_lse_tc2_supported(const FanLseParams<float>& p, int D, bool bwd) {
    const int lam = fan_lse_tc2_lam(p);
    if (lam == -2) return false;
    const i64 L = lam >= 0 ? p.rd.size[lam] : 1;
    const i64 FP = L * p.F, NG = (FP + T2_TILES * 128 - 1) / (T2_TILES * 128);
    const i64 n_u = p.n_rho / L;
    if (!bwd && p.psum_rows > 0 && p.psum_rows < NG) return false;
    // the adjoint writes one partial per fan group: only the planner-committed compact gS layout [users, NG, kappa]
    // is covered completely (nothing zero-fills adjoint tensors any more: plan.py build_backward)
    if (bwd && p.gs_compact == 0) return false;
    if (p.qn && (lam >= 0 && (p.q_lstride[lam] != 0 || p.q_sstride[lam] != 0))) return false;
    if (FP < 96 || NG > T2_MAXG || (bwd && NG > L) || (bwd && p.gs_compact > 0 && p.gs_compact != NG) || p.Kk > 32 || n_u < 16 || p.nb > TC_NB) return false;
    if (p.rd.nd - (lam >= 0 ? 1 : 0) > T2_ND) return false;
    const i64 lim = (i64)1 << 31;
    i64 vspan = (i64)p.Kk * p.v_k + 32 * p.v_ev, ospan = FP * (p.o_f > 0 ? p.o_f : 1), bspan = 0;
    for (int k = 0; k < p.rd.nd; ++k) {
        vspan += (i64)p.rd.size[k] * p.vstride[k];
        ospan += (i64)p.rd.size[k] * p.ostride[k];
        for (int i = 0; i < p.nb; ++i) { i64 b = (i64)p.rd.size[k] * p.bstride[i][k] + (i64)p.Kk * p.b_k[i]; if (b > bspan) bspan = b; }
    }
    if (vspan >= lim || ospan >= lim || bspan >= lim || p.n_rho * (i64)p.Kk >= lim) return false;
    switch (D) { case 2: case 4: case 6: case 8: case 12: case 16: case 18: return true; }
    return false;
}

template <int D>
static int launch_fan_lse_tc2_D(const FanLseParams<float>& p, bool bwd, cudaStream_t stream, int sm_count) {
    constexpr int KT = (2 * D + 1 + 7) / 8 * 8;
    constexpr size_t OPER = (size_t)(KT / 4) * T2_N * 16;
    constexpr size_t TAIL = T2_TILES * 128 * 12 + 2 * 4 * T2_US * 32 * 4 + 32 * 4 + (2 * T2_STAGES + 2 * T2_ACC) * 8 + 16;
    const size_t smem = (T2_STAGES * 2) * OPER + TAIL;
    // the fp16-pair formulation of the forward (9 MMAs per tile and block instead of 15): ALAN_B200_TC_F16=1
    const char* e16 = getenv("ALAN_B200_TC_F16");
    const bool h16 = e16 && atoi(e16) == 1;
    const int lam = fan_lse_tc2_lam(p);
    Tc2Geom geo;
    memset(&geo, 0, sizeof(geo));
    for (int j = 0; j < T2_ND; ++j) geo.sz[j] = 1;
    geo.L = lam >= 0 ? p.rd.size[lam] : 1;
    // gS is [rd dims row-major, kappa]: element strides of the rd dims in units of Kk
    i64 sstride[AB_MAXD];
    { i64 acc = 1; for (int k = p.rd.nd - 1; k >= 0; --k) { sstride[k] = acc; acc *= p.rd.size[k]; } }
    const int n_user_dims = p.rd.nd - (lam >= 0 ? 1 : 0);
    int j = T2_ND - n_user_dims;
    bool ev2 = (D % 2 == 0) && p.v_ev == 1 && p.v_k % 2 == 0 && ((uintptr_t)p.v % 8 == 0);
    for (int k = 0; k < p.rd.nd; ++k) {
        if (k == lam) {
            geo.l_lam = (int)p.lstride[k]; geo.o_lam = (int)p.ostride[k]; geo.g_lam = (int)p.gstride[k]; geo.s_lam = (int)sstride[k];
            continue;
        }
        geo.sz[j] = p.rd.size[k];
        geo.vs[j] = (int)p.vstride[k]; geo.os[j] = (int)p.ostride[k]; geo.gs[j] = (int)p.gstride[k]; geo.ss[j] = (int)sstride[k];
        for (int i = 0; i < p.nb; ++i) geo.bs[i][j] = (int)p.bstride[i][k];
        geo.qls[j] = (int)p.q_lstride[k]; geo.qss[j] = (int)p.q_sstride[k];
        ev2 = ev2 && (p.vstride[k] % 2 == 0);
        ++j;
    }
    geo.qn = p.qn; geo.q_lev = (int)p.q_lev; geo.q_sev = (int)p.q_sev; geo.qc = p.qn ? (float)p.q_coeff : 0.f;
    for (int i = 0; i < p.nb; ++i) { geo.bk[i] = (int)p.b_k[i]; geo.bc[i] = p.bcoeff[i]; }
    if (p.gs_compact > 0) {
        // planner-committed layout [users (row-major), fan group, kappa]: nothing but the partials is ever stored
        i64 acc = p.gs_compact;
        for (int jj = T2_ND - 1; jj >= T2_ND - n_user_dims; --jj) { geo.ss[jj] = (int)acc; acc *= geo.sz[jj]; }
        geo.s_lam = 1;
    }
    geo.nb = p.nb;
    geo.g_f = (int)p.g_f;
    geo.vec2 = ev2 ? 1 : 0;
    geo.flat = 1;
    for (int jj = 0; jj < T2_ND - 1; ++jj) if (geo.sz[jj] != 1) geo.flat = 0;
    geo.n_u = (int)(p.n_rho / geo.L);
    geo.FP = geo.L * p.F;
    geo.NG = (geo.FP + T2_TILES * 128 - 1) / (T2_TILES * 128);
    if (bwd) {
        bool v4 = p.o_f == 1 && geo.o_lam == p.F && p.g_f == 1 && geo.g_lam == p.F && geo.FP % 4 == 0 &&
                  (uintptr_t)p.lse % 16 == 0 && (uintptr_t)p.gout % 16 == 0;
        for (int jj = 0; jj < T2_ND; ++jj) v4 = v4 && (geo.sz[jj] == 1 || (geo.os[jj] % 4 == 0 && geo.gs[jj] % 4 == 0));
        geo.lg_vec4 = v4 ? 1 : 0;
    }
    const i64 n_blocks = ((i64)geo.n_u + T2_US - 1) / T2_US;
    // CTAs per fan group in proportion to the group's tiles (the last group may be short), each at most n_blocks
    const int tiles_total = (geo.FP + 127) / 128;
    int blocks = 0;
    {
        // weight of a group = its tiles + a per-block cost measured at ~1.75 tiles (builder + block bookkeeping)
        int max_cta = sm_count;
        if (!bwd && p.psum_rows > 0 && max_cta > p.psum_rows) max_cta = p.psum_rows;
        int left_cta = max_cta > geo.NG ? max_cta : geo.NG, left_tiles = tiles_total;
        int left_w = 4 * tiles_total + T2_BLOCK_COST4 * geo.NG;
        for (int g = 0; g < geo.NG; ++g) {
            int tg = tiles_total - T2_TILES * g; if (tg > T2_TILES) tg = T2_TILES;
            int n = (int)(((i64)left_cta * (4 * tg + T2_BLOCK_COST4) + left_w / 2) / left_w);
            if (n < 1) n = 1;
            if (n > left_cta - (geo.NG - 1 - g)) n = left_cta - (geo.NG - 1 - g);
            left_cta -= n; left_tiles -= tg; left_w -= 4 * tg + T2_BLOCK_COST4;
            if (n > n_blocks) n = (int)n_blocks;
            geo.cta_lo[g] = blocks;
            blocks += n;
        }
        for (int g = geo.NG; g <= T2_MAXG; ++g) geo.cta_lo[g] = blocks;
    }
    if (bwd) {
        launch_fan_lse_tc2_adj<D>(p, geo, blocks, stream);                 // transposed product: fan_tc2b.cuh
    } else {
        static const cudaError_t attr_false = cudaFuncSetAttribute(fan_lse_tc2_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // once per process
        (void)attr_false;
        const char* est = getenv("ALAN_B200_TC_STAG");
        const bool stag = est && atoi(est) == 1;
        if constexpr (D == 18) {
            constexpr int KT16 = (2 * D + 3 + 15) / 16 * 16;
            constexpr size_t OPER16 = (size_t)(KT16 / 8) * T2_N * 16;
            const size_t smem16 = (T2_STAGES * 2) * OPER16 + TAIL;
#define T2_LAUNCH_VARIANT(H, S, SM)                                                                                       \
            {                                                                                                             \
                static const cudaError_t av = cudaFuncSetAttribute(fan_lse_tc2_kernel<D, false, H, S>,                     \
                                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM)); \
                (void)av;                                                                                                 \
                fan_lse_tc2_kernel<D, false, H, S><<<blocks, T2_WARPS * 32, (SM), stream>>>(p, geo);                        \
                return 0;                                                                                                 \
            }
            if (h16 && stag) T2_LAUNCH_VARIANT(true, true, smem16)
            if (h16) T2_LAUNCH_VARIANT(true, false, smem16)
            if (stag) T2_LAUNCH_VARIANT(false, true, smem)
#undef T2_LAUNCH_VARIANT
        }
        fan_lse_tc2_kernel<D, false><<<blocks, T2_WARPS * 32, smem, stream>>>(p, geo);
    }
    return 0;
}

static int launch_fan_lse_tc2(const FanLseParams<float>& p, int D, bool bwd, cudaStream_t stream, int sm_count) {
    switch (D) {
        case 2: return launch_fan_lse_tc2_D<2>(p, bwd, stream, sm_count);
        case 4: return launch_fan_lse_tc2_D<4>(p, bwd, stream, sm_count);
        case 6: return launch_fan_lse_tc2_D<6>(p, bwd, stream, sm_count);
        case 8: return launch_fan_lse_tc2_D<8>(p, bwd, stream, sm_count);
        case 12: return launch_fan_lse_tc2_D<12>(p, bwd, stream, sm_count);
        case 16: return launch_fan_lse_tc2_D<16>(p, bwd, stream, sm_count);
        case 18: return launch_fan_lse_tc2_D<18>(p, bwd, stream, sm_count);
    }
    return 1;
}

}  // namespace tc
