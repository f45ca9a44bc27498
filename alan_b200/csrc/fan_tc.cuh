// fan_tc.cuh -- fan_lse on the 5th-generation tensor cores (tcgen05 + TMEM), fp32 data, 3xTF32.
//
// Same contract as fan_lse2_kernel (fused.cuh):
//     out[rho, f] = LSE_eps_kappa( b[rho,kappa] - c[f] - sum_d (v - l)^2 w[f,d] ) + cadd
// but the d-contraction -- a genuinely dense  [f, d] x [d, kappa]  product per rho -- runs as
// tcgen05.mma.kind::tf32 with the accumulator in tensor memory.  The LSE reduces over kappa, so kappa
// must land on TMEM *columns* (registers of one thread after tcgen05.ld), not on lanes (threads).
// That is arranged by making the constant operand block-diagonal:
//
//   A (TMEM, written once per CTA)   rows  i = (rs, f)       rs in 0..3, f in 0..31      M = 128
//                                    cols  k = (rs', dd)     dd in 0..KB                  K = 4 KB
//                                    A[i,k] = [rs == rs'] * { -w[f,dd] log2e | 1 | -c[f] log2e }
//   B (smem, one stage per tile)     rows  n = kappa (32),   cols k = (rs', dd)
//                                    B[n,k] = { (v - l)^2 of rho = 4 tile + rs' | bias b | 1 }
//   D (TMEM) = A B^T                 D[(rs,f), kappa] = log2e * S[rho_rs, kappa, f]   complete, incl. bias
//
// so thread (rs, f) of the epilogue reads its 32 kappa values with ONE tcgen05.ld and does the whole
// max / exp2 / sum in registers.  fp32 accuracy comes from the 3xTF32 split  A_hi B_hi + A_lo B_hi +
// A_hi B_lo  (relative error ~2^-21 per product; the dropped lo*lo term is 2^-22).
//
// Roles (25 warps, one persistent CTA per SM; a supertile = 4 groups x 4 rho slots = 16 rho, N = 128 kappa columns):
//   epilogue  teams of four warps: warp % 4 = TMEM lane quadrant = rho slot rs, lane = f; team e takes the rho
//             groups [e GPT, (e + 1) GPT) of EVERY supertile (forward: 2 teams, adjoint: 4 teams)
//   MMA       one warp, one elected lane issues the 30 tcgen05.mma of a supertile; it also owns the TMEM allocation
//   builders  warp (gq, rs) writes the K-columns of rho slot rs of its group(s), lane = kappa (forward: 16 warps
//             with one rho each, adjoint: 8 warps with two rho each)
// Pipelines: smem stages full/empty (builders <-> MMA), TMEM accumulators tfull/tempty (MMA <-> epilogue);
// mbarriers, tcgen05.commit for the MMA-side arrivals; every wait is a bounded spin that traps.
#pragma once
#include "fused.cuh"

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the waiting warp sleeps until the phase completes (or the hint, in nanoseconds,
// runs out) instead of polling -- the polling loops of the builder / MMA warps were 16 % of all issued warp instructions
// of the dense forward kernel (ncu, round 2), taken from the sub-partitions' epilogue warps.
// (ALAN_B200_WAIT_HINT_NS, read when a plan is created, sets the hint: tuning aid.)
__constant__ unsigned g_wait_hint_ns = 100000u;
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(g_wait_hint_ns) : "memory");
    return ok != 0;
}
// Bounded wait (~2 s): a protocol bug must fault (trap), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin)
        if (spin > (1u << 22)) __trap();
}
#ifdef TC_DEBUG_SPIN
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, long long& acc) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
}
#define MBAR_WAIT(bar, par, acc) mbar_wait_t(bar, par, acc)
#else
#define MBAR_WAIT(bar, par, acc) mbar_wait(bar, par)
#endif
// one lane of a converged warp (the way the MMA issuer is elected: the surrounding code stays warp-uniform, so
// descriptors and TMEM addresses live in uniform registers instead of being re-broadcast for every UTCHMMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// round-to-nearest (ties away) to TF32's 10 mantissa bits: add half an ulp, clear the 13 low bits.  (cvt.rna.tf32.f32
// compiles to a six-instruction emulation on sm_100a; the operands here are finite.)
__device__ __forceinline__ uint32_t to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32, M = 128, N = 32, K = 8
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes; LBO = byte step between the two 16-byte K chunks
// of one MMA, SBO = byte step between 8-row groups (cute::UMMA::SmemDescriptor, version 1 = sm_100).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

#define TC_LD32(r, addr)                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                    \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                    \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"     \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),       \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
                 : "r"(addr) : "memory")

#define TC_ST8(addr, r, o)                                                                                    \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"              \
                 :: "r"(addr), "r"(r[o]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), \
                    "r"(r[o + 6]), "r"(r[o + 7]) : "memory")

// Geometry of the rho iteration space remapped by the host into four right-aligned dim slots (unused
// slots have extent 1, stride 0), all 32-bit: compile-time slot indices make every stride a constant-
// bank operand, and the multi-index of a role advances incrementally (add-with-carry) -- no divisions
// and no parameter-array walks in the per-tile loops.
#define TC_ND 4
#define TC_NB 4
struct TcGeom {
    int sz[TC_ND];
    int vs[TC_ND], ls[TC_ND], os[TC_ND], gs[TC_ND];
    int g_f;
    int bs[TC_NB][TC_ND];
    int bk[TC_NB];
    float bc[TC_NB];
    int nb;
    int vec2;
};

struct TcIdx {
    int i[TC_ND];
    __device__ __forceinline__ void set(unsigned lin, const TcGeom& g) {
#pragma unroll
        for (int k = TC_ND - 1; k >= 0; --k) {
            const unsigned sz = (unsigned)g.sz[k];
            const unsigned q = lin / sz;
            i[k] = (int)(lin - q * sz);
            lin = q;
        }
    }
    // this += step (both valid multi-indices); overflow of the top slot wraps (callers mask rho >= n_rho)
    __device__ __forceinline__ void add(const TcIdx& st, const TcGeom& g) {
        int carry = 0;
#pragma unroll
        for (int k = TC_ND - 1; k >= 0; --k) {
            const int v = i[k] + st.i[k] + carry;
            carry = v >= g.sz[k] ? 1 : 0;
            i[k] = carry ? v - g.sz[k] : v;
        }
    }
    __device__ __forceinline__ int dot(const int* st) const {
        int o = 0;
#pragma unroll
        for (int k = 0; k < TC_ND; ++k) o += i[k] * st[k];
        return o;
    }
};

constexpr int TC_G = 4;           // rho groups per supertile: N = 32 TC_G kappa columns per MMA (N = 128: below
                                  // that a tcgen05.mma still costs ~64 cycles, measured: N = 32 ran at 64 clk / MMA)
constexpr int TC_N = 32 * TC_G;
constexpr int TC_RHO = 4 * TC_G;  // rho per supertile
// Warp roles differ per direction (25 warps either way): the forward pass is bound by the MMA issuer and needs all
// 16 builder warps to keep it fed (2 epilogue teams suffice); the adjoint is bound by its heavier epilogue
// (exp2, multiply, cross-lane sum over f) and runs 4 epilogue teams with 8 builder warps building two rho each.
template <bool BWD> struct TcRoles {
    static constexpr int EPI = BWD ? 4 : 2;          // epilogue teams of 4 warps (warp % 4 = TMEM lane quadrant = rs)
    static constexpr int BPW = BWD ? 2 : 1;          // rho built per builder warp and supertile
    static constexpr int BW = 4 * TC_G / BPW;        // builder warps: warp (gq, rs) builds groups gq, gq + TC_G / BPW, ...
    static constexpr int MMA_WARP = 4 * EPI;
    static constexpr int WARPS = 4 * EPI + 1 + BW;
};
constexpr int TC_WARPS = 25;
static_assert(TcRoles<true>::WARPS == TC_WARPS && TcRoles<false>::WARPS == TC_WARPS, "role split must keep 25 warps");
constexpr int TC_STAGES = 2;      // smem stages of the B operand (80 KB each at D = 18)
constexpr int TC_ACC = 2;         // TMEM accumulator stages (multiple of TC_EPI)
constexpr int TC_TMEM_COLS = 512; // A_hi + A_lo (<= 80 each) + 2 x 128 accumulator columns; one CTA per SM

// NC = 16-byte K chunks per rho block: KB = 4 NC >= D + 2
template <int D, bool BWD>
__global__ void __launch_bounds__(TC_WARPS * 32, 1) fan_lse_tc_kernel(const __grid_constant__ FanLseParams<float> p, const __grid_constant__ TcGeom geo) {
    constexpr int NC = (D + 2 + 3) / 4, KB = 4 * NC, KT = 4 * KB;          // KT = K extent of the MMA (<= 80)
    constexpr int KSTEPS = KT / 8;
    constexpr int TC_EPI = TcRoles<BWD>::EPI, TC_BPW = TcRoles<BWD>::BPW, TC_BW = TcRoles<BWD>::BW,
                  TC_MMA_WARP = TcRoles<BWD>::MMA_WARP;
    constexpr uint32_t LBO = TC_N * 16, SBO = 8 * 16;                      // [chunk][TC_N rows][16 B]
    constexpr uint32_t OPER = 4 * NC * LBO;                                // bytes of one operand part of one stage
    constexpr uint32_t A_HI = 0, A_LO = KT, D_COL = 2 * KT;
    static_assert(2 * KT + TC_ACC * TC_N <= TC_TMEM_COLS, "TMEM budget");
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((128u >> 4) << 24);

    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char* stage_base = tc_smem;                                   // TC_STAGES x (B_hi | B_lo)
    uint64_t* bars = (uint64_t*)(tc_smem + TC_STAGES * 2 * OPER);
    uint64_t* full = bars;                                                 // [TC_STAGES]  builders -> MMA
    uint64_t* empty = bars + TC_STAGES;                                    // [TC_STAGES]  MMA -> builders
    uint64_t* tfull = bars + 2 * TC_STAGES;                                // [TC_ACC]     MMA -> epilogue
    uint64_t* tempty = bars + 2 * TC_STAGES + TC_ACC;                      // [TC_ACC]     epilogue -> MMA
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * TC_STAGES + 2 * TC_ACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float LS = 1.4426950408889634f;
    const int Kk = p.Kk;
    const unsigned n_rho = (unsigned)p.n_rho;
    const unsigned n_tiles = (n_rho + TC_RHO - 1) / TC_RHO;

    // zero the B stages once: padding K columns stay zero forever
    for (uint32_t i = threadIdx.x; i < TC_STAGES * 2 * OPER / 16; i += blockDim.x)
        reinterpret_cast<float4*>(stage_base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    // padding rows (kappa >= Kk) carry a hugely negative bias so that the epilogue needs no column mask
    {
        const float big = -1.0e30f, bh = __uint_as_float(to_tf32(big)), bl = big - bh;
        const int npad = 32 - Kk, per_stage = TC_G * 4 * npad;
        for (int i = threadIdx.x; i < TC_STAGES * per_stage; i += blockDim.x) {
            const int s = i / per_stage, r = i - s * per_stage, gr = r / npad, kz = Kk + r - gr * npad;
            const int g = gr >> 2, rs = gr & 3;
            const int off = ((rs * NC + D / 4) * TC_N + 32 * g + kz) * 4 + (D & 3);
            reinterpret_cast<float*>(stage_base + (size_t)s * 2 * OPER)[off] = bh;
            reinterpret_cast<float*>(stage_base + (size_t)s * 2 * OPER + OPER)[off] = bl;
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], TC_BW * 32); mbar_init(&empty[s], 1); }
        for (int a = 0; a < TC_ACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 128 * TC_EPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    long long dbg0 = 0, dbg1 = 0, dbg2 = 0; (void)dbg0; (void)dbg1; (void)dbg2;
    const long long dbg_t0 = clock64(); (void)dbg_t0;

    if (warp < 4) {
        // ---------------------------------------------------------------- A operand, once
        const int rs = warp, f = lane;
        float wv[KB];
#pragma unroll
        for (int dd = 0; dd < KB; ++dd) wv[dd] = 0.f;
        if (f < p.F) {
            float c = 0.f;
#pragma unroll
            for (int dd = 0; dd < D; ++dd) {
                const float sc = p.s[f * p.s_f + dd * p.s_ev];
                wv[dd] = -LS / (2.f * (sc * sc));
                c += logf(sc);
            }
            wv[D] = 1.f;
            wv[D + 1] = -(c + float(D) * float(HALF_LOG_2PI)) * LS;
        }
        const uint32_t lane_base = tmem + ((uint32_t)(32 * warp) << 16);
#pragma unroll
        for (int g = 0; g < KT / 8; ++g) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int kk = 8 * g + e;
                const int blk = kk / KB, dd = kk - blk * KB;
                float x = 0.f;
#pragma unroll
                for (int q = 0; q < KB; ++q) if (q == dd) x = wv[q];
                if (blk != rs) x = 0.f;
                hi[e] = to_tf32(x);
                lo[e] = __float_as_uint(x - __uint_as_float(hi[e]));
            }
            TC_ST8(lane_base + A_HI + 8 * g, hi, 0);
            TC_ST8(lane_base + A_LO + 8 * g, lo, 0);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp < 4 * TC_EPI) {
        // ---------------------------------------------------------------- epilogue (team = warp / 4)
        const int rs = warp & 3, f = lane, o_f = (int)p.o_f;
        // multi-index of rho = TC_RHO * tile + rs for this team's first tile, and its per-iteration step
        // every team works on every supertile: team e takes the rho groups [e GPT, (e + 1) GPT), so an accumulator
        // stage is held for GPT (not TC_G) row computations before it goes back to the MMA issuer
        constexpr int GPT = TC_G / TC_EPI;
        const int g0 = (warp >> 2) * GPT;
        TcIdx base, step, four;
        base.set(TC_RHO * blockIdx.x + 4 * g0 + rs, geo);
        step.set(TC_RHO * gridDim.x, geo);
        four.set(4, geo);
        // offsets and (adjoint) the raw lse / gout values are fetched ONE supertile ahead: their global-load
        // latency overlaps the previous supertile's work instead of sitting in front of every accumulator read
        int n_ooff[GPT];
        float n_lse[GPT], n_g[GPT];
        auto fetch = [&](unsigned tile) {
            TcIdx cur = base;
#pragma unroll
            for (int g = 0; g < GPT; ++g) {
                const unsigned rho = TC_RHO * tile + 4 * (g0 + g) + rs;
                const int o = cur.dot(geo.os);
                n_ooff[g] = o;
                n_lse[g] = INFINITY; n_g[g] = 0.f;                               // idle rows: weight 0
                if (BWD && rho < n_rho && f < p.F) {
                    n_lse[g] = p.lse[o + f * o_f];
                    n_g[g] = p.gout[cur.dot(geo.gs) + f * geo.g_f];
                }
                cur.add(four, geo);
            }
            base.add(step, geo);
        };
        if (blockIdx.x < n_tiles) fetch(blockIdx.x);
        for (unsigned it = 0; blockIdx.x + (i64)it * gridDim.x < n_tiles; ++it) {
            const unsigned tile = blockIdx.x + it * gridDim.x;
            const int a = it % TC_ACC;
            const uint32_t pa = (it / TC_ACC) & 1;
            int ooff[GPT];
            float lz[GPT], gz[GPT];
#pragma unroll
            for (int g = 0; g < GPT; ++g) {
                ooff[g] = n_ooff[g];
                lz[g] = BWD ? (n_lse[g] - p.cadd) * LS : 0.f;
                gz[g] = n_g[g];
            }
            if (tile + gridDim.x < n_tiles) fetch(tile + gridDim.x);
            MBAR_WAIT(&tfull[a], pa, dbg0);
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < GPT; ++g) {
                const unsigned rho = TC_RHO * tile + 4 * (g0 + g) + rs;
                uint32_t r[32];
#ifdef TC_DEBUG_SPIN
                const long long tA = clock64();
#endif
                TC_LD32(r, tmem + ((uint32_t)(32 * rs) << 16) + D_COL + TC_N * a + 32 * (g0 + g));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (g == GPT - 1) { tc_fence_before(); mbar_arrive(&tempty[a]); }
#ifdef TC_DEBUG_SPIN
                const long long tB = clock64(); dbg1 += tB - tA;
#endif
                if (!BWD) {
                    float m = __uint_as_float(r[0]);
#pragma unroll
                    for (int k = 1; k < 32; ++k) m = fmaxf(m, __uint_as_float(r[k]));
                    // packed subtract / accumulate (add.f32x2): half the FADD issue slots around the 32 ex2
                    const float2 nm2 = make_float2(-m, -m);
                    float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 32; k += 2) {
                        const float2 d2 = __fadd2_rn(make_float2(__uint_as_float(r[k]), __uint_as_float(r[k + 1])), nm2);
                        acc2 = __fadd2_rn(acc2, make_float2(FastExp<float>::ex(d2.x), FastExp<float>::ex(d2.y)));
                    }
                    const float sum = acc2.x + acc2.y;
                    if (rho < n_rho && f < p.F)
                        p.out[ooff[g] + f * o_f] = logf(sum + Eps<float>::v()) + m * 0.6931471805599453f + p.cadd;
                } else {
                    // weights of this (rho, f) row, then the sum over f (the 32 lanes of this warp): fixed-order
                    // butterfly reduce-scatter, lane j ends with the sum for kappa = j (31 shuffles for 32 columns)
                    const float2 nl2 = make_float2(-lz[g], -lz[g]), g2 = make_float2(gz[g], gz[g]);
                    float wv[32];
#pragma unroll
                    for (int k = 0; k < 32; k += 2) {
                        const float2 d0 = __fadd2_rn(make_float2(__uint_as_float(r[k]), __uint_as_float(r[k + 1])), nl2);
                        const float2 e0 = __fmul2_rn(make_float2(FastExp<float>::ex(d0.x), FastExp<float>::ex(d0.y)), g2);
                        wv[k] = e0.x; wv[k + 1] = e0.y;
                    }
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
                        for (int i = 0; i < off; ++i) {
                            const bool up = (lane & off) != 0;
                            const float send = up ? wv[i] : wv[i + off];
                            const float mine = up ? wv[i + off] : wv[i];
                            wv[i] = mine + __shfl_xor_sync(0xffffffffu, send, off);
                        }
                    }
                    if (rho < n_rho && lane < Kk) p.gS[(i64)rho * Kk + lane] = wv[0];
                }
#ifdef TC_DEBUG_SPIN
                dbg2 += clock64() - tB;
#endif
            }
        }
    } else if (warp == TC_MMA_WARP) {
        // ---------------------------------------------------------------- MMA issuer
        unsigned it = 0;
        for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int s = it % TC_STAGES, a = it % TC_ACC;
            const uint32_t ps = (it / TC_STAGES) & 1, pa = (it / TC_ACC) & 1;
            MBAR_WAIT(&full[s], ps, dbg0);
            MBAR_WAIT(&tempty[a], pa ^ 1, dbg1);
            tc_fence_after();
            const uint32_t bhi = smem_u32(stage_base + (size_t)s * 2 * OPER), blo = bhi + OPER;
            const uint32_t d = tmem + D_COL + TC_N * a;
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < KSTEPS; ++j) {
                    const uint64_t dh = smem_desc(bhi + j * 2 * LBO, LBO, SBO), dl = smem_desc(blo + j * 2 * LBO, LBO, SBO);
                    mma_tf32_ts(d, tmem + A_HI + 8 * j, dh, IDESC, j > 0 ? 1u : 0u);
                    mma_tf32_ts(d, tmem + A_LO + 8 * j, dh, IDESC, 1u);
                    mma_tf32_ts(d, tmem + A_HI + 8 * j, dl, IDESC, 1u);
                }
                tc_commit(&empty[s]);
                tc_commit(&tfull[a]);
            }
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------------------- builders: warp (g, rs), lane = kappa
        const int bw = warp - TC_MMA_WARP - 1, gq = bw >> 2, rs = bw & 3, kz = lane;
        const int vk = (int)p.v_k, lk = (int)p.l_k, vev = (int)p.v_ev, lev = (int)p.l_ev, nb = geo.nb, vec2 = geo.vec2;
        TcIdx cur[TC_BPW], step;
#pragma unroll
        for (int h = 0; h < TC_BPW; ++h) cur[h].set(TC_RHO * blockIdx.x + 4 * (gq + h * (TC_G / TC_BPW)) + rs, geo);
        step.set(TC_RHO * gridDim.x, geo);
        unsigned it = 0;
        for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int s = it % TC_STAGES;
            const uint32_t ps = (it / TC_STAGES) & 1;
            float t[TC_BPW][KB];
#pragma unroll
            for (int h = 0; h < TC_BPW; ++h) {
                const int g = gq + h * (TC_G / TC_BPW);
                const bool live = TC_RHO * tile + 4 * g + rs < n_rho && kz < Kk;
                const int voff = cur[h].dot(geo.vs), loff = cur[h].dot(geo.ls);
                int boff[TC_NB];
#pragma unroll
                for (int i = 0; i < TC_NB; ++i) boff[i] = cur[h].dot(geo.bs[i]);
                cur[h].add(step, geo);
#pragma unroll
                for (int dd = 0; dd < KB; ++dd) t[h][dd] = 0.f;
                if (live) {
                    float b = 0.f;
#pragma unroll
                    for (int i = 0; i < TC_NB; ++i) if (i < nb) b += geo.bc[i] * p.b[i][boff[i] + kz * geo.bk[i]];
                    const float* vp = p.v + voff + kz * vk;
                    const float* lp = p.l + loff + kz * lk;
                    if (vec2) {
#pragma unroll
                        for (int q = 0; q < D / 2; ++q) {
                            const float2 vv = *reinterpret_cast<const float2*>(vp + 2 * q);
                            const float2 ll = *reinterpret_cast<const float2*>(lp + 2 * q);
                            const float d0 = vv.x - ll.x, d1 = vv.y - ll.y;
                            t[h][2 * q] = d0 * d0; t[h][2 * q + 1] = d1 * d1;
                        }
                    } else {
#pragma unroll
                        for (int dd = 0; dd < D; ++dd) { const float df = vp[dd * vev] - lp[dd * lev]; t[h][dd] = df * df; }
                    }
                    t[h][D] = b * LS;
                    t[h][D + 1] = 1.f;
                }
            }
            MBAR_WAIT(&empty[s], ps ^ 1, dbg0);
            if (kz < Kk) {                       // rows of rho >= n_rho are written as zeros: finite, masked later
                float* bh = reinterpret_cast<float*>(stage_base + (size_t)s * 2 * OPER);
                float* bl = reinterpret_cast<float*>(stage_base + (size_t)s * 2 * OPER + OPER);
#pragma unroll
                for (int h = 0; h < TC_BPW; ++h) {
                    const int g = gq + h * (TC_G / TC_BPW);
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        // the tensor core reads only the 19 TF32 bits of each word: the raw fp32 value IS the "hi" part
                        // (truncated by the hardware) and lo = t - trunc(t) is its exact remainder
                        float4 hh, l;
                        hh.x = t[h][4 * c + 0]; l.x = hh.x - __uint_as_float(__float_as_uint(hh.x) & 0xFFFFE000u);
                        hh.y = t[h][4 * c + 1]; l.y = hh.y - __uint_as_float(__float_as_uint(hh.y) & 0xFFFFE000u);
                        hh.z = t[h][4 * c + 2]; l.z = hh.z - __uint_as_float(__float_as_uint(hh.z) & 0xFFFFE000u);
                        hh.w = t[h][4 * c + 3]; l.w = hh.w - __uint_as_float(__float_as_uint(hh.w) & 0xFFFFE000u);
                        const int off = ((rs * NC + c) * TC_N + 32 * g + kz) * 4;
                        *reinterpret_cast<float4*>(bh + off) = hh;
                        *reinterpret_cast<float4*>(bl + off) = l;
                    }
                }
            }
            fence_async_smem();
            mbar_arrive(&full[s]);
        }
    }

#ifdef TC_DEBUG_SPIN
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4 || warp == TC_MMA_WARP || warp == TC_MMA_WARP + 1 || warp == TC_MMA_WARP + 6))
        printf("bwd=%d warp %d total %lld wait0 %lld wait1/ldtm %lld compute %lld\n", (int)BWD, warp, clock64() - dbg_t0, dbg0, dbg1, dbg2);
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TC_TMEM_COLS) : "memory");
    }
}

template <int D>
static int launch_fan_lse_tc_D(const FanLseParams<float>& p, bool bwd, cudaStream_t stream, int sm_count) {
    constexpr int NC = (D + 2 + 3) / 4;
    const size_t smem = (size_t)TC_STAGES * 2 * (4 * NC * TC_N * 16) + (2 * TC_STAGES + 2 * TC_ACC) * 8 + 16;
    bool ev2 = (D % 2 == 0) && p.v_ev == 1 && p.l_ev == 1 && p.v_k % 2 == 0 && p.l_k % 2 == 0 &&
               ((uintptr_t)p.v % 8 == 0) && ((uintptr_t)p.l % 8 == 0);
    for (int k = 0; k < p.rd.nd && ev2; ++k) ev2 = (p.vstride[k] % 2 == 0) && (p.lstride[k] % 2 == 0);
    const i64 n_tiles = (p.n_rho + TC_RHO - 1) / TC_RHO;
    i64 blocks = n_tiles < (i64)sm_count ? n_tiles : (i64)sm_count;
    if (blocks < 1) blocks = 1;
    TcGeom geo;
    memset(&geo, 0, sizeof(geo));
    for (int j = 0; j < TC_ND; ++j) geo.sz[j] = 1;
    for (int k = 0; k < p.rd.nd; ++k) {
        const int j = TC_ND - p.rd.nd + k;
        geo.sz[j] = p.rd.size[k];
        geo.vs[j] = (int)p.vstride[k]; geo.ls[j] = (int)p.lstride[k]; geo.os[j] = (int)p.ostride[k]; geo.gs[j] = (int)p.gstride[k];
        for (int i = 0; i < p.nb; ++i) geo.bs[i][j] = (int)p.bstride[i][k];
    }
    for (int i = 0; i < p.nb; ++i) { geo.bk[i] = (int)p.b_k[i]; geo.bc[i] = p.bcoeff[i]; }
    geo.nb = p.nb;
    geo.g_f = (int)p.g_f;
    geo.vec2 = ev2 ? 1 : 0;
    if (bwd) {
        static const cudaError_t attr_true = cudaFuncSetAttribute(fan_lse_tc_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // once per process
        (void)attr_true;
        fan_lse_tc_kernel<D, true><<<(int)blocks, TC_WARPS * 32, smem, stream>>>(p, geo);
    } else {
        static const cudaError_t attr_false = cudaFuncSetAttribute(fan_lse_tc_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // once per process
        (void)attr_false;
        fan_lse_tc_kernel<D, false><<<(int)blocks, TC_WARPS * 32, smem, stream>>>(p, geo);
    }
    return 0;
}

// The tensor-core path covers the shapes of the hierarchical-Gaussian hot path: fp32, fan and kappa
// extents up to 32 (K <= 32), event extent up to 18; everything else runs fan_lse2_kernel.
static bool fan_lse_tc_supported(const FanLseParams<float>& p, int D) {
    if (p.F > 32 || p.Kk > 32 || p.F < 8 || p.n_rho < 64 || p.rd.nd > TC_ND || p.nb > TC_NB) return false;
    // 32-bit element offsets inside the kernel: every operand must span fewer than 2^31 elements
    const i64 lim = (i64)1 << 31;
    i64 vspan = (i64)p.Kk * p.v_k + 32 * p.v_ev, lspan = (i64)p.Kk * p.l_k + 32 * p.l_ev, ospan = 32 * p.o_f;
    i64 bspan = 0;
    for (int k = 0; k < p.rd.nd; ++k) {
        vspan += (i64)p.rd.size[k] * p.vstride[k]; lspan += (i64)p.rd.size[k] * p.lstride[k];
        ospan += (i64)p.rd.size[k] * p.ostride[k];
        for (int i = 0; i < p.nb; ++i) { i64 b = (i64)p.rd.size[k] * p.bstride[i][k] + (i64)p.Kk * p.b_k[i]; if (b > bspan) bspan = b; }
    }
    if (vspan >= lim || lspan >= lim || ospan >= lim || bspan >= lim || p.n_rho * (i64)p.Kk >= lim) return false;
    switch (D) { case 2: case 4: case 6: case 8: case 12: case 16: case 18: return true; }
    return false;
}

static int launch_fan_lse_tc(const FanLseParams<float>& p, int D, bool bwd, cudaStream_t stream, int sm_count) {
    switch (D) {
        case 2: return launch_fan_lse_tc_D<2>(p, bwd, stream, sm_count);
        case 4: return launch_fan_lse_tc_D<4>(p, bwd, stream, sm_count);
        case 6: return launch_fan_lse_tc_D<6>(p, bwd, stream, sm_count);
        case 8: return launch_fan_lse_tc_D<8>(p, bwd, stream, sm_count);
        case 12: return launch_fan_lse_tc_D<12>(p, bwd, stream, sm_count);
        case 16: return launch_fan_lse_tc_D<16>(p, bwd, stream, sm_count);
        case 18: return launch_fan_lse_tc_D<18>(p, bwd, stream, sm_count);
    }
    return 1;
}

}  // namespace tc
