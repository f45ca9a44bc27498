// kernels.cuh -- generic (any plate tree / any axes) kernels of the logPQ engine.
//
//   expr_fwd / expr_bwd   factor evaluation K1+K2 and its adjoint
//   reduce_*              log-semiring contraction K3, plate sum K4, adjoint weights K6
//   chain_*               Timeseries log-matmul chain K5 and its adjoint
//   sample_kernel         posterior K resampling K7
//   gather_kernel         sample gather K8
//
// All reductions are fixed-order (sequential per thread, xor-butterfly per warp, planner-
// chosen split counts): results are bit-reproducible run to run; no float atomics anywhere.
#pragma once
#include "vm.cuh"
#include <vector>

// ------------------------------------------------------------------------------------------
// leaf loads
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T load_leaf(const Opnd& L, i64 off, const int* idx) {
    const T* p = (const T*)L.ptr;
    if (L.mode == 0) return p[off];
    if (L.mode == 1) return idx[L.mdim] >= 1 ? p[off - L.stride[L.mdim]] : T(0);   // prev[t] = x[t-1]
    return idx[L.mdim] == 0 ? p[off] : T(0);                                         // init only at t = 0
}

// ------------------------------------------------------------------------------------------
// K1/K2: factor expression, summed over the event dims
//   reference: Dist.log_prob -> TorchDimDist.log_prob -> sum_non_dim (dist.py:297-302,
//   TorchDimDist.py:127-162); lambdas of dist.py:221-227 fused in.
// ------------------------------------------------------------------------------------------
template <typename T>
struct ExprParams {
    Dims d;                 // n_a = kept (output) dims, the rest are summed
    int n_leaves;
    Opnd leaf[AB_MAXL];
    VMProg<T> prog;
    T* out;
    int acc;
    T scale;
    i64 n_out, n_red;
};

// Specialisation selector: the overwhelmingly common factor program is "three loads + Normal"
// (every mean-field Gaussian Q and most priors).  Those run as straight-line code instead of through
// the VM loop; NORMAL3 = true means leaf order is (value, loc, scale) = nl[0..3).
struct Normal3 { int on; int nl[3]; };

template <typename T>
static Normal3 detect_normal3(const VMProg<T>& P) {
    Normal3 n; n.on = 0; n.nl[0] = n.nl[1] = n.nl[2] = 0;
    if (P.n_instr != 4) return n;
    int leaf_of_reg[AB_NREG];
    for (int r = 0; r < AB_NREG; ++r) leaf_of_reg[r] = -1;
    for (int i = 0; i < 3; ++i) {
        unsigned w0 = P.ins[i][0];
        if ((w0 & 0xff) != V_LOAD) return n;
        leaf_of_reg[(w0 >> 8) & 0xff] = (w0 >> 16) & 0xff;
    }
    unsigned w0 = P.ins[3][0], w1 = P.ins[3][1];
    if ((w0 & 0xff) != V_NORMAL || (int)((w0 >> 8) & 0xff) != P.res) return n;
    int ra = (w0 >> 16) & 0xff, rb = (w0 >> 24) & 0xff, rc = w1 & 0xff;
    if (leaf_of_reg[ra] < 0 || leaf_of_reg[rb] < 0 || leaf_of_reg[rc] < 0) return n;
    n.nl[0] = leaf_of_reg[ra]; n.nl[1] = leaf_of_reg[rb]; n.nl[2] = leaf_of_reg[rc];
    n.on = 1;
    return n;
}

// WARP = false: one thread per output, sequential over the summed event dims.
// WARP = true : one warp per output, lanes stride over the summed dims, fixed-order butterfly sum
//               (few outputs: the work is latency-bound unless the event loop is spread over lanes).
template <typename T, bool WARP, bool N3>
__device__ __forceinline__ void expr_fwd_body(const ExprParams<T>& p, const Normal3& n3, const i64 t0, const i64 tn) {
    T reg[N3 ? 1 : AB_NREG];
    T lv[AB_MAXL];
    int idx[AB_MAXD];
    i64 base[AB_MAXL];
    const int lane = WARP ? (threadIdx.x & 31) : 0, nl = WARP ? 32 : 1;
    for (i64 o = WARP ? (t0 >> 5) : t0; o < p.n_out; o += WARP ? (tn >> 5) : tn) {
        unravel(o, p.d, 0, p.d.n_a, idx);
        for (int l = 0; l < p.n_leaves; ++l) base[l] = dot_stride(p.leaf[l], idx, 0, p.d.n_a);
        T sum = T(0);
        if (N3 && p.d.nd - p.d.n_a == 1 && p.leaf[n3.nl[0]].mode == 0 && p.leaf[n3.nl[1]].mode == 0 && p.leaf[n3.nl[2]].mode == 0) {
            // one summed event dim, plain leaves: walk it with three constant strides (no index decoding)
            const int kd = p.d.n_a;
            const T* pv = (const T*)p.leaf[n3.nl[0]].ptr + base[n3.nl[0]];
            const T* pl = (const T*)p.leaf[n3.nl[1]].ptr + base[n3.nl[1]];
            const T* ps = (const T*)p.leaf[n3.nl[2]].ptr + base[n3.nl[2]];
            const i64 sv = p.leaf[n3.nl[0]].stride[kd], sl = p.leaf[n3.nl[1]].stride[kd], ss = p.leaf[n3.nl[2]].stride[kd];
#pragma unroll 6
            for (i64 r = lane; r < p.n_red; r += nl) sum += normal_lp(pv[r * sv], pl[r * sl], ps[r * ss]);   // several loads in flight
        } else
        for (i64 r = lane; r < p.n_red; r += nl) {
            unravel(r, p.d, p.d.n_a, p.d.nd, idx);
            if (N3) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const Opnd& L = p.leaf[n3.nl[q]];
                    lv[q] = load_leaf<T>(L, base[n3.nl[q]] + dot_stride(L, idx, p.d.n_a, p.d.nd), idx);
                }
                sum += normal_lp(lv[0], lv[1], lv[2]);
            } else {
                for (int l = 0; l < p.n_leaves; ++l)
                    lv[l] = load_leaf<T>(p.leaf[l], base[l] + dot_stride(p.leaf[l], idx, p.d.n_a, p.d.nd), idx);
                sum += vm_eval(p.prog, lv, reg);
            }
        }
        if (WARP) sum = warp_sum(sum);
        if (lane == 0) {
            T v = p.scale * sum;
            p.out[o] = p.acc ? p.out[o] + v : v;
        }
    }
}

template <typename T, bool WARP, bool N3>
__global__ void __launch_bounds__(256) expr_fwd_kernel(const __grid_constant__ ExprParams<T> p, const Normal3 n3) {
    expr_fwd_body<T, WARP, N3>(p, n3, (i64)blockIdx.x * blockDim.x + threadIdx.x, (i64)gridDim.x * blockDim.x);
}

// Large scalar-event expressions (nothing summed, plain leaves): each thread owns B consecutive cells of the
// innermost output dim, decodes their common outer index once and interprets every VM instruction once for
// the whole batch (vm_eval_batch).
template <typename T, int B>
__global__ void __launch_bounds__(256) expr_fwd_batch_kernel(const __grid_constant__ ExprParams<T> p) {
    T reg[AB_NREG][B];
    T lv[AB_MAXL][B];
    int idx[AB_MAXD];
    const int kin = p.d.n_a - 1;                      // innermost kept dim; its extent is a multiple of B
    const i64 n_groups = p.n_out / B;
    for (i64 gidx = (i64)blockIdx.x * blockDim.x + threadIdx.x; gidx < n_groups; gidx += (i64)gridDim.x * blockDim.x) {
        const i64 o = gidx * B;
        unravel(o, p.d, 0, p.d.n_a, idx);
        for (int l = 0; l < p.n_leaves; ++l) {
            const T* ptr = (const T*)p.leaf[l].ptr + dot_stride(p.leaf[l], idx, 0, p.d.n_a);
            const i64 st = p.leaf[l].stride[kin];
#pragma unroll
            for (int q = 0; q < B; ++q) lv[l][q] = ptr[q * st];
        }
        vm_eval_batch<T, B>(p.prog, lv, reg);
#pragma unroll
        for (int q = 0; q < B; ++q) {
            const T v = p.scale * reg[p.prog.res][q];
            p.out[o + q] = p.acc ? p.out[o + q] + v : v;
        }
    }
}

template <typename T>
static void launch_expr_fwd(const ExprParams<T>& p, cudaStream_t stream, int sm_count) {
    const Normal3 n3 = detect_normal3(p.prog);
    if (!n3.on && p.n_red == 1 && p.d.n_a >= 1 && p.n_out >= (1 << 16)) {
        bool plain = true;
        for (int l = 0; l < p.n_leaves; ++l) plain = plain && p.leaf[l].mode == 0;
        const int inner = p.d.size[p.d.n_a - 1];
        if (plain) {
            const i64 cap = (i64)sm_count * 8;
#define AB_BATCH(BB) if (inner % BB == 0) { i64 g = (p.n_out / BB + 255) / 256; \
                expr_fwd_batch_kernel<T, BB><<<(int)(g > cap ? cap : g), 256, 0, stream>>>(p); return; }
            AB_BATCH(8) AB_BATCH(5) AB_BATCH(4) AB_BATCH(3) AB_BATCH(2)
#undef AB_BATCH
        }
    }
    const bool warp = p.n_out < 16384 && p.n_red >= 8;
    i64 threads = p.n_out * (warp ? 32 : 1);
    i64 g = (threads + 255) / 256, cap = (i64)sm_count * 8;
    int grid = (int)(g > cap ? cap : (g < 1 ? 1 : g));
    if (warp) {
        if (n3.on) expr_fwd_kernel<T, true, true><<<grid, 256, 0, stream>>>(p, n3);
        else expr_fwd_kernel<T, true, false><<<grid, 256, 0, stream>>>(p, n3);
    } else {
        if (n3.on) expr_fwd_kernel<T, false, true><<<grid, 256, 0, stream>>>(p, n3);
        else expr_fwd_kernel<T, false, false><<<grid, 256, 0, stream>>>(p, n3);
    }
}

// Adjoint w.r.t. one leaf, gather style: one thread (or warp) per (leaf element, split); it
// loops over every iteration point that read the element.  No atomics.
template <typename T>
struct ExprBwdParams {
    Dims d;                 // n_a = dims the target leaf carries (thread index), rest are looped
    int n_leaves;
    Opnd leaf[AB_MAXL];
    Opnd gout;              // adjoint of the expression output (stride 0 on event dims)
    VMProg<T> prog;
    int target;
    T* gleaf;
    int acc;
    T scale;
    i64 n_kept, n_loop;
    int nsplit;
};

template <typename T, bool WARP, bool N3>
__device__ __forceinline__ void expr_bwd_body(const ExprBwdParams<T>& p, const Normal3& n3, const i64 t0, const i64 tn) {
    T reg[N3 ? 1 : AB_NREG];
    T adj[N3 ? 1 : AB_NREG];
    T lv[AB_MAXL];
    int idx[AB_MAXD];
    i64 base[AB_MAXL];
    const Opnd& tg = p.leaf[p.target];
    const i64 total = p.n_kept * p.nsplit;
    const int lane = WARP ? (threadIdx.x & 31) : 0, nl = WARP ? 32 : 1;
    // which argument of the Normal the target leaf is (it may be more than one)
    const bool tv = N3 && n3.nl[0] == p.target, tl = N3 && n3.nl[1] == p.target, ts = N3 && n3.nl[2] == p.target;
    for (i64 w = WARP ? (t0 >> 5) : t0; w < total; w += WARP ? (tn >> 5) : tn) {
        i64 e = w % p.n_kept;
        int s = (int)(w / p.n_kept);
        unravel(e, p.d, 0, p.d.n_a, idx);
        bool live = true;
        if (tg.mode == 1) {                       // leaf element t feeds iteration point t + 1
            idx[tg.mdim] += 1;
            live = idx[tg.mdim] < p.d.size[tg.mdim];
        }
        T sum = T(0);
        if (live) {
            for (int l = 0; l < p.n_leaves; ++l) base[l] = dot_stride(p.leaf[l], idx, 0, p.d.n_a);
            i64 gbase = dot_stride(p.gout, idx, 0, p.d.n_a);
            if (N3 && p.d.nd - p.d.n_a == 1 && p.leaf[n3.nl[0]].mode == 0 && p.leaf[n3.nl[1]].mode == 0 && p.leaf[n3.nl[2]].mode == 0) {
                // one looped dim, plain leaves: constant strides, no index decoding
                const int kd = p.d.n_a;
                const T* pv = (const T*)p.leaf[n3.nl[0]].ptr + base[n3.nl[0]];
                const T* pl = (const T*)p.leaf[n3.nl[1]].ptr + base[n3.nl[1]];
                const T* ps = (const T*)p.leaf[n3.nl[2]].ptr + base[n3.nl[2]];
                const T* pg = (const T*)p.gout.ptr + gbase;
                const i64 sv = p.leaf[n3.nl[0]].stride[kd], sl = p.leaf[n3.nl[1]].stride[kd], ss = p.leaf[n3.nl[2]].stride[kd],
                          sg = p.gout.stride[kd];
                // pointer increments instead of 64-bit index products, one reciprocal per point (none when the scale
                // does not vary along the loop): the loop was issue-bound at ~65 instructions per point (ncu)
                const i64 j0 = s + (i64)lane * p.nsplit, step = (i64)p.nsplit * nl;
                const T* qv = pv + j0 * sv; const T* ql = pl + j0 * sl; const T* qs = ps + j0 * ss; const T* qg = pg + j0 * sg;
                const i64 dv = step * sv, dl = step * sl, ds = step * ss, dg = step * sg;
                T iv = T(0), isc = T(0);
                if (ss == 0 && j0 < p.n_loop) { const T sc = *qs; iv = T(1) / (sc * sc); isc = iv * sc; }
#pragma unroll 4
                for (i64 j = j0; j < p.n_loop; j += step) {
                    if (ss != 0) { const T sc = *qs; iv = T(1) / (sc * sc); isc = iv * sc; }
                    const T df = *qv - *ql;
                    T g = T(0);
                    if (tv) g -= df * iv;
                    if (tl) g += df * iv;
                    if (ts) g += (df * df * iv - T(1)) * isc;
                    sum += g * *qg;
                    qv += dv; ql += dl; qs += ds; qg += dg;
                }
            } else
            for (i64 j = s + (i64)lane * p.nsplit; j < p.n_loop; j += (i64)p.nsplit * nl) {
                unravel(j, p.d, p.d.n_a, p.d.nd, idx);
                if (tg.mode == 2 && idx[tg.mdim] != 0) continue;
                T g;
                if (N3) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const Opnd& L = p.leaf[n3.nl[q]];
                        lv[q] = load_leaf<T>(L, base[n3.nl[q]] + dot_stride(L, idx, p.d.n_a, p.d.nd), idx);
                    }
                    const T sc = lv[2], df = lv[0] - lv[1], iv = T(1) / (sc * sc);
                    g = T(0);
                    if (tv) g -= df * iv;
                    if (tl) g += df * iv;
                    if (ts) g += (df * df * iv - T(1)) / sc;
                } else {
                    for (int l = 0; l < p.n_leaves; ++l)
                        lv[l] = load_leaf<T>(p.leaf[l], base[l] + dot_stride(p.leaf[l], idx, p.d.n_a, p.d.nd), idx);
                    vm_eval(p.prog, lv, reg);
                    g = vm_grad(p.prog, reg, adj, p.target);
                }
                T go = ((const T*)p.gout.ptr)[gbase + dot_stride(p.gout, idx, p.d.n_a, p.d.nd)];
                sum += g * go;
            }
        }
        if (WARP) sum = warp_sum(sum);
        if (lane == 0) {
            T v = p.scale * sum;
            if (p.nsplit > 1) p.gleaf[(i64)s * p.n_kept + e] = v;
            else p.gleaf[e] = p.acc ? p.gleaf[e] + v : v;
        }
    }
}

template <typename T, bool WARP, bool N3>
__global__ void __launch_bounds__(256) expr_bwd_kernel(const __grid_constant__ ExprBwdParams<T> p, const Normal3 n3) {
    expr_bwd_body<T, WARP, N3>(p, n3, (i64)blockIdx.x * blockDim.x + threadIdx.x, (i64)gridDim.x * blockDim.x);
}

template <typename T>
static void launch_expr_bwd(const ExprBwdParams<T>& p, cudaStream_t stream, int sm_count) {
    const Normal3 n3 = detect_normal3(p.prog);
    const i64 total = p.n_kept * p.nsplit;
    const bool warp = total < 16384 && p.n_loop / p.nsplit >= 8;
    i64 threads = total * (warp ? 32 : 1);
    i64 g = (threads + 255) / 256, cap = (i64)sm_count * 8;
    int grid = (int)(g > cap ? cap : (g < 1 ? 1 : g));
    if (warp) {
        if (n3.on) expr_bwd_kernel<T, true, true><<<grid, 256, 0, stream>>>(p, n3);
        else expr_bwd_kernel<T, true, false><<<grid, 256, 0, stream>>>(p, n3);
    } else {
        if (n3.on) expr_bwd_kernel<T, false, true><<<grid, 256, 0, stream>>>(p, n3);
        else expr_bwd_kernel<T, false, false><<<grid, 256, 0, stream>>>(p, n3);
    }
}

// ------------------------------------------------------------------------------------------
// K3/K4/K6: reductions over K axes / plate axes of a broadcast sum of factors
//   mode 0 SUM      out[o] = sum_r s(o,r)                          (logpq.py:149 plate sum)
//   mode 1 LSE_EPS  out[o] = log(sum_r exp(s - max_r s) + eps) + max_r s   (utils.py:207-222)
//   mode 2 LSE      out[o] = logsumexp_r s                         (logpq.py:139)
//   mode 3 WSUM     out[o] = sum_r gout(o,r) * exp(s(o,r) + cadd - lse(o,r))   (adjoint of LSE; cadd is the
//                   constant the forward op added after its LSE, so that exp(.) is the softmax weight)
//   with s(o,r) = sum_f coeff_f * F_f[o,r] (broadcast through zero strides).
// ------------------------------------------------------------------------------------------
enum { R_SUM = 0, R_LSE_EPS = 1, R_LSE = 2, R_WSUM = 3 };

template <typename T>
struct ReduceParams {
    Dims d;                 // n_a = output dims, rest reduced
    int mode;
    int nf;
    Opnd f[AB_MAXL];
    T coeff[AB_MAXL];
    Opnd lse_m, lse_lo, gout;   // WSUM only: (max, log(sum + eps)) of the forward LSE and the adjoint of its output
    T* m_out; T* lo_out;        // LSE modes, optional: the pair the adjoint needs (null when nothing differentiates the op)
    T* out;
    int acc;
    T scale, cadd;
    i64 n_out, n_red;
    int nsplit;
};

// offset contributed by the reduced index j; NRED = number of reduced dims known at compile time
// (1 or 2 on the hot paths), 0 = generic.
template <int NRED>
__device__ __forceinline__ i64 red_off(const Dims& d, const Opnd& o, i64 j) {
    if (NRED == 1) return j * o.stride[d.n_a];
    if (NRED == 2) {
        const unsigned s1 = (unsigned)d.size[d.n_a + 1];
        if ((j >> 31) == 0) {
            const unsigned q = (unsigned)j / s1;
            return (i64)q * o.stride[d.n_a] + (i64)((unsigned)j - q * s1) * o.stride[d.n_a + 1];
        }
        i64 q = j / s1;
        return q * o.stride[d.n_a] + (j - q * s1) * o.stride[d.n_a + 1];
    }
    i64 off = 0;
    if ((j >> 31) == 0) {
        unsigned l = (unsigned)j;
#pragma unroll 1
        for (int k = d.nd - 1; k >= d.n_a; --k) {
            const unsigned s = (unsigned)d.size[k];
            const unsigned q = l / s;
            off += (i64)(l - q * s) * o.stride[k];
            l = q;
        }
        return off;
    }
#pragma unroll 1
    for (int k = d.nd - 1; k >= d.n_a; --k) {
        int s = d.size[k];
        i64 q = j / s;
        off += (j - q * s) * o.stride[k];
        j = q;
    }
    return off;
}

template <typename T, int NRED>
__device__ __forceinline__ T factor_sum(const ReduceParams<T>& p, const i64* base, i64 j) {
    T s = T(0);
#pragma unroll 1
    for (int f = 0; f < p.nf; ++f)
        s += p.coeff[f] * ((const T*)p.f[f].ptr)[base[f] + red_off<NRED>(p.d, p.f[f], j)];
    return s;
}

// LSE result from (max m, shifted sum a); also stores the pair (m, lo) when the op is differentiated
template <typename T>
__device__ __forceinline__ T lse_finish(const ReduceParams<T>& p, i64 o, T m, T a, bool writer) {
    const T lo = p.mode == R_LSE_EPS ? ab_log(a + Eps<T>::v()) : ab_log(a);
    if (writer && p.m_out != nullptr) { p.m_out[o] = m; p.lo_out[o] = lo; }
    return lo + m;
}

template <typename T>
__device__ __forceinline__ void reduce_store(const ReduceParams<T>& p, i64 o, int s, T res) {
    if (p.nsplit > 1) { p.out[(i64)s * p.n_out + o] = res; return; }
    T v = p.scale * res + (p.mode == R_WSUM ? T(0) : p.cadd);
    p.out[o] = p.acc ? p.out[o] + v : v;
}

// one warp per (output, split); lanes stride over the reduced index.  For reduced extents up to
// 128 the lane keeps its (at most four) values in registers, so the factors are read once.
template <typename T, int NRED>
__device__ __forceinline__ void reduce_warp_body(const ReduceParams<T>& p, const i64 t0, const i64 tn) {
    int idx[AB_MAXD];
    i64 base[AB_MAXL];
    const int lane = threadIdx.x & 31;
    const i64 warp = t0 >> 5;
    const i64 nwarps = tn >> 5;
    const i64 total = p.n_out * p.nsplit;
    const i64 chunk = (p.n_red + p.nsplit - 1) / p.nsplit;
    for (i64 w = warp; w < total; w += nwarps) {
        i64 o = w % p.n_out;
        int s = (int)(w / p.n_out);
        i64 lo = (i64)s * chunk, hi = lo + chunk < p.n_red ? lo + chunk : p.n_red;
        unravel(o, p.d, 0, p.d.n_a, idx);
        for (int f = 0; f < p.nf; ++f) base[f] = dot_stride(p.f[f], idx, 0, p.d.n_a);
        T res;
        if (p.mode == R_SUM) {
            T a = T(0);
            for (i64 j = lo + lane; j < hi; j += 32) a += factor_sum<T, NRED>(p, base, j);
            res = warp_sum(a);
        } else if (p.mode == R_WSUM) {
            i64 mbase = dot_stride(p.lse_m, idx, 0, p.d.n_a), lbase = dot_stride(p.lse_lo, idx, 0, p.d.n_a),
                gbase = dot_stride(p.gout, idx, 0, p.d.n_a);
            T a = T(0);
            for (i64 j = lo + lane; j < hi; j += 32) {
                T sv = factor_sum<T, NRED>(p, base, j);
                T mm = ((const T*)p.lse_m.ptr)[mbase + red_off<NRED>(p.d, p.lse_m, j)];
                T l = ((const T*)p.lse_lo.ptr)[lbase + red_off<NRED>(p.d, p.lse_lo, j)];
                T g = ((const T*)p.gout.ptr)[gbase + red_off<NRED>(p.d, p.gout, j)];
                a += g * ab_exp((sv - mm) - l);          // softmax weight exp(s - m) / (sum + eps), as autograd forms it
            }
            res = warp_sum(a);
        } else if (hi - lo <= 128) {
            T v[4];
            T m = neg_inf<T>();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                i64 j = lo + lane + 32 * q;
                v[q] = j < hi ? factor_sum<T, NRED>(p, base, j) : neg_inf<T>();
                m = ab_max(m, v[q]);
            }
            m = warp_max(m);
            T a = T(0);
#pragma unroll
            for (int q = 0; q < 4; ++q) a += (lo + lane + 32 * q < hi) ? ab_exp(v[q] - m) : T(0);
            a = warp_sum(a);
            res = lse_finish(p, o, m, a, lane == 0);
        } else {
            T m = neg_inf<T>();
            for (i64 j = lo + lane; j < hi; j += 32) m = ab_max(m, factor_sum<T, NRED>(p, base, j));
            m = warp_max(m);
            T a = T(0);
            for (i64 j = lo + lane; j < hi; j += 32) a += ab_exp(factor_sum<T, NRED>(p, base, j) - m);
            a = warp_sum(a);
            res = lse_finish(p, o, m, a, lane == 0);
        }
        if (lane == 0) reduce_store(p, o, s, res);
    }
}

template <typename T, int NRED>
__global__ void __launch_bounds__(256) reduce_warp_kernel(const __grid_constant__ ReduceParams<T> p) {
    reduce_warp_body<T, NRED>(p, (i64)blockIdx.x * blockDim.x + threadIdx.x, (i64)gridDim.x * blockDim.x);
}

// one thread per (output, split): outputs contiguous in memory / small reduced extent / none
template <typename T, int NRED>
__device__ __forceinline__ void reduce_thread_body(const ReduceParams<T>& p, const i64 t0, const i64 tn) {
    int idx[AB_MAXD];
    i64 base[AB_MAXL];
    const i64 total = p.n_out * p.nsplit;
    const i64 chunk = (p.n_red + p.nsplit - 1) / p.nsplit;
    for (i64 w = t0; w < total; w += tn) {
        i64 o = w % p.n_out;
        int s = (int)(w / p.n_out);
        i64 lo = (i64)s * chunk, hi = lo + chunk < p.n_red ? lo + chunk : p.n_red;
        unravel(o, p.d, 0, p.d.n_a, idx);
        for (int f = 0; f < p.nf; ++f) base[f] = dot_stride(p.f[f], idx, 0, p.d.n_a);
        T res;
        if (p.mode == R_SUM) {
            // eight independent loads in flight, added in index order: same bits as the one-by-one loop
            T a = T(0);
            i64 j = lo;
            for (; j + 8 <= hi; j += 8) {
                T v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = factor_sum<T, NRED>(p, base, j + q);
#pragma unroll
                for (int q = 0; q < 8; ++q) a += v[q];
            }
            for (; j < hi; ++j) a += factor_sum<T, NRED>(p, base, j);
            res = a;
        } else if (p.mode == R_WSUM) {
            i64 mbase = dot_stride(p.lse_m, idx, 0, p.d.n_a), lbase = dot_stride(p.lse_lo, idx, 0, p.d.n_a),
                gbase = dot_stride(p.gout, idx, 0, p.d.n_a);
            T a = T(0);
            for (i64 j = lo; j < hi; ++j) {
                T sv = factor_sum<T, NRED>(p, base, j);
                T mm = ((const T*)p.lse_m.ptr)[mbase + red_off<NRED>(p.d, p.lse_m, j)];
                T l = ((const T*)p.lse_lo.ptr)[lbase + red_off<NRED>(p.d, p.lse_lo, j)];
                T g = ((const T*)p.gout.ptr)[gbase + red_off<NRED>(p.d, p.gout, j)];
                a += g * ab_exp((sv - mm) - l);
            }
            res = a;
        } else {
            T m = neg_inf<T>();
            for (i64 j = lo; j < hi; ++j) m = ab_max(m, factor_sum<T, NRED>(p, base, j));
            T a = T(0);
            for (i64 j = lo; j < hi; ++j) a += ab_exp(factor_sum<T, NRED>(p, base, j) - m);
            res = lse_finish(p, o, m, a, true);
        }
        reduce_store(p, o, s, res);
    }
}

template <typename T, int NRED>
__global__ void __launch_bounds__(256) reduce_thread_kernel(const __grid_constant__ ReduceParams<T> p) {
    reduce_thread_body<T, NRED>(p, (i64)blockIdx.x * blockDim.x + threadIdx.x, (i64)gridDim.x * blockDim.x);
}

template <typename T>
static bool reduce_uses_warps(const ReduceParams<T>& p, bool thread_hint) {
    const i64 per = (p.n_red + p.nsplit - 1) / p.nsplit;
    // small ops are bound by the latency of their dependent loads, not by coalescing: one thread per output walks
    // `per` cache-missing loads in a row (measured 27 us for the [30, 30] top-level contraction of cfg-5), a warp
    // per output issues them side by side
    if (p.n_out * p.nsplit * per <= 16384 && per >= 8) return true;
    return per >= 16 && (!thread_hint || (p.n_out * p.nsplit < 16384 && per >= 32));
}

template <typename T>
static void launch_reduce(const ReduceParams<T>& p, bool thread_hint, cudaStream_t stream, int sm_count) {
    const int nred = p.d.nd - p.d.n_a;
    const bool warp = reduce_uses_warps(p, thread_hint);
    i64 threads = p.n_out * p.nsplit * (warp ? 32 : 1);
    i64 g = (threads + 255) / 256, cap = (i64)sm_count * 8;
    int grid = (int)(g > cap ? cap : (g < 1 ? 1 : g));
    if (warp) {
        if (nred == 1) reduce_warp_kernel<T, 1><<<grid, 256, 0, stream>>>(p);
        else if (nred == 2) reduce_warp_kernel<T, 2><<<grid, 256, 0, stream>>>(p);
        else reduce_warp_kernel<T, 0><<<grid, 256, 0, stream>>>(p);
    } else {
        if (nred == 1) reduce_thread_kernel<T, 1><<<grid, 256, 0, stream>>>(p);
        else if (nred == 2) reduce_thread_kernel<T, 2><<<grid, 256, 0, stream>>>(p);
        else reduce_thread_kernel<T, 0><<<grid, 256, 0, stream>>>(p);
    }
}

// ------------------------------------------------------------------------------------------
// C1: cross-rank sum of small tensors over NVLink peer memory, inside the program (no host-issued collective).
// Every rank owns one symmetric buffer, mapped into every other rank's address space (peer[r]).  Per site:
//     header  { u32 epoch (local use only); u32 flags[8] (flags[r] is written by rank r) }          256 bytes
//     data    2 x n_total elements (double-buffered by epoch parity)
// One call = pack my pieces into data[e & 1] of MY buffer; fence; write e into flags[me] of EVERY rank's header;
// wait until all of my flags reached e; add all ranks' packs IN RANK ORDER (bit-identical sums on every rank) and
// unpack in place.  A rank can be at most one epoch ahead of a peer (it cannot pass barrier e + 1 before the peer
// has signalled e + 1, which the peer does after it has finished reading epoch e), so two data buffers suffice.
// Remote data is read with volatile loads (peer lines may sit in the local L1); the wait is a bounded spin that
// traps instead of hanging the GPU if a peer never arrives.  The message is a few KB: what matters is latency
// (two NVLink traversals), not bandwidth.  reference analogue: `prev_lpq + lp` over Split chunks (logpq.py:151-153).
// ------------------------------------------------------------------------------------------
#define AB_XR_MAXW 8
#define AB_XR_MAXP 16
#define AB_XR_HDR 256
template <typename T>
struct XReduceParams {
    int rank, world, n_pieces;
    i64 n_total;
    char* site[AB_XR_MAXW];               // this site's region in every rank's symmetric buffer (header, then data)
    T* piece[AB_XR_MAXP];
    i64 piece_n[AB_XR_MAXP];
};

__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <typename T>
__device__ __forceinline__ void xreduce_body(const XReduceParams<T>& p, const i64 t0, const i64 tn) {
    __shared__ unsigned s_epoch;
    unsigned* my_hdr = reinterpret_cast<unsigned*>(p.site[p.rank]);
    if (t0 == 0) s_epoch = my_hdr[0] + 1u;
    __syncthreads();
    const unsigned e = s_epoch;
    const i64 buf = (i64)(e & 1u) * p.n_total;
    T* mine = reinterpret_cast<T*>(p.site[p.rank] + AB_XR_HDR) + buf;
    i64 off = 0;
    for (int q = 0; q < p.n_pieces; ++q) {
        for (i64 i = t0; i < p.piece_n[q]; i += tn) mine[off + i] = p.piece[q][i];
        off += p.piece_n[q];
    }
    __threadfence_system();
    __syncthreads();
    if (t0 < p.world) {
        st_release_sys_u32(reinterpret_cast<unsigned*>(p.site[t0]) + 1 + p.rank, e);       // "rank me has published epoch e"
        const unsigned* f = my_hdr + 1 + t0;
        const long long c0 = clock64();
        while ((int)(ld_acquire_sys_u32(f) - e) < 0) {
            if (clock64() - c0 > (1LL << 33)) __trap();                                   // a peer never arrived (~4 s)
        }
    }
    __syncthreads();
    off = 0;
    for (int q = 0; q < p.n_pieces; ++q) {
        for (i64 i = t0; i < p.piece_n[q]; i += tn) {
            T v[AB_XR_MAXW];
#pragma unroll
            for (int r = 0; r < AB_XR_MAXW; ++r)                             // all remote loads in flight at once ...
                v[r] = r < p.world ? *reinterpret_cast<const volatile T*>(reinterpret_cast<const T*>(p.site[r] + AB_XR_HDR) + buf + off + i) : T(0);
            T a = T(0);
#pragma unroll
            for (int r = 0; r < AB_XR_MAXW; ++r) if (r < p.world) a += v[r];   // ... added in rank order
            p.piece[q][i] = a;
        }
        off += p.piece_n[q];
    }
    __syncthreads();
    if (t0 == 0) my_hdr[0] = e;
}

template <typename T>
__global__ void __launch_bounds__(512) xreduce_kernel(const __grid_constant__ XReduceParams<T> p) {
    xreduce_body<T>(p, threadIdx.x, blockDim.x);
}

// ------------------------------------------------------------------------------------------
// Reduction sequences (plan.py ReduceSeqOp): a run of consecutive SMALL reductions -- the top-level contractions, their
// adjoints, the cross-rank sum between them -- in one single-CTA launch, __syncthreads() between the ops.  The
// parameter blocks stay in kernel-parameter space and are read with a uniform index (no staging copy: gathering a
// 1.4 KB block with one word per thread serialises on the constant cache, which is what made the general small-op
// sequences below cost as much as the launches they replaced).
// ------------------------------------------------------------------------------------------
#define AB_RSEQ_MAX 12
#define AB_RSEQ_THREADS 512
#define RC_MAXD 4
#define RC_MAXF 4
enum { RS_REDUCE = 0, RS_XREDUCE = 2 };

// A reduction in ~0.3 KB instead of the 1.4 KB of ReduceParams: <= 4 (coalesced) dims, <= 4 factors, 32-bit strides.
// Why it matters: kernel parameters live in the constant bank and arrive COLD; a small op walks its parameter block
// through dependent constant-cache misses (~0.3 us each), which is where a microsecond-scale op spends its time.
template <typename T>
struct RCompact {
    T* out; T* m_out; T* lo_out;
    const T* f[RC_MAXF]; const T* lse_m; const T* lse_lo; const T* gout;
    int fs[RC_MAXF][RC_MAXD]; int ms[RC_MAXD], ls[RC_MAXD], gs[RC_MAXD];
    int size[RC_MAXD];
    int nd, n_a, mode, nf, acc, nsplit, warp, kind;
    int n_out, n_red;
    T coeff[RC_MAXF]; T scale, cadd;
};

template <typename T>
struct RSeqParams {
    int n, pad;
    RCompact<T> op[AB_RSEQ_MAX];
    XReduceParams<T> x;                   // at most one cross-rank reduction per sequence
};

template <typename T>
__device__ __forceinline__ T rc_fsum(const RCompact<T>& p, const int* base, const int* ir) {
    T s = T(0);
#pragma unroll 1
    for (int f = 0; f < p.nf; ++f) {
        int off = base[f];
        for (int k = p.n_a; k < p.nd; ++k) off += ir[k] * p.fs[f][k];
        s += p.coeff[f] * p.f[f][off];
    }
    return s;
}
__device__ __forceinline__ void rc_unravel(int l, const int* size, int lo, int hi, int* idx) {
    for (int k = hi - 1; k >= lo; --k) { const int q = l / size[k]; idx[k] = l - q * size[k]; l = q; }
}
__device__ __forceinline__ int rc_dot(const int* idx, const int* st, int lo, int hi) {
    int off = 0;
    for (int k = lo; k < hi; ++k) off += idx[k] * st[k];
    return off;
}

// Same arithmetic, in the same order, as reduce_thread_body / reduce_warp_body (bit-identical results).
template <typename T>
__device__ void rc_body(const RCompact<T>& p, const int t0, const int tn) {
    const bool warp = p.warp != 0;
    const int lane = warp ? (t0 & 31) : 0, L = warp ? 32 : 1;
    const int worker = warp ? (t0 >> 5) : t0, nworkers = warp ? (tn >> 5) : tn;
    const int total = p.n_out * p.nsplit;
    const int chunk = (p.n_red + p.nsplit - 1) / p.nsplit;
    int idx[RC_MAXD], base[RC_MAXF];
    for (int w = worker; w < total; w += nworkers) {
        const int o = w % p.n_out, s = w / p.n_out;
        const int lo = s * chunk, hi = lo + chunk < p.n_red ? lo + chunk : p.n_red;
        rc_unravel(o, p.size, 0, p.n_a, idx);
        for (int f = 0; f < p.nf; ++f) base[f] = rc_dot(idx, p.fs[f], 0, p.n_a);
        T res;
        if (p.mode == R_SUM) {
            T a = T(0);
            for (int j = lo + lane; j < hi; j += L) { rc_unravel(j, p.size, p.n_a, p.nd, idx); a += rc_fsum(p, base, idx); }
            res = warp ? warp_sum(a) : a;
        } else if (p.mode == R_WSUM) {
            const int mb = rc_dot(idx, p.ms, 0, p.n_a), lb = rc_dot(idx, p.ls, 0, p.n_a), gb = rc_dot(idx, p.gs, 0, p.n_a);
            T a = T(0);
            for (int j = lo + lane; j < hi; j += L) {
                rc_unravel(j, p.size, p.n_a, p.nd, idx);
                const T sv = rc_fsum(p, base, idx);
                const T mm = p.lse_m[mb + rc_dot(idx, p.ms, p.n_a, p.nd)];
                const T l = p.lse_lo[lb + rc_dot(idx, p.ls, p.n_a, p.nd)];
                const T g = p.gout[gb + rc_dot(idx, p.gs, p.n_a, p.nd)];
                a += g * ab_exp((sv - mm) - l);
            }
            res = warp ? warp_sum(a) : a;
        } else {
            T m = neg_inf<T>();
            for (int j = lo + lane; j < hi; j += L) { rc_unravel(j, p.size, p.n_a, p.nd, idx); m = ab_max(m, rc_fsum(p, base, idx)); }
            if (warp) m = warp_max(m);
            T a = T(0);
            for (int j = lo + lane; j < hi; j += L) { rc_unravel(j, p.size, p.n_a, p.nd, idx); a += ab_exp(rc_fsum(p, base, idx) - m); }
            if (warp) a = warp_sum(a);
            const T lg = p.mode == R_LSE_EPS ? ab_log(a + Eps<T>::v()) : ab_log(a);
            if (lane == 0 && p.m_out != nullptr) { p.m_out[o] = m; p.lo_out[o] = lg; }
            res = lg + m;
        }
        if (lane == 0) {
            if (p.nsplit > 1) p.out[s * p.n_out + o] = res;
            else {
                const T v = p.scale * res + (p.mode == R_WSUM ? T(0) : p.cadd);
                p.out[o] = p.acc ? p.out[o] + v : v;
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(AB_RSEQ_THREADS) reduce_seq_kernel(const __grid_constant__ RSeqParams<T> sp) {
    __shared__ __align__(16) RCompact<T> s_op[AB_RSEQ_MAX];
    const int t0 = threadIdx.x, tn = blockDim.x, lane = t0 & 31, wid = t0 >> 5, nw = tn >> 5;
    {
        // parameter blocks -> shared memory: every warp fetches whole 64-byte lines with warp-uniform addresses (one
        // constant-cache transaction per word, the cold misses of different warps overlap)
        const unsigned* src = reinterpret_cast<const unsigned*>(&sp.op[0]);
        unsigned* dst = reinterpret_cast<unsigned*>(s_op);
        const int nwords = sp.n * (int)(sizeof(RCompact<T>) / 4);
        for (int line = wid; line * 16 < nwords; line += nw) {
#pragma unroll
            for (int w = 0; w < 16; ++w) {
                const int k = line * 16 + w;
                if (k < nwords) { const unsigned v = src[k]; if (lane == w) dst[k] = v; }
            }
        }
    }
    __syncthreads();
    for (int i = 0; i < sp.n; ++i) {
        if (i) __syncthreads();                   // same CTA: the previous op's global-memory results are visible
        if (s_op[i].kind == RS_XREDUCE) { __threadfence(); xreduce_body<T>(sp.x, (i64)t0, (i64)tn); continue; }
        rc_body<T>(s_op[i], t0, tn);
    }
}

// host: ReduceParams -> RCompact when it fits (else the sequence runs as separate launches)
template <typename T>
static bool rc_from(const ReduceParams<T>& p, bool warp, RCompact<T>& c) {
    memset(&c, 0, sizeof(c));
    if (p.d.nd > RC_MAXD || p.nf > RC_MAXF) return false;
    if (p.n_out * p.nsplit > 0x3fffffff || p.n_red > 0x3fffffff) return false;
    auto fits = [&](const Opnd& o, int* st) {
        if (o.mode != 0) return false;
        for (int k = 0; k < p.d.nd; ++k) { if (o.stride[k] > 0x3fffffff || o.stride[k] < -0x3fffffff) return false; st[k] = (int)o.stride[k]; }
        return true;
    };
    for (int f = 0; f < p.nf; ++f) { if (!fits(p.f[f], c.fs[f])) return false; c.f[f] = (const T*)p.f[f].ptr; c.coeff[f] = p.coeff[f]; }
    if (p.mode == R_WSUM) {
        if (!fits(p.lse_m, c.ms) || !fits(p.lse_lo, c.ls) || !fits(p.gout, c.gs)) return false;
        c.lse_m = (const T*)p.lse_m.ptr; c.lse_lo = (const T*)p.lse_lo.ptr; c.gout = (const T*)p.gout.ptr;
    }
    for (int k = 0; k < p.d.nd; ++k) c.size[k] = p.d.size[k];
    c.out = p.out; c.m_out = p.m_out; c.lo_out = p.lo_out;
    c.nd = p.d.nd; c.n_a = p.d.n_a; c.mode = p.mode; c.nf = p.nf; c.acc = p.acc; c.nsplit = p.nsplit;
    c.warp = warp ? 1 : 0; c.kind = RS_REDUCE;
    c.n_out = (int)p.n_out; c.n_red = (int)p.n_red; c.scale = p.scale; c.cadd = p.cadd;
    return true;
}

// ------------------------------------------------------------------------------------------
// Small-op sequences.  A plate tree has dozens of ops whose whole iteration space is a few thousand
// points (global latents, top-level contractions, their adjoints): as separate launches each costs a
// launch latency plus a drain, several microseconds for nanoseconds of work.  Consecutive small ops of
// a program are therefore executed by ONE launch of one 1024-thread CTA that runs them back to back,
// __syncthreads() between ops (same CTA, so global-memory results of one op are visible to the next).
// The op bodies are the very same device functions the stand-alone kernels call.
// ------------------------------------------------------------------------------------------
enum { SK_EXPR = 0, SK_EXPR_BWD = 1, SK_REDUCE = 2, SK_FILL = 3, SK_XREDUCE = 4 };
#define AB_SEQ_MAX 16
#define AB_SEQ_POINTS 4096        // an op is "small" when its iteration space has at most this many points

template <typename T>
struct SeqOp {
    int kind;
    int warp;                    // lanes over the reduced / looped index
    Normal3 n3;
    union {
        ExprParams<T> e;
        ExprBwdParams<T> b;
        ReduceParams<T> r;
        XReduceParams<T> x;
        struct { void* ptr; i64 nbytes; } f;
    };
    __host__ __device__ SeqOp() {}
};

// Resident tensors of a sequence: the small tensors its ops pass to one another (and their small inputs) live in
// SHARED memory for the duration of the launch.  Without this every op of the sequence pays global-memory round
// trips for operands its predecessor has just produced (~1 us each with a flushed L2: measured, a 9-op sequence
// cost MORE than 9 launches inside a CUDA graph); with it an op costs a few hundred cycles.  The host finds the
// candidates (operand footprint <= AB_RES_TENSOR bytes, AB_RES_TOTAL in all), replaces their pointers in the op
// parameters by tags (AB_RES_TAG << 48 | offset), and the kernel resolves the tags against its shared-memory
// base, loads the tensors that are read before they are (completely) written, and writes back every tensor the
// sequence wrote -- later programs (the adjoint pass) and the big kernels read them from global memory.
// (Writing real shared-memory addresses into the parameters on the host does not work: nvcc assumes that pointers
// arriving in kernel parameters are global and emits LDG/STG for them, which fault on the shared window.  The copy
// of each op's parameter block into shared memory, where the tags are resolved, is what this variant pays: with it
// an op costs ~3 us, so sequences stay opt-in -- see alan_b200_plan_create.)
#define AB_RES_MAX 56
#define AB_RES_TAG 0xA1B2ull
#define AB_RES_TENSOR (32 * 1024)
#define AB_RES_TOTAL (160 * 1024)
struct ResEnt { void* g; int nbytes; int off; int flags; int pad; };      // flags: 1 = load at start, 2 = store at end

template <typename T>
struct SeqParams {
    int n, n_res, smem_bytes, pad;
    ResEnt res[AB_RES_MAX];
    SeqOp<T> op[AB_SEQ_MAX];
    __host__ __device__ SeqParams() : n(0), n_res(0), smem_bytes(0), pad(0) {}
};

static_assert(sizeof(SeqParams<double>) <= 32000, "SeqParams travels as a kernel parameter (32 764-byte limit)");

template <typename P>
__device__ __forceinline__ void res_fix(P*& p, unsigned char* base) {
    const unsigned long long v = (unsigned long long)p;
    if ((v >> 48) == AB_RES_TAG) p = reinterpret_cast<P*>(base + (v & 0xFFFFFFFFull));
}

template <typename T>
__device__ void seq_resolve(SeqOp<T>& o, unsigned char* base) {
    switch (o.kind) {
        case SK_EXPR:
            for (int l = 0; l < o.e.n_leaves; ++l) res_fix(o.e.leaf[l].ptr, base);
            res_fix(o.e.out, base);
            break;
        case SK_EXPR_BWD:
            for (int l = 0; l < o.b.n_leaves; ++l) res_fix(o.b.leaf[l].ptr, base);
            res_fix(o.b.gout.ptr, base); res_fix(o.b.gleaf, base);
            break;
        case SK_REDUCE:
            for (int f = 0; f < o.r.nf; ++f) res_fix(o.r.f[f].ptr, base);
            res_fix(o.r.lse_m.ptr, base); res_fix(o.r.lse_lo.ptr, base); res_fix(o.r.gout.ptr, base);
            res_fix(o.r.out, base); res_fix(o.r.m_out, base); res_fix(o.r.lo_out, base);
            break;
        case SK_FILL: res_fix(o.f.ptr, base); break;
        case SK_XREDUCE:
            for (int q = 0; q < o.x.n_pieces; ++q) res_fix(o.x.piece[q], base);
            break;
    }
}

#define AB_SEQ_THREADS 512
template <typename T>
__global__ void __launch_bounds__(AB_SEQ_THREADS) small_seq_kernel(const __grid_constant__ SeqParams<T> sp) {
    extern __shared__ __align__(16) unsigned char seq_smem[];
    constexpr int CUR = (int)((sizeof(SeqOp<T>) + 15) & ~(size_t)15);
    SeqOp<T>& cur = *reinterpret_cast<SeqOp<T>*>(seq_smem);          // the running op's parameters, tags resolved
    unsigned char* base = seq_smem + CUR;
    const i64 t0 = threadIdx.x, tn = blockDim.x;
    for (int e = 0; e < sp.n_res; ++e) {
        if (!(sp.res[e].flags & 1)) continue;
        const unsigned* g = reinterpret_cast<const unsigned*>(sp.res[e].g);
        unsigned* d = reinterpret_cast<unsigned*>(base + sp.res[e].off);
        for (int k = (int)t0; k < sp.res[e].nbytes / 4; k += (int)tn) d[k] = g[k];
    }
    for (int i = 0; i < sp.n; ++i) {
        __syncthreads();                                             // the previous op is done with `cur` and its results are visible
        {
            const unsigned* src = reinterpret_cast<const unsigned*>(&sp.op[i]);
            unsigned* dst = reinterpret_cast<unsigned*>(seq_smem);
            for (int k = (int)t0; k < (int)(sizeof(SeqOp<T>) / 4); k += (int)tn) dst[k] = src[k];
        }
        __syncthreads();
        if (t0 == 0) seq_resolve(cur, base);
        __syncthreads();
        const SeqOp<T>& o = cur;
        switch (o.kind) {
            case SK_EXPR:
                if (o.warp) { if (o.n3.on) expr_fwd_body<T, true, true>(o.e, o.n3, t0, tn); else expr_fwd_body<T, true, false>(o.e, o.n3, t0, tn); }
                else { if (o.n3.on) expr_fwd_body<T, false, true>(o.e, o.n3, t0, tn); else expr_fwd_body<T, false, false>(o.e, o.n3, t0, tn); }
                break;
            case SK_EXPR_BWD:
                if (o.warp) { if (o.n3.on) expr_bwd_body<T, true, true>(o.b, o.n3, t0, tn); else expr_bwd_body<T, true, false>(o.b, o.n3, t0, tn); }
                else { if (o.n3.on) expr_bwd_body<T, false, true>(o.b, o.n3, t0, tn); else expr_bwd_body<T, false, false>(o.b, o.n3, t0, tn); }
                break;
            case SK_REDUCE: {
                const int nred = o.r.d.nd - o.r.d.n_a;
                if (o.warp) { if (nred == 1) reduce_warp_body<T, 1>(o.r, t0, tn); else reduce_warp_body<T, 0>(o.r, t0, tn); }
                else { if (nred == 1) reduce_thread_body<T, 1>(o.r, t0, tn); else reduce_thread_body<T, 0>(o.r, t0, tn); }
                break;
            }
            case SK_FILL: {
                unsigned* w = (unsigned*)o.f.ptr;                       // 4-byte granularity (all our tensors)
                for (i64 k = t0; k < o.f.nbytes / 4; k += tn) w[k] = 0u;
                break;
            }
            case SK_XREDUCE:
                __threadfence();                                        // earlier ops' results (other CTAs' writes were
                xreduce_body<T>(o.x, t0, tn);                           // completed by the launch boundary) -> pack
                break;
        }
    }
    __syncthreads();
    for (int e = 0; e < sp.n_res; ++e) {
        if (!(sp.res[e].flags & 2)) continue;
        unsigned* g = reinterpret_cast<unsigned*>(sp.res[e].g);
        const unsigned* d = reinterpret_cast<const unsigned*>(base + sp.res[e].off);
        for (int k = (int)t0; k < sp.res[e].nbytes / 4; k += (int)tn) g[k] = d[k];
    }
}

// Host side: decide which operands of the collected ops become resident and tag their pointers.
template <typename T>
struct SeqResidency {
    struct Use { void* base; size_t bytes; bool first_write; size_t first_bytes; bool written; };
    std::vector<Use> uses;
    Use& at(const void* p) {
        for (auto& u : uses) if (u.base == p) return u;
        uses.push_back(Use{const_cast<void*>(p), 0, false, 0, false});
        uses.back().first_bytes = (size_t)-1;
        return uses.back();
    }
    void touch(const void* p, size_t bytes, bool write) {
        if (!p) return;
        Use& u = at(p);
        if (u.first_bytes == (size_t)-1) { u.first_write = write; u.first_bytes = bytes; }
        if (bytes > u.bytes) u.bytes = bytes;
        if (write) u.written = true;
    }
    static size_t fp(const Opnd& o, const Dims& d) {
        i64 n = 1;
        for (int k = 0; k < d.nd; ++k) n += (i64)(d.size[k] - 1) * (o.stride[k] < 0 ? -o.stride[k] : o.stride[k]);
        return (size_t)n * sizeof(T);
    }
    template <typename F> static void each(SeqOp<T>& o, F&& f) {      // f(pointer reference, footprint bytes, is write, reads first)
        switch (o.kind) {
            case SK_EXPR:
                for (int l = 0; l < o.e.n_leaves; ++l) f(const_cast<void*&>(o.e.leaf[l].ptr), fp(o.e.leaf[l], o.e.d), false, true);
                f(reinterpret_cast<void*&>(o.e.out), (size_t)o.e.n_out * sizeof(T), true, o.e.acc != 0);
                break;
            case SK_EXPR_BWD:
                for (int l = 0; l < o.b.n_leaves; ++l) f(const_cast<void*&>(o.b.leaf[l].ptr), fp(o.b.leaf[l], o.b.d), false, true);
                f(const_cast<void*&>(o.b.gout.ptr), fp(o.b.gout, o.b.d), false, true);
                f(reinterpret_cast<void*&>(o.b.gleaf), (size_t)o.b.n_kept * o.b.nsplit * sizeof(T), true, o.b.acc != 0 && o.b.nsplit == 1);
                break;
            case SK_REDUCE:
                for (int q = 0; q < o.r.nf; ++q) f(const_cast<void*&>(o.r.f[q].ptr), fp(o.r.f[q], o.r.d), false, true);
                if (o.r.mode == R_WSUM) {
                    f(const_cast<void*&>(o.r.lse_m.ptr), fp(o.r.lse_m, o.r.d), false, true);
                    f(const_cast<void*&>(o.r.lse_lo.ptr), fp(o.r.lse_lo, o.r.d), false, true);
                    f(const_cast<void*&>(o.r.gout.ptr), fp(o.r.gout, o.r.d), false, true);
                }
                f(reinterpret_cast<void*&>(o.r.out), (size_t)o.r.n_out * o.r.nsplit * sizeof(T), true, o.r.acc != 0 && o.r.nsplit == 1);
                if (o.r.m_out) {
                    f(reinterpret_cast<void*&>(o.r.m_out), (size_t)o.r.n_out * sizeof(T), true, false);
                    f(reinterpret_cast<void*&>(o.r.lo_out), (size_t)o.r.n_out * sizeof(T), true, false);
                }
                break;
            case SK_FILL: f(o.f.ptr, (size_t)o.f.nbytes, true, false); break;
            case SK_XREDUCE:
                for (int q = 0; q < o.x.n_pieces; ++q) f(reinterpret_cast<void*&>(o.x.piece[q]), (size_t)o.x.piece_n[q] * sizeof(T), true, true);
                break;
        }
    }
    // returns the dynamic shared-memory bytes of the launch
    int build(SeqParams<T>& sp) {
        uses.clear();
        for (int i = 0; i < sp.n; ++i)
            each(sp.op[i], [&](void*& p, size_t bytes, bool write, bool reads_first) {
                if (write && reads_first) touch(p, bytes, false);
                touch(p, bytes, write);
            });
        constexpr int CUR = (int)((sizeof(SeqOp<T>) + 15) & ~(size_t)15);
        size_t off = 0;
        sp.n_res = 0;
        std::vector<int> slot(uses.size(), -1);
        for (size_t k = 0; k < uses.size(); ++k) {
            const Use& u = uses[k];
            const size_t need = (u.bytes + 15) & ~(size_t)15;
            if (u.bytes == 0 || u.bytes > AB_RES_TENSOR || off + need > AB_RES_TOTAL || sp.n_res == AB_RES_MAX) continue;
            if (((unsigned long long)u.base >> 48) != 0 || ((unsigned long long)u.base & 3)) continue;
            ResEnt& e = sp.res[sp.n_res];
            e.g = u.base; e.nbytes = (int)((u.bytes + 3) & ~(size_t)3); e.off = (int)off; e.pad = 0;
            const bool covered = u.first_write && u.first_bytes >= u.bytes;      // completely written before any read
            e.flags = (covered ? 0 : 1) | (u.written ? 2 : 0);
            slot[k] = sp.n_res++;
            off += need;
        }
        for (int i = 0; i < sp.n; ++i)
            each(sp.op[i], [&](void*& p, size_t, bool, bool) {
                if (!p) return;
                for (size_t k = 0; k < uses.size(); ++k)
                    if (uses[k].base == p && slot[k] >= 0) {
                        p = reinterpret_cast<void*>((AB_RES_TAG << 48) | (unsigned long long)sp.res[slot[k]].off);
                        return;
                    }
            });
        sp.smem_bytes = CUR + (int)off;
        return sp.smem_bytes;
    }
};

// ------------------------------------------------------------------------------------------
// K5: Timeseries chain.  One CTA per (pair, outer).  reference utils.py:478-510.
//   C[i,k] = log( sum_j exp(A[i,j]-a_i) exp(B[j,k]-b_k) + eps ) + a_i + b_k
//   a_i = max_j A[i,j], b_k = max_j B[j,k]; the odd tail of a level is carried unreduced.
// Dynamic smem: 2*K*K + 2*K elements.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) chain_level_kernel(const T* __restrict__ X, T* __restrict__ Y,
                                                          int n_in, int n_outl, int K) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* EA = (T*)smem_raw;
    T* EB = EA + K * K;
    T* am = EB + K * K;
    T* bm = am + K;
    const int i_out = blockIdx.x, outer = blockIdx.y;
    const i64 KK = (i64)K * K;
    const T* Xo = X + (i64)outer * n_in * KK;
    T* Yo = Y + (i64)outer * n_outl * KK + (i64)i_out * KK;
    const int npairs = n_in / 2;
    if (i_out >= npairs) {                      // carried remainder (last element of this level)
        const T* src = Xo + (i64)(n_in - 1) * KK;
        for (int e = threadIdx.x; e < KK; e += blockDim.x) Yo[e] = src[e];
        return;
    }
    const T* A = Xo + (i64)(2 * i_out) * KK;
    const T* B = A + KK;
    for (int e = threadIdx.x; e < KK; e += blockDim.x) { EA[e] = A[e]; EB[e] = B[e]; }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        T m = neg_inf<T>();
        for (int j = 0; j < K; ++j) m = ab_max(m, EA[i * K + j]);
        am[i] = m;
        T n = neg_inf<T>();
        for (int j = 0; j < K; ++j) n = ab_max(n, EB[j * K + i]);
        bm[i] = n;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < KK; e += blockDim.x) {
        EA[e] = ab_exp(EA[e] - am[e / K]);
        EB[e] = ab_exp(EB[e] - bm[e % K]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < KK; e += blockDim.x) {
        int i = e / K, k = e % K;
        T acc = T(0);
        for (int j = 0; j < K; ++j) acc += EA[i * K + j] * EB[j * K + k];
        Yo[e] = ab_log(acc + Eps<T>::v()) + am[i] + bm[k];
    }
}

// out[outer, i] = logsumexp_k X[outer, i, k]   (no eps; torch.logsumexp, logpq.py:139)
template <typename T>
__global__ void chain_final_kernel(const T* __restrict__ X, T* __restrict__ out, i64 n_rows, int K) {
    for (i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (i64)gridDim.x * blockDim.x) {
        const T* x = X + r * K;
        T m = neg_inf<T>();
        for (int k = 0; k < K; ++k) m = ab_max(m, x[k]);
        T mm = (m == neg_inf<T>()) ? T(0) : m;
        T a = T(0);
        for (int k = 0; k < K; ++k) a += ab_exp(x[k] - mm);
        out[r] = ab_log(a) + mm;
    }
}

template <typename T>
__global__ void chain_final_bwd_kernel(const T* __restrict__ X, const T* __restrict__ out,
                                       const T* __restrict__ gout, T* __restrict__ gX, i64 n_rows, int K) {
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n_rows * K; e += (i64)gridDim.x * blockDim.x) {
        i64 r = e / K;
        gX[e] = gout[r] * ab_exp(X[e] - out[r]);
    }
}

// Adjoint of one chain level, one CTA per (output element, outer).  Includes the path through
// the amax shifts (weight eps/(P+eps), split evenly among ties as torch.amax does): the shifted
// product P can be << 1, so that term is visible at 1e-5 (SURVEY.md "hard parts").
// Dynamic smem: 3*K*K + 4*K elements.
template <typename T>
__global__ void __launch_bounds__(256) chain_level_bwd_kernel(const T* __restrict__ X, const T* __restrict__ gY,
                                                              T* __restrict__ gX, int n_in, int n_outl, int K) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* EA = (T*)smem_raw;
    T* EB = EA + K * K;
    T* D = EB + K * K;      // gY / (P + eps)
    T* am = D + K * K;
    T* bm = am + K;
    T* ga = bm + K;         // adjoint of a_i
    T* gb = ga + K;         // adjoint of b_k
    const int i_out = blockIdx.x, outer = blockIdx.y;
    const i64 KK = (i64)K * K;
    const T* Xo = X + (i64)outer * n_in * KK;
    T* gXo = gX + (i64)outer * n_in * KK;
    const T* g = gY + (i64)outer * n_outl * KK + (i64)i_out * KK;
    const int npairs = n_in / 2;
    if (i_out >= npairs) {
        T* dst = gXo + (i64)(n_in - 1) * KK;
        for (int e = threadIdx.x; e < KK; e += blockDim.x) dst[e] = g[e];
        return;
    }
    const T* A = Xo + (i64)(2 * i_out) * KK;
    const T* B = A + KK;
    T* gA = gXo + (i64)(2 * i_out) * KK;
    T* gB = gA + KK;
    for (int e = threadIdx.x; e < KK; e += blockDim.x) { EA[e] = A[e]; EB[e] = B[e]; }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        T m = neg_inf<T>(), n = neg_inf<T>();
        for (int j = 0; j < K; ++j) { m = ab_max(m, EA[i * K + j]); n = ab_max(n, EB[j * K + i]); }
        am[i] = m; bm[i] = n;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < KK; e += blockDim.x) {
        EA[e] = ab_exp(EA[e] - am[e / K]);
        EB[e] = ab_exp(EB[e] - bm[e % K]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < KK; e += blockDim.x) {
        int i = e / K, k = e % K;
        T acc = T(0);
        for (int j = 0; j < K; ++j) acc += EA[i * K + j] * EB[j * K + k];
        D[e] = g[e] / (acc + Eps<T>::v());
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        T sa = T(0), sb = T(0);
        for (int k = 0; k < K; ++k) { sa += D[i * K + k]; sb += D[k * K + i]; }
        ga[i] = sa * Eps<T>::v();
        gb[i] = sb * Eps<T>::v();
    }
    __syncthreads();
    for (int e = threadIdx.x; e < KK; e += blockDim.x) {
        int i = e / K, j = e % K;
        // gA[i,j]
        T s = T(0);
        for (int k = 0; k < K; ++k) s += D[i * K + k] * EB[j * K + k];
        T v = s * EA[e];
        if (A[e] == am[i]) { int ties = 0; for (int jj = 0; jj < K; ++jj) ties += (A[i * K + jj] == am[i]); v += ga[i] / T(ties); }
        gA[e] = v;
        // gB[i,j] viewed as B[j'=i, k=j]
        int jr = i, k = j;
        T s2 = T(0);
        for (int ii = 0; ii < K; ++ii) s2 += EA[ii * K + jr] * D[ii * K + k];
        T v2 = s2 * EB[e];
        if (B[e] == bm[k]) { int ties = 0; for (int jj = 0; jj < K; ++jj) ties += (B[jj * K + k] == bm[k]); v2 += gb[k] / T(ties); }
        gB[e] = v2;
    }
}

// ------------------------------------------------------------------------------------------
// K7: posterior resampling of one contraction step (reference reduce_Ks.py:51-75).
// One thread per (n, plate cell).  lp_j = sum_f coeff_f F_f gathered at the already-sampled
// parent indices; p_j = exp(lp_j - max) in the factor dtype; index = first j with cumsum_j (float64) >= u * total;
// j is unravelled row-major over the step's K axes (unravel_index.py:97-100).
// ------------------------------------------------------------------------------------------
#define AB_MAXK 4
#define AB_MAXG 6
#define AB_MAXIDX 16
template <typename T>
struct SampleParams {
    Dims d;                              // batch dims: [N, plates...]; n_a = nd
    int nk; int ksize[AB_MAXK]; i64 ktotal;
    int nf;
    Opnd f[AB_MAXL];                     // strides over batch dims
    T coeff[AB_MAXL];
    i64 kstride[AB_MAXL][AB_MAXK];
    int ng[AB_MAXL];
    i64 gstride[AB_MAXL][AB_MAXG];
    int gsel[AB_MAXL][AB_MAXG];          // which idx tensor
    int n_idx;
    const i64* idxptr[AB_MAXIDX];
    i64 idxstride[AB_MAXIDX][AB_MAXD];
    const double* u;
    i64 ustride[AB_MAXD];
    i64* out[AB_MAXK];
    i64 n_batch;
};

template <typename T>
__device__ __forceinline__ T sample_lp(const SampleParams<T>& p, const i64* base, i64 j) {
    int kidx[AB_MAXK];
    for (int k = p.nk - 1; k >= 0; --k) { kidx[k] = (int)(j % p.ksize[k]); j /= p.ksize[k]; }
    T s = T(0);
    for (int f = 0; f < p.nf; ++f) {
        i64 off = base[f];
        for (int k = 0; k < p.nk; ++k) off += (i64)kidx[k] * p.kstride[f][k];
        s += p.coeff[f] * ((const T*)p.f[f].ptr)[off];
    }
    return s;
}

template <typename T>
__global__ void __launch_bounds__(128) sample_kernel(const __grid_constant__ SampleParams<T> p) {
    int idx[AB_MAXD];
    i64 base[AB_MAXL];
    i64 gv[AB_MAXIDX];
    for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < p.n_batch; b += (i64)gridDim.x * blockDim.x) {
        unravel(b, p.d, 0, p.d.nd, idx);
        for (int t = 0; t < p.n_idx; ++t) {
            i64 off = 0;
            for (int k = 0; k < p.d.nd; ++k) off += (i64)idx[k] * p.idxstride[t][k];
            gv[t] = p.idxptr[t][off];
        }
        for (int f = 0; f < p.nf; ++f) {
            i64 off = dot_stride(p.f[f], idx, 0, p.d.nd);
            for (int g = 0; g < p.ng[f]; ++g) off += gv[p.gsel[f][g]] * p.gstride[f][g];
            base[f] = off;
        }
        i64 uoff = 0;
        for (int k = 0; k < p.d.nd; ++k) uoff += (i64)idx[k] * p.ustride[k];
        const double u = p.u[uoff];
        T m = neg_inf<T>();
        for (i64 j = 0; j < p.ktotal; ++j) m = ab_max(m, sample_lp(p, base, j));
        // p_j = exp(lp_j - max) in the FACTOR dtype, as the reference forms it before torch.multinomial
        // (reduce_Ks.py:62-66); the cumulative sum and the comparison with u are in float64
        double total = 0.0;
        for (i64 j = 0; j < p.ktotal; ++j) total += (double)ab_exp(sample_lp(p, base, j) - m);
        const double thr = u * total;
        double c = 0.0;
        i64 pick = p.ktotal - 1;
        for (i64 j = 0; j < p.ktotal; ++j) {
            c += (double)ab_exp(sample_lp(p, base, j) - m);
            if (!(c < thr)) { pick = j; break; }
        }
        for (int k = p.nk - 1; k >= 0; --k) { p.out[k][b] = pick % p.ksize[k]; pick /= p.ksize[k]; }
    }
}

// ------------------------------------------------------------------------------------------
// K8: gather samples at resampled indices (reference Sample.py:359-381).
// x: [outer, K, inner]; idx: [N, outer / outer_div]; out: [N, outer, inner]; 4- or 8-byte elements.
// ------------------------------------------------------------------------------------------
template <typename E>
__global__ void gather_kernel(const E* __restrict__ x, const i64* __restrict__ idx, E* __restrict__ out,
                              i64 N, i64 outer, i64 K, i64 inner, i64 outer_div) {
    const i64 total = N * outer * inner;
    const i64 og = outer / outer_div;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        i64 in = e % inner;
        i64 o = (e / inner) % outer;
        i64 n = e / (inner * outer);
        i64 k = idx[n * og + o / outer_div];
        out[e] = x[(o * K + k) * inner + in];
    }
}

// ------------------------------------------------------------------------------------------
// widen: byte-typed inputs (binary features, 0/1 observations, small counts) travel over PCIe as bytes and become
// the working dtype here, 16 bytes per thread per pass (the reference keeps such inputs as float tensors on the
// host: examples/models/movielens/movielens.py:11-22; their values are exact in either type).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) widen_u8_kernel(const unsigned char* __restrict__ src, T* __restrict__ dst, i64 n) {
    const i64 nv = n / 16;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (i64)gridDim.x * blockDim.x) {
        const uint4 v = s4[i];
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        T* o = dst + i * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (sizeof(T) == 4) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(o) + 4 * q) =
                    make_float4((float)(w[q] & 255u), (float)((w[q] >> 8) & 255u), (float)((w[q] >> 16) & 255u), (float)(w[q] >> 24));
            } else {
#pragma unroll
                for (int b = 0; b < 4; ++b) o[4 * q + b] = (T)((w[q] >> (8 * b)) & 255u);
            }
        }
    }
    for (i64 i = nv * 16 + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        dst[i] = (T)src[i];
}
