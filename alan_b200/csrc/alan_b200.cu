// alan_b200.cu -- plan executor and C ABI (include/alan_b200.h) of the logPQ engine.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include "../../include/alan_b200.h"
#include "kernels.cuh"
#include "fused.cuh"
#include "fan_tc.cuh"
#include "fan_tc2.cuh"
#include "fan_tc2b.cuh"
#include "qfactor.cuh"
#include "sampling.cuh"
#include "normal_poly.cuh"
#include "chain.cuh"
#include "qem.cuh"
#include "mvn.cuh"
#include <type_traits>
#include <cstdlib>

#include <string>
#include <vector>
#include <map>
#include <cstdio>
#include <cstring>
#include <mutex>

#define AB_MAGIC 0x0A1AB200
#define AB_VERSION 1

enum { SP_WS = 0, SP_INPUT = 1, SP_OUTPUT = 2, SP_AUX = 3 };
enum { OP_FILL = 1, OP_EXPR = 2, OP_EXPR_BWD = 3, OP_REDUCE = 4, OP_CHAIN = 5, OP_CHAIN_BWD = 6, OP_SAMPLE = 7,
       OP_NORMAL_FAN = 8, OP_COPY = 9, OP_DOT = 10, OP_FAN_LSE = 11, OP_BERN_DOT = 12, OP_FAN_BWD = 13, OP_XREDUCE = 14, OP_NORMAL_Q_BWD = 15, OP_PERM = 16, OP_KGATHER = 17,
       OP_TS_SAMPLE = 18, OP_DEPS = 19, OP_NORMAL_POLY_SUM = 20, OP_PASTE = 21, OP_MVN_PREP = 22, OP_RSEQ = 23 };

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }

// One instantiated CUDA graph of a program for one binding of pointers (inputs, outputs, aux,
// workspace) and one stream.  A program is a fixed sequence of launches, so after the first call the
// whole sequence is replayed with a single cudaGraphLaunch: no per-op host work, no launch gaps.
struct GraphEntry {
    std::vector<const void*> key;
    cudaGraphExec_t exec;
    unsigned long long last_use;
};

struct alan_b200_plan {
    std::vector<int32_t> blob;
    int dtype, n_inputs, n_programs, n_fwd, n_bwd, sample_prog;
    size_t ws_bytes;
    std::vector<int> prog_start, prog_nops;
    int sm_count;
    // graph cache (mutable state behind a const handle: guarded by mu)
    mutable std::mutex mu, mu_par;
    mutable std::vector<std::vector<GraphEntry>> graphs;
    mutable std::vector<int> n_out, n_aux;          // per program, learnt on the first run (-1 = unknown)
    mutable unsigned long long tick = 0;
    bool use_graphs = true;
    bool graph_auto = false;       // replay only bindings that repeat (see alan_b200_plan_create)
    mutable std::vector<int> miss_streak;   // per program: captures since the last replay hit (auto mode gives up at 4)
    bool use_seq = false;          // run consecutive small ops as one launch (ALAN_B200_SEQ=0 at plan creation: off)
    i64 seq_points = AB_SEQ_POINTS;   // an op is small up to this many iteration points (ALAN_B200_SEQ_POINTS)
    bool seq_resident = true;      // small tensors of a sequence live in shared memory for the launch (ALAN_B200_SEQ_RESIDENT=0: global)
    bool seq_bigsum = false;        // fixed-order sums of partial rows with few outputs also count as small (ALAN_B200_SEQ_BIGSUM=0: no)
    bool use_tc = true;            // fan_lse on tcgen05 where the shape allows (ALAN_B200_NO_TC=1 at plan creation: FFMA2 kernel)
    bool use_tc2 = true;           // dense formulation (fan_tc2.cuh) where the loc is independent of the value's axes
                                   // (ALAN_B200_TC_BLOCKDIAG=1 at plan creation: block-diagonal kernel only)
    // cross-rank reductions inside programs (OP_XREDUCE): the ranks' symmetric buffers as mapped in this process and
    // the byte offset of every reduction site in them (alan_b200_plan_set_comm / alan_b200_comm_bytes)
    int comm_rank = 0, comm_world = 1;
    char* comm_peer[AB_XR_MAXW] = {nullptr};
    std::vector<size_t> site_off, site_elems;
    size_t comm_bytes = 0;
    // graphs cannot be captured on / launched into the legacy default stream: calls that arrive on it are
    // forwarded to this private stream, ordered by a pair of events
    mutable cudaStream_t side = nullptr;
    mutable cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    // Independent ops of a program on parallel branches: while a program is being CAPTURED into a CUDA graph, an op is
    // enqueued on one of a few side streams unless it depends (OP_DEPS table written by the planner) on the op before
    // it, so the instantiated graph carries the program's real dependency DAG instead of a chain -- the dozens of
    // microsecond-scale ops of the global latents and of the top-level contraction run beside the big kernels and
    // beside each other (ALAN_B200_PAR=0 at plan creation: chain).
    bool use_par = true;
    mutable std::vector<cudaStream_t> par_streams;
    mutable std::vector<cudaEvent_t> par_events;
    ~alan_b200_plan() {
        for (auto st : par_streams) cudaStreamDestroy(st);
        for (auto ev : par_events) cudaEventDestroy(ev);
        for (auto& v : graphs) for (auto& g : v) cudaGraphExecDestroy(g.exec);
        if (ev_in) cudaEventDestroy(ev_in);
        if (ev_out) cudaEventDestroy(ev_out);
        if (side) cudaStreamDestroy(side);
    }
};

struct Reader {
    const int32_t* p;
    int32_t i32() { return *p++; }
    i64 i64v() { uint32_t lo = (uint32_t)p[0]; uint32_t hi = (uint32_t)p[1]; p += 2; return (i64)(((uint64_t)hi << 32) | lo); }
    double f64() { i64 v = i64v(); double d; memcpy(&d, &v, 8); return d; }
};

struct Ctx {
    const void* const* inputs;
    void* const* outputs;
    const void* const* aux;
    char* ws;
    cudaStream_t stream;
    int sm_count;
    mutable int max_out = -1, max_aux = -1;        // highest output / aux slot the program touched
};

static void* tref(Reader& r, const Ctx& c) {
    int space = r.i32();
    i64 v = r.i64v();
    switch (space) {
        case SP_WS: return c.ws + v;
        case SP_INPUT: return c.inputs ? (void*)c.inputs[v] : nullptr;
        case SP_OUTPUT: if ((int)v > c.max_out) c.max_out = (int)v; return c.outputs ? c.outputs[v] : nullptr;
        case SP_AUX: if ((int)v > c.max_aux) c.max_aux = (int)v; return c.aux ? (void*)c.aux[v] : nullptr;
    }
    return nullptr;
}

static int grid_for(i64 n, int block, const Ctx& c, int per_sm = 8) {
    i64 g = (n + block - 1) / block;
    i64 cap = (i64)c.sm_count * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

static void read_dims(Reader& r, Dims& d, i64& na_total, i64& nb_total) {
    d.n_a = r.i32();
    int nb = r.i32();
    d.nd = d.n_a + nb;
    na_total = 1; nb_total = 1;
    for (int k = 0; k < d.nd; ++k) {
        d.size[k] = r.i32();
        if (k < d.n_a) na_total *= d.size[k]; else nb_total *= d.size[k];
    }
}

static void read_opnd(Reader& r, const Ctx& c, Opnd& o, int nd, bool with_mode) {
    o.ptr = tref(r, c);
    o.mode = 0; o.mdim = 0;
    if (with_mode) { o.mode = r.i32(); o.mdim = r.i32(); }
    for (int k = 0; k < nd; ++k) o.stride[k] = r.i64v();
    for (int k = nd; k < AB_MAXD; ++k) o.stride[k] = 0;
}

template <typename T>
static void read_prog(Reader& r, VMProg<T>& P) {
    P.n_instr = r.i32();
    for (int i = 0; i < P.n_instr; ++i) { P.ins[i][0] = (unsigned)r.i32(); P.ins[i][1] = (unsigned)r.i32(); }
    int nc = r.i32();
    for (int i = 0; i < nc; ++i) P.consts[i] = (T)r.f64();
    P.res = r.i32();
}

template <typename T>
static void parse_reduce(Reader& r, const Ctx& c, ReduceParams<T>& p, int& thread_hint) {
    p.mode = r.i32();
    p.out = (T*)tref(r, c);
    p.acc = r.i32();
    p.scale = (T)r.f64();
    p.cadd = (T)r.f64();
    p.nsplit = r.i32();
    thread_hint = r.i32();
    p.m_out = nullptr; p.lo_out = nullptr;
    if (p.mode == R_LSE_EPS || p.mode == R_LSE) {
        if (r.i32()) { p.m_out = (T*)tref(r, c); p.lo_out = (T*)tref(r, c); }
    }
    read_dims(r, p.d, p.n_out, p.n_red);
    p.nf = r.i32();
    for (int f = 0; f < p.nf; ++f) {
        p.coeff[f] = (T)r.f64();
        read_opnd(r, c, p.f[f], p.d.nd, false);
    }
    if (p.mode == R_WSUM) {
        read_opnd(r, c, p.lse_m, p.d.nd, false);
        read_opnd(r, c, p.lse_lo, p.d.nd, false);
        read_opnd(r, c, p.gout, p.d.nd, false);
    }
}

// returns an error message or nullptr
template <typename T>
static const char* parse_xreduce(Reader& r, const Ctx& c, const alan_b200_plan* plan, XReduceParams<T>& x, bool dry) {
    memset(&x, 0, sizeof(x));
    const int site = r.i32();
    x.n_pieces = r.i32();
    for (int q = 0; q < x.n_pieces; ++q) { x.piece[q] = (T*)tref(r, c); x.piece_n[q] = r.i64v(); x.n_total += x.piece_n[q]; }
    if (dry) return nullptr;
    if (plan->comm_world < 2 || !plan->comm_peer[0])
        return "this plan reduces across ranks inside its programs (fused collectives): call "
               "alan_b200_plan_set_comm with the ranks' symmetric buffers first";
    if (site < 0 || site >= (int)plan->site_off.size()) return "xreduce: unknown site";
    x.rank = plan->comm_rank; x.world = plan->comm_world;
    for (int q = 0; q < x.world; ++q) x.site[q] = plan->comm_peer[q] + plan->site_off[site];
    return nullptr;
}

#define AB_PAR_STREAMS 6
template <typename T>
static int run_ops(const alan_b200_plan* plan, int program, const Ctx& c0, bool count_only, int* launches,
                   std::vector<cudaEvent_t>* events = nullptr) {
    Reader r{plan->blob.data() + plan->prog_start[program]};
    int nl = 0;
    Ctx c = c0;                                        // c.stream is switched per op when branches are captured
    // ---- parallel branches (see alan_b200_plan::use_par): only under stream capture, never for profiling / counting
    const int n_ops = plan->prog_nops[program];
    bool par = false;
    if (plan->use_par && !count_only && events == nullptr && !plan->use_seq && n_ops > 2 &&
        plan->blob[plan->prog_start[program]] == OP_DEPS) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(c0.stream, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusActive) par = true;
        else cudaGetLastError();
    }
    std::vector<std::vector<int>> deps;                // per op: earlier ops it must follow
    std::vector<int> op_stream;                        // stream slot of every op issued so far (0 = the caller's stream)
    int last_on[AB_PAR_STREAMS];                       // last op issued on each slot (-1: slot unused so far)
    cudaStream_t slot_stream[AB_PAR_STREAMS];
    cudaEvent_t ev_start = nullptr;
    std::unique_lock<std::mutex> par_lock(plan->mu_par, std::defer_lock);
    if (par) {
        par_lock.lock();                               // the plan's side streams and events serve one capture at a time
        while ((int)plan->par_streams.size() < AB_PAR_STREAMS - 1) {
            cudaStream_t st = nullptr;
            if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); par = false; break; }
            plan->par_streams.push_back(st);
        }
        while (par && (int)plan->par_events.size() < n_ops + AB_PAR_STREAMS + 1) {
            cudaEvent_t ev = nullptr;
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); par = false; break; }
            plan->par_events.push_back(ev);
        }
    }
    if (par) {
        slot_stream[0] = c0.stream;
        for (int k = 1; k < AB_PAR_STREAMS; ++k) slot_stream[k] = plan->par_streams[k - 1];
        for (int k = 0; k < AB_PAR_STREAMS; ++k) last_on[k] = -1;
        op_stream.assign(n_ops, 0);
        ev_start = plan->par_events[n_ops];
        cudaEventRecord(ev_start, c0.stream);
    }
    // event of op i = par_events[i], recorded on the op's stream right after its launches
    auto par_begin = [&](int op_i) {
        if (!par) return;
        const std::vector<int>* d = op_i < (int)deps.size() ? &deps[op_i] : nullptr;
        int slot = -1;
        if (!d) slot = 0;                              // no table entry: chain on the caller's stream
        else {
            // continue the branch of a dependency that is the tip of its stream (the latest such op) ...
            int best = -1;
            for (int dep : *d) if (last_on[op_stream[dep]] == dep && dep > best) best = dep;
            if (best >= 0) slot = op_stream[best];
            // ... or open a branch on an unused slot; failing that, queue behind the caller's stream
            if (slot < 0) for (int k = 0; k < AB_PAR_STREAMS; ++k) if (last_on[k] < 0) { slot = k; break; }
            if (slot < 0) slot = 0;
        }
        if (last_on[slot] < 0 && slot != 0) cudaStreamWaitEvent(slot_stream[slot], ev_start, 0);     // fork
        if (d) for (int dep : *d) if (op_stream[dep] != slot) cudaStreamWaitEvent(slot_stream[slot], plan->par_events[dep], 0);
        if (!d && op_i > 0)                            // unknown dependencies: after everything issued so far
            for (int k = 1; k < AB_PAR_STREAMS; ++k) if (last_on[k] >= 0) cudaStreamWaitEvent(slot_stream[0], plan->par_events[last_on[k]], 0);
        op_stream[op_i] = slot;
        last_on[slot] = op_i;
        c.stream = slot_stream[slot];
    };
    auto par_end = [&](int op_i) {
        if (!par) return;
        cudaEventRecord(plan->par_events[op_i], c.stream);
    };
    auto par_join = [&]() {
        if (!par) return;
        for (int k = 1; k < AB_PAR_STREAMS; ++k)
            if (last_on[k] >= 0) {
                cudaEventRecord(plan->par_events[n_ops + 1 + k], slot_stream[k]);
                cudaStreamWaitEvent(slot_stream[0], plan->par_events[n_ops + 1 + k], 0);
            }
        c.stream = c0.stream;
    };
    // an error return in the middle of a program must still bring the side streams back into the capture
    struct Joiner { decltype(par_join)& j; bool done; ~Joiner() { if (!done) j(); } } joiner{par_join, false};
    // consecutive small ops are collected and run by one launch of small_seq_kernel (kernels.cuh)
    const bool batching = plan->use_seq && events == nullptr;
    SeqParams<T> seq;
    auto flush = [&]() {
        if (seq.n == 0) return;
        if (count_only) ++nl;
        else {
            SeqResidency<T> rz;
            const int smem = plan->seq_resident ? rz.build(seq) : ((seq.n_res = 0), (seq.smem_bytes = (int)((sizeof(SeqOp<T>) + 15) & ~(size_t)15)));
            static const cudaError_t attr = cudaFuncSetAttribute(small_seq_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                                 (int)(AB_RES_TOTAL + 4096));     // once per process and dtype
            (void)attr;
            small_seq_kernel<T><<<1, AB_SEQ_THREADS, smem, c.stream>>>(seq);
        }
        seq.n = 0;
    };
    for (int op_i = 0; op_i < plan->prog_nops[program]; ++op_i) {
        const int32_t* op_begin = r.p;
        if (events) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, c.stream); events->push_back(e); }
        int code = r.i32();
        int nwords = r.i32();
        if (code == OP_DEPS) {
            // table of the program's op dependencies: n, then per op its count and the indices of the earlier ops
            const int n = r.i32();
            deps.assign(n, {});
            for (int k = 0; k < n; ++k) { const int m = r.i32(); for (int j = 0; j < m; ++j) deps[k].push_back(r.i32()); }
            r.p = op_begin + nwords;
            if (par) { op_stream[op_i] = 0; }
            continue;
        }
        par_begin(op_i);
        const bool seqable = code == OP_FILL || code == OP_EXPR || code == OP_EXPR_BWD || code == OP_REDUCE || code == OP_XREDUCE;
        if (!seqable) flush();
        if (count_only && !seqable) {
            if (code == OP_CHAIN || code == OP_CHAIN_BWD) {
                r.p = op_begin + 2; Reader q = r;
                // skip trefs to read the extents
                int ntref = (code == OP_CHAIN) ? 3 : 6;
                q.p += 3 * ntref;
                i64 outer_ = q.i64v(); i64 T_ = q.i64v(); i64 K_ = q.i64v();
                ChainPlan cp;
                if (outer_ <= 65535 && chain_plan<T>(T_, K_, cp)) nl += cp.n_phases;          // segment phases (chain.cuh)
                else { int lv = 0; for (i64 n = T_; n > 1; n = n / 2 + n % 2) lv++; nl += lv + 1; }
            } else if (code != OP_FILL && code != OP_COPY) nl += 1;
            r.p = op_begin + nwords;
            continue;
        }
        switch (code) {
            case OP_FILL: {
                void* dst = tref(r, c);
                i64 nbytes = r.i64v();
                if (batching && nbytes <= 4 * plan->seq_points && nbytes % 4 == 0) {
                    if (seq.n == AB_SEQ_MAX) flush();
                    SeqOp<T>& o = seq.op[seq.n++];
                    o.kind = SK_FILL; o.warp = 0; o.f.ptr = dst; o.f.nbytes = nbytes;
                    break;
                }
                flush();
                if (count_only) break;               // memsets are not counted as kernel launches
                cudaMemsetAsync(dst, 0, (size_t)nbytes, c.stream);
                break;
            }
            case OP_COPY: {
                void* dst = tref(r, c);
                void* src = tref(r, c);
                i64 nbytes = r.i64v();
                cudaMemcpyAsync(dst, src, (size_t)nbytes, cudaMemcpyDeviceToDevice, c.stream);
                break;
            }
            case OP_EXPR: {
                ExprParams<T> p;
                p.out = (T*)tref(r, c);
                p.acc = r.i32();
                p.scale = (T)r.f64();
                read_dims(r, p.d, p.n_out, p.n_red);
                p.n_leaves = r.i32();
                for (int l = 0; l < p.n_leaves; ++l) read_opnd(r, c, p.leaf[l], p.d.nd, true);
                read_prog(r, p.prog);
                if (batching && p.n_out * p.n_red <= plan->seq_points) {
                    if (seq.n == AB_SEQ_MAX) flush();
                    SeqOp<T>& o = seq.op[seq.n++];
                    o.kind = SK_EXPR; o.warp = (p.n_red >= 8) ? 1 : 0; o.n3 = detect_normal3(p.prog); o.e = p;
                    break;
                }
                flush();
                if (count_only) { ++nl; break; }
                launch_expr_fwd<T>(p, c.stream, c.sm_count);
                break;
            }
            case OP_EXPR_BWD: {
                ExprBwdParams<T> p;
                p.gleaf = (T*)tref(r, c);
                p.acc = r.i32();
                p.scale = (T)r.f64();
                p.target = r.i32();
                p.nsplit = r.i32();
                read_dims(r, p.d, p.n_kept, p.n_loop);
                read_opnd(r, c, p.gout, p.d.nd, false);
                p.n_leaves = r.i32();
                for (int l = 0; l < p.n_leaves; ++l) read_opnd(r, c, p.leaf[l], p.d.nd, true);
                read_prog(r, p.prog);
                if (batching && p.n_kept * p.n_loop <= plan->seq_points) {
                    if (seq.n == AB_SEQ_MAX) flush();
                    SeqOp<T>& o = seq.op[seq.n++];
                    o.kind = SK_EXPR_BWD; o.warp = (p.n_loop / p.nsplit >= 8) ? 1 : 0; o.n3 = detect_normal3(p.prog); o.b = p;
                    break;
                }
                flush();
                if (count_only) { ++nl; break; }
                launch_expr_bwd<T>(p, c.stream, c.sm_count);
                break;
            }
            case OP_RSEQ: {
                // a run of small reductions (and at most one cross-rank sum) as one single-CTA launch (kernels.cuh)
                const int n = r.i32();
                if (n < 1 || n > AB_RSEQ_MAX) return fail("reduction sequence: bad op count");
                RSeqParams<T> sp;
                memset(&sp, 0, sizeof(sp));
                sp.n = n;
                std::vector<ReduceParams<T>> full(n);
                std::vector<int> hint(n, 0);
                bool has_x = false, compact = true;
                for (int k = 0; k < n; ++k) {
                    const int32_t* sub = r.p;
                    const int sub_code = r.i32();
                    const int sub_words = r.i32();
                    if (sub_code == OP_REDUCE) {
                        parse_reduce<T>(r, c, full[k], hint[k]);
                        compact = rc_from(full[k], reduce_uses_warps(full[k], hint[k] != 0), sp.op[k]) && compact;
                    } else if (sub_code == OP_XREDUCE) {
                        if (has_x) return fail("reduction sequence: more than one cross-rank reduction");
                        has_x = true;
                        const char* err = parse_xreduce<T>(r, c, plan, sp.x, count_only);
                        if (err) return fail(err);
                        sp.op[k].kind = RS_XREDUCE;
                    } else return fail("reduction sequence: unsupported member op");
                    r.p = sub + sub_words;
                }
                if (compact) { reduce_seq_kernel<T><<<1, AB_RSEQ_THREADS, 0, c.stream>>>(sp); break; }
                for (int k = 0; k < n; ++k) {                     // a member exceeds the compact form: one launch each
                    if (sp.op[k].kind == RS_XREDUCE) xreduce_kernel<T><<<1, 512, 0, c.stream>>>(sp.x);
                    else launch_reduce<T>(full[k], hint[k] != 0, c.stream, c.sm_count);
                }
                break;
            }
            case OP_REDUCE: {
                ReduceParams<T> p;
                int thread_hint = 0;
                parse_reduce<T>(r, c, p, thread_hint);
                // small: few points, or a plain fixed-order sum of partial rows with few outputs (one thread per output
                    // walks the rows: e.g. the 160 x 900 partial rows of the fused plate sum)
                if (batching && (p.n_out * p.n_red <= plan->seq_points ||
                                 (plan->seq_bigsum && p.mode == R_SUM && p.nsplit == 1 && p.n_out <= 1024 && p.n_out * p.n_red <= 64 * AB_SEQ_POINTS))) {
                    if (seq.n == AB_SEQ_MAX) flush();
                    SeqOp<T>& o = seq.op[seq.n++];
                    o.kind = SK_REDUCE; o.warp = reduce_uses_warps(p, thread_hint != 0) ? 1 : 0; o.r = p;
                    break;
                }
                flush();
                if (count_only) { ++nl; break; }
                launch_reduce<T>(p, thread_hint != 0, c.stream, c.sm_count);
                break;
            }
            case OP_XREDUCE: {
                XReduceParams<T> x;
                {
                    const char* err = parse_xreduce<T>(r, c, plan, x, count_only);
                    if (err) return fail(err);
                }
                if (count_only) { if (!batching) ++nl; else { if (seq.n == AB_SEQ_MAX) flush(); seq.n++; } break; }
                if (!batching) { xreduce_kernel<T><<<1, 512, 0, c.stream>>>(x); break; }      // stand-alone: a 0.4 KB parameter block
                if (seq.n == AB_SEQ_MAX) flush();
                SeqOp<T>& o = seq.op[seq.n++];
                o.kind = SK_XREDUCE; o.warp = 0; o.x = x;
                break;
            }
            case OP_CHAIN: {
                const T* ms = (const T*)tref(r, c);
                T* levels = (T*)tref(r, c);
                T* out = (T*)tref(r, c);
                i64 outer = r.i64v(), Tn = r.i64v(), K = r.i64v();
                if (launch_chain_fwd<T>(ms, levels, out, outer, Tn, K, c.stream) == 0) break;     // segment phases (chain.cuh)
                size_t smem = (size_t)(2 * K * K + 2 * K) * sizeof(T);
                if (smem > 200 * 1024) return fail("chain: K x K does not fit shared memory");
                if (outer > 65535) return fail("chain: more than 65535 independent chains");
                if (smem > 48 * 1024)
                    cudaFuncSetAttribute(chain_level_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                const T* X = ms;
                T* Y = levels;
                i64 n = Tn;
                while (n > 1) {
                    i64 no = n / 2 + n % 2;
                    chain_level_kernel<T><<<dim3((unsigned)no, (unsigned)outer), 256, smem, c.stream>>>(X, Y, (int)n, (int)no, (int)K);
                    X = Y;
                    Y += outer * no * K * K;
                    n = no;
                }
                chain_final_kernel<T><<<grid_for(outer * K, 128, c), 128, 0, c.stream>>>(X, out, outer * K, (int)K);
                break;
            }
            case OP_CHAIN_BWD: {
                const T* ms = (const T*)tref(r, c);
                const T* levels = (const T*)tref(r, c);
                const T* out = (const T*)tref(r, c);
                const T* gout = (const T*)tref(r, c);
                T* glevels = (T*)tref(r, c);
                T* gms = (T*)tref(r, c);
                i64 outer = r.i64v(), Tn = r.i64v(), K = r.i64v();
                if (launch_chain_bwd<T>(ms, levels, out, gout, glevels, gms, outer, Tn, K, c.stream) == 0) break;
                size_t smem = (size_t)(3 * K * K + 4 * K) * sizeof(T);
                if (smem > 200 * 1024) return fail("chain adjoint: K x K does not fit shared memory");
                if (outer > 65535) return fail("chain adjoint: more than 65535 independent chains");
                if (smem > 48 * 1024)
                    cudaFuncSetAttribute(chain_level_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                // level table
                std::vector<i64> ns; std::vector<i64> offs;
                { i64 n = Tn, off = 0; while (n > 1) { i64 no = n / 2 + n % 2; ns.push_back(n); offs.push_back(off); off += outer * no * K * K; n = no; } }
                int L = (int)ns.size();
                // final: X_last is levels[offs[L-1]] (or ms when T == 1)
                const T* Xlast = L ? levels + offs[L - 1] : ms;
                T* gXlast = L ? glevels + offs[L - 1] : gms;
                chain_final_bwd_kernel<T><<<grid_for(outer * K * K, 256, c), 256, 0, c.stream>>>(Xlast, out, gout, gXlast, outer * K, (int)K);
                for (int l = L - 1; l >= 0; --l) {
                    i64 n = ns[l], no = n / 2 + n % 2;
                    const T* X = l ? levels + offs[l - 1] : ms;
                    T* gX = l ? glevels + offs[l - 1] : gms;
                    const T* gY = glevels + offs[l];
                    chain_level_bwd_kernel<T><<<dim3((unsigned)no, (unsigned)outer), 256, smem, c.stream>>>(X, gY, gX, (int)n, (int)no, (int)K);
                }
                break;
            }
            case OP_SAMPLE: {
                SampleParams<T> p;
                memset(&p, 0, sizeof(p));
                int nb = r.i32();
                p.d.nd = nb; p.d.n_a = nb;
                p.n_batch = 1;
                for (int k = 0; k < nb; ++k) { p.d.size[k] = r.i32(); p.n_batch *= p.d.size[k]; }
                p.nk = r.i32();
                p.ktotal = 1;
                for (int k = 0; k < p.nk; ++k) { p.ksize[k] = r.i32(); p.ktotal *= p.ksize[k]; }
                p.nf = r.i32();
                for (int f = 0; f < p.nf; ++f) {
                    p.coeff[f] = (T)r.f64();
                    read_opnd(r, c, p.f[f], nb, false);
                    for (int k = 0; k < p.nk; ++k) p.kstride[f][k] = r.i64v();
                    p.ng[f] = r.i32();
                    for (int g = 0; g < p.ng[f]; ++g) { p.gstride[f][g] = r.i64v(); p.gsel[f][g] = r.i32(); }
                }
                p.n_idx = r.i32();
                for (int t = 0; t < p.n_idx; ++t) {
                    p.idxptr[t] = (const i64*)tref(r, c);
                    for (int k = 0; k < nb; ++k) p.idxstride[t][k] = r.i64v();
                }
                p.u = (const double*)tref(r, c);
                for (int k = 0; k < nb; ++k) p.ustride[k] = r.i64v();
                for (int k = 0; k < p.nk; ++k) p.out[k] = (i64*)tref(r, c);
                sample_kernel<T><<<grid_for(p.n_batch, 128, c), 128, 0, c.stream>>>(p);
                break;
            }
            case OP_NORMAL_FAN: {
                FanParams<T> p;
                memset(&p, 0, sizeof(p));
                p.out = (T*)tref(r, c);
                int D = r.i32();
                int nrd = r.i32();
                p.rd.nd = nrd; p.rd.n_a = nrd;
                p.n_rows = 1;
                for (int k = 0; k < nrd; ++k) { p.rd.size[k] = r.i32(); p.n_rows *= p.rd.size[k]; }
                for (int k = 0; k < nrd; ++k) p.vstride[k] = r.i64v();
                for (int k = 0; k < nrd; ++k) p.lstride[k] = r.i64v();
                for (int k = 0; k < nrd; ++k) p.ostride[k] = r.i64v();
                p.v = (const T*)tref(r, c); p.v_ev = r.i64v();
                p.l = (const T*)tref(r, c); p.l_ev = r.i64v();
                p.s = (const T*)tref(r, c); p.s_f = r.i64v(); p.s_ev = r.i64v();
                p.F = r.i32();
                p.o_f = r.i64v();
                if (launch_fan<T>(p, D, c.stream, c.sm_count)) return fail("normal_fan: unsupported event extent");
                break;
            }
            case OP_FAN_LSE: {
                FanLseParams<T> p;
                memset(&p, 0, sizeof(p));
                int bwd = r.i32();
                T* o = (T*)tref(r, c);
                if (bwd) { p.lse = o; p.gout = (const T*)tref(r, c); p.gS = (T*)tref(r, c); } else p.out = o;
                int D = r.i32();
                int nrd = r.i32();
                p.rd.nd = nrd; p.rd.n_a = nrd;
                p.n_rho = 1;
                for (int k = 0; k < nrd; ++k) { p.rd.size[k] = r.i32(); p.n_rho *= p.rd.size[k]; }
                for (int k = 0; k < nrd; ++k) p.vstride[k] = r.i64v();
                for (int k = 0; k < nrd; ++k) p.lstride[k] = r.i64v();
                for (int k = 0; k < nrd; ++k) p.ostride[k] = r.i64v();
                p.Kk = r.i32(); p.v_k = r.i64v(); p.l_k = r.i64v();
                p.v = (const T*)tref(r, c); p.v_ev = r.i64v();
                p.l = (const T*)tref(r, c); p.l_ev = r.i64v();
                p.s = (const T*)tref(r, c); p.s_f = r.i64v(); p.s_ev = r.i64v();
                p.F = r.i32(); p.o_f = r.i64v();
                p.nb = r.i32();
                for (int i = 0; i < p.nb; ++i) {
                    p.bcoeff[i] = (T)r.f64();
                    p.b[i] = (const T*)tref(r, c);
                    for (int k = 0; k < nrd; ++k) p.bstride[i][k] = r.i64v();
                    p.b_k[i] = r.i64v();
                }
                p.cadd = (T)r.f64();
                p.qn = r.i32();
                if (p.qn) {
                    p.q_l = (const T*)tref(r, c); p.q_s = (const T*)tref(r, c);
                    for (int k = 0; k < nrd; ++k) p.q_lstride[k] = r.i64v();
                    for (int k = 0; k < nrd; ++k) p.q_sstride[k] = r.i64v();
                    p.q_lev = r.i64v(); p.q_sev = r.i64v();
                    p.q_coeff = (T)r.f64();
                }
                if (!bwd) {
                    p.psum_rows = r.i32();
                    if (p.psum_rows > 0) { p.psum = (T*)tref(r, c); p.ps_lam = r.i64v(); p.ps_f = r.i64v(); p.ps_row = r.i64v(); }
                }
                if (bwd) {
                    for (int k = 0; k < nrd; ++k) p.gstride[k] = r.i64v();
                    p.g_f = r.i64v();
                    p.gs_compact = r.i32();
                }
                int rc = -1;
                if constexpr (std::is_same<T, float>::value) {
                    if (plan->use_tc && plan->use_tc2 && tc::fan_lse_tc2_supported(p, D, bwd != 0))
                        rc = tc::launch_fan_lse_tc2(p, D, bwd != 0, c.stream, c.sm_count);
                    else if (p.psum_rows > 0)
                        return fail("fan_lse: the plan fuses the plate sum into the dense tensor-core kernel but that kernel is "
                                    "disabled or does not cover this shape; rebuild the plan with ALAN_B200_NO_TC / "
                                    "ALAN_B200_TC_BLOCKDIAG set the way the run is");
                    else if (p.gs_compact > 0)
                        return fail("fan_lse adjoint: the plan commits to the dense tensor-core kernel (compact gS layout) but that "
                                    "kernel is disabled or does not cover this shape; rebuild the plan with ALAN_B200_NO_TC / "
                                    "ALAN_B200_TC_BLOCKDIAG set the way the run is");
                    else if (p.qn)
                        return fail("fan_lse: the plan evaluates the Q factor inside the dense tensor-core kernel but that kernel is "
                                    "disabled or does not cover this shape; rebuild the plan with ALAN_B200_NO_TC / "
                                    "ALAN_B200_TC_BLOCKDIAG set the way the run is");
                    else if (plan->use_tc && tc::fan_lse_tc_supported(p, D)) rc = tc::launch_fan_lse_tc(p, D, bwd != 0, c.stream, c.sm_count);
                }
                if (rc < 0 && p.qn) return fail("fan_lse: the inline Q factor needs the fp32 dense tensor-core kernel");
                if (rc < 0 && (p.gs_compact > 0 || p.psum_rows > 0)) return fail("fan_lse: compact gS layout / fused plate sum need the fp32 dense tensor-core kernel");
                if (rc < 0) rc = launch_fan_lse<T>(p, D, bwd != 0, c.stream, c.sm_count);
                if (rc) return fail(rc == 1 ? "fan_lse: unsupported event extent" : rc == 2 ? "fan_lse: tile does not fit shared memory" : "fan_lse: strides exceed 32-bit tile addressing");
                break;
            }
            case OP_DOT: {
                DotParams<T> p;
                p.out = (T*)tref(r, c);
                read_dims(r, p.d, p.n_out, p.n_red);
                if (p.d.nd - p.d.n_a > 1) return fail("dot: more than one reduced dim");
                read_opnd(r, c, p.a, p.d.nd, false);
                read_opnd(r, c, p.b, p.d.nd, false);
                if (p.d.nd == p.d.n_a) { p.a.stride[p.d.n_a] = 0; p.b.stride[p.d.n_a] = 0; }
                dot_kernel<T><<<grid_for(p.n_out, 256, c), 256, 0, c.stream>>>(p);
                break;
            }
            case OP_FAN_BWD: {
                FanBwdParams<T> q;
                memset(&q, 0, sizeof(q));
                FanParams<T>& p = q.f;
                int which = r.i32();
                p.out = (T*)tref(r, c);                      // G: adjoint of the factor
                q.R = (T*)tref(r, c);
                q.partial = (T*)tref(r, c);
                q.partial_w = (T*)tref(r, c);
                q.n_cta = r.i32();
                int D = r.i32();
                int nrd = r.i32();
                p.rd.nd = nrd; p.rd.n_a = nrd;
                p.n_rows = 1;
                for (int k = 0; k < nrd; ++k) { p.rd.size[k] = r.i32(); p.n_rows *= p.rd.size[k]; }
                for (int k = 0; k < nrd; ++k) p.vstride[k] = r.i64v();
                for (int k = 0; k < nrd; ++k) p.lstride[k] = r.i64v();
                for (int k = 0; k < nrd; ++k) p.ostride[k] = r.i64v();
                p.v = (const T*)tref(r, c); p.v_ev = r.i64v();
                p.l = (const T*)tref(r, c); p.l_ev = r.i64v();
                p.s = (const T*)tref(r, c); p.s_f = r.i64v(); p.s_ev = r.i64v();
                p.F = r.i32();
                p.o_f = r.i64v();
                if (launch_fan_bwd<T>(q, D, which, c.stream, c.sm_count)) return fail("fan_bwd: unsupported event extent");
                break;
            }
            case OP_NORMAL_POLY_SUM: {
                NormalPolyParams<T> p;
                memset(&p, 0, sizeof(p));
                p.out = (T*)tref(r, c);
                p.cadd = (T)r.f64();
                p.n_row = r.i32(); p.n_k = r.i32(); p.n_z = r.i32();
                p.d.nd = p.n_row + p.n_k + p.n_z; p.d.n_a = p.n_row + p.n_k;
                if (p.d.nd > AB_MAXD) return fail("normal_poly_sum: too many dims");
                p.rows = p.ks = p.zs = 1;
                for (int k = 0; k < p.d.nd; ++k) {
                    p.d.size[k] = r.i32();
                    if (k < p.n_row) p.rows *= p.d.size[k]; else if (k < p.d.n_a) p.ks *= p.d.size[k]; else p.zs *= p.d.size[k];
                }
                for (int k = 0; k < p.d.n_a; ++k) p.ostride[k] = r.i64v();
                p.n_zleaf = r.i32();
                if (p.n_zleaf > NP_MAXLEAF) return fail("normal_poly_sum: too many leaves");
                for (int l = 0; l < p.n_zleaf; ++l) read_opnd(r, c, p.zleaf[l], p.d.nd, false);
                p.n_kleaf = r.i32();
                if (p.n_kleaf > NP_MAXLEAF) return fail("normal_poly_sum: too many leaves");
                for (int l = 0; l < p.n_kleaf; ++l) read_opnd(r, c, p.kleaf[l], p.d.nd, false);
                p.n_zt = r.i32();
                if (p.n_zt > NP_MAXTERM) return fail("normal_poly_sum: too many terms");
                for (int t = 0; t < p.n_zt; ++t) { p.zt[t].coeff = r.f64(); p.zt[t].z[0] = r.i32(); p.zt[t].z[1] = r.i32(); p.zt[t].k[0] = r.i32(); p.zt[t].k[1] = r.i32(); }
                p.n_kt = r.i32();
                if (p.n_kt > NP_MAXTERM) return fail("normal_poly_sum: too many terms");
                for (int t = 0; t < p.n_kt; ++t) { p.kt[t].coeff = r.f64(); p.kt[t].z[0] = r.i32(); p.kt[t].z[1] = r.i32(); p.kt[t].k[0] = r.i32(); p.kt[t].k[1] = r.i32(); }
                p.scale_leaf = r.i32(); p.scale_const = r.f64();
                if (launch_normal_poly_sum<T>(p, c.stream, c.sm_count)) return fail("normal_poly_sum: unsupported shape");
                break;
            }
            case OP_PASTE: {
                const T* src = (const T*)tref(r, c);
                T* dst = (T*)tref(r, c);
                PasteParams p;
                memset(&p, 0, sizeof(p));
                p.nd = r.i32();
                if (p.nd > AB_MAXD) return fail("paste: too many dims");
                p.total = 1;
                for (int k = 0; k < p.nd; ++k) { p.size[k] = r.i32(); p.ss[k] = r.i64v(); p.ds[k] = r.i64v(); p.total *= p.size[k]; }
                paste_kernel<T><<<grid_for(p.total, 256, c), 256, 0, c.stream>>>(src, dst, p);
                break;
            }
            case OP_MVN_PREP: {
                const T* S = (const T*)tref(r, c);
                T* L = (T*)tref(r, c);
                T* W = (T*)tref(r, c);
                T* cst = (T*)tref(r, c);
                const i64 n_mat = r.i64v();
                const int d = r.i32(), mode = r.i32();
                const int rank = r.i32();
                const T* Dg = rank > 0 ? (const T*)tref(r, c) : nullptr;
                if (d < 1 || d > AB_MVN_MAXD) return fail("MultivariateNormal: event size must be between 1 and 64");
                const size_t smem = 2 * (size_t)d * d * sizeof(T);
                static const cudaError_t attr = cudaFuncSetAttribute(mvn_prep_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                                     (int)(2 * AB_MVN_MAXD * AB_MVN_MAXD * sizeof(T)));
                (void)attr;
                mvn_prep_kernel<T><<<grid_for(n_mat, 1, c), 32, smem, c.stream>>>(S, L, W, cst, n_mat, d, mode, Dg, rank);
                break;
            }
            case OP_PERM: {
                const double* u = (const double*)tref(r, c);
                i64* out = (i64*)tref(r, c);
                const i64 rows = r.i64v();
                const int K = r.i32(), mode = r.i32();
                if ((size_t)K * 8 * sizeof(double) > 48 * 1024) return fail("perm: K too large");
                perm_kernel<<<grid_for(rows, 8, c), 256, (size_t)K * 8 * sizeof(double), c.stream>>>(u, out, rows, K, mode);
                break;
            }
            case OP_KGATHER: {
                const T* x = (const T*)tref(r, c);
                const i64* perm = (const i64*)tref(r, c);
                T* out = (T*)tref(r, c);
                const i64 outer = r.i64v(), K = r.i64v(), inner = r.i64v();
                kgather_kernel<T><<<grid_for(outer * K * inner, 256, c), 256, 0, c.stream>>>(x, perm, out, outer, K, inner);
                break;
            }
            case OP_TS_SAMPLE: {
                TsSampleParams<T> p;
                memset(&p, 0, sizeof(p));
                p.e.out = (T*)tref(r, c);
                p.e.acc = 0; p.e.scale = T(1);
                read_dims(r, p.e.d, p.e.n_out, p.e.n_red);
                p.e.n_leaves = r.i32();
                for (int l = 0; l < p.e.n_leaves; ++l) read_opnd(r, c, p.e.leaf[l], p.e.d.nd, true);
                read_prog(r, p.e.prog);
                p.prev_leaf = r.i32(); p.t_dim = r.i32(); p.k_dim = r.i32();
                p.init = (const T*)tref(r, c);
                if (r.i32()) p.perm = (const i64*)tref(r, c);
                p.n_outer = r.i64v(); p.Tn = r.i32(); p.Kn = r.i32(); p.En = r.i32();
                const size_t smem = (size_t)2 * p.Kn * p.En * sizeof(T);
                if (smem > 200 * 1024) return fail("ts_sample: K x event extent does not fit shared memory");
                if (smem > 48 * 1024)
                    cudaFuncSetAttribute(ts_sample_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                int threads = p.Kn * p.En; threads = (threads + 31) / 32 * 32; if (threads > 1024) threads = 1024;
                i64 blocks = p.n_outer < (i64)c.sm_count * 4 ? p.n_outer : (i64)c.sm_count * 4;
                ts_sample_kernel<T><<<(int)(blocks < 1 ? 1 : blocks), threads, smem, c.stream>>>(p);
                break;
            }
            case OP_NORMAL_Q_BWD: {
                NormalQBwdParams<T> p;
                memset(&p, 0, sizeof(p));
                const int D = r.i32();
                const int nu = r.i32();
                p.ud.nd = nu; p.ud.n_a = nu;
                p.n_users = 1;
                for (int k = 0; k < nu; ++k) { p.ud.size[k] = r.i32(); p.n_users *= p.ud.size[k]; }
                for (int k = 0; k < nu; ++k) p.vstride[k] = r.i64v();
                for (int k = 0; k < nu; ++k) p.lstride[k] = r.i64v();
                for (int k = 0; k < nu; ++k) p.sstride[k] = r.i64v();
                for (int k = 0; k < nu; ++k) p.gstride[k] = r.i64v();
                p.Kk = r.i32(); p.v_k = r.i64v(); p.v_ev = r.i64v(); p.l_ev = r.i64v(); p.s_ev = r.i64v();
                p.v = (const T*)tref(r, c); p.l = (const T*)tref(r, c); p.s = (const T*)tref(r, c);
                p.scale_is_exp = r.i32();
                p.G = (const T*)tref(r, c); p.g_s = r.i64v(); p.g_k = r.i64v(); p.S = r.i32();
                p.coeff = (T)r.f64();
                if (r.i32()) { p.gl = (T*)tref(r, c); p.acc_l = r.i32(); }
                if (r.i32()) { p.gs = (T*)tref(r, c); p.acc_s = r.i32(); }
                int rc = launch_normal_q_bwd<T>(p, D, c.stream, c.sm_count);
                if (rc) return fail(rc == 1 ? "normal_q_bwd: unsupported event extent" : "normal_q_bwd: tile does not fit shared memory");
                break;
            }
            case OP_BERN_DOT: {
                BernDotParams<T> p;
                memset(&p, 0, sizeof(p));
                p.out = (T*)tref(r, c);
                p.cadd = (T)r.f64();
                int D = r.i32();
                read_dims(r, p.d, p.n_out, p.n_red);
                read_opnd(r, c, p.a, p.d.nd, false); p.a_ev = r.i64v();
                read_opnd(r, c, p.b, p.d.nd, false); p.b_ev = r.i64v();
                read_opnd(r, c, p.y, p.d.nd, false);
                p.side = r.i32();
                if (p.side) {
                    p.qout = (T*)tref(r, c);
                    read_opnd(r, c, p.ql, p.d.nd, false); p.ql_ev = r.i64v();
                    read_opnd(r, c, p.qs, p.d.nd, false); p.qs_ev = r.i64v();
                }
                if (launch_bern_dot<T>(p, D, c.stream, c.sm_count)) return fail("bern_dot_sum: unsupported event extent");
                break;
            }
            default:
                return fail("unknown opcode " + std::to_string(code));
        }
        if (r.p != op_begin + nwords)
            return fail("blob/executor mismatch in op " + std::to_string(op_i) + " (code " + std::to_string(code) +
                        "): consumed " + std::to_string((long)(r.p - op_begin)) + " of " + std::to_string(nwords) + " words");
        par_end(op_i);
    }
    flush();
    par_join();
    joiner.done = true;
    c0.max_out = c.max_out > c0.max_out ? c.max_out : c0.max_out;
    c0.max_aux = c.max_aux > c0.max_aux ? c.max_aux : c0.max_aux;
    if (events) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, c.stream); events->push_back(e); }
    if (launches) *launches = nl;
    if (!count_only) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(std::string("CUDA launch error: ") + cudaGetErrorString(e));
    }
    return 0;
}

// Pipe-peak micro-benchmarks (roofline denominators measured in the same process as the bench): 8 independent
// chains per thread, 2 x 1024-thread CTAs per SM, best of 5 by CUDA events.  which = 0: MUFU.EX2, 1: FFMA.
template <int MODE>
__global__ void __launch_bounds__(1024) pipe_peak_kernel(float* out, float seed, int iters) {
    float a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const float m = seed * 0.5f, c = 0.001f;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
        } else {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" {

int alan_b200_abi_version(void) { return AB_VERSION; }

int alan_b200_pipe_peak(int which, void* scratch, size_t scratch_bytes, double* ops_per_s, void* stream) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail("pipe_peak: no CUDA device");
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 2, block = 1024, iters = 2048;
    if (scratch_bytes < (size_t)grid * block * sizeof(float)) return fail("pipe_peak: scratch too small");
    if (which != 0 && which != 1) return fail("pipe_peak: which must be 0 (MUFU.EX2) or 1 (FFMA)");
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0, st);
        if (which == 0) pipe_peak_kernel<0><<<grid, block, 0, st>>>((float*)scratch, 1.0f + 1e-4f * r, iters);
        else pipe_peak_kernel<1><<<grid, block, 0, st>>>((float*)scratch, 1.0f + 1e-4f * r, iters);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("pipe_peak: ") + cudaGetErrorString(e));
    *ops_per_s = 8.0 * iters * (double)grid * block / (best * 1e-3);
    return 0;
}
const char* alan_b200_last_error(void) { return g_err.c_str(); }

int alan_b200_plan_create(const int32_t* blob, size_t n_words, alan_b200_plan** out) {
    if (!blob || n_words < 10) return fail("plan blob too short");
    if (blob[0] != AB_MAGIC) return fail("plan blob: bad magic");
    if (blob[1] != AB_VERSION) return fail("plan blob: version mismatch");
    alan_b200_plan* p = new alan_b200_plan();
    p->blob.assign(blob, blob + n_words);
    p->dtype = blob[2];
    p->n_inputs = blob[3];
    p->n_programs = blob[4];
    p->ws_bytes = (size_t)(((uint64_t)(uint32_t)blob[6] << 32) | (uint32_t)blob[5]);
    p->n_fwd = blob[7];
    p->n_bwd = blob[8];
    p->sample_prog = blob[9];
    if (p->dtype != 0 && p->dtype != 1) { delete p; return fail("plan blob: dtype must be 0 (f32) or 1 (f64)"); }
    if (n_words < (size_t)(10 + 2 * p->n_programs)) { delete p; return fail("plan blob truncated"); }
    for (int i = 0; i < p->n_programs; ++i) {
        p->prog_start.push_back(blob[10 + 2 * i]);
        p->prog_nops.push_back(blob[10 + 2 * i + 1]);
        if ((size_t)p->prog_start.back() > n_words) { delete p; return fail("plan blob: program offset out of range"); }
    }
    // cross-rank reduction sites: header + two data buffers each, 256-byte aligned
    for (int i = 0; i < p->n_programs; ++i) {
        const int32_t* q = blob + p->prog_start[i];
        for (int k = 0; k < p->prog_nops[i]; ++k) {
            if ((size_t)(q - blob) + 2 > n_words || q[1] < 2) { delete p; return fail("plan blob: malformed op table"); }
            if (q[0] == OP_XREDUCE) {
                const int site = q[2], np = q[3];
                size_t n = 0;
                for (int j = 0; j < np; ++j) {
                    const int32_t* w = q + 4 + 5 * j + 3;                       // tref = 3 words, then the 64-bit count
                    n += (size_t)(((uint64_t)(uint32_t)w[1] << 32) | (uint32_t)w[0]);
                }
                if (site >= (int)p->site_elems.size()) p->site_elems.resize(site + 1, 0);
                if (n > p->site_elems[site]) p->site_elems[site] = n;
            }
            q += q[1];
        }
    }
    {
        const size_t item = p->dtype == 0 ? 4 : 8;
        size_t off = 0;
        for (size_t sgl = 0; sgl < p->site_elems.size(); ++sgl) {
            p->site_off.push_back(off);
            off += AB_XR_HDR + ((2 * p->site_elems[sgl] * item + 255) / 256) * 256;
        }
        p->comm_bytes = off;
    }
    p->graphs.resize(p->n_programs);
    p->n_out.assign(p->n_programs, -1);
    p->n_aux.assign(p->n_programs, -1);
    // CUDA-graph replay of a program.  Measured on B200: a replay costs ~20 us of launch latency per program (two
    // programs per step), and saves the 2-4 us gap between each pair of dependent launches: cfg-2 (300 users, 33
    // launches of a few us) 0.26 -> 0.30 ms per step, cfg-5 (10 000 users) 0.75 -> 0.61 ms.  Default = auto: plans
    // with a large workspace (>= 32 MB: their kernels run for tens of microseconds) capture a program on the first
    // call with a binding of pointers and replay it when the binding returns; a caller that passes fresh buffers on
    // every call stops paying for captures after four in a row without a replay.  ALAN_B200_GRAPH=1 forces replay from the first call, ALAN_B200_GRAPH=0 turns it off.
    {
        const char* g = getenv("ALAN_B200_GRAPH");
        const bool big = p->ws_bytes >= ((size_t)32 << 20);
        p->use_graphs = g ? (g[0] != '0') : big;
        p->graph_auto = (g == nullptr) && big;
    }
    p->miss_streak.assign(p->n_programs, 0);
    p->use_tc = getenv("ALAN_B200_NO_TC") == nullptr;
    p->use_tc2 = getenv("ALAN_B200_TC_BLOCKDIAG") == nullptr;
    // Consecutive small ops of a program can run as ONE single-CTA launch (small_seq_kernel, its small tensors
    // resident in shared memory).  Measured on B200 (cfg-2 / cfg-5 step, whole step replayed as one CUDA graph):
    // separate launches 0.143 / 0.495 ms, sequences 0.205 / 0.562 ms, resident sequences 0.219 / 0.577 ms -- inside
    // a graph a tiny kernel costs ~1.7 us all in, less than the same op costs inside the one-CTA interpreter.
    // Hence opt-in (ALAN_B200_SEQ=1); cross-rank reductions always run through it (they are single-CTA by nature).
    { const char* sq = getenv("ALAN_B200_SEQ"); p->use_seq = (sq && sq[0] == '1'); }
    { const char* sq = getenv("ALAN_B200_PAR"); p->use_par = !(sq && sq[0] == '0'); }
    { const char* sq = getenv("ALAN_B200_SEQ_POINTS"); if (sq) p->seq_points = atoll(sq); }
    { const char* sq = getenv("ALAN_B200_SEQ_BIGSUM"); p->seq_bigsum = (sq && sq[0] == '1'); }
    { const char* sq = getenv("ALAN_B200_SEQ_RESIDENT"); p->seq_resident = !(sq && sq[0] == '0'); }
    int dev = 0;
    p->sm_count = 148;
    if (const char* e = getenv("ALAN_B200_WAIT_HINT_NS")) {          // tuning aid: mbarrier.try_wait suspend-time hint (fan_tc.cuh)
        const unsigned v = (unsigned)atoi(e);
        cudaMemcpyToSymbol(tc::g_wait_hint_ns, &v, sizeof(v));
    }
    if (cudaGetDevice(&dev) == cudaSuccess) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) p->sm_count = n;
    } else {
        cudaGetLastError();
    }
    *out = p;
    return 0;
}

void alan_b200_plan_destroy(alan_b200_plan* p) { delete p; }

size_t alan_b200_comm_bytes(const alan_b200_plan* p) { return p->comm_bytes; }

int alan_b200_plan_set_comm(alan_b200_plan* p, int rank, int world, void* const* peer_buffers, size_t bytes) {
    if (!p) return fail("null plan");
    if (world < 1 || world > AB_XR_MAXW) return fail("set_comm: between 1 and 8 ranks");
    if (rank < 0 || rank >= world) return fail("set_comm: rank out of range");
    if (bytes < p->comm_bytes) return fail("set_comm: symmetric buffers are smaller than alan_b200_comm_bytes(plan)");
    std::lock_guard<std::mutex> g(p->mu);
    p->comm_rank = rank; p->comm_world = world;
    for (int r = 0; r < world; ++r) {
        if (!peer_buffers[r]) return fail("set_comm: null peer buffer");
        p->comm_peer[r] = (char*)peer_buffers[r];
    }
    for (auto& v : p->graphs) { for (auto& ge : v) cudaGraphExecDestroy(ge.exec); v.clear(); }     // captured pointers are stale
    return 0;
}
size_t alan_b200_workspace_bytes(const alan_b200_plan* p) { return p->ws_bytes; }
int alan_b200_num_inputs(const alan_b200_plan* p) { return p->n_inputs; }
int alan_b200_num_programs(const alan_b200_plan* p) { return p->n_programs; }

int alan_b200_program_launches(const alan_b200_plan* p, int program) {
    if (program < 0 || program >= p->n_programs) return -1;
    Ctx c{};
    int n = 0;
    if (p->dtype == 0) run_ops<float>(p, program, c, true, &n); else run_ops<double>(p, program, c, true, &n);
    return n;
}

static int run_direct(const alan_b200_plan* p, int program, const Ctx& c) {
    return p->dtype == 0 ? run_ops<float>(p, program, c, false, nullptr) : run_ops<double>(p, program, c, false, nullptr);
}

static int run_graphed(const alan_b200_plan* p, int program, const Ctx& c, const void* const* inputs, void* const* outputs,
                       const void* const* aux, void* ws);

static int run_generic(const alan_b200_plan* p, int program, const void* const* inputs, void* const* outputs,
                       const void* const* aux, void* ws, void* stream) {
    if (!p) return fail("null plan");
    if (program < 0 || program >= p->n_programs) return fail("program index out of range");
    Ctx c{inputs, outputs, aux, (char*)ws, (cudaStream_t)stream, p->sm_count};
    if (!p->use_graphs) return run_direct(p, program, c);
    const cudaStream_t user = (cudaStream_t)stream;
    const bool legacy = user == nullptr || user == cudaStreamLegacy || user == cudaStreamPerThread;
    if (!legacy) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(user, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
            cudaGetLastError();
            return run_direct(p, program, c);
        }
        return run_graphed(p, program, c, inputs, outputs, aux, ws);
    }
    {
        std::lock_guard<std::mutex> g(p->mu);
        if (!p->side) {
            if (cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&p->ev_out, cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                p->side = nullptr;
            }
        }
    }
    if (!p->side) return run_direct(p, program, c);
    cudaEventRecord(p->ev_in, user);
    cudaStreamWaitEvent(p->side, p->ev_in, 0);
    c.stream = p->side;
    int rc = run_graphed(p, program, c, inputs, outputs, aux, ws);
    cudaEventRecord(p->ev_out, p->side);
    cudaStreamWaitEvent(user, p->ev_out, 0);
    return rc;
}

static int run_graphed(const alan_b200_plan* p, int program, const Ctx& c, const void* const* inputs, void* const* outputs,
                       const void* const* aux, void* ws) {
    void* stream = (void*)c.stream;
    std::unique_lock<std::mutex> lock(p->mu);
    const int no = p->n_out[program], na = p->n_aux[program];
    std::vector<const void*> key;
    if (no >= 0 && na >= 0) {
        key.reserve(2 + p->n_inputs + no + na);
        key.push_back(ws); key.push_back(stream);
        for (int i = 0; i < p->n_inputs; ++i) key.push_back(inputs[i]);
        for (int i = 0; i < no; ++i) key.push_back(outputs[i]);
        for (int i = 0; i < na; ++i) key.push_back(aux[i]);
        for (auto& g : p->graphs[program]) {
            if (g.key == key) {
                g.last_use = ++p->tick;
                p->miss_streak[program] = 0;
                cudaGraphExec_t exec = g.exec;
                lock.unlock();
                cudaError_t e = cudaGraphLaunch(exec, c.stream);
                if (e != cudaSuccess) return fail(std::string("cudaGraphLaunch: ") + cudaGetErrorString(e));
                return 0;
            }
        }
    }
    if (p->graph_auto && p->miss_streak[program] >= 4) {
        // auto mode: this caller binds fresh buffers on every call -- captures would never be replayed
        lock.unlock();
        return run_direct(p, program, c);
    }
    ++p->miss_streak[program];
    // miss: capture this call's launches, instantiate, remember, launch
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        lock.unlock();
        return run_direct(p, program, c);
    }
    int rc = run_direct(p, program, c);
    cudaError_t e = cudaStreamEndCapture(c.stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess || !graph) { cudaGetLastError(); lock.unlock(); return run_direct(p, program, c); }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { cudaGetLastError(); lock.unlock(); return run_direct(p, program, c); }
    p->n_out[program] = c.max_out + 1;
    p->n_aux[program] = c.max_aux + 1;
    key.clear();
    key.push_back(ws); key.push_back(stream);
    for (int i = 0; i < p->n_inputs; ++i) key.push_back(inputs[i]);
    for (int i = 0; i <= c.max_out; ++i) key.push_back(outputs[i]);
    for (int i = 0; i <= c.max_aux; ++i) key.push_back(aux[i]);
    auto& vec = p->graphs[program];
    if (vec.size() >= 8) {                              // small LRU: training loops rebind a handful of buffers
        size_t victim = 0;
        for (size_t i = 1; i < vec.size(); ++i) if (vec[i].last_use < vec[victim].last_use) victim = i;
        cudaGraphExecDestroy(vec[victim].exec);
        vec.erase(vec.begin() + victim);
    }
    vec.push_back(GraphEntry{key, exec, ++p->tick});
    lock.unlock();
    e = cudaGraphLaunch(exec, c.stream);
    if (e != cudaSuccess) return fail(std::string("cudaGraphLaunch: ") + cudaGetErrorString(e));
    return 0;
}

int alan_b200_profile(const alan_b200_plan* p, int program, const void* const* inputs, void* const* outputs,
                      const void* const* aux, void* ws, void* stream, float* ms_per_op, int max_ops) {
    if (!p) return -1;
    if (program < 0 || program >= p->n_programs) { fail("program index out of range"); return -1; }
    Ctx c{inputs, outputs, aux, (char*)ws, (cudaStream_t)stream, p->sm_count};
    std::vector<cudaEvent_t> ev;
    int rc = p->dtype == 0 ? run_ops<float>(p, program, c, false, nullptr, &ev)
                           : run_ops<double>(p, program, c, false, nullptr, &ev);
    cudaStreamSynchronize(c.stream);
    int n = (int)ev.size() - 1;
    for (int i = 0; i < n && i < max_ops; ++i) cudaEventElapsedTime(&ms_per_op[i], ev[i], ev[i + 1]);
    for (auto e : ev) cudaEventDestroy(e);
    return rc ? -1 : n;
}

int alan_b200_run(const alan_b200_plan* p, int program, const void* const* inputs, void* const* outputs,
                  void* ws, void* stream) {
    return run_generic(p, program, inputs, outputs, nullptr, ws, stream);
}

int alan_b200_logpq_fwd(const alan_b200_plan* p, int segment, const void* const* inputs, void* lp_out,
                        void* ws, void* stream) {
    if (!p) return fail("null plan");
    if (segment < 0 || segment >= p->n_fwd) return fail("forward segment out of range");
    void* outs[1] = {lp_out};
    return run_generic(p, segment, inputs, outs, nullptr, ws, stream);
}

int alan_b200_logpq_bwd(const alan_b200_plan* p, int segment, const void* const* inputs, const void* grad_lp,
                        void* const* grads_out, void* ws, void* stream) {
    if (!p) return fail("null plan");
    if (segment < 0 || segment >= p->n_bwd) return fail("backward segment out of range");
    const void* aux[1] = {grad_lp};
    return run_generic(p, p->n_fwd + segment, inputs, grads_out, aux, ws, stream);
}

int alan_b200_resample(const alan_b200_plan* p, const void* const* inputs, const double* const* uniforms,
                       int64_t* const* idx_out, void* ws, void* stream) {
    if (!p) return fail("null plan");
    if (p->sample_prog < 0) return fail("plan was built without a resampling program");
    return run_generic(p, p->sample_prog, inputs, (void* const*)idx_out, (const void* const*)uniforms, ws, stream);
}

int alan_b200_gather(const void* x, const int64_t* idx, void* out, int elem_bytes, int64_t N, int64_t outer,
                     int64_t K, int64_t inner, int64_t outer_div, void* stream) {
    i64 total = N * outer * inner;
    if (total <= 0) return 0;
    int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    if (elem_bytes == 4)
        gather_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (const i64*)idx, (float*)out, N, outer, K, inner, outer_div);
    else if (elem_bytes == 8)
        gather_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)x, (const i64*)idx, (double*)out, N, outer, K, inner, outer_div);
    else return fail("gather: element size must be 4 or 8 bytes");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("CUDA launch error: ") + cudaGetErrorString(e));
    return 0;
}

int alan_b200_widen_u8(const void* src, void* dst, int64_t n, int dtype, void* stream) {
    if (n <= 0) return 0;
    if (((uintptr_t)src % 16) || ((uintptr_t)dst % 16)) return fail("widen_u8: source and destination must be 16-byte aligned");
    i64 g = (n / 16 + 255) / 256;
    const int grid = (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
    if (dtype == 0) widen_u8_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)src, (float*)dst, n);
    else if (dtype == 1) widen_u8_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)src, (double*)dst, n);
    else return fail("widen_u8: dtype must be 0 (f32) or 1 (f64)");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("CUDA launch error: ") + cudaGetErrorString(e));
    return 0;
}

extern "C++" template <typename T>
int qem_update_impl(int family, i64 n, double lr, const void* new0, const void* new1, void* mean0, void* mean1,
                    void* param0, void* param1, cudaStream_t st) {
    QemParams<T> p;
    p.family = family; p.n = n; p.lr = (T)lr; p.one_minus_lr = (T)(1.0 - lr);
    p.m0 = (const T*)new0; p.m1 = (const T*)new1; p.e0 = (T*)mean0; p.e1 = (T*)mean1; p.p0 = (T*)param0; p.p1 = (T*)param1;
    i64 g = (n + 255) / 256;
    qem_update_kernel<T><<<(int)(g > 148 * 8 ? 148 * 8 : g), 256, 0, st>>>(p);
    return 0;
}

int alan_b200_qem_update(int family, int64_t n, double lr, const void* new0, const void* new1, void* mean0, void* mean1,
                         void* param0, void* param1, int dtype, void* stream) {
    if (n <= 0) return 0;
    if (family < QF_NORMAL || family > QF_BETA) return fail("qem_update: unknown family");
    const bool two_stats = family == QF_NORMAL || family == QF_GAMMA || family == QF_BETA;
    const bool two_params = two_stats;
    if (!new0 || !mean0 || !param0 || (two_stats && (!new1 || !mean1)) || (two_params && !param1))
        return fail("qem_update: missing moment / mean / parameter buffer for this family");
    if (!two_stats) { new1 = nullptr; mean1 = nullptr; }
    if (dtype == 0) qem_update_impl<float>(family, n, lr, new0, new1, mean0, mean1, param0, param1, (cudaStream_t)stream);
    else if (dtype == 1) qem_update_impl<double>(family, n, lr, new0, new1, mean0, mean1, param0, param1, (cudaStream_t)stream);
    else return fail("qem_update: dtype must be 0 (f32) or 1 (f64)");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("CUDA launch error: ") + cudaGetErrorString(e));
    return 0;
}

// ---- unit-level ops --------------------------------------------------------------------

extern "C++" template <typename T>
int lse_eps_impl(const void* x, void* out, i64 n_out, i64 n_red, cudaStream_t st) {
    ReduceParams<T> p;
    memset(&p, 0, sizeof(p));
    p.d.nd = 2; p.d.n_a = 1; p.d.size[0] = (int)n_out; p.d.size[1] = (int)n_red;
    p.mode = R_LSE_EPS; p.nf = 1; p.f[0].ptr = x; p.f[0].stride[0] = n_red; p.f[0].stride[1] = 1;
    p.coeff[0] = T(1); p.out = (T*)out; p.acc = 0; p.scale = T(1); p.cadd = T(0);
    p.n_out = n_out; p.n_red = n_red; p.nsplit = 1;
    i64 blocks = (n_out * 32 + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8; if (blocks < 1) blocks = 1;
    reduce_warp_kernel<T, 1><<<(int)blocks, 256, 0, st>>>(p);
    return 0;
}

int alan_b200_lse_eps(const void* x, void* out, int64_t n_out, int64_t n_red, int dtype, void* stream) {
    if (n_out <= 0 || n_red <= 0) return fail("lse_eps: empty input");
    if (dtype == 0) lse_eps_impl<float>(x, out, n_out, n_red, (cudaStream_t)stream);
    else if (dtype == 1) lse_eps_impl<double>(x, out, n_out, n_red, (cudaStream_t)stream);
    else return fail("lse_eps: dtype must be 0 or 1");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("CUDA launch error: ") + cudaGetErrorString(e));
    return 0;
}

int64_t alan_b200_chain_scratch_elems(int64_t outer, int64_t T, int64_t K) {
    i64 n = T, tot = 0;
    while (n > 1) { i64 no = n / 2 + n % 2; tot += outer * no * K * K; n = no; }
    return tot > 0 ? tot : 1;
}

extern "C++" template <typename T>
int chain_impl(const void* ms_, void* levels_, void* out_, i64 outer, i64 Tn, i64 K, cudaStream_t st) {
    if (launch_chain_fwd<T>((const T*)ms_, (T*)levels_, (T*)out_, outer, Tn, K, st) == 0) return 0;
    size_t smem = (size_t)(2 * K * K + 2 * K) * sizeof(T);
    if (smem > 200 * 1024) return 1;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(chain_level_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const T* X = (const T*)ms_;
    T* Y = (T*)levels_;
    i64 n = Tn;
    while (n > 1) {
        i64 no = n / 2 + n % 2;
        chain_level_kernel<T><<<dim3((unsigned)no, (unsigned)outer), 256, smem, st>>>(X, Y, (int)n, (int)no, (int)K);
        X = Y; Y += outer * no * K * K; n = no;
    }
    i64 rows = outer * K;
    chain_final_kernel<T><<<(int)((rows + 127) / 128), 128, 0, st>>>(X, (T*)out_, rows, (int)K);
    return 0;
}

int alan_b200_logmmexp_chain(const void* ms, void* levels, void* out, int64_t outer, int64_t T, int64_t K,
                             int dtype, void* stream) {
    if (outer <= 0 || T <= 0 || K <= 0) return fail("logmmexp_chain: empty input");
    int rc = dtype == 0 ? chain_impl<float>(ms, levels, out, outer, T, K, (cudaStream_t)stream)
           : dtype == 1 ? chain_impl<double>(ms, levels, out, outer, T, K, (cudaStream_t)stream) : 2;
    if (rc == 1) return fail("logmmexp_chain: K too large for shared memory");
    if (rc == 2) return fail("logmmexp_chain: dtype must be 0 or 1");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("CUDA launch error: ") + cudaGetErrorString(e));
    return 0;
}

extern "C++" template <typename T>
void normal_bcast_impl(const void* v, const void* loc, const void* scale, void* out, i64 nc, i64 ne,
                              const int64_t* vs, const int64_t* ls, const int64_t* ss, cudaStream_t st) {
    ExprParams<T> p;
    memset(&p, 0, sizeof(p));
    p.d.nd = 2; p.d.n_a = 1; p.d.size[0] = (int)nc; p.d.size[1] = (int)ne;
    p.n_leaves = 3;
    const void* ptrs[3] = {v, loc, scale};
    const int64_t* strs[3] = {vs, ls, ss};
    for (int l = 0; l < 3; ++l) { p.leaf[l].ptr = ptrs[l]; p.leaf[l].stride[0] = strs[l][0]; p.leaf[l].stride[1] = strs[l][1]; }
    p.prog.n_instr = 4;
    for (int l = 0; l < 3; ++l) { p.prog.ins[l][0] = V_LOAD | (l << 8) | (l << 16); p.prog.ins[l][1] = 0; }
    p.prog.ins[3][0] = V_NORMAL | (3 << 8) | (0 << 16) | (1u << 24); p.prog.ins[3][1] = 2;
    p.prog.res = 3;
    p.out = (T*)out; p.acc = 0; p.scale = T(1); p.n_out = nc; p.n_red = ne;
    launch_expr_fwd<T>(p, st, 148);
}

int alan_b200_normal_logpdf_bcast(const void* value, const void* loc, const void* scale, void* out,
                                  int64_t n_cells, int64_t n_event, const int64_t* vs, const int64_t* ls,
                                  const int64_t* ss, int dtype, void* stream) {
    if (n_cells <= 0 || n_event <= 0) return fail("normal_logpdf_bcast: empty input");
    if (dtype == 0) normal_bcast_impl<float>(value, loc, scale, out, n_cells, n_event, vs, ls, ss, (cudaStream_t)stream);
    else if (dtype == 1) normal_bcast_impl<double>(value, loc, scale, out, n_cells, n_event, vs, ls, ss, (cudaStream_t)stream);
    else return fail("normal_logpdf_bcast: dtype must be 0 or 1");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("CUDA launch error: ") + cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
