// vm.cuh -- register VM that evaluates one traced factor expression per iteration point.
//
// The reference resolves distribution arguments by calling Python lambdas on
// first-class-dim tensors and then calls torch.distributions.X(...).log_prob on the
// fully broadcast operands (src/alan/dist.py:211-232, TorchDimDist.py:127-162),
// materialising every intermediate.  Here the traced lambda AND the log-density are
// one instruction stream that is uniform across the grid (no divergence) and runs
// in registers/local memory per iteration point; only the event-summed factor cell
// is ever written.  The same stream is walked backwards for the adjoint.
#pragma once
#include "common.cuh"

enum VOp {
    V_LOAD = 0, V_CONST = 1, V_ADD = 2, V_SUB = 3, V_MUL = 4, V_DIV = 5, V_NEG = 6, V_EXP = 7, V_LOG = 8,
    V_SIGMOID = 9, V_SQUARE = 10, V_SQRT = 11, V_RECIP = 12, V_SOFTPLUS = 13, V_TANH = 14, V_ABS = 15,
    V_LOG1P = 16, V_POW = 17, V_LGAMMA = 18, V_MOV = 19, V_LT = 20, V_COS = 21, V_SIN = 22,
    V_NORMAL = 32, V_BERN_LOGITS = 33, V_BERN_PROBS = 34, V_LOGNORMAL = 35, V_LAPLACE = 36,
    V_EXPONENTIAL = 37, V_GAMMA = 38, V_BETA = 39, V_POISSON = 40, V_CAUCHY = 41, V_HALFNORMAL = 42,
    V_UNIFORM = 43, V_STUDENTT = 44, V_NEGBIN_LOGITS = 45, V_NEGBIN_PROBS = 46, V_BINOM_LOGITS = 47, V_BINOM_PROBS = 48
};

template <typename T>
struct VMProg {
    int n_instr;
    int res;
    unsigned ins[AB_MAXI][2];   // word0: op | dst<<8 | a<<16 | b<<24 ; word1: c | d<<8
    T consts[AB_MAXC];
};

#define HALF_LOG_2PI 0.91893853320467274178
#define LOG_PI 1.14472988584940017414
#define LOG_2 0.69314718055994530942

__device__ __forceinline__ float  ab_lgamma(float x)  { return lgammaf(x); }
__device__ __forceinline__ double ab_lgamma(double x) { return lgamma(x); }
__device__ __forceinline__ float  ab_tanh(float x)  { return tanhf(x); }
__device__ __forceinline__ double ab_tanh(double x) { return tanh(x); }
__device__ __forceinline__ float  ab_sqrt(float x)  { return sqrtf(x); }
__device__ __forceinline__ double ab_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float  ab_pow(float x, float y)   { return powf(x, y); }
__device__ __forceinline__ double ab_pow(double x, double y) { return pow(x, y); }

template <typename T> __device__ __forceinline__ T ab_sigmoid(T x) { return T(1) / (T(1) + ab_exp(-x)); }
// log_sigmoid(x) = min(x,0) - log1p(exp(-|x|))   (ATen log_sigmoid_forward)
template <typename T> __device__ __forceinline__ T ab_logsigmoid(T x) {
    return (x < T(0) ? x : T(0)) - ab_log1p(ab_exp(-ab_abs(x)));
}
template <typename T> __device__ __forceinline__ T ab_softplus(T x) {      // torch softplus, threshold 20
    return x > T(20) ? x : ab_log1p(ab_exp(x));
}
template <typename T> __device__ __forceinline__ T ab_xlogy(T x, T y) { return x == T(0) ? T(0) : x * ab_log(y); }

template <typename T> __device__ T ab_digamma(T x) {
    T r = 0;
    while (x < T(6)) { r -= T(1) / x; x += T(1); }
    T f = T(1) / (x * x);
    return r + ab_log(x) - T(0.5) / x
           - f * (T(1.0 / 12) - f * (T(1.0 / 120) - f * (T(1.0 / 252) - f * (T(1.0 / 240) - f * T(1.0 / 132)))));
}

// -(1-y)*x + log_sigmoid(x)  ==  -binary_cross_entropy_with_logits(x, y)  (torch bernoulli.py:121-125)
template <typename T> __device__ __forceinline__ T bern_logits_lp(T v, T x) {
    return -((T(1) - v) * x - ab_logsigmoid(x));
}
template <typename T> __device__ __forceinline__ T probs_to_logits(T p) {   // torch distributions/utils.py:101-137
    T e = Eps<T>::v();
    T pc = p < e ? e : (p > T(1) - e ? T(1) - e : p);
    return ab_log(pc) - ab_log1p(-pc);
}
template <typename T> __device__ __forceinline__ T normal_lp(T v, T loc, T scale) {   // torch normal.py:92-101
    T d = v - loc;
    return -(d * d) / (T(2) * (scale * scale)) - ab_log(scale) - T(HALF_LOG_2PI);
}

// torch negative_binomial.py:114-130: value a, total_count n, logits x
template <typename T> __device__ __forceinline__ T negbin_lp(T a, T n, T x) {
    T unnorm = n * ab_logsigmoid(-x) + a * ab_logsigmoid(x);
    T norm = -ab_lgamma(n + a) + ab_lgamma(T(1) + a) + ab_lgamma(n);
    if (n + a == T(0)) norm = T(0);
    return unnorm - norm;
}
// torch binomial.py:127-147: value k, total_count n, logits x
template <typename T> __device__ __forceinline__ T binom_lp(T k, T n, T x) {
    T normalize = n * (x > T(0) ? x : T(0)) + n * ab_log1p(ab_exp(-ab_abs(x))) - ab_lgamma(n + T(1));
    return k * x - ab_lgamma(k + T(1)) - ab_lgamma(n - k + T(1)) - normalize;
}

template <typename T>
__device__ __forceinline__ T vm_eval(const VMProg<T>& P, const T* leaf, T* reg) {
#pragma unroll 1
    for (int i = 0; i < P.n_instr; ++i) {
        unsigned w0 = P.ins[i][0], w1 = P.ins[i][1];
        int op = w0 & 0xff, dst = (w0 >> 8) & 0xff, ia = (w0 >> 16) & 0xff, ib = (w0 >> 24) & 0xff;
        int ic = w1 & 0xff, id = (w1 >> 8) & 0xff;
        T a = reg[ia < AB_NREG ? ia : 0], b = reg[ib < AB_NREG ? ib : 0];
        T y;
        switch (op) {
            case V_LOAD: y = leaf[ia]; break;
            case V_CONST: y = P.consts[ia]; break;
            case V_ADD: y = a + b; break;
            case V_SUB: y = a - b; break;
            case V_MUL: y = a * b; break;
            case V_DIV: y = a / b; break;
            case V_NEG: y = -a; break;
            case V_EXP: y = ab_exp(a); break;
            case V_LOG: y = ab_log(a); break;
            case V_SIGMOID: y = ab_sigmoid(a); break;
            case V_SQUARE: y = a * a; break;
            case V_SQRT: y = ab_sqrt(a); break;
            case V_RECIP: y = T(1) / a; break;
            case V_SOFTPLUS: y = ab_softplus(a); break;
            case V_TANH: y = ab_tanh(a); break;
            case V_ABS: y = ab_abs(a); break;
            case V_LOG1P: y = ab_log1p(a); break;
            case V_POW: y = ab_pow(a, b); break;
            case V_LGAMMA: y = ab_lgamma(a); break;
            case V_MOV: y = a; break;
            case V_LT: y = a < b ? T(1) : T(0); break;       // sampling transforms (Bernoulli draw = u < p); no gradient
            case V_COS: y = ab_cos(a); break;
            case V_SIN: y = ab_sin(a); break;
            case V_NORMAL: y = normal_lp(a, b, reg[ic]); break;
            case V_BERN_LOGITS: y = bern_logits_lp(a, b); break;
            case V_BERN_PROBS: y = bern_logits_lp(a, probs_to_logits(b)); break;
            case V_LOGNORMAL: { T lx = ab_log(a); y = normal_lp(lx, b, reg[ic]) - lx; break; }
            case V_LAPLACE: { T s = reg[ic]; y = -ab_log(T(2) * s) - ab_abs(a - b) / s; break; }
            case V_EXPONENTIAL: y = ab_log(b) - b * a; break;
            case V_GAMMA: { T r = reg[ic];
                y = ab_xlogy(b, r) + ab_xlogy(b - T(1), a) - r * a - ab_lgamma(b); break; }
            case V_BETA: { T c0 = reg[ic];
                y = ab_xlogy(b - T(1), a) + ab_xlogy(c0 - T(1), T(1) - a)
                    + ab_lgamma(b + c0) - ab_lgamma(b) - ab_lgamma(c0); break; }
            case V_POISSON: y = ab_xlogy(a, b) - b - ab_lgamma(a + T(1)); break;
            case V_CAUCHY: { T s = reg[ic]; T z = (a - b) / s;
                y = -T(LOG_PI) - ab_log(s) - ab_log1p(z * z); break; }
            case V_HALFNORMAL: y = (a >= T(0)) ? normal_lp(a, T(0), b) + T(LOG_2) : neg_inf<T>(); break;
            case V_UNIFORM: { T hi = reg[ic];
                y = (b <= a && hi > a) ? -ab_log(hi - b) : neg_inf<T>(); break; }
            case V_STUDENTT: { T loc = reg[ic], s = reg[id]; T df = b; T z = (a - loc) / s;
                T Z = ab_log(s) + T(0.5) * ab_log(df) + T(0.5 * LOG_PI) + ab_lgamma(T(0.5) * df)
                      - ab_lgamma(T(0.5) * (df + T(1)));
                y = -T(0.5) * (df + T(1)) * ab_log1p(z * z / df) - Z; break; }
            case V_NEGBIN_LOGITS: y = negbin_lp(a, b, reg[ic]); break;
            case V_NEGBIN_PROBS: y = negbin_lp(a, b, probs_to_logits(reg[ic])); break;
            case V_BINOM_LOGITS: y = binom_lp(a, b, reg[ic]); break;
            case V_BINOM_PROBS: y = binom_lp(a, b, probs_to_logits(reg[ic])); break;
            default: y = T(0);
        }
        reg[dst] = y;
    }
    return reg[P.res];
}

// Batched interpreter: one decode + one dispatch per VM instruction for B iteration points (the B consecutive
// cells of the innermost output dim a thread owns).  Instruction decode and the switch are the dominant cost of
// the scalar interpreter, so large factor expressions (radon's K^4 regression likelihood, Timeseries transitions)
// run several times faster here; arithmetic per point is identical to vm_eval (same expressions, same order).
#define VMB(EXPR) { _Pragma("unroll") for (int q = 0; q < B; ++q) { const T a = A[q], b = Bq[q], c = Cq[q], d = Dq[q]; (void)a; (void)b; (void)c; (void)d; Y[q] = (EXPR); } } break
template <typename T, int B>
__device__ __forceinline__ void vm_eval_batch(const VMProg<T>& P, const T (*leaf)[B], T (*reg)[B]) {
#pragma unroll 1
    for (int i = 0; i < P.n_instr; ++i) {
        const unsigned w0 = P.ins[i][0], w1 = P.ins[i][1];
        const int op = w0 & 0xff, dst = (w0 >> 8) & 0xff, ia = (w0 >> 16) & 0xff, ib = (w0 >> 24) & 0xff;
        const int ic = w1 & 0xff, id = (w1 >> 8) & 0xff;
        const T* A = reg[ia < AB_NREG ? ia : 0];
        const T* Bq = reg[ib < AB_NREG ? ib : 0];
        const T* Cq = reg[ic < AB_NREG ? ic : 0];
        const T* Dq = reg[id < AB_NREG ? id : 0];
        T Y[B];
        switch (op) {
            case V_LOAD: { _Pragma("unroll") for (int q = 0; q < B; ++q) Y[q] = leaf[ia][q]; } break;
            case V_CONST: { _Pragma("unroll") for (int q = 0; q < B; ++q) Y[q] = P.consts[ia]; } break;
            case V_ADD: VMB(a + b);
            case V_SUB: VMB(a - b);
            case V_MUL: VMB(a * b);
            case V_DIV: VMB(a / b);
            case V_NEG: VMB(-a);
            case V_EXP: VMB(ab_exp(a));
            case V_LOG: VMB(ab_log(a));
            case V_SIGMOID: VMB(ab_sigmoid(a));
            case V_SQUARE: VMB(a * a);
            case V_SQRT: VMB(ab_sqrt(a));
            case V_RECIP: VMB(T(1) / a);
            case V_SOFTPLUS: VMB(ab_softplus(a));
            case V_TANH: VMB(ab_tanh(a));
            case V_ABS: VMB(ab_abs(a));
            case V_LOG1P: VMB(ab_log1p(a));
            case V_POW: VMB(ab_pow(a, b));
            case V_LGAMMA: VMB(ab_lgamma(a));
            case V_MOV: VMB(a);
            case V_LT: VMB(a < b ? T(1) : T(0));
            case V_COS: VMB(ab_cos(a));
            case V_SIN: VMB(ab_sin(a));
            case V_NORMAL: VMB(normal_lp(a, b, c));
            case V_BERN_LOGITS: VMB(bern_logits_lp(a, b));
            case V_BERN_PROBS: VMB(bern_logits_lp(a, probs_to_logits(b)));
            case V_LOGNORMAL: VMB(normal_lp(ab_log(a), b, c) - ab_log(a));
            case V_LAPLACE: VMB(-ab_log(T(2) * c) - ab_abs(a - b) / c);
            case V_EXPONENTIAL: VMB(ab_log(b) - b * a);
            case V_GAMMA: VMB(ab_xlogy(b, c) + ab_xlogy(b - T(1), a) - c * a - ab_lgamma(b));
            case V_BETA: VMB(ab_xlogy(b - T(1), a) + ab_xlogy(c - T(1), T(1) - a) + ab_lgamma(b + c) - ab_lgamma(b) - ab_lgamma(c));
            case V_POISSON: VMB(ab_xlogy(a, b) - b - ab_lgamma(a + T(1)));
            case V_CAUCHY: VMB(-T(LOG_PI) - ab_log(c) - ab_log1p(((a - b) / c) * ((a - b) / c)));
            case V_HALFNORMAL: VMB((a >= T(0)) ? normal_lp(a, T(0), b) + T(LOG_2) : neg_inf<T>());
            case V_UNIFORM: VMB((b <= a && c > a) ? -ab_log(c - b) : neg_inf<T>());
            case V_STUDENTT: VMB(-T(0.5) * (b + T(1)) * ab_log1p(((a - c) / d) * ((a - c) / d) / b)
                                 - (ab_log(d) + T(0.5) * ab_log(b) + T(0.5 * LOG_PI) + ab_lgamma(T(0.5) * b)
                                    - ab_lgamma(T(0.5) * (b + T(1)))));
            case V_NEGBIN_LOGITS: VMB(negbin_lp(a, b, c));
            case V_NEGBIN_PROBS: VMB(negbin_lp(a, b, probs_to_logits(c)));
            case V_BINOM_LOGITS: VMB(binom_lp(a, b, c));
            case V_BINOM_PROBS: VMB(binom_lp(a, b, probs_to_logits(c)));
            default: { _Pragma("unroll") for (int q = 0; q < B; ++q) Y[q] = T(0); } break;
        }
#pragma unroll
        for (int q = 0; q < B; ++q) reg[dst][q] = Y[q];
    }
}
#undef VMB

// Reverse sweep.  reg[] must hold the forward values.  Returns d(result)/d(leaf `target`)
// (sum over every LOAD of that leaf).  adj[] is scratch of AB_NREG entries.
template <typename T>
__device__ __forceinline__ T vm_grad(const VMProg<T>& P, const T* reg, T* adj, int target) {
#pragma unroll 1
    for (int r = 0; r < AB_NREG; ++r) adj[r] = T(0);
    adj[P.res] = T(1);
    T out = T(0);
#pragma unroll 1
    for (int i = P.n_instr - 1; i >= 0; --i) {
        unsigned w0 = P.ins[i][0], w1 = P.ins[i][1];
        int op = w0 & 0xff, dst = (w0 >> 8) & 0xff, ia = (w0 >> 16) & 0xff, ib = (w0 >> 24) & 0xff;
        int ic = w1 & 0xff, id = (w1 >> 8) & 0xff;
        T g = adj[dst];
        adj[dst] = T(0);          // registers are single-assignment per live range
        if (g == T(0)) continue;
        T a = reg[ia < AB_NREG ? ia : 0], b = reg[ib < AB_NREG ? ib : 0], y = reg[dst];
        switch (op) {
            case V_LOAD: if (ia == target) out += g; break;
            case V_CONST: break;
            case V_ADD: adj[ia] += g; adj[ib] += g; break;
            case V_SUB: adj[ia] += g; adj[ib] -= g; break;
            case V_MUL: adj[ia] += g * b; adj[ib] += g * a; break;
            case V_DIV: adj[ia] += g / b; adj[ib] -= g * a / (b * b); break;
            case V_NEG: adj[ia] -= g; break;
            case V_EXP: adj[ia] += g * y; break;
            case V_LOG: adj[ia] += g / a; break;
            case V_SIGMOID: adj[ia] += g * y * (T(1) - y); break;
            case V_SQUARE: adj[ia] += g * T(2) * a; break;
            case V_SQRT: adj[ia] += g / (T(2) * y); break;
            case V_RECIP: adj[ia] -= g * y * y; break;
            case V_SOFTPLUS: adj[ia] += g * (a > T(20) ? T(1) : ab_sigmoid(a)); break;
            case V_TANH: adj[ia] += g * (T(1) - y * y); break;
            case V_ABS: adj[ia] += g * (a > T(0) ? T(1) : (a < T(0) ? T(-1) : T(0))); break;
            case V_LOG1P: adj[ia] += g / (T(1) + a); break;
            case V_POW: adj[ia] += g * b * ab_pow(a, b - T(1));
                        if (a > T(0)) adj[ib] += g * y * ab_log(a); break;
            case V_LGAMMA: adj[ia] += g * ab_digamma(a); break;
            case V_MOV: adj[ia] += g; break;
            case V_COS: adj[ia] -= g * ab_sin(a); break;
            case V_SIN: adj[ia] += g * ab_cos(a); break;
            case V_NORMAL: { T s = reg[ic]; T d = a - b; T iv = T(1) / (s * s);
                adj[ia] -= g * d * iv; adj[ib] += g * d * iv; adj[ic] += g * (d * d * iv - T(1)) / s; break; }
            case V_BERN_LOGITS: { T sg = ab_sigmoid(b);
                adj[ia] += g * b; adj[ib] += g * (a - sg); break; }
            case V_BERN_PROBS: { T e = Eps<T>::v();
                T x = probs_to_logits(b); T sg = ab_sigmoid(x);
                adj[ia] += g * x;
                if (b >= e && b <= T(1) - e) adj[ib] += g * (a - sg) / (b * (T(1) - b));
                break; }
            case V_LOGNORMAL: { T s = reg[ic]; T lx = ab_log(a); T d = lx - b; T iv = T(1) / (s * s);
                adj[ia] += g * (-d * iv - T(1)) / a; adj[ib] += g * d * iv;
                adj[ic] += g * (d * d * iv - T(1)) / s; break; }
            case V_LAPLACE: { T s = reg[ic]; T d = a - b; T sgn = d > T(0) ? T(1) : (d < T(0) ? T(-1) : T(0));
                adj[ia] -= g * sgn / s; adj[ib] += g * sgn / s;
                adj[ic] += g * (-T(1) / s + ab_abs(d) / (s * s)); break; }
            case V_EXPONENTIAL: adj[ia] -= g * b; adj[ib] += g * (T(1) / b - a); break;
            case V_GAMMA: { T r = reg[ic];
                adj[ia] += g * ((b - T(1)) / a - r);
                adj[ib] += g * (ab_log(r) + ab_log(a) - ab_digamma(b));
                adj[ic] += g * (b / r - a); break; }
            case V_BETA: { T c0 = reg[ic];
                adj[ia] += g * ((b - T(1)) / a - (c0 - T(1)) / (T(1) - a));
                T dg = ab_digamma(b + c0);
                adj[ib] += g * (ab_log(a) + dg - ab_digamma(b));
                adj[ic] += g * (ab_log(T(1) - a) + dg - ab_digamma(c0)); break; }
            case V_POISSON: adj[ib] += g * (a / b - T(1)); break;
            case V_CAUCHY: { T s = reg[ic]; T z = (a - b) / s; T q = T(2) * z / (T(1) + z * z);
                adj[ia] -= g * q / s; adj[ib] += g * q / s; adj[ic] += g * (q * z - T(1)) / s; break; }
            case V_HALFNORMAL: { T iv = T(1) / (b * b);
                adj[ia] -= g * a * iv; adj[ib] += g * (a * a * iv - T(1)) / b; break; }
            case V_UNIFORM: { T hi = reg[ic]; T w = T(1) / (hi - b); adj[ib] += g * w; adj[ic] -= g * w; break; }
            case V_STUDENTT: { T loc = reg[ic], s = reg[id]; T df = b; T z = (a - loc) / s;
                T q = (df + T(1)) * z / (df + z * z);
                adj[ia] -= g * q / s; adj[ic] += g * q / s; adj[id] += g * (q * z - T(1)) / s;
                T u = z * z / df;
                adj[ib] += g * (-T(0.5) * ab_log1p(u) + T(0.5) * (df + T(1)) * u / (df * (T(1) + u))
                                - T(0.5) / df - T(0.5) * ab_digamma(T(0.5) * df)
                                + T(0.5) * ab_digamma(T(0.5) * (df + T(1)))); break; }
            case V_NEGBIN_LOGITS: case V_NEGBIN_PROBS: {      // value a (discrete: no gradient), total_count b, logits / probs reg[ic]
                T pr = reg[ic], e = Eps<T>::v();
                T x = (op == V_NEGBIN_PROBS) ? probs_to_logits(pr) : pr;
                T sg = ab_sigmoid(x);
                T gx = g * (a * (T(1) - sg) - b * sg);       // d/dx [ n logsig(-x) + a logsig(x) ]
                if (op == V_NEGBIN_PROBS) { if (pr >= e && pr <= T(1) - e) adj[ic] += gx / (pr * (T(1) - pr)); }
                else adj[ic] += gx;
                adj[ib] += g * (ab_logsigmoid(-x) + ab_digamma(b + a) - ab_digamma(b));
                break; }
            case V_BINOM_LOGITS: case V_BINOM_PROBS: {        // value a, total_count b (discrete), logits / probs reg[ic]
                T pr = reg[ic], e = Eps<T>::v();
                T x = (op == V_BINOM_PROBS) ? probs_to_logits(pr) : pr;
                T gx = g * (a - b * ab_sigmoid(x));
                if (op == V_BINOM_PROBS) { if (pr >= e && pr <= T(1) - e) adj[ic] += gx / (pr * (T(1) - pr)); }
                else adj[ic] += gx;
                break; }
            default: break;
        }
    }
    return out;
}
