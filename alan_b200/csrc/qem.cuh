// QEM parameter update.  reference: BoundPlate._update_qem_moving_avg / _update_qem_convparams
// (src/alan/BoundPlate.py:256-296) and the mean -> conventional parameter conversions of src/alan/conversions.py:46-296.
//
// One elementwise pass over the parameters of ONE latent variable, in place:
//     mean_s <- mean_s * (1 - lr) + lr * moment_s          for the family's sufficient statistics s (<= 2)
//     conventional parameters <- mean2conv(mean_0, mean_1)
// Closed forms: Normal, Bernoulli, Poisson, Exponential, HalfNormal.  Newton iterations with digamma / trigamma: Gamma
// (Minka's generalised Newton, 6 steps) and Beta (the 2-component Dirichlet fit: 5 fixed-point steps through
// inverse_digamma, then 6 Newton steps), iteration counts as upstream.  digamma / trigamma follow torch's own
// algorithms (ATen Math.h calc_digamma / calc_trigamma: recurrence to x >= 10 / 6 steps, then the asymptotic series),
// so the iterates track the reference's to rounding.
#pragma once
#include "common.cuh"

enum { QF_NORMAL = 0, QF_BERNOULLI = 1, QF_POISSON = 2, QF_EXPONENTIAL = 3, QF_HALFNORMAL = 4, QF_GAMMA = 5, QF_BETA = 6 };

template <typename T>
__device__ T qem_digamma(T x) {
    const T PI = T(3.14159265358979323846);
    if (x == T(0)) return copysign(T(INFINITY), -x);
    T refl = T(0);
    if (x < T(0)) {
        if (x == trunc(x)) return T(NAN);
        const T r = x - trunc(x);
        refl = -PI / tan(PI * r);
        x = T(1) - x;
    }
    T result = T(0);
    while (x < T(10)) { result -= T(1) / x; x += T(1); }
    if (x == T(10)) return refl + result + T(2.25175258906672110764);
    T y = T(0);
    if (x < T(1.0e17)) {
        const T z = T(1) / (x * x);
        T p = T(8.33333333333333333333E-2);
        p = p * z + T(-2.10927960927960927961E-2);
        p = p * z + T(7.57575757575757575758E-3);
        p = p * z + T(-4.16666666666666666667E-3);
        p = p * z + T(3.96825396825396825397E-3);
        p = p * z + T(-8.33333333333333333333E-3);
        p = p * z + T(8.33333333333333333333E-2);
        y = z * p;
    }
    return refl + result + log(x) - (T(0.5) / x) - y;
}

template <typename T>
__device__ T qem_trigamma(T x) {
    const T PI = T(3.14159265358979323846);
    T sign = T(1), result = T(0);
    if (x < T(0.5)) {
        sign = T(-1);
        const T s = sin(PI * x);
        result -= (PI * PI) / (s * s);
        x = T(1) - x;
    }
    for (int i = 0; i < 6; ++i) { result += T(1) / (x * x); x += T(1); }
    const T ixx = T(1) / (x * x);
    result += (T(1) + T(1) / (T(2) * x) + ixx * (T(1) / T(6) - ixx * (T(1) / T(30) - ixx * (T(1) / T(42))))) / x;
    return sign * result;
}

// conversions.py:8-35: x with digamma(x) = y
template <typename T>
__device__ T qem_inverse_digamma(T y, T digamma_one) {
    T x = (y > T(-2.22)) ? exp(y) + T(0.5) : -(T(1) / (y - digamma_one));
    for (int i = 0; i < 6; ++i) x = x - (qem_digamma(x) - y) / qem_trigamma(x);
    return x;
}

template <typename T> struct TinyOf;
template <> struct TinyOf<float> { static __device__ float v() { return 1.17549435e-38f; } };
template <> struct TinyOf<double> { static __device__ double v() { return 2.2250738585072014e-308; } };

template <typename T>
struct QemParams {
    int family;
    i64 n;
    T lr, one_minus_lr;
    const T* m0; const T* m1;       // fresh moments (sample.moments of the sufficient statistics)
    T* e0; T* e1;                   // moving-average mean parameters, updated in place
    T* p0; T* p1;                   // conventional parameters, overwritten
};

template <typename T>
__global__ void __launch_bounds__(256) qem_update_kernel(const __grid_constant__ QemParams<T> p) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (i64)gridDim.x * blockDim.x) {
        T a = p.e0[i] * p.one_minus_lr + p.lr * p.m0[i];
        p.e0[i] = a;
        T b = T(0);
        if (p.e1 != nullptr) { b = p.e1[i] * p.one_minus_lr + p.lr * p.m1[i]; p.e1[i] = b; }
        switch (p.family) {
            case QF_NORMAL: {                                       // conversions.py:86-101
                p.p0[i] = a;
                T s = sqrt(b - a * a);
                if (s < TinyOf<T>::v()) s = TinyOf<T>::v();         // clamp(min=tiny); NaN stays NaN
                p.p1[i] = s;
                break;
            }
            case QF_BERNOULLI: case QF_POISSON: p.p0[i] = a; break; // :53-84
            case QF_EXPONENTIAL: p.p0[i] = T(1) / a; break;         // :103-114
            case QF_HALFNORMAL: p.p0[i] = sqrt(a); break;           // :281-293
            case QF_GAMMA: {                                        // :189-226 (a = E log x, b = E x)
                const T diff = a - log(b);
                T alpha = -T(0.5) / diff;
                for (int it = 0; it < 6; ++it) {
                    const T num = diff + log(alpha) - qem_digamma(alpha);
                    const T den = T(1) - alpha * qem_trigamma(alpha);
                    alpha = alpha * (T(1) / (T(1) + num / den));
                }
                p.p0[i] = alpha;
                p.p1[i] = alpha / b;
                break;
            }
            case QF_BETA: {                                         // :160-186 -> :117-158 (a = E log x, b = E log(1 - x))
                const T d1 = qem_digamma(T(1));
                T a0 = T(1), a1 = T(1);
                for (int it = 0; it < 5; ++it) {
                    const T ds = qem_digamma(a0 + a1);
                    const T n0 = qem_inverse_digamma(ds + a, d1), n1 = qem_inverse_digamma(ds + b, d1);
                    a0 = n0; a1 = n1;
                }
                for (int it = 0; it < 6; ++it) {
                    const T sum = a0 + a1;
                    const T ds = qem_digamma(sum);
                    const T g0 = ds - qem_digamma(a0) + a, g1 = ds - qem_digamma(a1) + b;
                    const T z = qem_trigamma(sum);
                    const T q0 = -qem_trigamma(a0), q1 = -qem_trigamma(a1);
                    const T bb = (g0 / q0 + g1 / q1) / (T(1) / z + (T(1) / q0 + T(1) / q1));
                    a0 = a0 - (g0 - bb) / q0;
                    a1 = a1 - (g1 - bb) / q1;
                }
                p.p0[i] = a0;
                p.p1[i] = a1;
                break;
            }
        }
    }
}
