// common.cuh -- shared definitions for the alan_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>

#define AB_MAXD 10      // iteration dims per op (named axes + event dims)
#define AB_MAXL 10      // leaves / factors per op
#define AB_MAXI 32      // VM instructions per expression
#define AB_MAXC 16      // VM constants per expression
#define AB_NREG 32      // VM registers

typedef long long i64;

template <typename T> struct Eps;
template <> struct Eps<float>  { static __host__ __device__ float  v() { return 1.1920928955078125e-07f; } };
template <> struct Eps<double> { static __host__ __device__ double v() { return 2.220446049250313e-16; } };

template <typename T> __device__ __forceinline__ T neg_inf();
template <> __device__ __forceinline__ float  neg_inf<float>()  { return -INFINITY; }
template <> __device__ __forceinline__ double neg_inf<double>() { return -INFINITY; }

// exp/log in the tensor dtype.  fp32 uses the accurate libdevice expf/logf (<= 2 ulp),
// NOT the fast intrinsics: parity with the reference is 1e-5 relative on sums of
// thousands of terms.
__device__ __forceinline__ float  ab_exp(float x)  { return expf(x); }
__device__ __forceinline__ double ab_exp(double x) { return exp(x); }
__device__ __forceinline__ float  ab_log(float x)  { return logf(x); }
__device__ __forceinline__ double ab_log(double x) { return log(x); }
__device__ __forceinline__ float  ab_log1p(float x)  { return log1pf(x); }
__device__ __forceinline__ double ab_log1p(double x) { return log1p(x); }
__device__ __forceinline__ float  ab_max(float a, float b)   { return fmaxf(a, b); }
__device__ __forceinline__ double ab_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float  ab_abs(float a)  { return fabsf(a); }
__device__ __forceinline__ double ab_abs(double a) { return fabs(a); }
__device__ __forceinline__ float  ab_cos(float a)  { return cosf(a); }
__device__ __forceinline__ double ab_cos(double a) { return cos(a); }
__device__ __forceinline__ float  ab_sin(float a)  { return sinf(a); }
__device__ __forceinline__ double ab_sin(double a) { return sin(a); }

template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = ab_max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Fixed-order butterfly: every lane ends with the same bits, independent of timing.
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct Dims {
    int nd;                 // total dims
    int n_a;                // leading dims mapped to the "thread/output" index
    int size[AB_MAXD];
};

struct Opnd {
    const void* ptr;
    i64 stride[AB_MAXD];
    int mode;               // 0 plain, 1 shifted-by-one along mdim, 2 only-at-index-0 of mdim
    int mdim;
};

// linear index over dims [lo, hi) -> idx[lo..hi)
__device__ __forceinline__ void unravel(i64 lin, const Dims& d, int lo, int hi, int* idx) {
    if ((lin >> 31) == 0) {                 // the common case: 32-bit divisions (a 64-bit one costs ~3x as much)
        unsigned l = (unsigned)lin;
#pragma unroll 1
        for (int k = hi - 1; k >= lo; --k) {
            const unsigned s = (unsigned)d.size[k];
            const unsigned q = l / s;
            idx[k] = (int)(l - q * s);
            l = q;
        }
        return;
    }
#pragma unroll 1
    for (int k = hi - 1; k >= lo; --k) {
        int s = d.size[k];
        i64 q = lin / s;
        idx[k] = (int)(lin - q * s);
        lin = q;
    }
}

__device__ __forceinline__ i64 dot_stride(const Opnd& o, const int* idx, int lo, int hi) {
    i64 off = 0;
#pragma unroll 1
    for (int k = lo; k < hi; ++k) off += (i64)idx[k] * o.stride[k];
    return off;
}
