// MultivariateNormal preparation.  reference: torch.distributions.MultivariateNormal as called by
// TorchDimDist.log_prob / .sample (src/alan/TorchDimDist.py:66,127-162; src/alan/dist.py:323-359).
//
// One warp per matrix (d <= 64, shared memory): from the covariance / precision / scale_tril argument S[n_mat, d, d]
//     L = scale_tril   (Sigma = L L^T; the draw is loc + L eps)
//     W = L^{-1}       (the Mahalanobis term is |W (x - loc)|^2: a matrix-vector expression of the factor VM)
//     c = -sum_i log L_ii - d/2 log(2 pi)
// so that  log N(x; loc, Sigma) = c - 1/2 |W (x - loc)|^2  is evaluated by the ordinary expression kernels over the
// (plate, K) cells, which never see a matrix factorisation.  The precision route follows torch's
// `_precision_to_scale_tril` (Cholesky of the flipped matrix, flipped back and transposed = W, then L = W^{-1}).
#pragma once
#include "common.cuh"

enum { MVN_COV = 0, MVN_PREC = 1, MVN_TRIL = 2, MVN_LOWRANK = 3 };   // 3: S = cov_factor [d, r], covariance = S S^T + diag(Dg)
#define AB_MVN_MAXD 64

// in-place lower Cholesky of the symmetric A[d][d] (row-major, leading dimension d) by one warp
template <typename T>
__device__ void mvn_cholesky(T* A, int d, int lane) {
    for (int j = 0; j < d; ++j) {
        T s = T(0);
        for (int k = lane; k < j; k += 32) s += A[j * d + k] * A[j * d + k];
        s = warp_sum(s);
        const T djj = sqrt(A[j * d + j] - s);
        __syncwarp();
        if (lane == 0) A[j * d + j] = djj;
        for (int i = j + 1 + lane; i < d; i += 32) {
            T acc = A[i * d + j];
            for (int k = 0; k < j; ++k) acc -= A[i * d + k] * A[j * d + k];
            A[i * d + j] = acc / djj;
        }
        __syncwarp();
    }
    for (int e = lane; e < d * d; e += 32) if (e % d > e / d) A[e] = T(0);
    __syncwarp();
}

// B = A^{-1} for lower-triangular A (forward substitution, one lane per column)
template <typename T>
__device__ void mvn_tri_inverse(const T* A, T* B, int d, int lane) {
    for (int c = lane; c < d; c += 32) {
        for (int i = 0; i < c; ++i) B[i * d + c] = T(0);
        for (int i = c; i < d; ++i) {
            T acc = (i == c) ? T(1) : T(0);
            for (int k = c; k < i; ++k) acc -= A[i * d + k] * B[k * d + c];
            B[i * d + c] = acc / A[i * d + i];
        }
    }
    __syncwarp();
}

template <typename T>
__global__ void __launch_bounds__(32) mvn_prep_kernel(const T* __restrict__ S, T* __restrict__ L, T* __restrict__ W,
                                                      T* __restrict__ c, i64 n_mat, int d, int mode,
                                                      const T* __restrict__ Dg, int r) {
    extern __shared__ __align__(16) unsigned char mvn_smem[];
    T* A = reinterpret_cast<T*>(mvn_smem);            // [d][d]
    T* B = A + d * d;                                 // [d][d]
    const int lane = threadIdx.x;
    for (i64 m = blockIdx.x; m < n_mat; m += gridDim.x) {
        const T* Sm = S + m * d * d;
        if (mode == MVN_LOWRANK) {
            // LowRankMultivariateNormal: covariance = F F^T + diag(Dg), F = cov_factor [d, r] of this matrix
            const T* F = S + m * d * r;
            for (int e = lane; e < d * d; e += 32) {
                const int i = e / d, j = e % d;
                T acc = (i == j) ? Dg[m * d + i] : T(0);
                for (int k = 0; k < r; ++k) acc += F[i * r + k] * F[j * r + k];
                A[e] = acc;
            }
        } else if (mode == MVN_PREC) { for (int e = lane; e < d * d; e += 32) A[e] = Sm[(d - 1 - e / d) * d + (d - 1 - e % d)]; }
        else { for (int e = lane; e < d * d; e += 32) A[e] = Sm[e]; }
        __syncwarp();
        if (mode == MVN_TRIL) { for (int e = lane; e < d * d; e += 32) if (e % d > e / d) A[e] = T(0); __syncwarp(); }
        else mvn_cholesky(A, d, lane);
        T *Lp = A, *Wp = B;
        if (mode == MVN_PREC) {
            // W[i][j] = Lf[d-1-j][d-1-i] (flip, then transpose), L = W^{-1}
            for (int e = lane; e < d * d; e += 32) B[e] = A[(d - 1 - e % d) * d + (d - 1 - e / d)];
            __syncwarp();
            mvn_tri_inverse(B, A, d, lane);
            Lp = A; Wp = B;
        } else {
            mvn_tri_inverse(A, B, d, lane);
        }
        T ld = T(0);
        for (int i = lane; i < d; i += 32) ld += log(Lp[i * d + i]);
        ld = warp_sum(ld);
        for (int e = lane; e < d * d; e += 32) { L[m * d * d + e] = Lp[e]; W[m * d * d + e] = Wp[e]; }
        if (lane == 0) c[m] = -ld - T(0.5) * T(d) * T(1.8378770664093454835606594728112);
        __syncwarp();
    }
}
