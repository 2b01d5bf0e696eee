// 2-D DCT / IDCT on n x n blocks held in shared memory as double, one thread per 1-D transform.
// The reference computes dct(dct(X, axis=0, 'ortho'), axis=1, 'ortho') in float64 (Encoder.py:779-784, 810-817);
// the transform is computed in FP64 here as well so that the only coefficients that can differ from SciPy are
// exact rounding ties.  dct1d / idct1d are the single place that defines the arithmetic order.
#pragma once
#include "so_common.cuh"

// Orthonormal DCT-II matrices C[k][n] = s_k * cos(pi*k*(2n+1)/(2N)) for N = 2, 4, 8, 16 (host-filled, FP64).
// layout: offset(N) = {2:0, 4:4, 8:20, 16:84}; total 340 entries.
__constant__ double c_dct[340];
// inverse scan position: c_scanpos[offset(N) + u*N + v] = index of (u,v) in the anti-diagonal scan of
// entropy_encoder_block (Encoder.py:1095-1123)
__constant__ uint16_t c_scanpos[340];

__device__ __forceinline__ constexpr int tbl_off(int n) { return n == 2 ? 0 : (n == 4 ? 4 : (n == 8 ? 20 : 84)); }

// in-place forward transform of N values at v[0], v[stride], ...
template <int N>
__device__ __forceinline__ void dct1d(double* v, int stride) {
    double x[N], y[N];
#pragma unroll
    for (int n = 0; n < N; ++n) x[n] = v[n * stride];
    const double* C = c_dct + tbl_off(N);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double s = 0.0;
#pragma unroll
        for (int n = 0; n < N; ++n) s = fma(C[k * N + n], x[n], s);
        y[k] = s;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) v[k * stride] = y[k];
}

template <int N>
__device__ __forceinline__ void idct1d(double* v, int stride) {
    double x[N], y[N];
#pragma unroll
    for (int n = 0; n < N; ++n) x[n] = v[n * stride];
    const double* C = c_dct + tbl_off(N);
#pragma unroll
    for (int n = 0; n < N; ++n) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) s = fma(C[k * N + n], x[k], s);
        y[n] = s;
    }
#pragma unroll
    for (int n = 0; n < N; ++n) v[n * stride] = y[n];
}

// 2-D transform of the BS x BS tile `ws` (row pitch BS+1 doubles) viewed as (BS/N)^2 independent N x N blocks.
// Must be called by all threads of the CTA (contains __syncthreads); axis 0 (columns) first, like the reference.
template <int BS, int N, bool INVERSE>
__device__ __forceinline__ void transform2d(double* ws, int t) {
    constexpr int P = BS + 1;
    constexpr int PARTS = BS / N;            // 1 (whole block) or 2 (four sub-blocks)
    __syncthreads();
    if (t < BS * PARTS) {                    // column `col`, rows part*N .. part*N+N-1
        const int col = t % BS, part = t / BS;
        double* p = ws + part * N * P + col;
        if (INVERSE) idct1d<N>(p, P); else dct1d<N>(p, P);
    }
    __syncthreads();
    if (t < BS * PARTS) {                    // row `row`, columns part*N ..
        const int row = t % BS, part = t / BS;
        double* p = ws + row * P + part * N;
        if (INVERSE) idct1d<N>(p, 1); else dct1d<N>(p, 1);
    }
    __syncthreads();
}
