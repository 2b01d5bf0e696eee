// 2-D DCT / IDCT on n x n blocks held in shared memory as double, one thread per 1-D transform.
// The reference computes dct(dct(X, axis=0, 'ortho'), axis=1, 'ortho') in float64 and rounds (Encoder.py:779-784,
// 810-817).  Exact rounding ties are common, so the 1-D transforms replay SciPy's own sequence of
// IEEE-754 double operations; FP64 is cheap here (B200: ~64 DFMA/clk/SM, the transform is <1 % of a frame).
#pragma once
#include "so_common.cuh"

#include "so_dct_ducc.cuh"

// inverse scan position: c_scanpos[offset(N) + u*N + v] = index of (u,v) in the anti-diagonal scan of
// entropy_encoder_block (Encoder.py:1095-1123); offset(N) = {2:0, 4:4, 8:20, 16:84}
__constant__ uint16_t c_scanpos[340];

__device__ __forceinline__ constexpr int tbl_off(int n) { return n == 2 ? 0 : (n == 4 ? 4 : (n == 8 ? 20 : 84)); }

// in-place 1-D transforms of N values at v[0], v[stride], ...: the generated straight-line programs that reproduce
// scipy.fftpack.dct / idct (norm='ortho') bit for bit (tools/dctgen/gen_dct.py, tests/test_dct_model.py)
template <int N>
__device__ __forceinline__ void dct1d(double* v, int stride) {
    if constexpr (N == 2) ducc_dct2_2(v, stride);
    else if constexpr (N == 4) ducc_dct2_4(v, stride);
    else if constexpr (N == 8) ducc_dct2_8(v, stride);
    else ducc_dct2_16(v, stride);
}

template <int N>
__device__ __forceinline__ void idct1d(double* v, int stride) {
    if constexpr (N == 2) ducc_dct3_2(v, stride);
    else if constexpr (N == 4) ducc_dct3_4(v, stride);
    else if constexpr (N == 8) ducc_dct3_8(v, stride);
    else ducc_dct3_16(v, stride);
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
