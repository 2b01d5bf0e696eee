/* CPU build of the generated straight-line DCT (tools/dctgen/gen_dct.py).  TEST INFRASTRUCTURE ONLY: lets the tests
 * check, without a GPU, that the exact program compiled into the CUDA library reproduces scipy.fftpack bit for bit.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o oracle/_build/libdct_ducc.so oracle/dct_ducc_c.c */
#include "dct_ducc_generated.h"

static void line(double* v, long stride, int n, int inverse) {
    if (!inverse) {
        if (n == 2) ducc_dct2_2(v, stride); else if (n == 4) ducc_dct2_4(v, stride);
        else if (n == 8) ducc_dct2_8(v, stride); else ducc_dct2_16(v, stride);
    } else {
        if (n == 2) ducc_dct3_2(v, stride); else if (n == 4) ducc_dct3_4(v, stride);
        else if (n == 8) ducc_dct3_8(v, stride); else ducc_dct3_16(v, stride);
    }
}

/* in-place 1-D transforms of `count` contiguous vectors of length n */
void ducc_1d(double* data, long count, int n, int inverse) {
    for (long i = 0; i < count; ++i) line(data + i * n, 1, n, inverse);
}

/* in-place 2-D transforms of `count` contiguous n x n blocks: axis 0 (columns) first, then axis 1 (Encoder.py:781) */
void ducc_2d(double* data, long count, int n, int inverse) {
    for (long b = 0; b < count; ++b) {
        double* p = data + b * n * n;
        for (int c = 0; c < n; ++c) line(p + c, n, n, inverse);
        for (int r = 0; r < n; ++r) line(p + r * n, 1, n, inverse);
    }
}
