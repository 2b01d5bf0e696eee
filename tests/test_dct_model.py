"""The generated straight-line DCT (same program as the CUDA kernels use) must equal scipy.fftpack bit for bit."""
import ctypes

import numpy as np
import pytest
from scipy.fftpack import dct, idct

from oracle.build_c import build_dct


@pytest.fixture(scope="module")
def lib():
    l = ctypes.CDLL(build_dct())
    for f in (l.ducc_1d, l.ducc_2d):
        f.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int]
        f.restype = None
    return l


@pytest.mark.parametrize("n", [2, 4, 8, 16])
def test_1d_bit_exact(lib, n):
    rng = np.random.default_rng(n)
    for kind in ("int", "float", "halves"):
        if kind == "int":
            x = rng.integers(-4080, 4081, (200000, n)).astype(np.float64)
        elif kind == "float":
            x = rng.normal(0, 500, (200000, n))
        else:
            x = rng.integers(-4080, 4081, (200000, n)) / 2.0
        for inverse, ref in ((0, dct(x, axis=1, norm="ortho")), (1, idct(x, axis=1, norm="ortho"))):
            y = x.copy()
            lib.ducc_1d(y.ctypes.data, len(y), n, inverse)
            assert np.array_equal(y.view(np.uint64), ref.view(np.uint64)), (kind, inverse)


@pytest.mark.parametrize("n", [2, 4, 8, 16])
def test_2d_residual_blocks_bit_exact(lib, n):
    rng = np.random.default_rng(100 + n)
    # integer residual blocks (the encoder's input) including flat ones, which produce exact rounding ties
    x = rng.integers(-255, 256, (50000, n, n)).astype(np.float64)
    x[::7] = np.round(x[::7] / 64) * 8
    ref = dct(dct(x, axis=1, norm="ortho"), axis=2, norm="ortho")
    y = x.copy()
    lib.ducc_2d(y.ctypes.data, len(y), n, 0)
    assert np.array_equal(y.view(np.uint64), ref.view(np.uint64))
    assert np.array_equal(np.round(y), np.round(ref))
    # dequantised coefficient blocks -> IDCT
    c = (rng.integers(-40, 41, (50000, n, n)) * (rng.random((50000, n, n)) < 0.2) * 16).astype(np.float64)
    ref = idct(idct(c, axis=1, norm="ortho"), axis=2, norm="ortho")
    y = c.copy()
    lib.ducc_2d(y.ctypes.data, len(y), n, 1)
    assert np.array_equal(y.view(np.uint64), ref.view(np.uint64))
