"""Deterministic synthetic Y-only sequences shared by the tests, the golden generator and bench.py.

All generators return ``uint8 [F, H, W]``.  The textured patterns translate / zoom so motion search has signal,
plus seeded Gaussian noise.  ``bright`` biases the mean above 128 so that the reference's uint8 wrap in half-pel
interpolation (quirk Q1, ``Encoder.py:390-397``) is exercised.
"""
from __future__ import annotations

import numpy as np


def translating(F: int, H: int, W: int, seed: int = 0, bright: bool = False, noise: float = 3.0) -> np.ndarray:
    """The survey generator: ``clip((sin((x+2t)/7)+cos((y+t)/5))*50+128+N(0,3))``; seed 0 reproduces SURVEY §4."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    out = np.empty((F, H, W), dtype=np.uint8)
    base = 190.0 if bright else 128.0
    amp = 30.0 if bright else 50.0
    for t in range(F):
        f = (np.sin((xx + 2 * t) / 7.0) + np.cos((yy + t) / 5.0)) * amp + base + rng.normal(0, noise, (H, W))
        out[t] = np.clip(f, 0, 255).astype(np.uint8)
    return out


def zooming(F: int, H: int, W: int, seed: int = 1, noise: float = 2.0) -> np.ndarray:
    """Radial texture that zooms about the frame centre by 1 % per frame, plus a diagonal drift."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    cy, cx = (H - 1) / 2.0, (W - 1) / 2.0
    out = np.empty((F, H, W), dtype=np.uint8)
    for t in range(F):
        s = 1.0 + 0.01 * t
        u = (xx - cx) / s + 1.5 * t
        v = (yy - cy) / s - 0.5 * t
        f = 128 + 45 * np.sin(u / 6.0) * np.cos(v / 9.0) + 35 * np.sin((u + v) / 13.0) + rng.normal(0, noise, (H, W))
        out[t] = np.clip(f, 0, 255).astype(np.uint8)
    return out


def scene_cut(F: int, H: int, W: int, seed: int = 2, cut_at: int = 2) -> np.ndarray:
    """Translating texture whose content is replaced by unrelated noise-heavy texture from frame ``cut_at`` on."""
    a = translating(F, H, W, seed=seed)
    rng = np.random.default_rng(seed + 100)
    b = zooming(F, H, W, seed=seed + 7, noise=12.0)
    b = np.clip(b.astype(np.int32) + rng.integers(-20, 21, size=b.shape), 0, 255).astype(np.uint8)
    a[cut_at:] = b[cut_at:]
    return a


def flat_ties(F: int, H: int, W: int, seed: int = 3) -> np.ndarray:
    """Piece-wise constant frames with few grey levels: almost every SAD ties, stressing the argmin order."""
    rng = np.random.default_rng(seed)
    out = np.empty((F, H, W), dtype=np.uint8)
    tile = rng.integers(0, 4, size=(F, (H + 3) // 4, (W + 3) // 4)) * 60 + 20
    for t in range(F):
        out[t] = np.kron(tile[t], np.ones((4, 4), dtype=np.int64))[:H, :W].astype(np.uint8)
    out[1:] = np.where(rng.random((F - 1, H, W)) < 0.5, out[:-1], out[1:])
    return out


GENERATORS = {"translating": translating, "zooming": zooming, "scene_cut": scene_cut, "flat_ties": flat_ties}


def make(kind: str, F: int, H: int, W: int, **kw) -> np.ndarray:
    return GENERATORS[kind](F, H, W, **kw)
