"""Dense-array <-> reference ``encoded_package`` list conversions.  TEST INFRASTRUCTURE ONLY.

Packed layout (also the layout of the C-ABI outputs, see ``include/streamoptima_b200.h``):
  split  u8  [F, nblk]              0 = whole block, 1 = four sub-blocks (Z order)
  mv     i16 [F, nblk, 4, 3]        P frame: (dx, dy, ref) of the whole block in slot 0 (split=0) or of the four
                                    sub-blocks (split=1).  I frame: scalar mode/offset in component 0.
  levels i16 [F, H, W]              quantised coefficients, each (sub-)block stored at its pixel position
The package format is the one built at ``Encoder.py:1877-1888``.
"""
from __future__ import annotations

import numpy as np


def package_to_arrays(frame_types, mvs_per_frame, levels_per_frame, H, W, bs):
    F = len(frame_types)
    nbx, nby = W // bs, H // bs
    nblk = nbx * nby
    sub = bs // 2
    split = np.zeros((F, nblk), np.uint8)
    mv = np.zeros((F, nblk, 4, 3), np.int16)
    lev = np.zeros((F, H, W), np.int16)
    for f in range(F):
        for b in range(nblk):
            s, m = mvs_per_frame[f][b]
            ls, l = levels_per_frame[f][b]
            assert s == ls
            split[f, b] = s
            y, x = (b // nbx) * bs, (b % nbx) * bs
            if s == 0:
                if frame_types[f] == 0:
                    mv[f, b, 0, 0] = m
                else:
                    mv[f, b, 0] = m
                lev[f, y:y + bs, x:x + bs] = np.asarray(l)
            else:
                for k in range(4):
                    if frame_types[f] == 0:
                        mv[f, b, k, 0] = m[k]
                    else:
                        mv[f, b, k] = m[k]
                    yy, xx = y + (k // 2) * sub, x + (k % 2) * sub
                    lev[f, yy:yy + sub, xx:xx + sub] = np.asarray(l[k])
    return split, mv, lev


def arrays_to_package(frame_types, split, mv, lev, bs):
    """Inverse of :func:`package_to_arrays` (plain Python ints / int ndarrays like the reference produces)."""
    F, H, W = lev.shape
    nbx = W // bs
    sub = bs // 2
    mvs_per_frame, levels_per_frame = [], []
    for f in range(F):
        fm, fl = [], []
        for b in range(split.shape[1]):
            y, x = (b // nbx) * bs, (b % nbx) * bs
            if split[f, b] == 0:
                if frame_types[f] == 0:
                    fm.append((0, int(mv[f, b, 0, 0])))
                else:
                    fm.append((0, tuple(int(v) for v in mv[f, b, 0])))
                fl.append((0, lev[f, y:y + bs, x:x + bs].astype(np.int64)))
            else:
                if frame_types[f] == 0:
                    fm.append((1, [int(mv[f, b, k, 0]) for k in range(4)]))
                else:
                    fm.append((1, [tuple(int(v) for v in mv[f, b, k]) for k in range(4)]))
                fl.append((1, [lev[f, y + (k // 2) * sub:y + (k // 2) * sub + sub,
                                   x + (k % 2) * sub:x + (k % 2) * sub + sub].astype(np.int64) for k in range(4)]))
        mvs_per_frame.append(fm)
        levels_per_frame.append(fl)
    return mvs_per_frame, levels_per_frame
