"""CPU models of two arithmetic claims the exhaustive-search kernels rest on (no GPU needed):

* the 32-bit / compact 64-bit keys of csrc/so_me_ring2.cuh order candidates exactly like the reference's sequential replace rule
  (/root/reference/Encoder.py:688-715 with ``is_better_mv`` :771: smaller SAD, then smaller |dx|+|dy|, then smaller reference index,
  otherwise the candidate visited first in the scan order ref, dx, dy -- i.e. the lexicographic minimum of (SAD, L1, ref, dx, dy),
  SURVEY.md appendix A4) and decode back to (SAD, ref, dx, dy) (``me_get`` format 2 in csrc/so_kernels.cuh);
* the lower bound of the pruned search (csrc/so_me_sea.cuh): SAD >= 64 * D - 252 with D the byte-wise distance of the quantised
  8x8 quadrant sums, hence a candidate with D > (U + 252) >> 6 cannot have SAD <= U.
"""
import itertools

import numpy as np
import pytest


def key32(sad, dx, dy, R):
    return (sad << 16) | ((abs(dx) + abs(dy)) << 8) | ((dx + R) << 1) | (1 if dy > 0 else 0)


def key64_compact(m, ref):
    """mr2_key64: bytes (low to high) m.b0, ref, m.b1, m.b2 | m.b3"""
    b = [(m >> (8 * i)) & 0xFF for i in range(4)]
    lo = b[0] | (ref << 8) | (b[1] << 16) | (b[2] << 24)
    return (b[3] << 32) | lo


def decode_compact(key, R):
    """me_get, packed == 2"""
    lo, hi = key & 0xFFFFFFFF, key >> 32
    l1 = (lo >> 16) & 0xFF
    dx = ((lo >> 1) & 0x7F) - R
    ady = l1 - abs(dx)
    return ((hi & 0xFF) << 8) | (lo >> 24), (lo >> 8) & 0xFF, dx, ady if lo & 1 else -ady


@pytest.mark.parametrize("R", [16, 32])
def test_key_order_equals_reference_tie_break(R):
    """Within one reference the winner is the lexicographic minimum of (SAD, |dx|+|dy|, dx, dy).  For equal (L1, dx) only the sign
    of dy is left to decide, which is what the key's last bit stores."""
    cands = [(dx, dy) for dx in range(-R, R + 1) for dy in range(-R, R + 1)]
    by_ref_rule = sorted(cands, key=lambda c: (abs(c[0]) + abs(c[1]), c[0], c[1]))
    by_key = sorted(cands, key=lambda c: key32(1234, c[0], c[1], R))
    assert by_ref_rule == by_key
    keys = [key32(1234, dx, dy, R) for dx, dy in cands]
    assert len(set(keys)) == len(keys) and max(keys) < 0xFFFFFFFF


def test_key_low_byte_is_additive():
    """The fold adds `dx * -254` (dx < 0) or `dx * 258` (dx >= 0) and the vertical addend `|dy| << 8 | 2R + (dy > 0)` to SAD << 16."""
    for R in (16, 32):
        for dx, dy in itertools.product(range(-R, R + 1), repeat=2):
            horiz = (dx * (-254 if dx < 0 else 258)) & 0xFFFFFFFF
            vert = (abs(dy) << 8) + 2 * R + (1 if dy > 0 else 0)
            assert ((777 << 16) + horiz + vert) & 0xFFFFFFFF == key32(777, dx, dy, R)


def test_compact_key64_orders_by_reference_and_round_trips():
    rng = np.random.default_rng(0)
    R = 32
    items = []
    for _ in range(4000):
        sad, ref = int(rng.integers(0, 65281)), int(rng.integers(0, 8))
        dx, dy = int(rng.integers(-R, R + 1)), int(rng.integers(-R, R + 1))
        k = key64_compact(key32(sad, dx, dy, R), ref)
        assert k < (1 << 64) - 1
        assert decode_compact(k, R) == (sad, ref, dx, dy)
        items.append(((sad, abs(dx) + abs(dy), ref, dx, dy), k))
    assert [t for t, _ in sorted(items)] == [t for t, _ in sorted(items, key=lambda it: it[1])]


def _quadrant_bytes(block):
    return np.array([int(block[y:y + 8, x:x + 8].sum()) >> 6 for y in (0, 8) for x in (0, 8)])


@pytest.mark.parametrize("seed", range(4))
def test_quadrant_bound_never_exceeds_the_sad(seed):
    rng = np.random.default_rng(seed)
    for _ in range(3000):
        kind = rng.integers(0, 3)
        if kind == 0:
            a, b = rng.integers(0, 256, (16, 16)), rng.integers(0, 256, (16, 16))
        elif kind == 1:          # nearly equal blocks: the regime where candidates survive
            a = rng.integers(0, 256, (16, 16))
            b = np.clip(a + rng.integers(-3, 4, (16, 16)), 0, 255)
        else:                    # flat blocks with an offset: the bound is tight up to the quantisation slack
            a = np.full((16, 16), int(rng.integers(0, 256)))
            b = np.full((16, 16), int(rng.integers(0, 256)))
        sad = int(np.abs(a - b).sum())
        d = int(np.abs(_quadrant_bytes(a) - _quadrant_bytes(b)).sum())
        assert 64 * d - 252 <= sad
        # the filter: with U the exact SAD of any candidate, everything with SAD <= U passes D <= (U + 252) >> 6
        for u in (sad, sad + 1, sad + 63, sad + 1000):
            assert d <= (u + 252) >> 6
