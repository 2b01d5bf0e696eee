"""Host side of the packed symbol streams (no GPU): the streams the sequence encodes deliver instead of raw levels
(``so_set_symbol_output``) are formatted into the reference's residual text, written as the two bitstream files and
turned back into levels -- all pinned to the reference's own outputs (tests/golden), with corrupt streams rejected."""
import os

import numpy as np
import pytest

from oracle import codec_oracle as co
from streamoptima_b200 import _native
from tests.golden_util import case_names, load_case


def _lib():
    from streamoptima_b200 import build
    build.build()
    return _native.load()


def pack_symbols(split, levels, bs):
    """Frame-by-frame packed streams from golden levels with the ORACLE's entropy_encoder_block (Encoder.py:1086-1131)."""
    F, H, W = levels.shape
    nbx, sub = W // bs, bs // 2
    sym, pos, cnt = [], np.zeros(F, np.uint64), np.zeros(F, np.uint32)
    for f in range(F):
        pos[f] = len(sym)
        for b in range(split.shape[1]):
            y, x = (b // nbx) * bs, (b % nbx) * bs
            if split[f, b] == 0:
                sym += co.rle_symbols(levels[f, y:y + bs, x:x + bs])
            else:
                for k in range(4):
                    yy, xx = y + (k // 2) * sub, x + (k % 2) * sub
                    sym += co.rle_symbols(levels[f, yy:yy + sub, xx:xx + sub])
        cnt[f] = len(sym) - int(pos[f])
    return np.asarray(sym if sym else [0], np.int16), pos, cnt


@pytest.mark.parametrize("name", [n for n in case_names() if not n.startswith("w1920")])
def test_packed_symbols_to_text_files_and_levels(name, tmp_path):
    lib = _lib()
    frames, enc, g = load_case(name)
    F, H, W = frames.shape
    bs = enc["block_size"]
    split = np.ascontiguousarray(g["split"], np.uint8)
    sym, pos, cnt = pack_symbols(split, g["levels"], bs)
    nblk = split.shape[1]
    # per-frame text
    res_lines = g["res_text"].splitlines()
    for f in range(F):
        cap = 1 << 22
        buf = bytearray(cap)
        sy = np.ascontiguousarray(sym[int(pos[f]):int(pos[f]) + int(cnt[f])]) if cnt[f] else np.zeros(1, np.int16)
        n = lib.so_format_residual_frame_packed(split[f].ctypes.data, sy.ctypes.data, int(cnt[f]), nblk, bs,
                                                (_native.C.c_char * cap).from_buffer(buf), cap)
        assert n >= 0 and bytes(buf[:n]).decode() == res_lines[f]
    # inverse RLE of the whole sequence
    lev = np.full((F, H, W), 77, np.int16)
    for threads in (1, 3):
        assert lib.so_symbols_to_levels(split.ctypes.data, sym.ctypes.data, pos.ctypes.data, cnt.ctypes.data, F, W, H, bs,
                                        lev.ctypes.data, threads) == 0
        np.testing.assert_array_equal(lev, g["levels"])
    # both files
    ft = np.ascontiguousarray(g["frame_types"], np.uint8)
    mv = np.ascontiguousarray(g["mv"], np.int16)
    rc_on = (g["qp_rows"] >= 0).any()
    qp = np.ascontiguousarray(g["qp_rows"], np.int32) if rc_on else None
    mvf, rsf = tmp_path / "mv.txt", tmp_path / "res.txt"
    rc = lib.so_write_bitstream_files_symbols(ft.ctypes.data, split.ctypes.data, mv.ctypes.data, sym.ctypes.data, pos.ctypes.data,
                                              cnt.ctypes.data, qp.ctypes.data if qp is not None else None, F, W, H, bs,
                                              os.fsencode(mvf), os.fsencode(rsf), 2)
    assert rc == 0
    assert open(mvf).read() == g["mv_text"]
    assert open(rsf).read() == g["res_text"]


def test_corrupt_symbol_streams_are_rejected():
    lib = _lib()
    frames, enc, g = load_case("s_vbs")
    F, H, W = frames.shape
    bs = enc["block_size"]
    split = np.ascontiguousarray(g["split"], np.uint8)
    sym, pos, cnt = pack_symbols(split, g["levels"], bs)
    lev = np.zeros((F, H, W), np.int16)

    def to_levels(s, p, c):
        return lib.so_symbols_to_levels(split.ctypes.data, s.ctypes.data, p.ctypes.data, c.ctypes.data, F, W, H, bs, lev.ctypes.data, 1)

    assert to_levels(sym, pos, cnt) == 0
    short = cnt.copy(); short[1] -= 1                       # a frame that ends inside a list
    assert to_levels(sym, pos, short) == -1
    long_ = cnt.copy(); long_[0] += 1                       # symbols left over after the last block
    assert to_levels(sym, pos, long_) == -1
    bad = sym.copy(); bad[int(pos[1])] = bs * bs + 5        # a zero run longer than the block
    assert to_levels(bad, pos, cnt) == -1
    bad = sym.copy(); bad[int(pos[1])] = -(bs * bs + 5)     # a non-zero run longer than the block
    assert to_levels(bad, pos, cnt) == -1
    rng = np.random.default_rng(3)                          # mutation fuzz: any return code, no crash / out-of-bounds write
    for _ in range(300):
        bad = sym.copy()
        for i in rng.integers(0, bad.size, 3):
            bad[i] = rng.integers(-300, 300)
        assert to_levels(bad, pos, cnt) in (0, -1)


def test_results_own_their_buffers():
    """EncodeResult semantics without a GPU: buffers are recycled only when nothing references them any more."""
    from streamoptima_b200.Encoder import EncodeResult, _Lease, _PinnedPool
    pool = _PinnedPool()

    def make(fill):
        lease = _Lease(pool)
        r = EncodeResult(a=lease.array((4, 8), np.int16), b=lease.array((16,), np.uint8))
        r.lease = lease
        r["a"][:] = fill
        return r

    r1 = make(5)
    keep = r1["a"][1]                 # a view survives the result
    ptr_a, ptr_b = r1["a"].ctypes.data, r1["b"].ctypes.data
    del r1
    r2 = make(9)                      # b's buffer is recycled, a's is not: `keep` still reads it
    assert r2["b"].ctypes.data == ptr_b and r2["a"].ctypes.data != ptr_a
    assert (keep == 5).all()
    del keep, r2
    r3 = make(1)                      # now everything is reusable
    assert {r3["a"].ctypes.data, r3["b"].ctypes.data} <= {ptr_a, ptr_b} | {r3["a"].ctypes.data}
    assert len(pool.pending) == 0
