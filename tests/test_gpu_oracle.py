"""GPU parity against the CPU oracle on seeded inputs, at configurations the goldens do not cover -- in particular
search range 16 with 16x16 blocks, which is the DIRECT TMA staging path of the exhaustive search (BASELINE configs 2-5)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import codec_oracle as co
from oracle.packing import package_to_arrays
from streamoptima_b200 import synth

pytestmark = pytest.mark.gpu

CASES = {
    "r16_int": dict(gen=("translating", dict(F=3, H=96, W=128, seed=41)),
                    enc=dict(block_size=16, search_range=16, Qp=3, intra_dur=8)),
    "r16_fme_nref2": dict(gen=("zooming", dict(F=4, H=96, W=128, seed=42)),
                          enc=dict(block_size=16, search_range=16, Qp=2, intra_dur=8, FMEEnable=True, nRefFrames=2)),
    "r16_fme_nref4_ties": dict(gen=("flat_ties", dict(F=6, H=64, W=96, seed=43)),
                               enc=dict(block_size=16, search_range=16, Qp=0, intra_dur=8, FMEEnable=True, nRefFrames=4)),
    "r16_vbs_fme_bright": dict(gen=("translating", dict(F=3, H=96, W=128, seed=44, bright=True)),
                               enc=dict(block_size=16, search_range=16, Qp=4, intra_dur=8, FMEEnable=True, nRefFrames=2,
                                        VBSEnable=True, lam=0.02)),
    "r16_vbs_int": dict(gen=("zooming", dict(F=3, H=96, W=128, seed=50)),
                        enc=dict(block_size=16, search_range=16, Qp=2, intra_dur=8, VBSEnable=True, lam=0.01)),
    "r16_vbs_fme_nref1_ties": dict(gen=("flat_ties", dict(F=3, H=96, W=128, seed=51)),
                                   enc=dict(block_size=16, search_range=16, Qp=1, intra_dur=8, VBSEnable=True, lam=0.05, FMEEnable=True)),
    "r16_vbs_fme_nref3_rc": dict(gen=("translating", dict(F=5, H=96, W=128, seed=52)),
                                 enc=dict(block_size=16, search_range=16, Qp=3, intra_dur=8, VBSEnable=True, lam=0.02, FMEEnable=True,
                                          nRefFrames=3, RCFlag=1, targetBR="900 kbps",
                                          qp_rate_tables=[[9000, 7000, 5200, 3900, 2800, 1900, 1300, 900, 600, 400, 250, 100],
                                                          [6000, 4600, 3400, 2500, 1800, 1200, 800, 560, 380, 250, 160, 60]])),
    "r32_fme_vbs_chunked": dict(gen=("translating", dict(F=3, H=96, W=128, seed=53)),
                                enc=dict(block_size=16, search_range=32, Qp=2, intra_dur=8, FMEEnable=True, nRefFrames=2,
                                         VBSEnable=True, lam=0.02)),
    "r48_int_chunked": dict(gen=("flat_ties", dict(F=3, H=96, W=160, seed=54)),
                            enc=dict(block_size=16, search_range=48, Qp=1, intra_dur=8)),
    "r20_i8_chunked": dict(gen=("zooming", dict(F=3, H=64, W=96, seed=55)),
                           enc=dict(block_size=8, search_range=20, Qp=3, intra_dur=8, FMEEnable=True)),
    "r17_i16_vbs_chunked": dict(gen=("translating", dict(F=3, H=64, W=96, seed=56)),
                                enc=dict(block_size=16, search_range=17, Qp=2, intra_dur=8, VBSEnable=True, lam=0.02)),
    "r8_i8_fme": dict(gen=("translating", dict(F=3, H=64, W=96, seed=45)),
                      enc=dict(block_size=8, search_range=8, Qp=1, intra_dur=8, FMEEnable=True)),
    "r5_i16": dict(gen=("zooming", dict(F=3, H=64, W=96, seed=46)),
                   enc=dict(block_size=16, search_range=5, Qp=5, intra_dur=8, nRefFrames=2)),
    "r7_i8_vbs": dict(gen=("translating", dict(F=3, H=64, W=96, seed=47)),
                      enc=dict(block_size=8, search_range=7, Qp=2, intra_dur=8, VBSEnable=True, lam=0.03)),
    "w_not_mult16": dict(gen=("translating", dict(F=3, H=48, W=72, seed=48)),
                         enc=dict(block_size=8, search_range=3, Qp=2, intra_dur=8, FMEEnable=True, nRefFrames=2)),
    "r0": dict(gen=("translating", dict(F=3, H=32, W=48, seed=49)),
               enc=dict(block_size=8, search_range=0, Qp=2, intra_dur=8)),
}


def _encode_gpu(frames, enc):
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    F, H, W = frames.shape
    e = dict(enc)
    c = Y_Video_codec(H, W, F, e.pop("block_size"), e.pop("search_range"), e.pop("Qp"), e.pop("intra_dur"), 0,
                      y_only_frame_arr=frames, **e)
    psnr = c.encode()
    return c, psnr


@pytest.mark.parametrize("name", list(CASES))
def test_matches_oracle(name):
    kind, gkw = CASES[name]["gen"]
    enc = CASES[name]["enc"]
    frames = synth.make(kind, **gkw)
    F, H, W = frames.shape
    c, psnr = _encode_gpu(frames, enc)
    p = c.encoded_package.packed
    o = co.OracleCodec(H, W, F, y_only_frame_arr=frames, **enc).encode()
    split, mv, lev = package_to_arrays(o["frame_types"], o["mvs"], o["levels"], H, W, enc["block_size"])
    assert c.encoded_package["frame_type_seq"] == o["frame_types"]
    np.testing.assert_array_equal(p["split"], split)
    np.testing.assert_array_equal(p["mv"], mv)
    np.testing.assert_array_equal(p["levels"], lev)
    np.testing.assert_array_equal(p["recon"], o["recon"])
    np.testing.assert_allclose(psnr, o["psnr"], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(p["qsize"], np.asarray(o["qsize"], np.uint32))
    np.testing.assert_array_equal(p["row_sizes"], np.asarray(o["row_sizes"], np.uint32))


_SCRIPT = r"""
import sys, hashlib, numpy as np
sys.path.insert(0, %r)
from streamoptima_b200 import synth
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
F, H, W = 4, 1088, 1920
frames = synth.translating(F, H, W, seed=3)
c = Y_Video_codec(H, W, F, 16, 16, 4, 30, 0, nRefFrames=4, FMEEnable=True, y_only_frame_arr=frames)
c.encode()
p = c.encoded_package.packed
print(hashlib.sha256(p["mv"].tobytes() + p["levels"].tobytes() + p["recon"].tobytes()).hexdigest())
"""


def test_full_size_direct_equals_expand_staging():
    """BASELINE config 2 geometry (1920x1088, i=16, r=16, half-pel, 4 refs): the DIRECT TMA staging used by the bench
    and the EXPAND staging (the one the small goldens exercise) must produce identical MVs, levels and reconstruction."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for extra in ({}, {"SO_ME_NO_DIRECT": "1"}):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, "-c", _SCRIPT % root], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1]


_RING_SCRIPT = r"""
import sys, hashlib, numpy as np
sys.path.insert(0, %r)
from streamoptima_b200 import synth
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
h = hashlib.sha256()
for (F, H, W, kw) in ((6, 1088, 1920, dict(nRefFrames=4, FMEEnable=True)),
                      (4, 1088, 1920, dict(nRefFrames=3, FMEEnable=True, VBSEnable=True, lam=0.02)),
                      (3, 544, 976, dict(nRefFrames=1)),                                  # integer search, odd number of block columns
                      (4, 272, 400, dict(nRefFrames=3, VBSEnable=True, lam=0.02)),
                      (4, 160, 1936, dict(nRefFrames=2, FMEEnable=True))):
    frames = synth.zooming(F, H, W, seed=21)
    c = Y_Video_codec(H, W, F, 16, 16, 3, 30, 0, y_only_frame_arr=frames, **kw)
    c.encode()
    p = c.encoded_package.packed
    for k in ("split", "mv", "levels", "recon"):
        h.update(np.ascontiguousarray(p[k]).tobytes())
print(h.hexdigest())
"""


def test_ring_kernel_v2_equals_v1():
    """The chunked item-ring search kernel (homogeneous chunks, only the valid vertical groups of edge rows, per-item merge with
    match + REDUX) against the first item-ring kernel (SO_ME_RING_V1), which the strip goldens and the oracle tests pinned
    in round 1: 1080p half-pel with 4 references, fused VBS search, integer search, widths whose pair count is not a
    multiple of the chunk size."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for extra in ({}, {"SO_ME_RING_V1": "1"}):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, "-c", _RING_SCRIPT % root], capture_output=True, text=True, env=env, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1]


def test_full_size_crop_matches_oracle_interior():
    """1080p P frame against the oracle on a crop: blocks whose whole search window lies inside the crop must get the
    same motion vector when the crop is encoded on its own with the same reference pixels (integer search, 1 ref)."""
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    F, H, W = 2, 1088, 1920
    frames = synth.translating(F, H, W, seed=5)
    c = Y_Video_codec(H, W, F, 16, 16, 4, 30, 0, y_only_frame_arr=frames)
    c.encode()
    p = c.encoded_package.packed
    recon0 = p["recon"][0]
    # oracle full search of frame 1 against the GPU's reconstruction of frame 0 on a 256x160 crop at (512, 320)
    y0, x0, ch, cw = 320, 512, 160, 256
    mv, sad = co.full_search(frames[1, y0:y0 + ch, x0:x0 + cw].astype(np.int64), [recon0[y0:y0 + ch, x0:x0 + cw]], 16, 16, False)
    nbx = W // 16
    for by in range(1, ch // 16 - 2):
        for bx in range(1, cw // 16 - 2):
            gb = (y0 // 16 + by) * nbx + (x0 // 16 + bx)
            got = tuple(int(v) for v in p["mv"][1, gb, 0])
            assert got == tuple(int(v) for v in mv[by, bx]), (by, bx)


def test_batched_units_equal_single_sequences_r16():
    """Three independent sequences in one context (batched launches: the item-ring search kernel indexes units through
    the plane / output offsets) must equal the three sequences encoded one by one; half-pel and integer search."""
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    U, F, H, W = 3, 4, 96, 128
    seqs = np.stack([synth.make(k, F=F, H=H, W=W, seed=70 + i) for i, k in enumerate(("translating", "zooming", "flat_ties"))])
    for bs, kw in ((16, dict(FMEEnable=True, nRefFrames=3)), (16, dict(nRefFrames=2)),
                   (16, dict(fast_me=True, FMEEnable=True, VBSEnable=True, lam=0.02, nRefFrames=2)),      # table-driven fast-ME chain,
                   (8, dict(fast_me=True, FMEEnable=True, nRefFrames=3))):                                # one walker per unit
        cb = Y_Video_codec(H, W, F, bs, 16, 2, 8, 0, **kw)
        ob = cb.encode_arrays(seqs)
        got = {k: np.array(ob[k]) for k in ("split", "mv", "levels", "recon", "row_sizes")}
        ob2 = cb.encode_arrays(seqs)                    # again: the fast-ME tables are now centred on the previous run's predictors
        for k in got:
            np.testing.assert_array_equal(got[k], ob2[k], err_msg=f"second run {k} {kw}")
        for u in range(U):
            cs = Y_Video_codec(H, W, F, bs, 16, 2, 8, 0, **kw)
            o1 = cs.encode_arrays(seqs[u])
            for k in got:
                np.testing.assert_array_equal(got[k][u], o1[k][0], err_msg=f"{k} unit {u} {kw}")


_INTRA_SCRIPT = r"""
import sys, hashlib, numpy as np
sys.path.insert(0, %r)
from streamoptima_b200 import synth
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
h = hashlib.sha256()
for (F, H, W, r, kw) in ((3, 272, 480, 16, dict(VBSEnable=True, lam=0.02)), (2, 96, 256, 40, dict(VBSEnable=True, lam=0.01)), (2, 1088, 1920, 16, {})):
    frames = synth.zooming(F, H, W, seed=9)
    c = Y_Video_codec(H, W, F, 16, r, 3, 1, 0, y_only_frame_arr=frames, **kw)      # I_Period 1: every frame is intra
    c.encode()
    p = c.encoded_package.packed
    for k in ("split", "mv", "levels", "recon"):
        h.update(np.ascontiguousarray(p[k]).tobytes())
print(h.hexdigest())
"""


def test_intra_fast16_kernels_equal_generic():
    """The warp-per-block intra search / shared-memory row chain for 16x16 blocks against the generic intra kernels
    (which the goldens pin at every block size): all-intra sequences with VBS, a range above 32 (two candidate chunks) and a
    full 1080p frame."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for extra in ({}, {"SO_INTRA_GENERIC": "1"}):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, "-c", _INTRA_SCRIPT % root], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1]


_FAST_SCRIPT = r"""
import sys, hashlib, numpy as np
sys.path.insert(0, %r)
from streamoptima_b200 import synth
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
h = hashlib.sha256()
for (F, H, W, bs, kw) in ((4, 272, 480, 16, dict(fast_me=True, FMEEnable=True, nRefFrames=4, VBSEnable=True, lam=0.015)),
                          (3, 144, 176, 16, dict(fast_me=True, nRefFrames=2)),
                          (3, 272, 480, 16, dict(fast_me=True, FMEEnable=True, ParallelMode=2)),
                          (3, 1088, 1920, 16, dict(fast_me=True, FMEEnable=True, nRefFrames=2)),
                          (5, 288, 352, 8, dict(fast_me=True, FMEEnable=True, nRefFrames=3, VBSEnable=True, lam=0.02)),
                          (4, 144, 176, 8, dict(fast_me=True, nRefFrames=8)),
                          (3, 288, 352, 8, dict(fast_me=True, FMEEnable=True, ParallelMode=2))):
    frames = synth.zooming(F, H, W, seed=12)
    c = Y_Video_codec(H, W, F, bs, 16, 4, 8, 0, y_only_frame_arr=frames, **kw)
    c.encode()
    p = c.encoded_package.packed
    for k in ("split", "mv", "levels", "recon"):
        h.update(np.ascontiguousarray(p[k]).tobytes())
# a scene cut in a multi-chunk frame: the predictors leave the tables' windows in mid-sequence (block-by-block walk of those chunks)
for (F, H, W, bs, kw) in ((6, 544, 960, 16, dict(fast_me=True, FMEEnable=True, nRefFrames=2)),
                          (5, 272, 480, 8, dict(fast_me=True, nRefFrames=1, VBSEnable=True, lam=0.02))):
    frames = synth.make("scene_cut", F=F, H=H, W=W, seed=31, cut_at=3)
    c = Y_Video_codec(H, W, F, bs, 16, 3, 30, 0, y_only_frame_arr=frames, **kw)
    c.encode()
    p = c.encoded_package.packed
    for k in ("split", "mv", "levels", "recon"):
        h.update(np.ascontiguousarray(p[k]).tobytes())
print(h.hexdigest())
"""


def test_fast_me16_kernel_equals_generic():
    """Fast-ME pipeline for 16x16 and 8x8 blocks (transition tables, the chain as a scan over composed tables, cooperative
    step, results from the scan or from the parallel search kernel) against the generic chain kernel that the goldens pin:
    half-pel + up to 8 refs + VBS, integer, ParallelMode 2, a 1080p chain and scene cuts in multi-chunk frames.  The
    one-warp block-by-block walker (SO_FAST_CHAIN_WALK) and the always-search variant (SO_FAST_ALWAYS_ME16) must agree too."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for extra in ({}, {"SO_FAST_GENERIC": "1"}, {"SO_FAST_CHAIN_WALK": "1"}, {"SO_FAST_ALWAYS_ME16": "1"}):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, "-c", _FAST_SCRIPT % root], capture_output=True, text=True, env=env, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert len(set(outs)) == 1, outs


_SIMPLE_SCRIPT = r"""
import sys, hashlib, numpy as np
sys.path.insert(0, %r)
from streamoptima_b200 import synth
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
h = hashlib.sha256()
for (F, H, W, bs, r, kw) in ((4, 96, 128, 16, 16, dict(FMEEnable=True, nRefFrames=3, VBSEnable=True, lam=0.02)),
                             (3, 96, 128, 16, 16, dict(nRefFrames=2)),
                             (3, 64, 96, 8, 5, dict(FMEEnable=True, VBSEnable=True, lam=0.03)),
                             (3, 64, 96, 16, 24, dict(FMEEnable=True))):
    frames = synth.flat_ties(F, H, W, seed=14)
    c = Y_Video_codec(H, W, F, bs, r, 2, 8, 0, y_only_frame_arr=frames, **kw)
    c.encode()
    p = c.encoded_package.packed
    for k in ("split", "mv", "levels", "recon"):
        h.update(np.ascontiguousarray(p[k]).tobytes())
print(h.hexdigest())
"""


def test_packed_search_kernels_equal_plain_search():
    """The word-packed TMA search kernels (item ring, stage-based, fused VBS, search chunks) against the plain
    one-warp-per-block search (me_simple_kernel, SO_ME_SIMPLE=1) on tie-heavy input: identical vectors, levels, frames."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for extra in ({}, {"SO_ME_SIMPLE": "1"}):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, "-c", _SIMPLE_SCRIPT % root], capture_output=True, text=True, env=env, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1]
