"""The pruned exhaustive search (SO_FLAG_SEA / ``Y_Video_codec.sea_prune``, csrc/so_me_sea.cuh) must give exactly the results of
the plain exhaustive search of find_best_match (/root/reference/Encoder.py:678-717): it only skips candidates whose SAD
lower bound exceeds the exact SAD of a predictor candidate.  Checked against the reference-generated strip golden, the CPU
oracle, and the plain search kernel on full-size frames, tie-heavy input, scene cuts and batched units."""
import numpy as np
import pytest

from streamoptima_b200 import synth
from tests.golden_util import load_case

pytestmark = pytest.mark.gpu
KEYS = ("split", "mv", "levels", "recon")


def _encode(frames, sea, bs=16, r=16, qp=4, intra_dur=8, **kw):
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    F, H, W = frames.shape
    c = Y_Video_codec(H, W, F, bs, r, qp, intra_dur, 0, y_only_frame_arr=frames, **kw)
    c.sea_prune = sea
    c.encode()
    p = {k: np.array(c.encoded_package.packed[k]) for k in KEYS}
    stats = c._ctx.sea_stats()
    return p, stats


def test_strip_golden_from_the_reference():
    """1920x128, i=16, r=16 half-pel, nRefFrames=4: output of the unmodified reference (oracle/gen_golden.py)."""
    frames, enc, g = load_case("w1920_fme_nref4")
    e = dict(enc)
    p, st = _encode(frames, True, bs=e.pop("block_size"), r=e.pop("search_range"), qp=e.pop("Qp"), intra_dur=e.pop("intra_dur"), **e)
    for k in KEYS:
        np.testing.assert_array_equal(p[k], g[k], err_msg=k)
    assert st["p_frames"] > 0 and st["exact_sads"] > 0


@pytest.mark.parametrize("kw", [dict(nRefFrames=2, FMEEnable=True), dict(nRefFrames=1), dict(nRefFrames=3, FMEEnable=True, ParallelMode=2)],
                         ids=["fme_nref2", "int_nref1", "pm2_fme_nref3"])
def test_small_frames_match_oracle(kw):
    from oracle import codec_oracle as co
    from oracle.packing import package_to_arrays
    F, H, W = 4, 96, 128
    frames = synth.translating(F, H, W, seed=11, bright=bool(kw.get("FMEEnable")))
    p, st = _encode(frames, True, qp=3, **kw)
    o = co.OracleCodec(H, W, F, 16, 16, 3, 8, 0, y_only_frame_arr=frames, **kw).encode()
    split, mv, lev = package_to_arrays(o["frame_types"], o["mvs"], o["levels"], H, W, 16)
    np.testing.assert_array_equal(p["split"], split)
    np.testing.assert_array_equal(p["mv"], mv)
    np.testing.assert_array_equal(p["levels"], lev)
    np.testing.assert_array_equal(p["recon"], o["recon"])
    assert st["p_frames"] > 0


@pytest.mark.parametrize("kind,F,H,W,kw", [
    ("translating", 7, 1088, 1920, dict(nRefFrames=4, FMEEnable=True)),          # the bench geometry (BASELINE configs[1])
    ("zooming", 4, 544, 976, dict(nRefFrames=1)),                                 # integer search, odd number of block columns
    ("zooming", 4, 160, 1936, dict(nRefFrames=2, FMEEnable=True)),
    ("flat_ties", 5, 272, 400, dict(nRefFrames=3, FMEEnable=True)),               # almost every SAD ties: the argmin order
    ("flat_ties", 4, 272, 400, dict(nRefFrames=2)),
    ("scene_cut", 5, 272, 640, dict(nRefFrames=2, FMEEnable=True)),               # no usable bound after the cut
    ("translating", 4, 2160, 3840, dict(nRefFrames=1)),                           # BASELINE configs[4] frame size
], ids=["c2_1080p", "int_odd_cols", "wide_fme", "ties_fme", "ties_int", "scene_cut", "c5_4k"])
def test_equals_plain_exhaustive_search(kind, F, H, W, kw):
    frames = synth.make(kind, F=F, H=H, W=W, seed=31)
    a, _ = _encode(frames, False, intra_dur=30, **kw)
    b, st = _encode(frames, True, intra_dur=30, **kw)
    for k in KEYS:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert st["p_frames"] == F - 1
    # the point of the exercise: far fewer exact SADs than candidates (33 x 33 offsets per block, reference and phase plane)
    if kind == "translating":
        nph = 4 if kw.get("FMEEnable") else 1
        cand = (H // 16) * (W // 16) * 1089 * nph * sum(min(f, kw["nRefFrames"]) for f in range(1, F))
        assert st["exact_sads"] < 0.1 * cand, (st, cand)


def test_batched_units_equal_single_sequences():
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    U, F, H, W = 3, 4, 96, 256
    seqs = np.stack([synth.make(k, F=F, H=H, W=W, seed=80 + i) for i, k in enumerate(("translating", "zooming", "flat_ties"))])
    for kw in (dict(FMEEnable=True, nRefFrames=3), dict(nRefFrames=2)):
        c = Y_Video_codec(H, W, F, 16, 16, 3, 8, 0, **kw)
        c.sea_prune = True
        res = c.encode_arrays(seqs)
        for u in range(U):
            one, _ = _encode(seqs[u], False, qp=3, **kw)
            for k in KEYS:
                np.testing.assert_array_equal(np.asarray(res[k])[u], one[k], err_msg=f"unit {u} {k}")


def test_per_frame_seam_two_units_one_frame_apart():
    """The pruned search through the per-frame C ABI (so_encode_inter with SO_FLAG_SEA), teacher-forced against the reference strip:
    unit 1 runs one frame behind unit 0, so the predictor memory and the quadrant planes of the two chains differ at every call."""
    from tests.test_gpu_seam import Seam, _assert_frame
    frames, enc, g = load_case("w1920_fme_nref4")
    F, H, W = frames.shape
    s = Seam(H, W, enc, max_batch=2, sea=True)
    s.reset(0)
    s.reset(1)
    for step in range(F + 1):
        for unit, f in ((0, step), (1, step - 1)):
            if not 0 <= f < F:
                continue
            out = s.encode(unit, frames[f], int(g["frame_types"][f]) == 0)
            _assert_frame(out, g, f)
            s.push(unit, g["recon"][f])
    assert s.ctx.sea_stats()["p_frames"] > 0
    s.ctx.close()


@pytest.mark.parametrize("kind", ["translating", "zooming", "scene_cut"])
def test_auto_mode_switches_between_the_two_searches_without_changing_the_output(kind):
    """``sea_prune = "auto"`` (SO_FLAG_SEA_AUTO) runs the plain kernel while pruning does not pay: whatever it decides per frame,
    the output is the plain search's."""
    F, H, W = 48, 272, 640
    frames = synth.make(kind, F=F, H=H, W=W, seed=17)
    kw = dict(nRefFrames=2, FMEEnable=True)
    a, _ = _encode(frames, False, intra_dur=16, **kw)
    b, st = _encode(frames, "auto", intra_dur=16, **kw)
    for k in KEYS:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert 0 < st["p_frames"] <= F - 3
