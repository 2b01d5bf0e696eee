"""Text-stream parser (CPU) and GPU decoder round trips."""
import os
import tempfile

import numpy as np
import pytest

from oracle.golden_cases import CASES
from streamoptima_b200 import decoder as dec
from tests.golden_util import case_names, load_case


def _write(tmp, g):
    mvf, rsf = os.path.join(tmp, "mv.txt"), os.path.join(tmp, "res.txt")
    open(mvf, "w").write(g["mv_text"])
    open(rsf, "w").write(g["res_text"])
    return mvf, rsf


def _decoder(frames, enc):
    F, H, W = frames.shape
    return dec.decoder(0, enc["intra_dur"], enc["block_size"], F, H, W, enc["Qp"], enc.get("nRefFrames", 1),
                       enc.get("FMEEnable", False), enc.get("lam"), enc.get("VBSEnable", False), False, enc.get("RCFlag"),
                       enc.get("targetBR"), 30, enc.get("qp_rate_tables"), ParallelMode=enc.get("ParallelMode", 0))


@pytest.mark.parametrize("name", case_names())
def test_parser_recovers_packed_arrays_from_reference_text(name):
    frames, enc, g = load_case(name)
    d = _decoder(frames, enc)
    with tempfile.TemporaryDirectory() as tmp:
        mvf, rsf = _write(tmp, g)
        ft, split, mv, lev, qps = d.parse_bitstream(mvf, rsf)                 # C++ parser (host threads)
        ft2, split2, mv2, lev2, qps2 = d.parse_bitstream_py(mvf, rsf)           # Python restatement
        with pytest.raises(ValueError):
            d.parse_bitstream(mvf, rsf, frames=len(g["frame_types"]) + 1)       # fewer lines than frames
    for a, b in ((ft, ft2), (split, split2), (mv, mv2), (lev, lev2)):
        np.testing.assert_array_equal(a, b)
    assert [list(q) for q in qps] == [list(q) for q in qps2]
    np.testing.assert_array_equal(ft, g["frame_types"])
    np.testing.assert_array_equal(split, g["split"])
    np.testing.assert_array_equal(mv, g["mv"])
    np.testing.assert_array_equal(lev, g["levels"])
    for f, q in enumerate(qps):
        assert list(q) == [int(v) for v in g["qp_rows"][f] if v >= 0]


@pytest.mark.gpu
@pytest.mark.parametrize("name", case_names())
def test_gpu_decoder_reproduces_reference_reconstruction(name):
    """The reference text streams decode to the reference encoder's reconstruction.  For nRefFrames > 1 the reference's
    own decoder resets its list at I frames and then indexes out of range (quirk Q7), so there the encoder's list
    semantics are used."""
    frames, enc, g = load_case(name)
    d = _decoder(frames, enc)
    with tempfile.TemporaryDirectory() as tmp:
        mvf, rsf = _write(tmp, g)
        ft, split, mv, lev, qps = d.parse_bitstream(mvf, rsf)
    nref = enc.get("nRefFrames", 1)
    out = d.decode_arrays(ft, split, mv, lev, qps if (enc.get("RCFlag") or 0) > 0 else None, reset_at_intra=(nref == 1))
    np.testing.assert_array_equal(out, g["recon"])


@pytest.mark.gpu
def test_full_size_encode_decode_round_trip():
    """1080p, config-2/3 style (i=16, r=16, half-pel, VBS, 2 refs): decode(encode(x)) equals the encoder's reconstruction."""
    from streamoptima_b200 import synth
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    F, H, W = 5, 1088, 1920
    frames = synth.zooming(F, H, W, seed=8)
    kw = dict(nRefFrames=2, FMEEnable=True, VBSEnable=True, lam=0.02)
    c = Y_Video_codec(H, W, F, 16, 16, 5, 4, 0, y_only_frame_arr=frames, **kw)
    c.encode()
    p = c.encoded_package.packed
    d = dec.decoder(0, 4, 16, F, H, W, 5, 2, True, 0.02, True)
    out = d.decode_arrays(p["frame_types"], p["split"], p["mv"], p["levels"], None, reset_at_intra=False)
    np.testing.assert_array_equal(out, p["recon"])
    psnr = [10 * np.log10(255.0 ** 2 / np.mean((out[i].astype(float) - frames[i]) ** 2)) for i in range(F)]
    np.testing.assert_allclose(psnr, c.encoded_package["PSNR per frame"], atol=1e-9)


@pytest.mark.parametrize("name", ["s_default", "s_vbs", "s_fme_bright", "s_rc1", "s_fast", "s_pm2", "s_intra_only_vbs"])
def test_reference_decoder_accepts_the_golden_text(name):
    """Format authority check, build container only: the UNCHANGED reference decoder.py parses the golden text streams
    (the format our C formatters emit byte-identically, tests/test_host_text.py) and reproduces the golden frames."""
    from oracle import reference_harness as rh
    if not rh.reference_available():
        pytest.skip("reference sources only exist in the build container")
    frames, enc, g = load_case(name)
    F, H, W = frames.shape
    out = rh.decode_with_reference(g["mv_text"].splitlines(), g["res_text"].splitlines(), H=H, W=W, F=F,
                                   block_size=enc["block_size"], Qp=enc["Qp"], intra_dur=enc["intra_dur"],
                                   nRefFrames=enc.get("nRefFrames", 1), FMEEnable=enc.get("FMEEnable", False), lam=enc.get("lam"),
                                   VBSEnable=enc.get("VBSEnable", False), RCFlag=enc.get("RCFlag"), targetBR=enc.get("targetBR"),
                                   qp_rate_tables=enc.get("qp_rate_tables"), ParallelMode=enc.get("ParallelMode", 0))
    np.testing.assert_array_equal(out, g["recon"])


@pytest.mark.gpu
def test_baseline_config2_full_length_round_trip_and_determinism():
    """BASELINE config 2 at full size and length (1920x1088, 300 frames, i=16, r=16 half-pel, 4 references, I_Period 30):
    size-independent properties the oracle cannot check in reasonable time -- decode(encode) reproduces the encoder's
    reconstruction bit for bit, two encodes are identical (the search merges with atomics: the result must not depend on
    their order), frame types follow I_Period, and the symbol streams account for exactly quantized_sized symbols."""
    import torch
    from bench import synth_frames_torch
    from streamoptima_b200 import decoder as dec
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    F, H, W = 300, 1088, 1920
    frames = synth_frames_torch(F, H, W, seed=7, device=torch.device("cuda", 0)).cpu().numpy()
    c = Y_Video_codec(H, W, F, 16, 16, 4, 30, 0, nRefFrames=4, FMEEnable=True)
    o1 = c.encode_arrays(frames)                                             # results own their arrays: o1 survives the
    o2 = c.encode_arrays(frames, want_levels=False, want_symbols=True)        # second encode on the same codec
    qsize = np.array(o2["stats"]["qsize"][0])
    np.testing.assert_array_equal(o1["mv"][0], o2["mv"][0])
    np.testing.assert_array_equal(o1["recon"][0], o2["recon"][0])
    # the second encode delivered packed symbols instead of raw levels: the levels rebuilt from them are the first encode's
    assert (o2["sym_count"][0] == qsize).all() and o2["sym_needed"] == int(qsize.sum())
    np.testing.assert_array_equal(o1["levels"][0], o2["levels"][0])
    assert [int(t) for t in o1["frame_types"][0]] == [0 if f % 30 == 0 else 1 for f in range(F)]
    assert (o1["row_sizes"][0].sum(axis=1) == qsize).all()
    offsets, symbols, base = c.symbol_streams()
    assert (np.diff(base.astype(np.int64)) == qsize).all() and symbols.size == int(qsize.sum())
    d = dec.decoder(0, 30, 16, F, H, W, 4, 4, True, None, False)
    out = d.decode_arrays(o1["frame_types"][0], o1["split"][0], o1["mv"][0], o1["levels"][0], None, reset_at_intra=False)
    np.testing.assert_array_equal(out, o1["recon"][0])
    mse = ((out.astype(np.float64) - frames) ** 2).mean(axis=(1, 2))
    assert (10 * np.log10(255.0 ** 2 / mse) > 30).all()                      # QP 4: a sane reconstruction on every frame


def test_parser_rejects_malformed_and_hostile_streams(tmp_path):
    """so_parse_bitstream_files must fail cleanly (ValueError), never index outside its arrays: run lengths larger than
    the block, absurd digit strings, bad frame types, vectors that do not fit int16, truncated lines -- plus a seeded
    mutation fuzz of a valid stream (any outcome but a crash / out-of-bounds write is fine)."""
    name = "s_vbs_fme_nref2"
    frames, enc, g = load_case(name)
    d = _decoder(frames, enc)
    mv_lines = g["mv_text"].splitlines()
    res_lines = g["res_text"].splitlines()

    def parse(mv, res):
        mvf, rsf = tmp_path / "mv.txt", tmp_path / "res.txt"
        mvf.write_text("\n".join(mv) + "\n")
        rsf.write_text("\n".join(res) + "\n")
        return d.parse_bitstream(str(mvf), str(rsf))

    parse(mv_lines, res_lines)                                        # sanity: the unmodified stream parses

    def res_with(first_block):                                        # replace the first block of frame 0
        rest = res_lines[0].split(";", 1)[1]
        return [first_block + ";" + rest] + res_lines[1:]

    bad_res = [
        res_with("0'([4294967295, -1, 5])"),                          # zero run that truncates to a negative int
        res_with("0'([2147483647, 2147483647, -1, 5])"),               # repeated runs driving the position negative
        res_with("0'([300, -1, 7])"),                                 # run past the end of a 16x16 block
        res_with("0'([-300, 1, 2, 3])"),                              # non-zero run longer than the block
        res_with("0'([99999999999999999999999999, 0])"),              # overflows long
        res_with("0'([-1, 70000, 0])"),                               # level outside int16
        res_with("0'([-1, 5, 0)"),                                    # missing ']'
        res_with("2'([0])"),                                          # split flag that is neither 0 nor 1
        [res_lines[0][: len(res_lines[0]) // 2]] + res_lines[1:],     # truncated line
    ]
    for res in bad_res:
        with pytest.raises(ValueError):
            parse(mv_lines, res)
    t, body = mv_lines[1].split("|", 1)
    first, rest = body.split(";", 1)
    bad_mv = [
        ["7|" + mv_lines[0].split("|", 1)[1]] + mv_lines[1:],                        # frame type
        [mv_lines[0], t + "|0'(40000, 0, 0);" + rest] + mv_lines[2:],                # vector outside int16
        [mv_lines[0], t + "|0'(0, 0, 99);" + rest] + mv_lines[2:],                   # reference index
        [mv_lines[0], t + "|0'(99999999999999999999, 0, 0);" + rest] + mv_lines[2:],
        [mv_lines[0], t + "|" + first] + mv_lines[2:],                               # too few blocks
        [mv_lines[0].replace("|", "", 1)] + mv_lines[1:],
    ]
    for mv in bad_mv:
        with pytest.raises(ValueError):
            parse(mv, res_lines)
    # mutation fuzz: random single-character edits of both streams; must return or raise ValueError, nothing else
    rng = np.random.default_rng(7)
    alphabet = "0123456789-,;'()[]|@ "
    for _ in range(300):
        mv, res = list(mv_lines), list(res_lines)
        for lines in (mv, res):
            f = int(rng.integers(len(lines)))
            s = lines[f]
            for _ in range(int(rng.integers(1, 4))):
                p = int(rng.integers(len(s)))
                kind = int(rng.integers(3))
                ch = alphabet[int(rng.integers(len(alphabet)))]
                s = s[:p] + (ch + s[p + 1:] if kind == 0 else (ch + s[p:] if kind == 1 else s[p + 1:]))
            lines[f] = s
        try:
            parse(mv, res)
        except ValueError:
            pass
