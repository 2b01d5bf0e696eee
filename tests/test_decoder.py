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
        ft, split, mv, lev, qps = d.parse_bitstream(mvf, rsf)
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
