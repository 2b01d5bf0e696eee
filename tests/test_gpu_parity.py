"""GPU parity: the CUDA encoder (through the C ABI) must reproduce the reference bit for bit on the golden cases.

Bar (BASELINE.json north_star): motion vectors, modes, quantised levels, reconstruction and both text streams
bit-exact; PSNR within 0.01 dB (here: identical, it is computed from an exact integer SSE).
"""
import numpy as np
import pytest

from tests.golden_util import case_names, load_case

pytestmark = pytest.mark.gpu


def _codec(frames, enc):
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    F, H, W = frames.shape
    e = dict(enc)
    return Y_Video_codec(H, W, F, e.pop("block_size"), e.pop("search_range"), e.pop("Qp"), e.pop("intra_dur"), 0,
                         y_only_frame_arr=frames, **e)


@pytest.mark.parametrize("name", case_names())
def test_free_running_bit_exact(name):
    frames, enc, g = load_case(name)
    c = _codec(frames, enc)
    psnr = c.encode()
    pkg = c.encoded_package
    p = pkg.packed
    assert pkg["frame_type_seq"] == g["frame_types"].tolist()
    np.testing.assert_array_equal(p["split"], g["split"])
    np.testing.assert_array_equal(p["mv"], g["mv"])
    np.testing.assert_array_equal(p["levels"], g["levels"])
    np.testing.assert_array_equal(p["recon"], g["recon"])
    np.testing.assert_allclose(psnr, g["psnr"], rtol=0, atol=1e-9)
    mae = np.array([m if np.isfinite(m) else -1.0 for m in pkg["MAE per Frame"]])
    np.testing.assert_allclose(mae, g["mae"], rtol=1e-15, atol=0)
    for f, q in enumerate(pkg["Qp_per_row_per_frame"]):
        assert list(q) == [int(v) for v in g["qp_rows"][f] if v >= 0]
    mv_lines, res_lines = c.bitstream_lines()
    assert "".join(l + "\n" for l in mv_lines) == g["mv_text"]
    assert "".join(l + "\n" for l in res_lines) == g["res_text"]


def test_package_structure_like_reference():
    frames, enc, g = load_case("s_vbs")
    c = _codec(frames, enc)
    c.encode()
    pkg = c.encoded_package
    assert set(pkg.keys()) == {"block size", "num frames", "height in pixels", "width in pixels", "search range",
                               "PSNR per frame", "SSIM per frame", "MAE per Frame", "MVS per Frame", "approx residual",
                               "Qp_per_row_per_frame", "frame_type_seq"}
    split, mv = pkg["MVS per Frame"][1][8]
    assert split in (0, 1)
    assert pkg["approx residual"][0][0][1].shape == (8, 8)
