"""Pin the CPU oracle (oracle/codec_oracle.py) against outputs of the unmodified reference (tests/golden)."""
import numpy as np
import pytest

from oracle import codec_oracle as co
from oracle.packing import package_to_arrays
from tests.golden_util import case_names, load_case

SLOW = {"cif_fme_nref4_vbs", "w1920_fme_nref4", "w1920_fme_nref4_vbs"}      # minutes in the NumPy port: SO_SLOW=1 replays them


@pytest.mark.parametrize("name", case_names())
def test_oracle_matches_reference(name):
    if name in SLOW and not pytest.importorskip("os").environ.get("SO_SLOW"):
        pytest.skip("slow case; set SO_SLOW=1")
    frames, enc, g = load_case(name)
    F, H, W = frames.shape
    bs = enc["block_size"]
    codec = co.OracleCodec(H, W, F, y_only_frame_arr=frames, **enc)
    out = codec.encode()
    assert out["frame_types"] == g["frame_types"].tolist()
    split, mv, lev = package_to_arrays(out["frame_types"], out["mvs"], out["levels"], H, W, bs)
    np.testing.assert_array_equal(split, g["split"])
    np.testing.assert_array_equal(mv, g["mv"])
    np.testing.assert_array_equal(lev, g["levels"])
    np.testing.assert_array_equal(out["recon"], g["recon"])
    for f, q in enumerate(out["qp_rows"]):
        np.testing.assert_array_equal(np.asarray(q, np.int16), g["qp_rows"][f][:len(q)])
        assert (g["qp_rows"][f][len(q):] == -1).all()
    np.testing.assert_allclose(out["psnr"], g["psnr"], rtol=0, atol=1e-9)
    mae = np.array([m if np.isfinite(m) else -1.0 for m in out["mae"]])
    np.testing.assert_allclose(mae, g["mae"], rtol=1e-15, atol=0)
    rc = enc.get("RCFlag")
    mv_text = "".join(co.mv_text_frame(t, m, q, W // bs, rc) + "\n"
                      for t, m, q in zip(out["frame_types"], out["mvs"], out["qp_rows"]))
    res_text = "".join(co.res_text_frame(l) + "\n" for l in out["levels"])
    assert mv_text == g["mv_text"]
    assert res_text == g["res_text"]


def test_rle_known_answers():
    # SURVEY.md appendix A3 (checked against Encoder.py:1086-1131)
    z = np.zeros((4, 4), np.int64)
    assert co.rle_symbols(z) == [0]
    a = z.copy(); a[0, 0] = 5
    assert co.rle_symbols(a) == [-1, 5, 0]
    b = z.copy(); b[0, 1] = 3; b[1, 0] = -2; b[3, 3] = 7
    assert co.rle_symbols(b) == [1, -2, 3, -2, 12, -1, 7]
    assert list(co.scan_order(4)) == [0, 1, 4, 2, 5, 8, 3, 6, 9, 12, 7, 10, 13, 11, 14, 15]
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.choice([2, 4, 8, 16]))
        blk = rng.integers(-3, 4, (n, n)) * (rng.random((n, n)) < 0.3)
        assert co.rle_length(blk) == len(co.rle_symbols(blk))


def test_q_matrix_known_answer():
    assert co.q_matrix(4, 1).tolist() == [[2, 2, 2, 4], [2, 2, 4, 8], [2, 4, 8, 8], [4, 8, 8, 8]]
