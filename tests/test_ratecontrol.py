"""Two-pass rate control driver and ROI QP maps (extensions: the reference has no code for either -- parity unpinned;
what is checked is internal consistency: tables are monotone, pass 2 obeys them, decode(encode) round-trips)."""
import numpy as np
import pytest

from streamoptima_b200 import ratecontrol


def test_roi_map_geometry():
    m = ratecontrol.roi_qp_map(2, 64, 96, 16, 6, 2, lambda f: (16 * f, 16, 16 * f + 48, 48))
    assert m.shape == (2, 24)
    g = m.reshape(2, 4, 6)
    assert (g[0, 1:3, 0:3] == 2).all() and g[0].sum() == 6 * 24 - 4 * 6
    assert (g[1, 1:3, 1:4] == 2).all() and g[1, 1, 0] == 6


@pytest.mark.gpu
def test_two_pass_encode_meets_tables_and_round_trips():
    from streamoptima_b200 import synth, decoder as dec
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    F, H, W = 6, 96, 128
    frames = synth.translating(F, H, W, seed=61)
    kw = dict(block_size=16, search_range=4, intra_dur=4, nRefFrames=2, FMEEnable=True)
    codec, tables = ratecontrol.two_pass_encode(frames, "300 kbps", kw, rc_flag=1)
    for t in tables:                                   # coarser quantisation never costs more symbols on average
        assert all(t[i] >= t[i + 1] for i in range(len(t) - 1)), t
    pkg = codec.encoded_package
    rows = pkg["Qp_per_row_per_frame"][0]
    assert len(rows) == H // 16 and all(0 <= q < 12 for q in rows)
    budget = codec.bitrate_per_row
    assert tables[0][rows[0]] < budget and (rows[0] == 0 or tables[0][rows[0] - 1] >= budget)   # first QP under the budget
    d = dec.decoder(0, 4, 16, F, H, W, 4, 2, True, None, False, RCFlag=1)
    p = pkg.packed
    out = d.decode_arrays(p["frame_types"], p["split"], p["mv"], p["levels"], pkg["Qp_per_row_per_frame"], reset_at_intra=False)
    np.testing.assert_array_equal(out, p["recon"])


@pytest.mark.gpu
def test_roi_qp_map_changes_only_quantisation_and_round_trips():
    from streamoptima_b200 import synth, decoder as dec
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    F, H, W, bs = 4, 96, 128, 16
    frames = synth.zooming(F, H, W, seed=62)
    qmap = ratecontrol.roi_qp_map(F, H, W, bs, 7, 1, lambda f: (32 + 8 * f, 16, 96 + 8 * f, 80))
    base = Y_Video_codec(H, W, F, bs, 8, 7, 8, 0, y_only_frame_arr=frames, FMEEnable=True)
    base.encode()
    roi = Y_Video_codec(H, W, F, bs, 8, 7, 8, 0, y_only_frame_arr=frames, FMEEnable=True)
    roi.roi_qp_map = qmap
    psnr_roi = roi.encode()
    pb, pr = base.encoded_package.packed, roi.encoded_package.packed
    # a uniform map equal to the base QP is the base encode
    uni = Y_Video_codec(H, W, F, bs, 8, 7, 8, 0, y_only_frame_arr=frames, FMEEnable=True)
    uni.roi_qp_map = np.full_like(qmap, 7)
    uni.encode()
    np.testing.assert_array_equal(uni.encoded_package.packed["levels"], pb["levels"])
    np.testing.assert_array_equal(uni.encoded_package.packed["recon"], pb["recon"])
    # the ROI is reconstructed better than in the base encode, frame 0 (intra: no drift) block-exactly
    g = qmap.reshape(F, H // bs, W // bs)[0]
    err_b = (pb["recon"][0].astype(int) - frames[0]) ** 2
    err_r = (pr["recon"][0].astype(int) - frames[0]) ** 2
    mask = np.kron(g == 1, np.ones((bs, bs), bool))
    assert err_r[mask].mean() < err_b[mask].mean()
    assert sum(psnr_roi) > sum(base.encoded_package["PSNR per frame"])
    d = dec.decoder(0, 8, bs, F, H, W, 7, 1, True, None, False)
    out = d.decode_arrays(pr["frame_types"], pr["split"], pr["mv"], pr["levels"], None, qp_map=qmap)
    np.testing.assert_array_equal(out, pr["recon"])
