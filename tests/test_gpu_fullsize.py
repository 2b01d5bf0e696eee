"""BASELINE configs 3-5 at their full frame sizes: size-independent properties (the oracle needs minutes per 1080p frame)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frames(F, H, W, seed):
    import torch
    from bench import synth_frames_torch
    return synth_frames_torch(F, H, W, seed=seed, device=torch.device("cuda", 0)).cpu().numpy()


def test_config3_vbs_two_pass_1080p_round_trip():
    """C3: 1080p, VBS + RD decision, two-pass rate control (pass-1 tables measured on the sequence itself)."""
    from streamoptima_b200 import ratecontrol, decoder as dec
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    F, H, W = 12, 1088, 1920
    frames = _frames(F, H, W, 11)
    kw = dict(block_size=16, search_range=16, intra_dur=30, nRefFrames=4, FMEEnable=True, VBSEnable=True, lam=0.02)
    codec, tables = ratecontrol.two_pass_encode(frames, "40 mbps", kw, rc_flag=1, qps=range(0, 12, 2))
    pkg = codec.encoded_package
    p = pkg.packed
    assert p["split"].any() and not p["split"].all()                        # the RD decision goes both ways on this content
    assert all(tables[0][i] >= tables[0][i + 1] for i in range(len(tables[0]) - 1))
    d = dec.decoder(0, 30, 16, F, H, W, 4, 4, True, 0.02, True, RCFlag=1)
    out = d.decode_arrays(p["frame_types"], p["split"], p["mv"], p["levels"], pkg["Qp_per_row_per_frame"], reset_at_intra=False)
    np.testing.assert_array_equal(out, p["recon"])


def test_config4_roi_map_parallel_mode_1080p_round_trip():
    """C4: 1080p, per-block QP map on a moving region (ROI extension) in block-parallel mode 2."""
    from streamoptima_b200 import ratecontrol, decoder as dec
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    F, H, W = 6, 1088, 1920
    frames = _frames(F, H, W, 12)
    qmap = ratecontrol.roi_qp_map(F, H, W, 16, 7, 2, lambda f: (400 + 32 * f, 300, 1000 + 32 * f, 800))
    c = Y_Video_codec(H, W, F, 16, 16, 7, 30, 0, y_only_frame_arr=frames, FMEEnable=True, nRefFrames=2, ParallelMode=2)
    c.roi_qp_map = qmap
    c.encode()
    p = c.encoded_package.packed
    d = dec.decoder(0, 30, 16, F, H, W, 7, 2, True, None, False, ParallelMode=2)
    out = d.decode_arrays(p["frame_types"], p["split"], p["mv"], p["levels"], None, reset_at_intra=False, qp_map=qmap)
    np.testing.assert_array_equal(out, p["recon"])
    g = qmap.reshape(F, H // 16, W // 16)[0]
    mask = np.kron(g == 2, np.ones((16, 16), bool))
    err = (p["recon"][0].astype(int) - frames[0]) ** 2
    assert err[mask].mean() < err[~mask].mean()                              # the region is coded finer than its surroundings


def test_config5_4k_streams_gop_sharded_equals_whole():
    """C5: 4K streams, I_Period 16, one reference: closed GOPs encoded as independent batched units (what each rank of a
    multi-GPU run does) equal the whole-sequence encode; decode round trip of one stream."""
    from streamoptima_b200 import sharding, decoder as dec
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    S, F, H, W = 2, 32, 2160, 3840
    streams = np.stack([_frames(F, H, W, 20 + s) for s in range(S)])
    whole = Y_Video_codec(H, W, F, 16, 16, 4, 16, 0)
    ow = {k: np.array(v) for k, v in whole.encode_arrays(streams).items() if k in ("mv", "recon", "levels", "split", "frame_types")}
    enc = Y_Video_codec(H, W, 16, 16, 16, 4, 16, 0)
    units = sharding.plan_units(S, F, 16, 1)
    assert len(units) == 4
    for rank in range(2):
        local, _ = sharding.encode_sharded(streams, lambda b: {k: np.array(v) for k, v in enc.encode_arrays(b).items()},
                                           intra_dur=16, n_ref_frames=1, rank=rank, world=2)
        for ui, out in local.items():
            u = units[ui]
            for k in ("mv", "recon", "levels"):
                np.testing.assert_array_equal(out[k], ow[k][u.stream, u.start:u.start + u.length], err_msg=f"{k} unit {ui}")
    d = dec.decoder(0, 16, 16, F, H, W, 4, 1, False, None, False)
    out = d.decode_arrays(ow["frame_types"][1], ow["split"][1], ow["mv"][1], ow["levels"][1], None)
    np.testing.assert_array_equal(out, ow["recon"][1])
