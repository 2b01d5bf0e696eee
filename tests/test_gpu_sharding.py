"""Closed-GOP / stream sharding on the GPU: with nRefFrames == 1 the GOPs of a stream are independent, so
encoding them as separate batched units (what every rank does with its share) must reproduce the whole-sequence encode
bit for bit; with nRefFrames > 1 a stream stays one unit (quirk Q7)."""
import numpy as np
import pytest

from streamoptima_b200 import sharding, synth

pytestmark = pytest.mark.gpu


def _codec(H, W, F, **kw):
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    return Y_Video_codec(H, W, F, 16, 16, 3, 4, 0, **kw)


@pytest.mark.parametrize("kw", [dict(), dict(FMEEnable=True), dict(FMEEnable=True, VBSEnable=True, lam=0.02)])
def test_gop_sharded_equals_whole_sequence(kw):
    S, F, H, W = 2, 10, 96, 128                       # I_Period 4 -> GOPs of 4, 4, 2 frames per stream
    streams = np.stack([synth.make(k, F=F, H=H, W=W, seed=90 + i) for i, k in enumerate(("translating", "zooming"))])
    whole = _codec(H, W, F, **kw)
    ow = {k: np.array(v) for k, v in whole.encode_arrays(streams).items() if k in ("split", "mv", "levels", "recon", "row_sizes")}
    units = sharding.plan_units(S, F, 4, 1)
    assert len(units) == 6
    enc = _codec(H, W, F, **kw)
    merged = {k: np.zeros_like(v) for k, v in ow.items()}
    gathered_all = None
    for rank in range(2):                              # the two ranks' shares, run one after the other on this GPU
        local, gathered = sharding.encode_sharded(streams, lambda b: {k: np.array(v) for k, v in enc.encode_arrays(b).items()},
                                                  intra_dur=4, n_ref_frames=1, rank=rank, world=2)
        for ui, out in local.items():
            u = units[ui]
            for k in merged:
                merged[k][u.stream, u.start:u.start + u.length] = out[k]
        gathered_all = gathered if gathered_all is None else {k: gathered_all[k] + gathered[k] for k in gathered}
    for k in ow:
        np.testing.assert_array_equal(merged[k], ow[k], err_msg=k)
    np.testing.assert_array_equal(gathered_all["row_sizes"], ow["row_sizes"])      # what the all-gather would deliver


def test_multi_reference_streams_are_single_units():
    assert [(u.stream, u.start, u.length) for u in sharding.plan_units(3, 32, 16, 4)] == [(0, 0, 32), (1, 0, 32), (2, 0, 32)]
