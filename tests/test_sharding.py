"""Host logic of the multi-GPU path: unit planning, round-robin assignment and the statistics all-gather (gloo, CPU)."""
import os
import socket

import numpy as np
import pytest

from streamoptima_b200 import sharding
from streamoptima_b200._native import STATS_DTYPE


def test_plan_units_closed_gops_only_for_one_reference():
    u = sharding.plan_units(2, 10, 4, 1)
    assert [(x.stream, x.start, x.length) for x in u] == [(0, 0, 4), (0, 4, 4), (0, 8, 2), (1, 0, 4), (1, 4, 4), (1, 8, 2)]
    assert [(x.stream, x.start, x.length) for x in sharding.plan_units(2, 10, 4, 3)] == [(0, 0, 10), (1, 0, 10)]
    assert len(sharding.plan_units(1, 10, 4, 1, parallel_mode=1)) == 1       # ParallelMode 1 has no I frames
    assert sharding.assign(u, 4) == [[0, 4], [1, 5], [2], [3]]


def _fake_encode(batch):
    """Deterministic stand-in for Y_Video_codec.encode_arrays: statistics are simple functions of the pixels."""
    U, L, H, W = batch.shape
    stats = np.zeros((U, L), dtype=STATS_DTYPE)
    stats["qsize"] = batch.reshape(U, L, -1).sum(axis=2) % 100000
    stats["sse"] = batch.reshape(U, L, -1).astype(np.uint64).max(axis=2)
    stats["frame_type"] = (np.arange(L)[None, :] % 4 != 0)
    rows = (batch.reshape(U, L, H // 8, 8, W).sum(axis=(3, 4)) % 1000).astype(np.uint32)
    return dict(stats=stats, row_sizes=rows)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    streams = rng.integers(0, 256, (3, 10, 16, 24), dtype=np.uint8)
    local, gathered = sharding.encode_sharded(streams, _fake_encode, intra_dur=4, n_ref_frames=1, rank=rank, world=world, dist=dist)
    q.put((rank, sorted(local.keys()), {k: v.copy() for k, v in gathered.items()}))
    dist.destroy_process_group()


def test_all_gather_of_statistics_world2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # reference: single rank
    rng = np.random.default_rng(0)
    streams = rng.integers(0, 256, (3, 10, 16, 24), dtype=np.uint8)
    _, want = sharding.encode_sharded(streams, _fake_encode, intra_dur=4, n_ref_frames=1)
    assert res[0][1] == [0, 2, 4, 6, 8] and res[1][1] == [1, 3, 5, 7]
    for r in res:
        for k in want:
            np.testing.assert_array_equal(r[2][k], want[k])


@pytest.mark.gpu
def test_gop_sharded_encode_equals_serial_encode():
    """nRefFrames == 1: encoding the closed GOPs as independent batched units gives the serial encoder's output."""
    from streamoptima_b200 import synth
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    F, H, W, ip = 12, 64, 96, 4
    frames = synth.translating(F, H, W, seed=9)
    kw = dict(FMEEnable=True, VBSEnable=True, lam=0.02)
    serial = Y_Video_codec(H, W, F, 8, 4, 3, ip, 0, y_only_frame_arr=frames, **kw)
    serial.encode()
    sp = serial.encoded_package.packed
    keep = {k: sp[k].copy() for k in ("split", "mv", "levels", "recon")}
    codec = Y_Video_codec(H, W, ip, 8, 4, 3, ip, 0, **kw)
    local, gathered = sharding.encode_sharded(frames[None], codec.encode_arrays, intra_dur=ip, n_ref_frames=1)
    assert sorted(local) == [0, 1, 2]
    for ui, out in local.items():
        sl = slice(ui * ip, ui * ip + ip)
        for k in ("split", "mv", "levels", "recon"):
            np.testing.assert_array_equal(out[k], keep[k][sl])
    np.testing.assert_array_equal(gathered["qsize"][0], sp["qsize"])
