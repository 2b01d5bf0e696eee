"""Throughput of the five BASELINE.json configurations (short runs; kernel-resident timing from so_last_timing and
end-to-end wall time of encode_arrays).  Development / documentation aid; bench.py stays the contract benchmark."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_frames_torch
from streamoptima_b200.Encoder import Y_Video_codec

Y_Video_codec.write_recon_yuv = False
dev = torch.device("cuda", 0)
TAB = [[90000, 70000, 52000, 39000, 28000, 19000, 13000, 9000, 6000, 4000, 2500, 1000],
       [60000, 46000, 34000, 25000, 18000, 12000, 8000, 5600, 3800, 2500, 1600, 600]]


def run(name, U, F, H, W, args, kw, reps=2):
    frames = torch.stack([synth_frames_torch(F, H, W, seed=u, device=dev) for u in range(U)]).cpu().numpy()
    c = Y_Video_codec(H, W, F, *args, 0, **kw)
    kw_e2e = dict(want_levels=False, want_recon=False, want_symbols=True)      # the residual travels as packed run-level symbols
    for _ in range(2):
        c.encode_arrays(frames, **kw_e2e)
    t0 = time.perf_counter()
    for _ in range(reps):
        c.encode_arrays(frames, **kw_e2e)
    wall = (time.perf_counter() - t0) / reps
    # kernels alone: the sequence(s) resident in HBM, CUDA events around so_seq_run
    from streamoptima_b200 import _native
    ctx = c._ctx
    _native.check(ctx.handle, ctx.lib.so_seq_upload(ctx.handle, frames.ctypes.data, U, F))
    _native.check(ctx.handle, ctx.lib.so_seq_sync(ctx.handle))
    dev_ms, t = 0.0, None
    for i in range(reps + 1):
        _native.check(ctx.handle, ctx.lib.so_seq_run(ctx.handle))
        t = ctx.last_timing()
        if i:
            dev_ms += t["device_ms"]
    out = dict(config=name, units=U, frames=F, size=f"{W}x{H}", fps_kernel=U * F / (dev_ms / reps / 1e3), fps_e2e=U * F / wall,
               me_ms_per_frame=t["me_ms"] / max(1, t["timed_frames"]) / U, launches=t["launches"])
    print(json.dumps(out), flush=True)
    c._ctx.close()


which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
if "c1" in which:
    run("C1 CIF i=8 r=2 QP6 I_Period 8 (x1 stream)", 1, 10, 288, 352, (8, 2, 6, 8), {})
    run("C1 CIF i=8 r=2 QP6 I_Period 8 (x64 streams batched)", 64, 10, 288, 352, (8, 2, 6, 8), {})
if "c2" in which:
    run("C2 1080p i=16 r=16 FME nRef4", 1, 60, 1088, 1920, (16, 16, 4, 30), dict(nRefFrames=4, FMEEnable=True))
if "c3" in which:
    run("C3 1080p VBS+RDO + table RC", 1, 60, 1088, 1920, (16, 16, 4, 30),
        dict(nRefFrames=4, FMEEnable=True, VBSEnable=True, lam=0.02, RCFlag=1, targetBR="20 mbps", qp_rate_tables=TAB))
if "c4" in which:
    run("C4 1080p ParallelMode 2 + row QPs", 1, 60, 1088, 1920, (16, 16, 4, 30),
        dict(nRefFrames=4, FMEEnable=True, ParallelMode=2, RCFlag=1, targetBR="20 mbps", qp_rate_tables=TAB))
    run("C4 1080p ParallelMode 1", 1, 30, 1088, 1920, (16, 16, 4, 30), dict(FMEEnable=True, ParallelMode=1))
if "c5" in which:
    run("C5 4K x8 streams, one GOP of 16 (I_Period 16), i=16 r=16", 8, 16, 2160, 3840, (16, 16, 4, 16), {})
    run("C5 4K x8 streams, half-pel", 8, 16, 2160, 3840, (16, 16, 4, 16), dict(FMEEnable=True))
