"""Development aid: where does the end-to-end step lose time relative to the resident step?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_frames_torch, CFG
from streamoptima_b200 import _native
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
dev = torch.device("cuda", 0)
F, H, W = 120, CFG["H"], CFG["W"]
ft = synth_frames_torch(F, H, W, 0, dev)
pin = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True); pin.copy_(ft); frames = pin.numpy()
c = Y_Video_codec(H, W, F, 16, 16, 4, 30, 0, nRefFrames=4, FMEEnable=True, y_only_frame_arr=frames)
for name, kw in (("all outputs", dict()), ("no levels", dict(want_levels=False)), ("no levels/recon", dict(want_levels=False, want_recon=False))):
    c.encode_arrays(frames, **kw)
    t0 = time.perf_counter(); c.encode_arrays(frames, **kw); torch.cuda.synchronize(); w = time.perf_counter() - t0
    print(name, "wall ms", round(w * 1e3, 1), c.last_timing, flush=True)
ctx = c._ctx; lib = ctx.lib
_native.check(ctx.handle, lib.so_seq_upload(ctx.handle, frames.ctypes.data, 1, F)); lib.so_seq_sync(ctx.handle)
for _ in range(2):
    t0 = time.perf_counter(); lib.so_seq_run(ctx.handle); t = ctx.last_timing(); w = time.perf_counter() - t0
    print("resident", "wall ms", round(w * 1e3, 1), t, flush=True)
