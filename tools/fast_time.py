"""Fast-ME (3x3 around the predictor, chained across blocks: Encoder.py:719-742, :581) timing at 1080p and CIF."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamoptima_b200 import synth
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
for (F, H, W, U, kw) in ((21, 288, 352, 1, dict(fast_me=True, FMEEnable=True, nRefFrames=1, bs=8)),
                         (20, 1088, 1920, 1, dict(fast_me=True, FMEEnable=True, nRefFrames=4)),
                         (20, 1088, 1920, 1, dict(fast_me=True, FMEEnable=True, nRefFrames=1, VBSEnable=True, lam=0.015)),
                         (21, 288, 352, 1, dict(fast_me=True, FMEEnable=True, nRefFrames=1, VBSEnable=True, lam=0.015)),
                         (21, 288, 352, 32, dict(fast_me=True, FMEEnable=True, nRefFrames=1, VBSEnable=True, lam=0.015))):
    import numpy as np
    frames = np.stack([synth.translating(F, H, W, seed=u) for u in range(U)])
    kw = dict(kw); bs = kw.pop("bs", 16)
    c = Y_Video_codec(H, W, F, bs, 16, 5, 21, 0, **kw)
    for _ in range(2):
        c.encode_arrays(frames)
    t = c.last_timing
    print(H, W, U, bs, kw, {k: round(v / (F if k == "device_ms" else max(1, t["timed_frames"])), 4) for k, v in t.items() if k.endswith("_ms")}, "ms/frame ->", round(U * F / t["device_ms"] * 1e3, 1), "fps")
    c._ctx.close()
