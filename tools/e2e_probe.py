"""Where the end-to-end time goes (C2 geometry): wall vs device for the three download modes of encode_arrays."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import synth_frames_torch
from streamoptima_b200.Encoder import Y_Video_codec
print("oracle/_ref:", os.listdir(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref"))
      if os.path.isdir(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")) else "ABSENT")
F, H, W = int(os.environ.get("F", 300)), 1088, 1920
dev = torch.device("cuda", 0)
fr = synth_frames_torch(F, H, W, 0, dev)
pin = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True); pin.copy_(fr); frames = pin.numpy(); del fr
Y_Video_codec.write_recon_yuv = False
c = Y_Video_codec(H, W, F, 16, 16, 4, 30, 0, nRefFrames=4, FMEEnable=True)
for name, kw in (("levels+recon", dict(want_levels=True, want_recon=True)),
                 ("symbols", dict(want_levels=False, want_recon=False, want_symbols=True)),
                 ("symbols+recon", dict(want_levels=False, want_recon=True, want_symbols=True)),
                 ("mv only", dict(want_levels=False, want_recon=False))):
    for qp in (0, 4, 8):
        c.const_init_Qp = qp
        ts = []
        for rep in range(3):
            t0 = time.perf_counter()
            out = c.encode_arrays(frames, **kw)
            ts.append(time.perf_counter() - t0)
        print(f"{name:14s} qp {qp}: wall {[round(t * 1e3, 1) for t in ts]} ms  device {c.last_timing['device_ms']:.1f} ms  "
              f"symbols/px {out.get('sym_needed', 0) / (F * H * W):.3f}", flush=True)
        del out
