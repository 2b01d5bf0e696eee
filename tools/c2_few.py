"""A few frames of the bench configuration (C2: 1080p, i=16, r=16 half-pel, 4 references) for profiling; VBS=1 adds the fused VBS search."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamoptima_b200 import synth
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
F, H, W = int(os.environ.get("F", 8)), 1088, 1920
frames = synth.translating(F, H, W, seed=0)
kw = dict(VBSEnable=True, lam=0.02) if os.environ.get("VBS") else {}
c = Y_Video_codec(H, W, F, 16, 16, 4, 30, 0, nRefFrames=4, FMEEnable=True, **kw)
c.encode_arrays(frames, want_levels=False, want_recon=False, want_symbols=True)
print(c.last_timing)
