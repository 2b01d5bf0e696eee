"""One fast-ME configuration (1080p, nRef 4, half-pel) for profiling the table-driven chain."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamoptima_b200 import synth
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
F, H, W = 10, 1088, 1920
frames = synth.translating(F, H, W, seed=0)
c = Y_Video_codec(H, W, F, 16, 16, 5, 21, 0, fast_me=True, FMEEnable=True, nRefFrames=4)
c.encode_arrays(frames)
print(c.last_timing)
