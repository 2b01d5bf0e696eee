"""A batched 4K launch set for profiling (C5 geometry: 8 units x 16 frames of 3840x2160, i=16, r=16 integer search, nRef=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import synth_frames_torch
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
U, F, H, W = int(os.environ.get("U", 8)), int(os.environ.get("F", 16)), 2160, 3840
dev = torch.device("cuda", 0)
frames = np.stack([synth_frames_torch(F, H, W, u, dev).cpu().numpy() for u in range(U)])
c = Y_Video_codec(H, W, F, 16, 16, 4, 16, 0)
for _ in range(int(os.environ.get("REPS", 1))):
    c.encode_arrays(frames, want_levels=False, want_recon=False, want_symbols=True)
print(c.last_timing)
