"""C3-like run (1080p VBS + table rate control) for profiling the fused VBS search kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_frames_torch
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
TAB = [[90000, 70000, 52000, 39000, 28000, 19000, 13000, 9000, 6000, 4000, 2500, 1000],
       [60000, 46000, 34000, 25000, 18000, 12000, 8000, 5600, 3800, 2500, 1600, 600]]
F, H, W = 12, 1088, 1920
frames = synth_frames_torch(F, H, W, seed=0, device=torch.device("cuda", 0)).cpu().numpy()
c = Y_Video_codec(H, W, F, 16, 16, 4, 30, 0, nRefFrames=4, FMEEnable=True, VBSEnable=True, lam=0.02, RCFlag=1, targetBR="20 mbps", qp_rate_tables=TAB)
c.encode_arrays(frames)
print(c.last_timing)
