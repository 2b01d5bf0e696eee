"""All-intra 1080p timing (intra search / finish+chain per frame) from the library's CUDA-event timers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamoptima_b200 import synth
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
F, H, W = 30, 1088, 1920
frames = synth.translating(F, H, W, seed=1)
c = Y_Video_codec(H, W, F, 16, 16, 4, 1, 0, y_only_frame_arr=frames)
for _ in range(3):
    c.encode_arrays(frames)
t = c.last_timing
print({k: (round(v / (F if k == "device_ms" else max(1, t["timed_frames"])), 4) if k.endswith("_ms") else v) for k, v in t.items()}, "per frame (ms)")
