"""Development aid: print producer/consumer timestamps of the ME kernel (library built with -DSO_ME_DEBUG)."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from streamoptima_b200 import synth, _native
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
F, H, W = 6, 1088, 1920
frames = synth.translating(F, H, W, seed=0)
c = Y_Video_codec(H, W, F, 16, 16, 4, 30, 0, nRefFrames=4, FMEEnable=True, y_only_frame_arr=frames)
c.encode()
print(c.last_timing)
lib = _native.load()
cudart = ctypes.CDLL("libcudart.so.12")
sym = ctypes.c_void_p.in_dll(lib, "g_me_dbg") if False else None
# read the __device__ symbol through a tiny helper exported by the debug build
buf = np.zeros(4096, np.int64)
lib.so_debug_read.argtypes = [ctypes.c_void_p]
lib.so_debug_read(buf.ctypes.data)
d = buf.reshape(128, 32)
base = d[2, 0]
print("stage: prod[wait_start, rawfull, empty, ready_signalled, tma_issued] | cons kb0 [fetch, ready, done] | cons kb10 [fetch, ready, done]  (clk rel.)")
for j in range(2, 40):
    r = d[j] - base
    print(j, r[0:5].tolist(), "|", r[5:8].tolist(), "|", r[8:11].tolist(), " expand=", r[3] - r[2], " T_kb0=", r[7] - r[6])
