"""Development aid: batched units (U>1) vs serial encode."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamoptima_b200 import synth, sharding
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
F, H, W, ip = 12, 64, 96, 4
frames = synth.translating(F, H, W, seed=9)
for kw in (dict(), dict(FMEEnable=True), dict(VBSEnable=True, lam=0.02), dict(FMEEnable=True, VBSEnable=True, lam=0.02), dict(fast_me=True)):
    serial = Y_Video_codec(H, W, F, 8, 4, 3, ip, 0, y_only_frame_arr=frames, **kw)
    serial.encode()
    sp = {k: v.copy() for k, v in serial.encoded_package.packed.items() if isinstance(v, np.ndarray)}
    codec = Y_Video_codec(H, W, ip, 8, 4, 3, ip, 0, **kw)
    out = codec.encode_arrays(np.stack([frames[g * ip:(g + 1) * ip] for g in range(3)]))
    msg = []
    for k in ("split", "mv", "levels", "recon"):
        got = out[k].reshape((F,) + out[k].shape[2:])
        eq = got == sp[k]
        bad = np.argwhere(~eq)
        msg.append(f"{k}: {len(bad)} diff" + (f" first {bad[:4].tolist()}" if len(bad) else ""))
    print(kw, " | ".join(msg), flush=True)
