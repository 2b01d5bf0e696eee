"""Small encodes that touch every kernel family; run as a plain smoke test of all kernel families (decode round trips); under compute-sanitizer where the pool allows it:
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from streamoptima_b200 import synth, decoder as dec
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
CASES = [
    (3, 96, 128, 16, 16, dict(FMEEnable=True, nRefFrames=2)),                              # item-ring kernel
    (3, 96, 128, 16, 16, dict(FMEEnable=True, nRefFrames=2, VBSEnable=True, lam=0.02)),    # fused VBS ring kernel
    (3, 96, 128, 16, 16, dict()),                                                          # integer, one phase
    (3, 64, 96, 8, 5, dict(FMEEnable=True, VBSEnable=True, lam=0.03)),                     # stage kernel, EXPAND staging
    (3, 64, 96, 16, 24, dict(FMEEnable=True)),                                             # search chunks
    (3, 32, 48, 4, 2, dict(VBSEnable=True, lam=0.02)),                                     # 2x2 sub-blocks: plain search
    (3, 96, 128, 16, 16, dict(fast_me=True, FMEEnable=True, nRefFrames=2, VBSEnable=True, lam=0.015)),   # fast ME chain
    (2, 96, 128, 16, 40, dict(VBSEnable=True, lam=0.01)),                                  # intra chunks (I_Period 1 below)
]
for n, (F, H, W, bs, r, kw) in enumerate(CASES):
    frames = synth.zooming(F, H, W, seed=n)
    ip = 1 if n == len(CASES) - 1 else 8
    c = Y_Video_codec(H, W, F, bs, r, 3, ip, 0, y_only_frame_arr=frames, **kw)
    c.encode()
    p = c.encoded_package.packed
    c.symbol_streams()
    d = dec.decoder(0, ip, bs, F, H, W, 3, kw.get("nRefFrames", 1), kw.get("FMEEnable", False), None, kw.get("VBSEnable", False))
    out = d.decode_arrays(p["frame_types"], p["split"], p["mv"], p["levels"], None, reset_at_intra=False)
    assert np.array_equal(out, p["recon"]), n
    print("case", n, "ok", flush=True)
print("all ok")
