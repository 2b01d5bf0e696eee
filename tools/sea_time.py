"""Pruned (successive elimination, SO_FLAG_SEA) against plain exhaustive search: kernels alone on resident frames (CUDA events of
so_seq_run) and end to end through encode_arrays.  CFG=c2 (default): 1080p, i=16, r=16 half-pel, 4 references; CFG=c5: 8 x 4K units,
integer search, 1 reference.  F = frames (default 60)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from streamoptima_b200 import synth, _native
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
cfg = os.environ.get("CFG", "c2")
F = int(os.environ.get("F", 60))
if cfg == "c2":
    U, H, W, kw, ip = 1, 1088, 1920, dict(nRefFrames=4, FMEEnable=True), 30
else:
    U, H, W, kw, ip = 8, 2160, 3840, dict(nRefFrames=1), 16
    F = min(F, 16)
kind = os.environ.get("KIND", "translating")
frames = np.stack([synth.make(kind, F=F, H=H, W=W, seed=u) for u in range(U)])
out = {"cfg": cfg, "kind": kind, "frames": F, "units": U}
ref = None
for sea in (False, True, "auto"):
    c = Y_Video_codec(H, W, F, 16, 16, 4, ip, 0, **kw)
    c.sea_prune = sea
    r = c.encode_arrays(frames, want_levels=False, want_recon=True, want_symbols=True)
    sig = [np.array(r[k]) for k in ("split", "mv", "recon")]
    if ref is None:
        ref = sig
    else:
        out["identical"] = out.get("identical", True) and all(np.array_equal(a, b) for a, b in zip(ref, sig))
    ctx = c._ctx
    _native.check(ctx.handle, ctx.lib.so_seq_upload(ctx.handle, frames.ctypes.data, U, F))
    _native.check(ctx.handle, ctx.lib.so_seq_sync(ctx.handle))
    ms = []
    for _ in range(4):
        _native.check(ctx.handle, ctx.lib.so_seq_run(ctx.handle))
        t = ctx.last_timing()
        ms.append(t["device_ms"])
    t0 = time.perf_counter()
    for _ in range(2):
        c.encode_arrays(frames, want_levels=False, want_recon=False, want_symbols=True)
    e2e = (time.perf_counter() - t0) / 2
    key = "auto" if sea == "auto" else ("sea" if sea else "plain")
    out[key] = {"device_ms": min(ms[1:]), "device_fps": U * F / (min(ms[1:]) / 1e3), "e2e_fps": U * F / e2e, "timing": t}
    if sea:
        out[key]["stats"] = ctx.sea_stats()
print(json.dumps(out))
