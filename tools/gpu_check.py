"""Development aid: run every golden case free-running through the CUDA encoder and summarise mismatches."""
import hashlib
import sys
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.golden_util import case_names, load_case
from streamoptima_b200.Encoder import Y_Video_codec

names = sys.argv[1:] or case_names()
Y_Video_codec.write_recon_yuv = False
for name in names:
    frames, enc, g = load_case(name)
    F, H, W = frames.shape
    try:
        c = Y_Video_codec(H, W, F, enc.pop("block_size"), enc.pop("search_range"), enc.pop("Qp"), enc.pop("intra_dur"), 0,
                          y_only_frame_arr=frames, **enc)
        psnr = c.encode()
    except Exception as e:
        print(f"{name}: EXC {type(e).__name__}: {e}")
        continue
    p = c.encoded_package.packed
    res = {}
    res["types"] = (p["frame_types"] == g["frame_types"]).all()
    for k in ("split", "mv", "levels", "recon"):
        eq = (p[k] == g[k])
        per_frame = [bool(eq[f].all()) for f in range(F)]
        res[k] = f"{int((~eq).sum())} diff; frames ok={''.join('1' if v else '0' for v in per_frame)}"
    mvl, rsl = c.bitstream_lines()
    mvt = "".join(l + "\n" for l in mvl); rst = "".join(l + "\n" for l in rsl)
    res["mv_text"] = mvt == g["mv_text"]
    res["res_text"] = rst == g["res_text"]
    res["dpsnr"] = float(np.max(np.abs(np.array(psnr) - g["psnr"])))
    mae = np.array([m if np.isfinite(m) else -1.0 for m in c.encoded_package["MAE per Frame"]])
    res["mae_ok"] = bool(np.allclose(mae, g["mae"], rtol=1e-15, atol=0))
    print(name, res, c.last_timing, flush=True)
