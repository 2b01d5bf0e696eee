"""Generate ``tests/golden/*.npz`` by running the unmodified reference (build container only).

TEST INFRASTRUCTURE ONLY.   usage:  python -m oracle.gen_golden [case ...]      (default: all missing cases)

Every fixture stores the reference's outputs for one case of ``oracle/golden_cases.py``: frame types, packed MVs /
split flags / levels (``oracle/packing.py``), reconstructed frames, per-row QPs, PSNR, MAE, the canonical MV and
residual text streams (``Encoder.py:1419-1542``, zlib-compressed) and sha256 digests of input and outputs.
Environment the goldens were produced with is recorded in ``tests/golden/MANIFEST.json``.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_harness as rh          # noqa: E402
from oracle.golden_cases import CASES                # noqa: E402
from oracle.packing import package_to_arrays         # noqa: E402
from streamoptima_b200 import synth                  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def generate(name: str) -> dict:
    case = CASES[name]
    kind, gkw = case["gen"]
    frames = synth.make(kind, **gkw)
    enc = dict(case["enc"])
    t0 = time.time()
    r = rh.run_reference(frames, **enc)
    dt = time.time() - t0
    F, H, W = frames.shape
    bs = enc["block_size"]
    split, mv, lev = package_to_arrays(r["frame_types"], r["mvs"], r["levels"], H, W, bs)
    rows = H // bs
    qp_rows = np.full((F, rows), -1, np.int16)
    for f, q in enumerate(r["qp_rows"]):
        if len(q):
            qp_rows[f, :len(q)] = q
    mv_text = "".join(l + "\n" for l in r["mv_text"]).encode()
    res_text = "".join(l + "\n" for l in r["res_text"]).encode()
    mae = np.array([m if np.isfinite(m) else -1.0 for m in r["mae"]], np.float64)
    out = dict(frame_types=np.array(r["frame_types"], np.uint8), split=split, mv=mv, levels=lev, recon=r["recon"],
               qp_rows=qp_rows, psnr=np.array(r["psnr"], np.float64), mae=mae,
               mv_text_z=np.frombuffer(zlib.compress(mv_text, 9), np.uint8),
               res_text_z=np.frombuffer(zlib.compress(res_text, 9), np.uint8),
               sha_input=np.array(sha(frames.tobytes())), sha_mv_text=np.array(sha(mv_text)),
               sha_res_text=np.array(sha(res_text)), sha_recon=np.array(sha(r["recon"].tobytes())))
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    return dict(seconds=round(dt, 1), sha_input=sha(frames.tobytes())[:16], sha_mv_text=sha(mv_text)[:16],
                sha_res_text=sha(res_text)[:16], sha_recon=sha(r["recon"].tobytes())[:16],
                psnr=[round(p, 4) for p in r["psnr"]], frame_types=r["frame_types"])


def main(argv):
    import scipy
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    man_path = os.path.join(GOLDEN_DIR, "MANIFEST.json")
    manifest = json.load(open(man_path)) if os.path.exists(man_path) else {"cases": {}}
    manifest["environment"] = dict(python=sys.version.split()[0], numpy=np.__version__, scipy=scipy.__version__,
                                   reference="/root/reference (Suyashagarw/StreamOptima, unmodified apart from the "
                                             "shims listed in oracle/reference_harness.py)")
    names = argv or [n for n in CASES if not os.path.exists(os.path.join(GOLDEN_DIR, n + ".npz"))]
    for n in names:
        info = generate(n)
        if os.path.exists(man_path):          # another generator process may have added cases meanwhile
            manifest["cases"].update(json.load(open(man_path)).get("cases", {}))
        manifest["cases"][n] = info
        print(n, info, flush=True)
        with open(man_path, "w") as f:
            json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:])
