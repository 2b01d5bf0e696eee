"""Randomised differential test: small random configurations, CUDA path vs the CPU oracle, everything bit for bit (packed
outputs, both text streams, device-generated symbols, decode round trip).
    python tools/fuzz_vs_oracle.py [n_cases] [seed] [big|ring]   (GPU box; about 0.1 s per case, 1 s with `big`)
`ring`: only the geometry of the bench kernels (16x16 blocks, r = 16: item-ring search kernel, fused VBS search, and -- on half of
the non-VBS cases -- the pruned search `sea_prune`), frames up to 208 x 304, several seconds of oracle per case."""
import os, sys, time, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import codec_oracle as co
from oracle.packing import package_to_arrays
from streamoptima_b200 import synth, decoder as dec
from streamoptima_b200.Encoder import Y_Video_codec
Y_Video_codec.write_recon_yuv = False
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
BIG = len(sys.argv) > 3 and sys.argv[3] == "big"       # frames up to 160 x 224: several CTAs / chunks per launch (slower oracle)
RING = len(sys.argv) > 3 and sys.argv[3] == "ring"
TAB = [[9000, 7000, 5200, 3900, 2800, 1900, 1300, 900, 600, 400, 250, 100], [6000, 4600, 3400, 2500, 1800, 1200, 800, 560, 380, 250, 160, 60]]
bad = skipped = 0
t0 = time.time()
for n in range(n_cases):
    bs = int(rng.choice([4, 8, 16, 16]))
    H = bs * int(rng.integers(2, max(3, (160 if BIG else 80) // bs))); W = bs * int(rng.integers(3, max(4, (224 if BIG else 112) // bs)))
    F = int(rng.integers(2, 5))
    r = int(rng.choice([0, 1, 2, 3, 4, 5, 7, 8, 16])) if bs == 16 else int(rng.choice([0, 1, 2, 3, 4, 6]))
    if rng.random() < 0.04: r = int(rng.choice([17, 20, 33]))                      # search chunks (ranges above 16)
    kw = dict(block_size=bs, search_range=r, Qp=int(rng.integers(0, {4: 10, 8: 11, 16: 12}[bs])), intra_dur=int(rng.integers(1, 6)))
    if rng.random() < 0.5: kw["FMEEnable"] = True
    if rng.random() < 0.5: kw["nRefFrames"] = int(rng.integers(2, 5)) if rng.random() < 0.85 else int(rng.integers(5, 9))
    if RING:
        bs, r = 16, 16
        H = 16 * int(rng.integers(2, 14)); W = 16 * int(rng.integers(3, 20))
        kw.update(block_size=16, search_range=16, Qp=int(rng.integers(0, 12)))
    if r > 16: F = min(F, 3)
    if rng.random() < 0.4: kw.update(VBSEnable=True, lam=float(rng.choice([0.005, 0.02, 0.3])))
    mode = rng.random()
    if mode < (0.0 if RING else 0.3): kw["fast_me"] = True
    pm = rng.random()
    if pm < 0.15 and not (kw.get("VBSEnable") and kw.get("fast_me")): kw["ParallelMode"] = 2
    elif pm < 0.25 and not (kw.get("VBSEnable") and kw.get("fast_me")): kw["ParallelMode"] = 1
    rc = rng.random()
    if rc < 0.2: kw.update(RCFlag=1, targetBR=f"{int(rng.integers(300, 1500))} kbps", qp_rate_tables=TAB)
    elif rc < 0.3: kw.update(RCFlag=2, targetBR=f"{int(rng.integers(300, 1500))} kbps", qp_rate_tables=TAB,
                             intra_thresh=int(rng.integers(200, 6000)))      # scene-cut re-encode (Encoder.py:1851-1856)
    kind = str(rng.choice(["translating", "zooming", "flat_ties"]))
    frames = synth.make(kind, F=F, H=H, W=W, seed=int(rng.integers(0, 1000)))
    try:
        try:
            o = co.OracleCodec(H, W, F, y_only_frame_arr=frames, **kw).encode()
        except TypeError:
            skipped += 1        # rate table has no QP under the row budget: the reference crashes the same way
            continue
        e = dict(kw)
        c = Y_Video_codec(H, W, F, e.pop("block_size"), e.pop("search_range"), e.pop("Qp"), e.pop("intra_dur"), 0, y_only_frame_arr=frames, **e)
        sea = bs == 16 and r == 16 and not kw.get("VBSEnable") and rng.random() < 0.5
        c.sea_prune = bool(sea)
        c.encode()
        p = c.encoded_package.packed
        split, mv, lev = package_to_arrays(o["frame_types"], o["mvs"], o["levels"], H, W, bs)
        ok = (np.array_equal(p["split"], split) and np.array_equal(p["mv"], mv) and np.array_equal(p["levels"], lev)
              and np.array_equal(p["recon"], o["recon"]) and c.encoded_package["frame_type_seq"] == o["frame_types"])
        if ok:      # text bitstreams (host formatters and device-generated symbols) and a decode round trip on the GPU
            rcf = kw.get("RCFlag")
            want_mv = [co.mv_text_frame(t, m, q, W // bs, rcf) for t, m, q in zip(o["frame_types"], o["mvs"], o["qp_rows"])]
            want_res = [co.res_text_frame(l) for l in o["levels"]]
            mv_lines, res_lines = c.bitstream_lines()
            ok = mv_lines == want_mv and res_lines == want_res and c.residual_lines_from_symbols() == want_res
            d = dec.decoder(0, kw["intra_dur"], bs, F, H, W, kw["Qp"], kw.get("nRefFrames", 1), kw.get("FMEEnable", False), kw.get("lam"),
                            kw.get("VBSEnable", False), RCFlag=rcf, ParallelMode=kw.get("ParallelMode", 0))
            qp = c.encoded_package["Qp_per_row_per_frame"] if (rcf or 0) > 0 else None
            out = d.decode_arrays(p["frame_types"], p["split"], p["mv"], p["levels"], qp, reset_at_intra=False)
            ok = ok and np.array_equal(out, p["recon"])
    except Exception:
        traceback.print_exc()
        ok = False
    if not ok:
        bad += 1
        print("MISMATCH", n, kind, (F, H, W), kw, "sea" if c.sea_prune else "", flush=True)
print(f"{n_cases} cases ({skipped} skipped: no QP fits the rate budget), {bad} mismatches, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
