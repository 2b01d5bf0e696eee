#!/usr/bin/env python
"""Benchmark of the StreamOptima B200 encode hot path (contract: see the task statement / DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  CPU arm: the reference's own encoder on the host cores

Workload (BASELINE.json configs[1], "C2"): synthetic 1080p (coded 1920x1088) Y sequence, 300 frames, i=16, r=16 exhaustive
half-pel search over nRefFrames=4, I_Period 30; step k encodes the whole sequence at QP = k mod 12 (the QP sweep).  One
step = one pass of the hot path over that 300-frame batch.  With N GPUs every rank encodes its own stream of the same shape
(weak scaling, no data-path collective; the per-frame statistics two-pass rate control consumes are all-gathered over NCCL
inside the timed region).  On top of the contract line the same JSON carries:
  * ``c5_sharded``  -- BASELINE configs[4]: a FIXED batch of 32 4K streams x 32 frames, I_Period 16, closed GOPs dealt to the
                       ranks (strong scaling; this is the multi-GPU number that can fail);
  * ``e2e_full_flow`` -- ``encode()`` + ``transmit_bitstream()`` of the C2 sequence: what the reference's main.py does;
  * ``c1``          -- BASELINE configs[0] (CIF) through ``encode()`` on the GPU next to the unmodified reference on the host.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(W=1920, H=1088, F=300, bs=16, r=16, nref=4, fme=True, intra_dur=30)
C1 = dict(W=352, H=288, F=10, bs=8, r=2, qp=6, intra_dur=8)
C5 = dict(W=3840, H=2160, F=32, S=32, bs=16, r=16, qp=4, intra_dur=16)
METRIC = "1080p_encode_frames_per_s"
# measured by tools/int_peak.cu on this pool's B200 (profiles/int_peak_r01.json): VABSDIFF4.U8.ACC issues at
# 64 lanes/clk/SM -> 18.33e12 lane-instructions/s at 1.965 GHz; plain IADD reaches 36.3e12 (both integer pipes)
INT_PEAK_FILE = os.path.join(ROOT, "profiles", "int_peak_r01.json")
# CPU sample of the C2 geometry: frames 0-4 (I + four P frames searching 2, 3, 4, 4 references: the last two are the steady
# state) of a small window, same i / r / half-pel / nRefFrames
CPU_SAMPLE = dict(H=64, W=96, F=5)


def synth_frames_torch(F, H, W, seed, device):
    """Translating texture + noise (streamoptima_b200/synth.py formula) generated on the GPU for speed."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    yy = torch.arange(H, device=device, dtype=torch.float32)[:, None]
    xx = torch.arange(W, device=device, dtype=torch.float32)[None, :]
    out = torch.empty((F, H, W), dtype=torch.uint8, device=device)
    for t in range(F):
        f = (torch.sin((xx + 2 * t) / 7.0) + torch.cos((yy + t) / 5.0)) * 50 + 128
        f = f + torch.randn((H, W), generator=g, device=device) * 3.0
        out[t] = f.clamp(0, 255).to(torch.uint8)
    return out


def synth_frames_numpy(F, H, W, seed):
    from streamoptima_b200 import synth
    return synth.translating(F, H, W, seed=seed)


def me_work(W, H, bs, r, fme, F, intra_dur, nref):
    """Algorithmic SAD pixel-ops W_me = sum over P frames, blocks, refs of N_valid * bs^2, using the
    validity tests of Encoder.py:695-698.  -> (total, exhaustive-search launches, work of one steady-state P frame)."""
    R = 2 * r if fme else r

    def count(pos, size_px):
        size = 2 * size_px - 1 if fme else size_px
        p0 = 2 * pos if fme else pos
        lo = -p0
        hi = size - bs - 1 - p0
        if fme:
            hi = min(hi, size - 3 * bs - 1 - p0)
        lo, hi = max(lo, -R), min(hi, R)
        return max(0, hi - lo + 1)
    nx = sum(count(x, W) for x in range(0, W, bs))
    ny = sum(count(y, H) for y in range(0, H, bs))
    per_ref = nx * ny * bs * bs
    total, launches = 0, 0
    nlist = 1                                   # ref_frames = [128 frame] (Encoder.py:1798); FIFO of nRefFrames (:1864-1867)
    for f in range(F):
        if f % intra_dur != 0:
            total += per_ref * nlist
            launches += 1
        if f < F - 1:
            nlist = min(nlist + 1, nref)
    return total, launches, per_ref * nref


def me_work_per_sequence(cfg):
    t, l, _ = me_work(cfg["W"], cfg["H"], cfg["bs"], cfg["r"], cfg["fme"], cfg["F"], cfg["intra_dur"], cfg["nref"])
    return t, l


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        self.index = index

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own Python encoder (byte code in oracle/_ref, made by oracle/build_ref.py in the build container)
# when it is there, else the oracle port.  The only place besides tests/ and smoke() that may execute oracle/.
# ---------------------------------------------------------------------------------------------------------------------
def _reference_kind():
    from oracle import build_ref
    return "reference" if build_ref.load_compiled_reference() is not None else "port"


def _encode_cpu(frames, bs, r, qp, intra_dur, **kw):
    """One whole encode() of ``frames`` with the unmodified reference (or the port).  -> seconds."""
    import contextlib
    import io
    from oracle import build_ref
    F, H, W = frames.shape
    ref = build_ref.load_compiled_reference()
    t0 = time.perf_counter()
    if ref is not None:
        enc_mod, _ = ref
        old = os.getcwd()
        with tempfile.TemporaryDirectory() as d:            # the reference writes ./yuv/y_only_reconstructed.yuv (Encoder.py:1894)
            os.makedirs(os.path.join(d, "yuv"))
            os.chdir(d)
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    c = enc_mod.Y_Video_codec(H, W, F, bs, r, qp, intra_dur, 0, y_only_frame_arr=frames, **kw)
                    if kw.get("nRefFrames", 1) > 1:
                        c.decoder.decode = lambda *a, **k: None     # the throw-away internal decode raises for nRef > 1 (quirk Q7)
                    c.encode()
            finally:
                os.chdir(old)
    else:
        from oracle import codec_oracle as co
        co.OracleCodec(H, W, F, bs, r, qp, intra_dur, 0, y_only_frame_arr=frames, **kw).encode()
    return time.perf_counter() - t0


def _cpu_worker(job):
    crop, cfg = job
    return _encode_cpu(crop, cfg["bs"], cfg["r"], 4, cfg["intra_dur"], nRefFrames=cfg["nref"], FMEEnable=cfg["fme"])


def cpu_sample(frames_np, cfg, procs=None):
    """The reference encoder on a bounded sample of the C2 workload: frames 0-4 of a CPU_SAMPLE window per worker process
    (the reference is single-threaded in ParallelMode 0; one independent window per host core uses the machine).  The
    sample's time is almost all candidate evaluations, so it is scaled to the full frame by SAD work:
    frames/s = (sample SAD work x workers / wall) / (SAD work of one steady-state 1080p P frame with 4 references).
    Returns (frames/s, description, seconds, workers)."""
    import multiprocessing as mp
    ch, cw, cf = CPU_SAMPLE["H"], CPU_SAMPLE["W"], CPU_SAMPLE["F"]
    H, W = cfg["H"], cfg["W"]
    ncpu = os.cpu_count() or 1
    procs = procs or max(1, min(ncpu, 64))
    spots = [(y, x) for y in range(0, H - ch + 1, ch) for x in range(0, W - cw + 1, cw)]
    jobs = []
    for i in range(procs):
        y, x = spots[(i * 7) % len(spots)]
        jobs.append((np.ascontiguousarray(frames_np[:cf, y:y + ch, x:x + cw]), cfg))
    kind = _reference_kind()
    t0 = time.perf_counter()
    if procs == 1:
        _cpu_worker(jobs[0])
    else:
        with mp.get_context("spawn").Pool(procs) as pool:      # spawn: the parent may hold a CUDA context
            pool.map(_cpu_worker, jobs, chunksize=1)
    dt = time.perf_counter() - t0
    w_sample, _, _ = me_work(cw, ch, cfg["bs"], cfg["r"], cfg["fme"], cf, cfg["intra_dur"], cfg["nref"])
    _, _, w_frame = me_work(W, H, cfg["bs"], cfg["r"], cfg["fme"], cfg["F"], cfg["intra_dur"], cfg["nref"])
    fps = procs * w_sample / dt / w_frame
    what = ("the UNMODIFIED reference encoder (oracle/_ref byte code of /root/reference, Y_Video_codec.encode())" if kind == "reference"
            else "oracle/codec_oracle.py (NumPy port of the reference; oracle/_ref absent)")
    desc = (f"{what}, single-threaded like ParallelMode 0, run as {procs} independent worker processes ({ncpu} host CPUs), each on "
            f"frames 0-{cf - 1} (I + P frames searching 2,3,4,4 references) of a {cw}x{ch} window of the C2 sequence, i={cfg['bs']} "
            f"r={cfg['r']} half-pel nRef={cfg['nref']} QP=4: {dt:.1f} s wall; scaled to the full frame by SAD work "
            f"({w_sample:.3e} pixel-SADs per window vs {w_frame:.3e} per steady-state 1080p P frame) -- labelled extrapolation")
    return fps, desc, dt, procs, kind


def cpu_c1_full():
    """BASELINE configs[0] run IN FULL by the reference on one host core (it is its own CPU-runnable case)."""
    frames = synth_frames_numpy(C1["F"], C1["H"], C1["W"], seed=0)
    dt = _encode_cpu(frames, C1["bs"], C1["r"], C1["qp"], C1["intra_dur"])
    return {"frames_per_s": C1["F"] / dt, "seconds": dt, "cores": 1, "kind": _reference_kind(),
            "what": "C1: CIF 352x288, 10 frames, i=8, r=2, QP=6, I_Period=8, whole encode() incl. its internal decode"}


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm pieces
# ---------------------------------------------------------------------------------------------------------------------
def gpu_c1(device):
    """C1 through the drop-in class on the GPU: wall time of encode() (host arrays in, package out)."""
    from streamoptima_b200.Encoder import Y_Video_codec
    frames = synth_frames_numpy(C1["F"], C1["H"], C1["W"], seed=0)
    c = Y_Video_codec(C1["H"], C1["W"], C1["F"], C1["bs"], C1["r"], C1["qp"], C1["intra_dur"], 0, y_only_frame_arr=frames)
    c.device = device
    c.encode()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        psnr = c.encode()
    dt = (time.perf_counter() - t0) / reps
    return {"frames_per_s": C1["F"] / dt, "seconds": dt, "psnr_first": psnr[0],
            "what": "C1 through Y_Video_codec.encode() on the GPU (wall, H2D + D2H inside), mean of 20 calls"}


def full_flow(codec, frames, qp):
    """encode() + transmit_bitstream() of the C2 sequence (what the reference's main.py:9-73 does with the class)."""
    codec.const_init_Qp = qp
    codec.y_only_f_arr = frames
    with tempfile.TemporaryDirectory() as d:
        mvf, rsf = os.path.join(d, "mv.txt"), os.path.join(d, "res.txt")
        t0 = time.perf_counter()
        codec.encode()
        t1 = time.perf_counter()
        codec.transmit_bitstream(mv_file=mvf, residual_file=rsf)
        t2 = time.perf_counter()
        nbytes = os.path.getsize(mvf) + os.path.getsize(rsf)
    F = frames.shape[0]
    return {"frames_per_s": F / (t2 - t0), "encode_s": t1 - t0, "transmit_bitstream_s": t2 - t1, "text_bytes": nbytes, "qp": qp,
            "what": "Y_Video_codec.encode() + transmit_bitstream() on the 300-frame C2 sequence, host arrays in, both text files "
                    "written (tmpfs/disk of the box), residual text formatted from the packed symbols on host threads"}


def sea_experiment(frames, cfg, device, steps=2):
    """VERDICT r1 item 6(c), reported on its own and never folded into `value` / `roofline`: the same C2 steps with the
    exhaustive search pruned by SAD lower bounds (Y_Video_codec.sea_prune -> SO_FLAG_SEA, csrc/so_me_sea.cuh).  The output is
    the plain search's bit for bit (checked here on split / MVs / statistics); what changes is how many candidates get an exact SAD."""
    from streamoptima_b200 import _native
    from streamoptima_b200.Encoder import Y_Video_codec
    F, H, W = frames.shape
    res = {}
    sig = {}
    for sea in (False, True):
        c = Y_Video_codec(H, W, F, cfg["bs"], cfg["r"], 4, cfg["intra_dur"], 0, nRefFrames=cfg["nref"], FMEEnable=cfg["fme"])
        c.device = device
        c.sea_prune = sea
        kw = dict(want_levels=False, want_recon=False, want_symbols=True)
        out = c.encode_arrays(frames, **kw)
        sig[sea] = (np.array(out["split"]), np.array(out["mv"]), np.array(out["stats"]["sse"]), np.array(out["stats"]["qsize"]))
        c.encode_arrays(frames, **kw)
        t0 = time.perf_counter()
        for k in range(steps):
            c.const_init_Qp = 4 + k
            c.encode_arrays(frames, **kw)
        e2e = (time.perf_counter() - t0) / steps
        ctx = c._ctx
        _native.check(ctx.handle, ctx.lib.so_seq_upload(ctx.handle, frames.ctypes.data, 1, F))
        _native.check(ctx.handle, ctx.lib.so_seq_sync(ctx.handle))
        ms = 0.0
        s0 = ctx.sea_stats()
        for k in range(steps):
            ctx.set_qp(4 + k)
            _native.check(ctx.handle, ctx.lib.so_seq_run(ctx.handle))
            ms += ctx.last_timing()["device_ms"]
        s1 = ctx.sea_stats()
        res["pruned" if sea else "plain"] = {"frames_per_s": steps * F / (ms / 1e3), "e2e_frames_per_s": F / e2e}
        if sea:
            w_me, me_l = me_work_per_sequence(cfg)
            res["exact_sads_per_p_frame"] = (s1["exact_sads"] - s0["exact_sads"]) / max(1, s1["p_frames"] - s0["p_frames"])
            res["candidates_per_p_frame"] = w_me / (cfg["bs"] * cfg["bs"]) / me_l          # what the plain search evaluates (algorithmic)
        c._ctx.close()
        c._ctx = None
    res["identical_output"] = all(np.array_equal(a, b) for a, b in zip(sig[False], sig[True]))
    res["speedup"] = res["pruned"]["frames_per_s"] / res["plain"]["frames_per_s"]
    res["what"] = ("C2 at QP 4-5, kernels alone (CUDA events of so_seq_run, frames resident) and end to end (encode_arrays), plain vs pruned "
                   "exhaustive search.  Content dependent: this is the translating texture of the bench; on the zooming texture the "
                   "pruned search is 3 % slower than the plain one, on C5 (integer search, 1 reference) it is 1.14x faster "
                   "(profiles/r02_sea_experiment.json).  Off by default; roofline.achieved above is the plain kernel only")
    return res


def c5_sharded(rank, world, local_rank, dist, reps=2):
    """BASELINE configs[4]: FIXED batch (strong scaling) of C5.S 4K streams x C5.F frames, I_Period 16, nRefFrames 1 (closed GOPs,
    quirk Q7), r=16 integer search; the S*F/16 GOPs are dealt round-robin to the ranks (streamoptima_b200/sharding.py), every
    rank encodes its GOPs as one batched call, statistics are all-gathered over NCCL.  -> dict (rank 0) or None."""
    import torch
    from streamoptima_b200 import sharding
    from streamoptima_b200.Encoder import Y_Video_codec
    dev = torch.device("cuda", local_rank)
    H, W, F, S, IP = C5["H"], C5["W"], C5["F"], C5["S"], C5["intra_dur"]
    units = sharding.plan_units(S, F, IP, 1)
    mine = sharding.assign(units, world)[rank]
    pinned = torch.empty((len(mine), IP, H, W), dtype=torch.uint8, pin_memory=True)
    cache = {}
    for k, ui in enumerate(mine):           # every rank synthesises only the GOPs it owns (same generator, seed = stream)
        u = units[ui]
        if u.stream not in cache:
            cache = {u.stream: synth_frames_torch(F, H, W, seed=u.stream, device=dev)}
        pinned[k].copy_(cache[u.stream][u.start:u.start + u.length])
    del cache
    torch.cuda.empty_cache()
    c = Y_Video_codec(H, W, IP, C5["bs"], C5["r"], C5["qp"], IP, 0)
    c.device = local_rank
    frames = pinned.numpy()
    kw = dict(want_levels=False, want_recon=False, want_symbols=True)
    out = None
    for _ in range(3):          # warm-up like the timed loop runs: the previous result is alive during the call, so two sets of
        out = c.encode_arrays(frames, **kw)      # pinned result buffers come into being here, not in the timed region

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(reps):
        out = c.encode_arrays(frames, **kw)
        d2h += out["sym_needed"] * 2 + out["split"].nbytes + out["mv"].nbytes + out["row_sizes"].nbytes + out["stats"].nbytes
        if world > 1:       # the statistics two-pass rate control consumes, all-gathered
            st = out["stats"]
            mine_t = torch.zeros((len(units) // world + 1, IP, 2), dtype=torch.int64, device=dev)
            mine_t[:len(mine)] = torch.from_numpy(np.stack([st["qsize"].astype(np.int64), st["sse"].astype(np.int64)], axis=-1)).to(dev)
            parts = [torch.empty_like(mine_t) for _ in range(world)]
            dist.all_gather(parts, mine_t)
    barrier()
    wall = time.perf_counter() - t0
    # kernels alone: the same batch resident in HBM (no copies inside the region), CUDA events of so_seq_run
    from streamoptima_b200 import _native
    ctx = c._ctx
    _native.check(ctx.handle, ctx.lib.so_seq_upload(ctx.handle, frames.ctypes.data, len(mine), IP))
    _native.check(ctx.handle, ctx.lib.so_seq_sync(ctx.handle))
    _native.check(ctx.handle, ctx.lib.so_seq_run(ctx.handle))
    ctx.last_timing()
    barrier()
    dev_ms = fin_ms = 0.0
    fin_launches = 0
    for _ in range(reps):
        _native.check(ctx.handle, ctx.lib.so_seq_run(ctx.handle))
        t = ctx.last_timing()
        dev_ms += t["device_ms"]; fin_ms += t["finish_inter_ms"]; fin_launches += t["finish_inter_launches"]
    barrier()
    tm = torch.tensor([dev_ms, wall * 1e3, float(d2h), fin_ms], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tm.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tm.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = sm = tm
    c._ctx.close()
    if rank != 0:
        return None
    total = S * F * reps
    e2e_s = float(mx[1]) / 1e3
    hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    # rank 0's inter finish kernel (transform / quantisation / RLE size / reconstruction), one launch = one P frame of all its units:
    # 5*H*W algorithmic bytes per frame and unit (cur 1 + predictor 1 + levels 2 + recon 1)
    fin_gbs = 5.0 * H * W * len(mine) * fin_launches / (fin_ms / 1e3) / 1e9 if fin_ms > 0 else None
    return {"workload": f"C5: {S} x 4K (3840x2160) streams x {F} frames, I_Period {IP}, i=16 r=16 integer search, nRef=1, QP {C5['qp']}: "
                        f"{len(units)} closed GOPs dealt round-robin to {world} rank(s), one batched call per rank",
            "scaling": "strong", "n_gpus": world, "units": len(units), "reps": reps,
            "device_frames_per_s": total / (float(mx[0]) / 1e3), "e2e_frames_per_s": total / e2e_s,
            "device_note": "kernels alone: the rank's GOPs resident in HBM, CUDA events around so_seq_run, max over ranks",
            "h2d_GBps_aggregate": total * H * W / e2e_s / 1e9, "d2h_GBps_aggregate": float(sm[2]) / e2e_s / 1e9,
            "d2h_bytes_per_frame": float(sm[2]) / total, "h2d_bytes_per_frame": H * W,
            "limiter": "e2e is bounded by the host->device copy of the raw frames: 8.3 MB in per 4K frame against %.1f MB out; h2d_GBps_aggregate / "
                       "n_gpus is the rate every rank's copies reach while all ranks share the host's memory system (one rank alone: ~21 GB/s)"
                       % (float(sm[2]) / total / 1e6),
            "roofline_transform_batched": {"bound": "hbm by bytes; instruction-bound in practice (FP64 replay of SciPy's DCT: ~1100 warp instructions per block, "
                                                    "DRAM traffic = algorithmic bytes, profiles/r02_inter_finish16_batched_ncu_summary.txt)",
                                           "kernel": "inter_finish16_kernel<false>, %d units (4K frames) per launch" % len(mine),
                                           "achieved": fin_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": fin_gbs / hbm_peak if fin_gbs else None,
                                           "algorithmic_bytes_per_frame": 5 * H * W, "launches_timed": fin_launches}}


def main():
    # the contract is ONE JSON line on stdout: anything libraries print there (NCCL's version banner, ...) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=CFG["F"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip c5_sharded / e2e_full_flow / c1 (the contract line only)")
    args = ap.parse_args()
    cfg = dict(CFG)
    cfg["F"] = args.frames
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "C2: synthetic 1080p (coded 1920x1088) Y, %d frames, i=16, r=16 exhaustive half-pel ME, nRefFrames=4, "
                          "I_Period=30, QP sweep (step k -> QP k mod 12)" % cfg["F"],
              "width": cfg["W"], "height": cfg["H"], "frames_per_step": cfg["F"], "block_size": cfg["bs"],
              "search_range": cfg["r"], "nRefFrames": cfg["nref"], "FMEEnable": True, "I_Period": cfg["intra_dur"],
              "parallelism": f"{args.gpus} independent stream(s), one per GPU" if args.gpus > 1 else "single GPU",
              "l2_policy": "inputs (%.0f MB/step) larger than the 126 MB L2" % (cfg["F"] * cfg["H"] * cfg["W"] / 1e6)}

    if args.impl == "reference":
        if rank != 0:
            return
        frames = synth_frames_numpy(CPU_SAMPLE["F"], cfg["H"], cfg["W"], seed=0)
        for _ in range(min(args.warmup, 1)):
            cpu_sample(frames, cfg)
        t0 = time.perf_counter()
        vals = [cpu_sample(frames, cfg) for _ in range(args.steps)]
        dt = time.perf_counter() - t0
        fps = float(np.mean([v[0] for v in vals]))
        emit({"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
              "config": config,
              "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": vals[0][3], "kind": vals[0][4], "sample": vals[0][1]},
              "c1_reference_full": cpu_c1_full(),
              "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    import torch
    import torch.distributed as dist
    from streamoptima_b200 import _native
    from streamoptima_b200.Encoder import Y_Video_codec

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    F, H, W = cfg["F"], cfg["H"], cfg["W"]
    frames_t = synth_frames_torch(F, H, W, seed=rank, device=dev)
    frames_pinned = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True)       # e2e inputs live in pinned host memory
    frames_pinned.copy_(frames_t)
    frames = frames_pinned.numpy()
    del frames_t
    torch.cuda.empty_cache()

    Y_Video_codec.write_recon_yuv = False
    codec = Y_Video_codec(H, W, F, cfg["bs"], cfg["r"], 0, cfg["intra_dur"], 0, nRefFrames=cfg["nref"], FMEEnable=cfg["fme"],
                          y_only_frame_arr=frames)
    codec.device = local_rank
    ctx = codec._context(cfg["bs"], cfg["r"], cfg["intra_dur"], max_batch=1)
    lib = ctx.lib

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stats_host = np.zeros((1, F), dtype=_native.STATS_DTYPE)

    def step_resident(k):
        ctx.set_qp(k % 12)
        _native.check(ctx.handle, lib.so_seq_run(ctx.handle))
        t = ctx.last_timing()                       # waits for the step's last event
        if world > 1:
            # pass-1 statistics (quantized_sized, SSE, type per frame) that two-pass rate control consumes: all-gathered
            # over NCCL/NVLink -- a few KB, the only collective of the path
            _native.check(ctx.handle, lib.so_seq_download(ctx.handle, None, None, None, None, None, stats_host.ctypes.data))
            mine = torch.from_numpy(np.stack([stats_host["qsize"][0].astype(np.int64), stats_host["sse"][0].astype(np.int64),
                                              stats_host["frame_type"][0].astype(np.int64)], axis=1)).to(dev)
            gathered = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine)
        return t

    # ---- kernel-only: inputs resident in HBM
    _native.check(ctx.handle, lib.so_seq_upload(ctx.handle, frames.ctypes.data, 1, F))
    _native.check(ctx.handle, lib.so_seq_sync(ctx.handle))
    for k in range(args.warmup):
        step_resident(k)
    barrier()
    sampler = ClockSampler(local_rank)          # rank 0 samples its own GPU; the other ranks do not spawn nvidia-smi
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    dev_ms = me_ms = tq_ms = xs_ms = fin_ms = 0.0
    launches = me_launches = xs_launches = timed_frames = fin_launches = 0
    for k in range(args.steps):
        t = step_resident(k)
        dev_ms += t["device_ms"]; me_ms += t["me_ms"]; tq_ms += t["tq_ms"]; xs_ms += t["search_ms"]
        fin_ms += t["finish_inter_ms"]; fin_launches += t["finish_inter_launches"]
        launches += t["launches"]; me_launches += t["me_launches"]; xs_launches += t["search_launches"]
        timed_frames += t["timed_frames"]
    barrier()
    wall = time.perf_counter() - t0

    # ---- end to end through the public API: host frames in, host results out, copies inside the timed region.  The
    # residual comes back as packed run-level symbols generated on the device (the text bitstream is formatted from them);
    # the reconstruction stays in HBM as the reference frame -- encode() fetches it once, for the .yuv the reference writes
    e2e_kw = dict(want_levels=False, want_recon=False, want_symbols=True)
    for k in range(max(3, min(args.warmup, 5))):      # QP 0 has the longest symbol streams: the pinned result buffers (two sets
        codec.const_init_Qp = 0                       # alternate, a result owns its buffers) reach their final size here
        warm = codec.encode_arrays(frames, **e2e_kw)
    del warm
    barrier()
    t1 = time.perf_counter()
    e2e_dev_ms = 0.0
    d2h = 0
    e2e_launches = 0
    for k in range(args.steps):
        codec.const_init_Qp = k % 12
        out = codec.encode_arrays(frames, **e2e_kw)
        _ = int(out["stats"]["sse"][0, -1])
        e2e_dev_ms += codec.last_timing["device_ms"]
        e2e_launches += codec.last_timing["launches"]
        d2h += out["sym_needed"] * 2 + out["split"].nbytes + out["mv"].nbytes + out["row_sizes"].nbytes + out["stats"].nbytes
    barrier()
    e2e_wall = time.perf_counter() - t1
    sampler.stop_flag = True
    if rank == 0:
        sampler.join()
    del out

    tm = torch.tensor([dev_ms, wall * 1e3, e2e_wall * 1e3, me_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max, e2e_ms_max, me_ms_max = [float(v) for v in tm.tolist()]
    line = None
    if rank == 0:
        total_frames = args.steps * F * world
        value = total_frames / (dev_ms_max / 1e3)
        w_me, me_l = me_work_per_sequence(cfg)
        peaks = json.load(open(INT_PEAK_FILE)) if os.path.exists(INT_PEAK_FILE) else {}
        peak = peaks.get("vabsdiff4_lane_Tops", 18.33)
        peak_iadd = peaks.get("iadd_lane_Tops", 36.3)
        # rank 0's own exhaustive-search launches, CUDA events around the launch on the context stream.  The library records
        # per-kernel events on every 8th frame only (they cost ~10 us of stream serialisation per frame), so the per-launch
        # duration is the mean over those xs_launches launches of the timed region.  Work per launch = W_me / launches of the
        # whole sequence (the first nRefFrames-1 P frames search fewer references, which makes this mean 0.5 % smaller than
        # the work of the timed launches: the reported fraction errs on the low side)
        avg_launch_ms = xs_ms / max(1, xs_launches)
        achieved = (w_me / 4 / me_l) / (avg_launch_ms / 1e3) / 1e12
        traffic, traffic_detail = None, None                            # DRAM bytes per ME launch from the committed ncu capture
        tpath = next((p for p in (os.path.join(ROOT, "profiles", n) for n in ("r02_me_traffic.json", "r01_me_traffic.json")) if os.path.exists(p)), None)
        if tpath and F >= 30:
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]      # bytes per launch, one ncu --set full capture
            algo = (cfg["nref"] + 1) * H * W         # the current frame and nRefFrames reference frames, once
            traffic_detail = {"unit": "B per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                              "algorithmic_bytes_per_launch": algo, "traffic_over_algorithmic": traffic / algo,
                              "design_footprint_bytes_per_launch": tj.get("algorithmic_bytes_per_launch"),
                              "source": tj["source"],
                              "note": "the search reads 16 planes per reference (4 half-pel phases x 4 byte-shifted copies, the price of "
                                      "16-byte aligned TMA boxes): %.1fx the algorithmic bytes, each plane once per launch; at < 3 %% of "
                                      "DRAM peak it costs no time in this ALU-bound kernel" % (traffic / algo)}
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "wall_ms_per_step": wall_ms_max / args.steps,
                "roofline": {"bound": "int32-alu (VABSDIFF4 pipe)", "kernel": "me_ring2_kernel<false>", "achieved": achieved,
                             "peak": peak, "unit": "T lane-instr/s (1 instr = 4 pixel SADs)", "frac": achieved / peak,
                             "peak_int32": peak_iadd, "frac_int32": achieved / peak_iadd,
                             "peak_source": "measured: tools/int_peak.cu, profiles/int_peak_r01.json (MEASURED_PEAKS.json has no integer figure); "
                                            "frac = against the VABSDIFF4.U8.ACC issue rate (64 lanes/clk/SM, the instruction that does the work), "
                                            "frac_int32 = against plain IADD on both integer pipes (128 lanes/clk/SM)",
                             "algorithmic_sad_pixel_ops_per_step": w_me, "launches_per_step": me_l,
                             "avg_launch_ms": avg_launch_ms, "launches_timed": xs_launches,
                             "me_share_of_step": avg_launch_ms * me_l * args.steps / dev_ms,
                             "traffic": traffic, "traffic_detail": traffic_detail},
                "roofline_transform": {"bound": "hbm", "kernel": "inter_finish16_kernel<false>", "achieved": 5.0 * H * W * fin_launches / (fin_ms / 1e3) / 1e9,
                                       "peak": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
                                       if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0,
                                       "unit": "GB/s", "launches_timed": fin_launches,
                                       "note": "5*H*W algorithmic bytes per P frame (SURVEY 8d) over the CUDA-event time of the finish kernel's timed launches; ONE 1080p frame per launch (1020 CTAs) -- c5_sharded.roofline_transform_batched is the same kernel on a batched launch"},
                "e2e": {"value": total_frames / (e2e_ms_max / 1e3), "unit": "frames/s", "h2d_bytes_per_step": F * H * W,
                        "d2h_bytes_per_step": d2h / args.steps, "wall_ms_per_step": e2e_ms_max / args.steps,
                        "device_ms_per_step_rank0": e2e_dev_ms / args.steps,
                        "api": "Y_Video_codec.encode_arrays(frames, want_levels=False, want_recon=False, want_symbols=True): pinned host frames in; "
                               "split / MVs / packed run-level symbols / row sizes / statistics out"},
                "gpu_launches": launches + e2e_launches, "clocks": sampler.summary()}
        line["roofline_transform"]["frac"] = line["roofline_transform"]["achieved"] / line["roofline_transform"]["peak"]
        if world == 1 and not args.no_cpu_baseline:
            fps, desc, dt, procs, kind = cpu_sample(frames, cfg)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": procs, "kind": kind, "sample": desc}
    if not args.no_extras:
        if rank == 0 and world == 1:
            for _ in range(2):                               # warm-up: a package owns its buffers and the previous package is alive during
                full_flow(codec, frames, 4)                  # the next encode(), so two sets of pinned buffers come into being here
            line["e2e_full_flow"] = full_flow(codec, frames, 4)
            line["c1"] = {"gpu": gpu_c1(local_rank)}
            line["sea_experiment"] = sea_experiment(frames, cfg, local_rank)
            if not args.no_cpu_baseline:
                line["c1"]["cpu"] = cpu_c1_full()
        # free the C2 buffers, then the fixed 4K batch sharded over all ranks
        codec._ctx.close()
        codec._ctx = None
        del frames, frames_pinned
        Y_Video_codec._pool.clear()
        torch.cuda.empty_cache()
        c5 = c5_sharded(rank, world, local_rank, dist)
        if rank == 0:
            line["c5_sharded"] = c5
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
