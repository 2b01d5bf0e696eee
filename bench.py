#!/usr/bin/env python
"""Benchmark of the StreamOptima B200 encode hot path (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  CPU arm: the oracle port of the reference's algorithm

Workload (BASELINE.json configs[1]): synthetic 1080p (coded 1920x1088) Y sequence, 300 frames, i=16, r=16 exhaustive
half-pel search over nRefFrames=4, I_Period 30; step k encodes the whole sequence at QP = k mod 12 (the QP sweep).
One step = one pass of the hot path over that 300-frame batch.  With N GPUs every rank encodes its own stream of the
same shape (weak scaling, no data-path collective) and the per-frame statistics that two-pass rate control consumes
are all-gathered over NCCL inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(W=1920, H=1088, F=300, bs=16, r=16, nref=4, fme=True, intra_dur=30)
METRIC = "1080p_encode_frames_per_s"
# measured by tools/int_peak.cu on this pool's B200 (profiles/int_peak_r01.json): VABSDIFF4.U8.ACC issues at
# 64 lanes/clk/SM -> 18.33e12 lane-instructions/s at 1.965 GHz; plain IADD reaches 36.3e12 (both integer pipes)
INT_PEAK_FILE = os.path.join(ROOT, "profiles", "int_peak_r01.json")


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


def me_work_per_sequence(cfg):
    """Algorithmic SAD pixel-ops W_me = sum over P frames, blocks, refs of N_valid * bs^2, using the
    validity tests of Encoder.py:695-698, and the number of exhaustive-search launches."""
    W, H, bs, r = cfg["W"], cfg["H"], cfg["bs"], cfg["r"]
    fme = cfg["fme"]
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
    nlist = 1
    for f in range(cfg["F"]):
        if f % cfg["intra_dur"] != 0:
            total += per_ref * nlist
            launches += 1
        if f < cfg["F"] - 1:
            nlist = min(nlist + 1, cfg["nref"]) if nlist < cfg["nref"] else cfg["nref"]
    return total, launches


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


def _cpu_worker(job):
    """One oracle encode of a crop (runs in a worker process)."""
    crop, cfg = job
    from oracle import codec_oracle as co
    cf, ch, cw = crop.shape
    t0 = time.perf_counter()
    co.OracleCodec(ch, cw, cf, cfg["bs"], cfg["r"], 4, cfg["intra_dur"], 0, nRefFrames=cfg["nref"], FMEEnable=cfg["fme"],
                   y_only_frame_arr=crop).encode()
    return time.perf_counter() - t0


def cpu_sample(frames_np, cfg, procs=None):
    """Oracle port of the reference's algorithm on a bounded sample of the same workload: 3 frames (I, P, P) of 640x256
    windows, same i / r / nRef / half-pel.  The reference is single-threaded (ParallelMode 0); to use the host's cores the
    sample runs one independent window per worker process.  Returns (frames/s extrapolated to the full frame, description,
    seconds, workers)."""
    import multiprocessing as mp
    ch, cw, cf = 256, 640, 3
    H, W = cfg["H"], cfg["W"]
    ncpu = os.cpu_count() or 1
    procs = procs or max(1, min(ncpu, 64))
    spots = [(y, x) for y in range(0, H - ch + 1, ch) for x in range(0, W - cw + 1, cw)]        # 4 x 3 distinct windows at 1080p
    jobs = []
    for i in range(procs):
        y, x = spots[i % len(spots)]
        jobs.append((np.ascontiguousarray(frames_np[:cf, y:y + ch, x:x + cw]), cfg))
    t0 = time.perf_counter()
    if procs == 1:
        _cpu_worker(jobs[0])
    else:
        with mp.get_context("spawn").Pool(procs) as pool:      # spawn: the parent may hold a CUDA context
            pool.map(_cpu_worker, jobs, chunksize=1)
    dt = time.perf_counter() - t0
    frac = (ch * cw) / (H * W)
    fps = procs * cf * frac / dt
    desc = (f"oracle/codec_oracle.py (NumPy/SciPy port of the reference; single-threaded like the reference's ParallelMode 0) run as "
            f"{procs} independent worker processes ({ncpu} host CPUs), each on frames 0-2 (I,P,P) of a {cw}x{ch} window, i={cfg['bs']} "
            f"r={cfg['r']} half-pel nRef={cfg['nref']} QP=4: {dt:.1f} s wall; frames/s = workers x 3 frames x window area / frame area "
            f"({frac:.4f}) / wall -- labelled extrapolation")
    return fps, desc, dt, procs


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
        frames = synth_frames_numpy(3, cfg["H"], cfg["W"], seed=0)
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
                          "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": vals[0][3], "kind": "port", "sample": vals[0][1]},
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
    dev_ms = me_ms = tq_ms = xs_ms = 0.0
    launches = me_launches = xs_launches = timed_frames = 0
    for k in range(args.steps):
        t = step_resident(k)
        dev_ms += t["device_ms"]; me_ms += t["me_ms"]; tq_ms += t["tq_ms"]; xs_ms += t["search_ms"]
        launches += t["launches"]; me_launches += t["me_launches"]; xs_launches += t["search_launches"]
        timed_frames += t["timed_frames"]
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    if rank == 0:
        sampler.join()

    # ---- end to end through the public API: host frames in, host results out, copies inside the timed region
    for k in range(min(args.warmup, 2)):
        codec.const_init_Qp = k % 12
        codec.encode_arrays(frames)
    barrier()
    t1 = time.perf_counter()
    e2e_dev_ms = 0.0
    for k in range(args.steps):
        codec.const_init_Qp = k % 12
        out = codec.encode_arrays(frames)
        _ = int(out["stats"]["sse"][0, -1])
        e2e_dev_ms += codec.last_timing["device_ms"]
    barrier()
    e2e_wall = time.perf_counter() - t1
    nblk = (H // cfg["bs"]) * (W // cfg["bs"])
    d2h = F * (H * W * 3 + nblk * (1 + 24) + (H // cfg["bs"]) * 4 + 32)

    tm = torch.tensor([dev_ms, wall * 1e3, e2e_wall * 1e3, me_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max, e2e_ms_max, me_ms_max = [float(v) for v in tm.tolist()]
    if rank == 0:
        total_frames = args.steps * F * world
        value = total_frames / (dev_ms_max / 1e3)
        w_me, me_l = me_work_per_sequence(cfg)
        peaks = json.load(open(INT_PEAK_FILE)) if os.path.exists(INT_PEAK_FILE) else {}
        peak = peaks.get("vabsdiff4_lane_Tops", 18.33)
        # rank 0's own exhaustive-search launches, CUDA events around the launch on the context stream.  The library records
        # per-kernel events on every 8th frame only (they cost ~10 us of stream serialisation per frame), so the per-launch
        # duration is the mean over those xs_launches launches of the timed region.  Work per launch = W_me / launches of the
        # whole sequence (the first nRefFrames-1 P frames search fewer references, which makes this mean 0.5 % smaller than
        # the work of the timed launches: the reported fraction errs on the low side)
        avg_launch_ms = xs_ms / max(1, xs_launches)
        achieved = (w_me / 4 / me_l) / (avg_launch_ms / 1e3) / 1e12
        traffic, traffic_detail = None, None                            # DRAM bytes per ME launch from the committed ncu capture
        tpath = os.path.join(ROOT, "profiles", "r01_me_traffic.json")
        if os.path.exists(tpath) and F >= 30:
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]      # bytes per launch, one ncu --set full capture
            traffic_detail = {"unit": "B per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                              "algorithmic_bytes_per_launch": tj.get("algorithmic_bytes_per_launch"), "source": tj["source"],
                              "note": tj.get("note")}
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "wall_ms_per_step": wall_ms_max / args.steps,
                "roofline": {"bound": "int32-alu (VABSDIFF4 pipe)", "kernel": "me_ring_kernel<false>", "achieved": achieved,
                             "peak": peak, "unit": "T lane-instr/s (1 instr = 4 pixel SADs)", "frac": achieved / peak,
                             "peak_source": "measured: tools/int_peak.cu vabsdiff4.add, profiles/int_peak_r01.json (MEASURED_PEAKS.json has no integer figure)",
                             "algorithmic_sad_pixel_ops_per_step": w_me, "launches_per_step": me_l,
                             "avg_launch_ms": avg_launch_ms, "launches_timed": xs_launches,
                             "me_share_of_step": avg_launch_ms * me_l * args.steps / dev_ms,
                             "traffic": traffic, "traffic_detail": traffic_detail},
                "roofline_transform": {"bound": "hbm", "achieved": 5.0 * H * W * timed_frames / (tq_ms / 1e3) / 1e9,
                                       "peak": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
                                       if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0,
                                       "unit": "GB/s", "note": "5*H*W algorithmic bytes per frame (SURVEY §8d) over the transform/quant/recon kernels of the frames that carry per-kernel events; latency-bound at one 1080p frame per launch"},
                "e2e": {"value": total_frames / (e2e_ms_max / 1e3), "unit": "frames/s", "h2d_bytes_per_step": F * H * W,
                        "d2h_bytes_per_step": d2h, "wall_ms_per_step": e2e_ms_max / args.steps,
                        "device_ms_per_step_rank0": e2e_dev_ms / args.steps},
                "gpu_launches": launches, "clocks": sampler.summary()}
        line["roofline_transform"]["frac"] = line["roofline_transform"]["achieved"] / line["roofline_transform"]["peak"]
        if world == 1 and not args.no_cpu_baseline:
            fps, desc, dt, procs = cpu_sample(frames, cfg)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": procs, "kind": "port", "sample": desc}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
