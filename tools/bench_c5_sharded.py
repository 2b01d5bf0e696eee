"""BASELINE config 5 on N GPUs: a batch of 4K streams, I_Period 16, closed GOPs dealt round-robin to the ranks
(streamoptima_b200/sharding.py), statistics all-gathered over NCCL.  Launch with torchrun like bench.py:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c5_sharded.py [--streams 8]
Prints one JSON line on rank 0 (frames/s over all ranks, device time = max over ranks)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import synth_frames_torch
from streamoptima_b200 import sharding
from streamoptima_b200.Encoder import Y_Video_codec

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=8)
ap.add_argument("--frames", type=int, default=32)
ap.add_argument("--fme", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
dev = torch.device("cuda", local_rank)
H, W, F, S, IP = 2160, 3840, args.frames, args.streams, 16
units = sharding.plan_units(S, F, IP, 1)
mine = sharding.assign(units, world)[rank]
# every rank synthesises only the GOPs it owns (same generator as bench.py, seed = stream)
streams = {}
for ui in mine:
    u = units[ui]
    if u.stream not in streams:
        streams[u.stream] = synth_frames_torch(F, H, W, seed=u.stream, device=dev).cpu().numpy()
batch = np.stack([streams[units[ui].stream][units[ui].start:units[ui].start + units[ui].length] for ui in mine])
pinned = torch.empty(batch.shape, dtype=torch.uint8, pin_memory=True)
pinned.numpy()[...] = batch
Y_Video_codec.write_recon_yuv = False
c = Y_Video_codec(H, W, IP, 16, 16, 4, IP, 0, FMEEnable=bool(args.fme))
c.device = local_rank
c.encode_arrays(pinned.numpy())
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
dev_ms = 0.0
for _ in range(args.reps):
    out = c.encode_arrays(pinned.numpy())
    dev_ms += c.last_timing["device_ms"]
    if world > 1:       # the statistics two-pass rate control consumes, all-gathered
        st = out["stats"]
        mine_t = torch.from_numpy(np.stack([st["qsize"].astype(np.int64), st["sse"].astype(np.int64)], axis=-1)).to(dev)
        parts = [torch.empty_like(mine_t) for _ in range(world)]
        dist.all_gather(parts, mine_t)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
tm = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
if rank == 0:
    total = S * F * args.reps
    print(json.dumps({"config": f"C5: {S} x 4K streams, {F} frames, I_Period 16, i=16 r=16 {'half-pel' if args.fme else 'integer'}, nRef=1, "
                                f"closed GOPs round-robin over {world} GPU(s)", "n_gpus": world, "units": len(units),
                      "fps_kernel": total / (float(tm[0]) / 1e3), "fps_e2e": total / (float(tm[1]) / 1e3)}))
if world > 1:
    dist.destroy_process_group()
