"""Multi-GPU sharding of the encode path: one process per GPU, no data-path collective.

The P-frame chain inside a GOP is serial, so work shards only where it is naturally independent:
  * separate streams, and
  * closed GOPs -- frames [k*I_Period, (k+1)*I_Period) -- which are independent exactly when ``nRefFrames == 1``
    (for nRefFrames > 1 the reference never resets its reference list at I frames, quirk Q7, Encoder.py:1864-1867, so
    P frames reach across the I frame and a stream is then one indivisible unit).
Units are dealt round-robin to ranks; each rank encodes its units as ONE batched call (``Y_Video_codec.encode_arrays``
with a leading unit axis -> the kernels see them as grid.y).  The only collective is an all-gather of the per-frame
statistics (``quantized_sized``, SSE, rows) that two-pass rate control and ROI maps consume: a few KB, NCCL over NVLink
on GPUs (``gloo`` in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Sequence

import numpy as np


@dataclass(frozen=True)
class Unit:
    stream: int
    start: int      # first frame (a multiple of I_Period)
    length: int


def plan_units(n_streams: int, n_frames: int, intra_dur: int, n_ref_frames: int, parallel_mode: int = 0) -> List[Unit]:
    """Independent work units in (stream, start) order."""
    closed = n_ref_frames == 1 and parallel_mode != 1
    units = []
    for s in range(n_streams):
        if closed:
            for g0 in range(0, n_frames, intra_dur):
                units.append(Unit(s, g0, min(intra_dur, n_frames - g0)))
        else:
            units.append(Unit(s, 0, n_frames))
    return units


def assign(units: Sequence[Unit], world: int) -> List[List[int]]:
    """Round-robin: unit k -> rank k % world.  Returns unit indices per rank."""
    return [list(range(r, len(units), world)) for r in range(world)]


def encode_sharded(streams: np.ndarray, encode_units: Callable[[np.ndarray], dict], intra_dur: int, n_ref_frames: int,
                   parallel_mode: int = 0, rank: int = 0, world: int = 1, dist=None, device="cpu"):
    """Encode ``streams`` u8 [S,F,H,W] sharded over ``world`` ranks.

    ``encode_units(batch u8 [U,L,H,W]) -> dict`` encodes U independent units of equal length L (e.g. a bound
    ``Y_Video_codec.encode_arrays``) and returns at least ``stats`` (structured, [U,L]) and ``row_sizes`` ([U,L,rows]).
    Returns ``(local, gathered)``: ``local`` maps unit index -> that unit's output dict (this rank's units only);
    ``gathered`` holds, for EVERY frame of every stream, ``qsize [S,F]``, ``sse [S,F]``, ``frame_type [S,F]`` and
    ``row_sizes [S,F,rows]`` -- identical on all ranks after the all-gather.
    """
    S, F, H, W = streams.shape
    units = plan_units(S, F, intra_dur, n_ref_frames, parallel_mode)
    mine = assign(units, world)[rank]
    local = {}
    # batch units of equal length into one call
    by_len = {}
    for ui in mine:
        by_len.setdefault(units[ui].length, []).append(ui)
    for L, idxs in by_len.items():
        batch = np.stack([streams[units[ui].stream, units[ui].start:units[ui].start + L] for ui in idxs])
        out = encode_units(np.ascontiguousarray(batch))
        for k, ui in enumerate(idxs):
            local[ui] = {key: (val[k].copy() if isinstance(val, np.ndarray) and val.ndim >= 1 and val.shape[0] == len(idxs) else val)
                         for key, val in out.items()}
    rows = next(iter(local.values()))["row_sizes"].shape[-1] if local else 0
    if dist is not None and world > 1:
        import torch
        t = torch.tensor([rows], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rows = int(t.item())
    # fixed-size contribution: [S, F, 3 + rows] int64, zero where another rank owns the frame
    contrib = np.zeros((S, F, 3 + rows), np.int64)
    for ui, out in local.items():
        u = units[ui]
        st = out["stats"]
        sl = (u.stream, slice(u.start, u.start + u.length))
        contrib[sl][:, 0] = st["qsize"]
        contrib[sl][:, 1] = st["sse"].astype(np.int64)
        contrib[sl][:, 2] = st["frame_type"]
        contrib[sl][:, 3:] = out["row_sizes"]
    if dist is not None and world > 1:
        import torch
        t = torch.from_numpy(contrib).to(device)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)                 # every frame is owned by exactly one rank: the sum is the union
        contrib = sum(p.cpu().numpy() for p in parts)
    gathered = dict(qsize=contrib[..., 0], sse=contrib[..., 1], frame_type=contrib[..., 2], row_sizes=contrib[..., 3:])
    return local, gathered
