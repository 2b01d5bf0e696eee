"""Two-pass rate control and ROI QP maps built on the reference's table-driven controller.

NOT a restatement of reference code: the reference README advertises multi-pass encoding and ROI, its sources contain
neither -- "parity unpinned".  What exists in the reference, and is reproduced bit-exactly by the encoder,
is the table-driven row-level controller (``RCFlag`` 1/2, ``qp_rate_tables``, Encoder.py:1576-1609) and the statistics
it would be fed with (``quantized_sized`` and the per-row sizes, Encoder.py:1627-1639).  This module closes the loop:

  pass 1  encode the sequence at every QP of a ladder with rate control off and measure, per frame type, the mean size
          of a block row in bits (8 bits per run-level symbol, the unit of ``calculate_RD_cost``, Encoder.py:1147).
          The ladder is sharded round-robin over ranks; the per-QP means are all-gathered (NCCL over NVLink on GPUs).
  pass 2  encode with ``RCFlag`` and the measured tables at the target bitrate.
"""
from __future__ import annotations

import numpy as np

from .Encoder import Y_Video_codec


def measure_rate_tables(frames, codec_kwargs, qps=range(12), rank=0, world=1, dist=None, device="cpu", make_codec=None):
    """-> [intra_table, inter_table]: mean bits per block row at each QP (ints, what ``qp_rate_tables`` expects).

    ``codec_kwargs``: positional/keyword parameters of the codec WITHOUT Qp and rate-control arguments:
    ``dict(block_size=, search_range=, intra_dur=, nRefFrames=, FMEEnable=, ...)``."""
    frames = np.asarray(frames)
    F, H, W = frames.shape
    qps = list(qps)
    kw = dict(codec_kwargs)
    bs, r, ip = kw.pop("block_size"), kw.pop("search_range"), kw.pop("intra_dur")
    sums = np.zeros((len(qps), 2, 2), np.float64)          # [qp][type][sum_bits_per_row, n_frames]
    for i in range(rank, len(qps), world):
        mk = make_codec or (lambda qp: Y_Video_codec(H, W, F, bs, r, qp, ip, 0, **kw))
        c = mk(qps[i])
        out = c.encode_arrays(frames, want_levels=False, want_recon=False)
        types = out["stats"]["frame_type"][0]
        rows = out["row_sizes"][0].astype(np.float64) * 8.0
        for t in (0, 1):
            sel = types == t
            if sel.any():
                sums[i, t, 0] = rows[sel].mean(axis=1).sum()
                sums[i, t, 1] = sel.sum()
    if dist is not None and world > 1:
        import torch
        tt = torch.from_numpy(sums).to(device)
        parts = [torch.empty_like(tt) for _ in range(world)]
        dist.all_gather(parts, tt)                          # every QP was measured by exactly one rank
        sums = sum(p.cpu().numpy() for p in parts)
    tables = [[0] * len(qps), [0] * len(qps)]
    for i in range(len(qps)):
        for t in (0, 1):
            n = sums[i, t, 1]
            tables[t][i] = int(round(sums[i, t, 0] / n)) if n else 0
    if all(v == 0 for v in tables[1]):
        tables[1] = list(tables[0])
    return tables


def two_pass_encode(frames, target_br, codec_kwargs, rc_flag=1, intra_thresh=None, frame_rate=30, init_qp=4, qps=range(12),
                    rank=0, world=1, dist=None, device="cpu"):
    """Pass 1 + pass 2.  Returns ``(codec, tables)``; ``codec.encoded_package`` holds the pass-2 result."""
    frames = np.asarray(frames)
    F, H, W = frames.shape
    tables = measure_rate_tables(frames, codec_kwargs, qps, rank, world, dist, device)
    kw = dict(codec_kwargs)
    bs, r, ip = kw.pop("block_size"), kw.pop("search_range"), kw.pop("intra_dur")
    codec = Y_Video_codec(H, W, F, bs, r, init_qp, ip, 0, y_only_frame_arr=frames, RCFlag=rc_flag, targetBR=target_br,
                          frame_rate=frame_rate, qp_rate_tables=tables, intra_thresh=intra_thresh, **kw)
    budget = codec.bitrate_per_row
    if not any(v < budget for v in tables[0]):
        raise ValueError(f"target bitrate too low: row budget {budget:.0f} bits, cheapest measured row {min(tables[0])} bits")
    codec.encode()
    return codec, tables


def roi_qp_map(F, H, W, block_size, base_qp, roi_qp, box_fn):
    """Per-block QP map [F, n_blocks]: ``roi_qp`` inside the (moving) region ``box_fn(f) -> (x0, y0, x1, y1)`` in pixels,
    ``base_qp`` elsewhere.  A block belongs to the region when its centre does."""
    nbx, nby = W // block_size, H // block_size
    cx = (np.arange(nbx) + 0.5) * block_size
    cy = (np.arange(nby) + 0.5) * block_size
    out = np.full((F, nby, nbx), base_qp, np.int32)
    for f in range(F):
        x0, y0, x1, y1 = box_fn(f)
        m = ((cy[:, None] >= y0) & (cy[:, None] < y1)) & ((cx[None, :] >= x0) & (cx[None, :] < x1))
        out[f][m] = roi_qp
    return out.reshape(F, nby * nbx)
