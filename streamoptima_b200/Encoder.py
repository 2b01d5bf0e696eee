"""Drop-in ``Y_Video_codec`` whose per-block hot path runs as sm_100a CUDA kernels.

Mirrors the reference class surface (``/root/reference/Encoder.py``): constructor (Encoder.py:24), ``encode()``
(:1790), ``transmit_bitstream()`` (:1544), ``encoded_package`` (:1877-1888).  Host Python keeps what the reference
keeps on the host -- parameter handling, the data-independent rate-control row QPs (:1576-1609), package building and
the text bitstream -- and crosses one C-ABI boundary (``include/streamoptima_b200.h``) for everything inside
``complete_intra_flow`` / ``complete_inter_flow``.  There is no CPU fallback.

Deliberate deviations from the reference (all documented in DESIGN.md):
  * ``intra_mode=1`` and ``ParallelMode=3`` raise ``NotImplementedError``: both crash / hang in the reference.
  * frames may have any size that is a multiple of ``block_size`` (the reference hard-codes 288x352 in its intra search,
    Encoder.py:1165,1248; at 288x352 the behaviour is identical).
  * the embedded throw-away ``decoder.decode`` call (Encoder.py:1873) is not made.
  * ``"SSIM per frame"`` is a list of NaN (skimage is not part of this path).
  * ``transmit_bitstream`` writes the parseable ``entropy_encoder_frame`` text to ``residual_file`` (the reference
    writes an unparseable NumPy repr there, quirk Q10).
"""
from __future__ import annotations

import math
import os
import sys
import weakref

import numpy as np

from . import _native


def generate_Q_matrix(i, QP):
    """Encoder.py:938-945."""
    a, b = np.mgrid[0:i, 0:i]
    s = a + b
    return np.where(s < i - 1, 2 ** QP, np.where(s == i - 1, 2 ** (QP + 1), 2 ** (QP + 2))).astype(int)


class _PinnedPool:
    """Pinned host buffers recycled across encodes (``cudaHostAlloc`` of GBs costs more than the encode itself).  PyTorch
    supplies the pinned memory; nothing else of it is used here.

    A buffer handed back by a dead result waits in ``pending`` until nothing references its ndarray any more (views the
    caller kept out of a result stay valid for as long as they are held: a result is never overwritten behind its
    owner's back); only then is it offered to the next encode."""

    def __init__(self, max_free=12):
        self.free, self.pending, self.max_free = [], [], max_free        # entries: (nbytes, tensor, uint8 ndarray over it)

    def _sweep(self):
        keep = []
        for it in self.pending:
            if sys.getrefcount(it[2]) <= 2:            # the tuple and getrefcount's argument: no view is alive
                if len(self.free) < self.max_free:
                    self.free.append(it)
            else:
                keep.append(it)
        self.pending = keep

    def take(self, nbytes):
        nbytes = max(int(nbytes), 1)
        self._sweep()
        best = None
        for i, (n, _, _) in enumerate(self.free):
            if nbytes <= n <= max(2 * nbytes, nbytes + (1 << 20)) and (best is None or n < self.free[best][0]):
                best = i
        if best is not None:
            return self.free.pop(best)
        import torch
        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        return (nbytes, t, t.numpy())

    def give(self, item):
        self.pending.append(item)

    def clear(self):
        self.free, self.pending = [], []


class _Lease:
    """The buffers one result owns; they go back to the pool when the result dies."""

    def __init__(self, pool):
        self.items = []
        self.pool = pool
        weakref.finalize(self, _Lease._release, pool, self.items)

    def array(self, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        item = self.pool.take(n)
        self.items.append(item)
        return item[2][:n].view(dtype).reshape(shape)

    @staticmethod
    def _release(pool, items):
        while items:
            pool.give(items.pop())


class EncodeResult(dict):
    """Outputs of one encode call: host arrays that belong to this object (``levels`` is rebuilt from the symbol
    streams on first access when the call did not download it)."""

    lease = None
    _levels_from_symbols = None

    def __missing__(self, key):
        if key == "levels" and self._levels_from_symbols is not None:
            lev = self._levels_from_symbols()
            dict.__setitem__(self, "levels", lev)
            return lev
        raise KeyError(key)

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default


class _Packed(dict):
    """``encoded_package.packed``: the arrays of one encoded sequence.  It owns the ``EncodeResult`` they live in (so a later
    encode on the same codec cannot touch them); ``levels`` is rebuilt from the symbol streams on first access."""

    def __init__(self, result, **arrays):
        super().__init__(arrays)
        self.result = result

    def __missing__(self, key):
        if key == "levels":
            lev = self.result["levels"][0]
            dict.__setitem__(self, "levels", lev)
            return lev
        raise KeyError(key)

    def __contains__(self, key):
        return key == "levels" or dict.__contains__(self, key)


class EncodedPackage(dict):
    """``encoded_package`` (Encoder.py:1877-1888) whose two heavy keys are materialised from the packed arrays on first
    access (building ~10^4 Python tuples / ndarrays per 1080p frame costs far more than encoding it)."""

    _LAZY = ("MVS per Frame", "approx residual")

    def __init__(self, eager, frame_types, split, mv, levels, bs, packed=None):
        super().__init__(eager)
        self._ft, self._split, self._mv, self._bs = frame_types, split, mv, bs
        # ``packed``: the arrays behind the package (split, mv, levels, recon, symbol streams ...).  ``levels`` may be lazy
        # (rebuilt from the symbol streams on first access), hence the indirection.
        self.packed = packed if packed is not None else dict(frame_types=frame_types, split=split, mv=mv, levels=levels)

    @property
    def _lev(self):
        return self.packed["levels"]

    def __missing__(self, key):
        if key not in self._LAZY:
            raise KeyError(key)
        self._materialise()
        return dict.__getitem__(self, key)

    def _materialise(self):
        F, H, W = self._lev.shape
        bs, sub = self._bs, self._bs // 2
        nbx, nby = W // bs, H // bs
        mvs_all, lev_all = [], []
        for f in range(F):
            intra = self._ft[f] == 0
            spf = self._split[f].tolist()
            mvl = self._mv[f].tolist()                                   # [nblk][4][3] python ints
            # all blocks of the frame as views of one int array: [nby, nbx, bs, bs] (and the four quadrants of each)
            blocks = self._lev[f].astype(int).reshape(nby, bs, nbx, bs).swapaxes(1, 2)
            fm, fl = [], []
            for b, sp in enumerate(spf):
                m = mvl[b]
                blk = blocks[b // nbx, b % nbx]
                if sp == 0:
                    fm.append((0, m[0][0] if intra else tuple(m[0])))
                    fl.append((0, blk))
                else:
                    fm.append((1, [m[k][0] for k in range(4)] if intra else [tuple(m[k]) for k in range(4)]))
                    fl.append((1, [blk[(k // 2) * sub:(k // 2) * sub + sub, (k % 2) * sub:(k % 2) * sub + sub] for k in range(4)]))
            mvs_all.append(fm)
            lev_all.append(fl)
        dict.__setitem__(self, "MVS per Frame", mvs_all)
        dict.__setitem__(self, "approx residual", lev_all)

    def __contains__(self, key):
        return key in self._LAZY or dict.__contains__(self, key)

    def keys(self):
        self._materialise() if not dict.__contains__(self, "MVS per Frame") else None
        return dict.keys(self)

    def items(self):
        self.keys()
        return dict.items(self)

    def values(self):
        self.keys()
        return dict.values(self)

    def __iter__(self):
        return iter(self.keys())

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default


class Y_Video_codec:
    """Same constructor as the reference (Encoder.py:24)."""

    write_recon_yuv = True      # encode() writes yuv/y_only_reconstructed.yuv like the reference (Encoder.py:1894)
    roi_qp_map = None           # EXTENSION (not in the reference): int [F, n_blocks] per-block QPs for the final quantisation
    device = 0                  # CUDA device ordinal used by encode()
    sea_prune = False           # EXTENSION, results unchanged: prune the exhaustive search with SAD lower bounds (SO_FLAG_SEA);
                                # "auto": only while it pays on the content (SO_FLAG_SEA_AUTO)

    def __init__(self, h_pixels, w_pixels, frames, block_size, search_range, Qp, intra_dur, intra_mode, lam=None,
                 VBSEnable=False, nRefFrames=1, yuv_file=None, y_only_frame_arr=None, fast_me=False, FMEEnable=False,
                 RCFlag=None, targetBR=None, frame_rate=30, qp_rate_tables=None, intra_thresh=None, ParallelMode=0):
        self.h_pixels, self.w_pixels, self.frames = h_pixels, w_pixels, frames
        self.block_size = block_size
        self.num_blocks_per_row = w_pixels / block_size
        self.sub_block_size = block_size // 2
        self.search_range = search_range
        self.Qp = Qp
        self.const_init_Qp = Qp
        self.intra_dur, self.intra_mode = intra_dur, intra_mode
        self.Q = generate_Q_matrix(block_size, Qp)
        self.Qpm1 = Qp - 1 if Qp > 0 else Qp
        self.Qm1 = generate_Q_matrix(self.sub_block_size, self.Qpm1)
        self.encoded_package = None
        self.encoded_package_f = False
        self.nRefFrames, self.fast_me, self.FMEEnable, self.VBSEnable = nRefFrames, fast_me, FMEEnable, VBSEnable
        self.lam = lam
        self.RCFlag = RCFlag
        self.target_bitrate = None
        self.bitrate_per_row = None
        self.frame_rate = frame_rate
        self.qr_rate_tables = qp_rate_tables
        self.intra_thresh = intra_thresh
        self.ParallelMode = ParallelMode
        if targetBR is not None:                                                 # Encoder.py:78-88
            tokens = targetBR.split(" ")
            num = int(tokens[0])
            if tokens[1] == "kbps":
                self.target_bitrate = num * 1024
            elif tokens[1] == "mbps":
                self.target_bitrate = num * 1048576
            else:
                self.target_bitrate = num
            self.bitrate_per_row = (self.target_bitrate // self.frame_rate) / (self.h_pixels / self.block_size)
        # Encoder.py:103-106.  With a file the luma planes are NOT read here: encode() hands the path to the library, which
        # reads them chunk by chunk while earlier chunks are being encoded (so_encode_yuv420_file); touching
        # ``y_only_f_arr`` still gives the array the reference would hold.
        self._yuv_file = yuv_file
        self._y_arr = None if yuv_file is not None else y_only_frame_arr
        self._ctx = None
        self._ctx_key = None
        self.last_timing = None

    @property
    def y_only_f_arr(self):
        if self._y_arr is None and self._yuv_file is not None:
            self._y_arr = self.read_yuv(self._yuv_file, self.h_pixels, self.w_pixels, self.frames)
        return self._y_arr

    @y_only_f_arr.setter
    def y_only_f_arr(self, value):
        self._y_arr = value

    # Encoder.py:110-126
    @staticmethod
    def read_yuv(raw_yuv_420_f, height, width, frames):
        size_y = width * height
        size_uv = int(size_y / 4)
        out = np.empty((frames, height, width), np.uint8)
        with open(raw_yuv_420_f, "rb") as f:
            for i in range(frames):
                out[i] = np.frombuffer(f.read(size_y), dtype=np.uint8).reshape(height, width)
                f.read(size_uv * 2)
        return out

    # Encoder.py:1576-1580
    def get_appropriate_Qp_value(self, frame_type, row_bit_budget):
        for Qp, bitrate in enumerate(self.qr_rate_tables[frame_type]):
            if bitrate < row_bit_budget:
                return Qp, bitrate

    def _rc_row_qps(self, rows):
        """Row QPs of Encoder.py:1599-1609 / 1668-1678: data-independent (quirk Q9), both flows index table 0."""
        qps = []
        budget = self.bitrate_per_row
        spent = 0
        for row in range(rows):
            budget = self.bitrate_per_row if row == 0 else self.bitrate_per_row + (budget - spent)
            qp, spent = self.get_appropriate_Qp_value(0, budget)      # TypeError when no QP fits, like the reference
            qps.append(qp)
        return qps

    def _sea_mode(self):
        return "auto" if self.sea_prune == "auto" else bool(self.sea_prune)

    def _context(self, block_size, search_range, intra_dur, max_batch=1):
        """Native context for the CURRENT attribute values: the reference reads ``self.*`` every frame, so everything a
        context is created from is part of the cache key (only ``const_init_Qp`` is re-pushed on reuse)."""
        rc_on = self.RCFlag is not None and self.RCFlag > 0
        # the reference fails with TypeError in these two cases (``lam * bits`` Encoder.py:1158, ``size > None`` :1852)
        if self.VBSEnable and self.lam is None:
            raise TypeError("VBSEnable needs lam (the RD cost multiplies it, Encoder.py:1158)")
        if rc_on and self.RCFlag > 1 and self.intra_thresh is None:
            raise TypeError("RCFlag=2 needs intra_thresh (scene-cut test, Encoder.py:1852)")
        row_qps = tuple(self._rc_row_qps(self.h_pixels // block_size)) if rc_on else ()
        key = (block_size, search_range, intra_dur, max_batch, self.device, self.h_pixels, self.w_pixels, self.nRefFrames,
               bool(self.FMEEnable), bool(self.fast_me), bool(self.VBSEnable), self.RCFlag or 0, self.ParallelMode,
               float(self.lam or 0.0), int(self.intra_thresh or 0), row_qps, self._sea_mode())
        if self._ctx is not None and self._ctx_key == key:
            self._ctx.set_qp(self.const_init_Qp)
            return self._ctx
        if self._ctx is not None:
            self._ctx.close()
        self._ctx = _native.Context(width=self.w_pixels, height=self.h_pixels, block_size=block_size, search_range=search_range,
                                    qp=self.const_init_Qp, intra_dur=intra_dur, n_ref_frames=self.nRefFrames,
                                    fme=self.FMEEnable, fast_me=self.fast_me, vbs=self.VBSEnable, rc_flag=self.RCFlag or 0,
                                    parallel_mode=self.ParallelMode, lam=self.lam or 0.0, intra_thresh=self.intra_thresh or 0,
                                    max_batch=max_batch, device=self.device, sea=self._sea_mode())
        self._ctx_key = key
        if rc_on:
            self._ctx.set_row_qps(list(row_qps))
        return self._ctx

    def encode_arrays(self, frames_u8, block_size=None, search_range=None, intra_dur=None, want_levels=True, want_recon=True,
                      qp_map=None, want_symbols=False):
        """Encode ``frames_u8`` ([F,H,W] or [U,F,H,W] for U independent sequences) and return the packed outputs.

        This is the call the reference-facing ``encode()`` is built on; inputs and outputs are HOST arrays and the
        host<->device copies happen inside ``so_encode_sequence``.  ``want_symbols``: the run-level symbol streams
        (Encoder.py:1086-1131) are generated on the device and downloaded packed (``symbols`` / ``sym_pos`` / ``sym_count``);
        with ``want_levels=False`` they replace the 2 B/px raw levels on the device-to-host path and ``levels`` is rebuilt
        from them on first access.  The returned arrays belong to the returned ``EncodeResult``.
        """
        block_size = block_size or self.block_size
        search_range = self.search_range if search_range is None else search_range
        intra_dur = intra_dur or self.intra_dur
        arr = np.asarray(frames_u8)
        if arr.dtype != np.uint8:
            raise TypeError("frames must be uint8")
        if arr.ndim == 3:
            arr = arr[None]
        U, F, H, W = arr.shape
        if (H, W) != (self.h_pixels, self.w_pixels):
            raise ValueError("frame size does not match the codec")
        if H % block_size or W % block_size:
            raise ValueError("frame dimensions must be multiples of the block size (Encoder.py:1382)")
        ctx = self._context(block_size, search_range, intra_dur, max_batch=U)
        qp_map = self.roi_qp_map if qp_map is None else qp_map
        ctx.set_block_qps(qp_map)            # ROI extension; None clears
        # the input is handed to the library as is (pinned or pageable): its upload is chunked and overlapped with the
        # encode of earlier chunks, so an extra staging copy would only add host time
        a_in = np.ascontiguousarray(arr)
        return self._run(ctx, U, F, want_levels, want_recon, want_symbols,
                         lambda o: ctx.lib.so_encode_sequence(ctx.handle, a_in.ctypes.data, U, F, *o))

    def encode_yuv_file(self, path, src_height=None, src_width=None, first_frame=0, n_frames=None, block_size=None,
                        search_range=None, intra_dur=None, want_levels=True, want_recon=True, qp_map=None, want_symbols=False):
        """Encode frames of a planar YUV 4:2:0 file without materialising them on the Python side (read_yuv + pad_hw,
        Encoder.py:110-126 / :140-155, fused with the encode: ``so_encode_yuv420_file``).  ``src_height`` / ``src_width``: luma
        size in the file (default: the codec's size); smaller sources are padded with 128 to the codec's size."""
        block_size = block_size or self.block_size
        search_range = self.search_range if search_range is None else search_range
        intra_dur = intra_dur or self.intra_dur
        F = self.frames if n_frames is None else n_frames
        sh = self.h_pixels if src_height is None else src_height
        sw = self.w_pixels if src_width is None else src_width
        if self.h_pixels % block_size or self.w_pixels % block_size:
            raise ValueError("frame dimensions must be multiples of the block size (Encoder.py:1382)")
        ctx = self._context(block_size, search_range, intra_dur, max_batch=1)
        ctx.set_block_qps(self.roi_qp_map if qp_map is None else qp_map)
        bpath = os.fsencode(path)
        return self._run(ctx, 1, F, want_levels, want_recon, want_symbols,
                         lambda o: ctx.lib.so_encode_yuv420_file(ctx.handle, bpath, sw, sh, first_frame, F, *o))

    _pool = _PinnedPool()
    _sym_guess = 0.35          # symbols per pixel the symbol buffer of this codec is sized for (grows to what its encodes needed)

    def _run(self, ctx, U, F, want_levels, want_recon, want_symbols, call):
        """Pinned output buffers (leased from the pool, owned by the result) + one library call ``call(output pointers)``."""
        H, W = self.h_pixels, self.w_pixels
        nblk, rows = ctx.nblk, ctx.rows
        C = _native.C
        lease = _Lease(self._pool)
        # pinned targets: a pageable target makes the per-chunk D2H synchronous and stalls the host behind the GPU
        split = lease.array((U, F, nblk), np.uint8)
        mv = lease.array((U, F, nblk, 4, 3), np.int16)
        row_sizes = lease.array((U, F, rows), np.uint32)
        stats = lease.array((U, F, 32), np.uint8).view(_native.STATS_DTYPE).reshape(U, F)
        levels = lease.array((U, F, H, W), np.int16) if want_levels else None
        recon = lease.array((U, F, H, W), np.uint8) if want_recon else None
        sym = None
        if want_symbols:
            cap = int(self._sym_guess * U * F * H * W) + 4 * nblk * U * F + 1024
            symbols = lease.array((cap,), np.int16)
            pos, cnt = np.zeros((U, F), np.uint64), np.zeros((U, F), np.uint32)
            sym = _native.so_symbol_out(symbols.ctypes.data, cap, pos.ctypes.data, cnt.ctypes.data, 0)
            _native.check(ctx.handle, ctx.lib.so_set_symbol_output(ctx.handle, C.byref(sym)))
        try:
            rc = call((split.ctypes.data, mv.ctypes.data, levels.ctypes.data if want_levels else None,
                       recon.ctypes.data if want_recon else None, row_sizes.ctypes.data, stats.ctypes.data))
        finally:
            if want_symbols:
                ctx.lib.so_set_symbol_output(ctx.handle, None)
        if want_symbols and rc == -3 and sym.needed > sym.capacity:
            # the guess was too small: everything else is complete and the symbols are still on the device -- fetch them into
            # a buffer of the right size (no second encode) and remember the density for the next call
            cap = int(sym.needed) + 1024
            symbols = lease.array((cap,), np.int16)
            sym.symbols, sym.capacity = symbols.ctypes.data, cap
            rc = ctx.lib.so_fetch_symbols(ctx.handle, C.byref(sym))
        _native.check(ctx.handle, rc)
        self.last_timing = ctx.last_timing()
        self._last_shape = (U, F)
        out = EncodeResult(split=split, mv=mv, recon=recon, row_sizes=row_sizes, stats=stats,
                           frame_types=stats["frame_type"].astype(np.uint8))
        out.lease = lease
        if want_levels:
            out["levels"] = levels
        if want_symbols:
            self._sym_guess = max(self._sym_guess, 1.15 * sym.needed / float(U * F * H * W))     # per codec: densities differ with QP
            out.update(symbols=symbols[:max(int(sym.needed), 1)], sym_pos=pos, sym_count=cnt, sym_needed=int(sym.needed))
            if not want_levels:
                bs = ctx.bs

                def rebuild(split=split, symbols=symbols, pos=pos, cnt=cnt):
                    lev = np.empty((U, F, H, W), np.int16)
                    rc2 = _native.load().so_symbols_to_levels(split.ctypes.data, symbols.ctypes.data, pos.ctypes.data, cnt.ctypes.data,
                                                       U * F, W, H, bs, lev.ctypes.data, 0)
                    if rc2 != 0:
                        raise RuntimeError("corrupt symbol stream")
                    return lev
                out._levels_from_symbols = rebuild
        elif not want_levels:
            out["levels"] = None
        return out

    @staticmethod
    def psnr_from_sse(sse, npx):
        mse = np.float64(sse) / np.float64(npx)
        if mse == 0:
            return float("inf")
        return float(10.0 * np.log10(255.0 ** 2 / mse))

    def encode(self, intra_mode=None, intra_dur=None, search_range=None, block_size=None, save_enc_pkg=True):
        """Encoder.py:1790-1898: returns the per-frame PSNR list and fills ``encoded_package``."""
        if search_range is None: search_range = self.search_range
        if block_size is None: block_size = self.block_size
        if intra_dur is None: intra_dur = self.intra_dur
        if intra_mode is None: intra_mode = self.intra_mode
        if intra_mode != 0:
            raise NotImplementedError("intra_mode=1 raises TypeError in the reference (Encoder.py:1399-1407)")
        if self.ParallelMode == 3:
            raise NotImplementedError("ParallelMode=3 is racy / crashes in the reference (Encoder.py:1712-1787)")
        if self._yuv_file is not None and self._y_arr is None:
            out = self.encode_yuv_file(self._yuv_file, block_size=block_size, search_range=search_range, intra_dur=intra_dur,
                                       want_levels=False, want_symbols=True)
        else:
            frames = np.ascontiguousarray(self.y_only_f_arr[:self.frames])
            # the residual travels as packed run-level symbols (generated on the device), not as 2 B/px raw levels: the text
            # bitstream is formatted from them and ``packed["levels"]`` / ``"approx residual"`` are rebuilt on demand
            out = self.encode_arrays(frames, block_size, search_range, intra_dur, want_levels=False, want_symbols=True)
        st = out["stats"][0]
        H, W = self.h_pixels, self.w_pixels
        nblk = (H // block_size) * (W // block_size)
        psnr_per_frame = [self.psnr_from_sse(int(s), H * W) for s in st["sse"]]
        mae_per_frame = [float("inf") if inf else (float(n) / float(d)) / nblk
                         for n, d, inf in zip(st["mae_num"], st["mae_den"], st["mae_inf"])]
        frame_types = [int(t) for t in st["frame_type"]]
        rc_on = self.RCFlag is not None and self.RCFlag > 0
        qp_rows = self._rc_row_qps(H // block_size) if rc_on else []
        eager = {"block size": block_size, "num frames": self.frames, "height in pixels": H, "width in pixels": W,
                 "search range": search_range, "PSNR per frame": psnr_per_frame,
                 "SSIM per frame": [float("nan")] * self.frames, "MAE per Frame": mae_per_frame,
                 "Qp_per_row_per_frame": [list(qp_rows) for _ in range(self.frames)], "frame_type_seq": frame_types}
        packed = _Packed(out, frame_types=np.asarray(frame_types, np.uint8), split=out["split"][0], mv=out["mv"][0],
                         recon=out["recon"][0], row_sizes=out["row_sizes"][0], qsize=st["qsize"].copy(),
                         symbols=out["symbols"], sym_pos=out["sym_pos"][0], sym_count=out["sym_count"][0])
        pkg = EncodedPackage(eager, packed["frame_types"], packed["split"], packed["mv"], None, block_size, packed=packed)
        self.encoded_package_f = True
        if save_enc_pkg:
            self.encoded_package = pkg
        self._last_package = pkg
        if self.write_recon_yuv:
            os.makedirs("yuv", exist_ok=True)
            with open("yuv/y_only_reconstructed.yuv", "wb") as f:
                f.write(out["recon"][0].tobytes())
        return psnr_per_frame

    # ---- text bitstream ----------------------------------------------------------------------- Encoder.py:1419-1573
    def differential_encoder_frame(self, frame_type, split, mv, qp_rows):
        """MV/QP text of one frame from packed arrays (without the ``"<type>|"`` prefix, like Encoder.py:1419)."""
        lib = _native.load()
        nblk = split.shape[0]
        split = np.ascontiguousarray(split, np.uint8)
        mv = np.ascontiguousarray(mv, np.int16)
        qp = np.ascontiguousarray(qp_rows, np.int32) if len(qp_rows) else None
        text = self._fmt(lambda cbuf, cap: lib.so_format_mv_frame(int(frame_type), split.ctypes.data, mv.ctypes.data, nblk,
                                                                  int(self.num_blocks_per_row),
                                                                  qp.ctypes.data if qp is not None else None, cbuf, cap),
                         64 + nblk * 48)
        return text.split("|", 1)[1]

    def entropy_encoder_frame(self, split, levels, block_size=None):
        lib = _native.load()
        block_size = block_size or self.block_size
        H, W = levels.shape
        split = np.ascontiguousarray(split, np.uint8)
        levels = np.ascontiguousarray(levels, np.int16)
        return self._fmt(lambda cbuf, cap: lib.so_format_residual_frame(split.ctypes.data, levels.ctypes.data, W, H, block_size,
                                                                        cbuf, cap), 1024 + H * W * 2)

    @staticmethod
    def _fmt(call, cap):
        """Run a C formatter into a growing buffer (it returns -(needed+1) when ``cap`` is too small)."""
        while True:
            buf = bytearray(cap)
            n = call((_native.C.c_char * cap).from_buffer(buf), cap)
            if n >= 0:
                return bytes(buf[:n]).decode()
            cap = -n + 16

    def symbol_streams(self):
        """Run-level symbols of the last encode, generated on the GPU (count -> device prefix scan -> emit).

        Returns ``(offsets u32 [U*F, 4*nblk+1], symbols i16 [total], sym_base u64 [U*F+1])``: the symbols of sub-block k of
        block b of frame f are ``symbols[sym_base[f] + offsets[f, 4*b+k] : sym_base[f] + offsets[f, 4*b+k+1]]``."""
        ctx = self._ctx
        lib = ctx.lib
        _native.check(ctx.handle, lib.so_seq_symbols(ctx.handle))
        U, F = self._last_shape
        n1 = ctx.nblk * 4 + 1
        offsets = np.empty((U * F, n1), np.uint32)
        base = np.empty(U * F + 1, np.uint64)
        needed = _native.C.c_uint64(0)
        rc = lib.so_seq_download_symbols(ctx.handle, offsets.ctypes.data, None, 0, base.ctypes.data, _native.C.byref(needed))
        if rc not in (0, -3):
            _native.check(ctx.handle, rc)
        symbols = np.empty(max(int(needed.value), 1), np.int16)
        _native.check(ctx.handle, lib.so_seq_download_symbols(ctx.handle, offsets.ctypes.data, symbols.ctypes.data, symbols.size,
                                                              base.ctypes.data, _native.C.byref(needed)))
        return offsets, symbols[:int(needed.value)], base

    def residual_lines_from_symbols(self):
        """Residual text of the last single-sequence encode formatted from the device-generated symbol streams."""
        lib = _native.load()
        offsets, symbols, base = self.symbol_streams()
        pkg = self.encoded_package if self.encoded_package is not None else self._last_package
        split = np.ascontiguousarray(pkg.packed["split"], np.uint8)
        lines = []
        for f in range(split.shape[0]):
            sy = symbols[int(base[f]):int(base[f + 1])]
            sy = np.ascontiguousarray(sy) if sy.size else np.zeros(1, np.int16)
            off = np.ascontiguousarray(offsets[f])
            sp = np.ascontiguousarray(split[f])
            lines.append(self._fmt(lambda cbuf, cap: lib.so_format_residual_frame_symbols(sp.ctypes.data, off.ctypes.data, sy.ctypes.data,
                                                                                         sp.shape[0], cbuf, cap), 1024 + sy.size * 8 + sp.shape[0] * 24))
        return lines

    def bitstream_lines(self):
        """-> (mv_lines, residual_lines) of the last encode, one string per frame (no trailing newline)."""
        pkg = self.encoded_package if self.encoded_package is not None else self._last_package
        p = pkg.packed
        mv_lines, res_lines = [], []
        for f in range(len(p["frame_types"])):
            t = int(p["frame_types"][f])
            mv_lines.append(str(t) + "|" + self.differential_encoder_frame(t, p["split"][f], p["mv"][f], pkg["Qp_per_row_per_frame"][f]))
            if dict.__contains__(p, "symbols"):
                res_lines.append(self._residual_line_packed(p, f, pkg["block size"]))
            else:
                res_lines.append(self.entropy_encoder_frame(p["split"][f], p["levels"][f], pkg["block size"]))
        return mv_lines, res_lines

    def _residual_line_packed(self, p, f, block_size):
        """Residual text of frame ``f`` straight from its packed symbol stream (``so_format_residual_frame_packed``)."""
        lib = _native.load()
        sp = np.ascontiguousarray(p["split"][f], np.uint8)
        n = int(p["sym_count"][f])
        sy = np.ascontiguousarray(p["symbols"][int(p["sym_pos"][f]):int(p["sym_pos"][f]) + n]) if n else np.zeros(1, np.int16)

        def call(cbuf, cap):
            r = lib.so_format_residual_frame_packed(sp.ctypes.data, sy.ctypes.data, n, sp.shape[0], block_size, cbuf, cap)
            if r == -2 ** 63:
                raise RuntimeError("corrupt symbol stream")
            return r
        return self._fmt(call, 1024 + n * 8 + sp.shape[0] * 24)

    def transmit_bitstream(self, intra_dur=None, block_size=None, mv_file=None, residual_file=None):
        if not self.encoded_package_f:
            print("[ERROR] No encoded package available, please run encode() first")
            return
        # both files are written by the library: frames are formatted in parallel on host threads (so_write_bitstream_files);
        # byte-identical to joining bitstream_lines() with newlines
        pkg = self.encoded_package if self.encoded_package is not None else self._last_package
        p = pkg.packed
        F = len(p["frame_types"])
        H, W = pkg["height in pixels"], pkg["width in pixels"]
        qp = pkg["Qp_per_row_per_frame"]
        qp_arr = np.ascontiguousarray(qp, np.int32) if len(qp) and len(qp[0]) else None
        lib = _native.load()
        ft, sp = np.ascontiguousarray(p["frame_types"], np.uint8), np.ascontiguousarray(p["split"], np.uint8)   # kept alive
        mvs = np.ascontiguousarray(p["mv"], np.int16)                                                           # across the call
        qpp = qp_arr.ctypes.data if qp_arr is not None else None
        if dict.__contains__(p, "symbols"):      # residual text from the packed symbol streams: no levels on the host
            sy = np.ascontiguousarray(p["symbols"], np.int16)
            pos, cnt = np.ascontiguousarray(p["sym_pos"], np.uint64), np.ascontiguousarray(p["sym_count"], np.uint32)
            rc = lib.so_write_bitstream_files_symbols(ft.ctypes.data, sp.ctypes.data, mvs.ctypes.data, sy.ctypes.data, pos.ctypes.data,
                                                      cnt.ctypes.data, qpp, F, W, H, pkg["block size"], os.fsencode(mv_file),
                                                      os.fsencode(residual_file), 0)
        else:
            lev = np.ascontiguousarray(p["levels"], np.int16)
            rc = lib.so_write_bitstream_files(ft.ctypes.data, sp.ctypes.data, mvs.ctypes.data, lev.ctypes.data, qpp, F, W, H,
                                              pkg["block size"], os.fsencode(mv_file), os.fsencode(residual_file), 0)
        if rc != 0:
            raise OSError(f"cannot write the bitstream files {mv_file!r} / {residual_file!r}")
        if os.path.isdir("files"):       # debug dump of the reference (Encoder.py:1559,1568); only when ./files exists
            with open("files/mvs_per_frame_raw.txt", "w") as f:
                pkg = self.encoded_package if self.encoded_package is not None else self._last_package
                for t, mvs in zip(pkg["frame_type_seq"], pkg["MVS per Frame"]):
                    f.write(str(t) + "|" + str(mvs) + "\n")
