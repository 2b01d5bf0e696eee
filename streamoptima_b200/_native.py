"""ctypes binding of ``libstreamoptima_b200.so`` (C ABI declared in ``include/streamoptima_b200.h``).

There is no CPU fallback: if the shared library is missing (run ``python -c "import __graft_entry__ as g; g.build()"``)
or no sm_100 device is visible, the calls raise :class:`NativeError`.
"""
from __future__ import annotations

import ctypes as C
import os

_LIB_NAME = "libstreamoptima_b200.so"
_HERE = os.path.dirname(os.path.abspath(__file__))

SO_FLAG_FME, SO_FLAG_FAST_ME, SO_FLAG_VBS, SO_FLAG_SEA, SO_FLAG_SEA_AUTO = 1, 2, 4, 8, 16
SO_MAX_REF = 8


class NativeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"streamoptima_b200 native error {code}: {msg}")
        self.code = code


class so_params(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("block_size", C.c_int32), ("search_range", C.c_int32),
                ("qp", C.c_int32), ("intra_dur", C.c_int32), ("n_ref_frames", C.c_int32), ("flags", C.c_uint32),
                ("rc_flag", C.c_int32), ("parallel_mode", C.c_int32), ("lam", C.c_double), ("intra_thresh", C.c_int64),
                ("max_batch", C.c_int32), ("reserved", C.c_int32)]


class so_frame_stats(C.Structure):
    _fields_ = [("sse", C.c_uint64), ("mae_num", C.c_uint64), ("mae_den", C.c_uint32), ("mae_inf", C.c_uint32),
                ("qsize", C.c_uint32), ("frame_type", C.c_uint32)]


class so_frame_out(C.Structure):
    _fields_ = [("split", C.c_void_p), ("mv", C.c_void_p), ("levels", C.c_void_p), ("recon", C.c_void_p),
                ("row_sizes", C.c_void_p), ("stats", C.c_void_p)]


class so_symbol_out(C.Structure):
    _fields_ = [("symbols", C.c_void_p), ("capacity", C.c_uint64), ("pos", C.c_void_p), ("count", C.c_void_p),
                ("needed", C.c_uint64)]


STATS_DTYPE = [("sse", "<u8"), ("mae_num", "<u8"), ("mae_den", "<u4"), ("mae_inf", "<u4"), ("qsize", "<u4"),
               ("frame_type", "<u4")]

# every symbol declared in include/streamoptima_b200.h
EXPORTS = ["so_abi_version", "so_last_error", "so_device_count", "so_ctx_create", "so_ctx_destroy", "so_set_qp", "so_set_row_qps", "so_set_block_qps",
           "so_ref_reset", "so_ref_push", "so_encode_intra", "so_encode_inter", "so_encode_sequence", "so_encode_yuv420_file", "so_seq_upload", "so_seq_run", "so_seq_download", "so_seq_sync", "so_decode_sequence", "so_seq_symbols", "so_seq_download_symbols",
           "so_format_residual_frame_symbols", "so_write_bitstream_files", "so_parse_bitstream_files",
           "so_set_symbol_output", "so_fetch_symbols", "so_format_residual_frame_packed", "so_symbols_to_levels",
           "so_write_bitstream_files_symbols",
           "so_last_timing", "so_last_me_launches", "so_last_search_timing", "so_last_finish_timing", "so_sea_stats",
           "so_format_mv_frame", "so_format_residual_frame"]

_lib = None


def lib_path() -> str:
    return os.path.join(_HERE, _LIB_NAME)


def load():
    """Load the shared library (once) and declare signatures.  Raises NativeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise NativeError(-100, f"{path} not found: build it with __graft_entry__.build(); there is no CPU fallback")
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.so_abi_version.restype = i32
    lib.so_last_error.restype = C.c_char_p
    lib.so_last_error.argtypes = [vp]
    lib.so_device_count.restype = i32
    lib.so_ctx_create.argtypes = [C.POINTER(vp), C.POINTER(so_params), i32]
    lib.so_ctx_destroy.argtypes = [vp]
    lib.so_ctx_destroy.restype = None
    lib.so_set_qp.argtypes = [vp, i32]
    lib.so_set_block_qps.argtypes = [vp, vp, i32]
    lib.so_set_row_qps.argtypes = [vp, C.POINTER(C.c_int32), i32]
    lib.so_ref_reset.argtypes = [vp, i32, vp]
    lib.so_ref_push.argtypes = [vp, i32, vp, vp]
    lib.so_encode_intra.argtypes = [vp, i32, vp, C.POINTER(so_frame_out), vp]
    lib.so_encode_inter.argtypes = [vp, i32, vp, C.POINTER(so_frame_out), vp]
    lib.so_encode_sequence.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.so_seq_upload.argtypes = [vp, vp, i32, i32]
    lib.so_seq_run.argtypes = [vp]
    lib.so_seq_download.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.so_seq_sync.argtypes = [vp]
    lib.so_seq_symbols.argtypes = [vp]
    lib.so_seq_download_symbols.argtypes = [vp, vp, vp, C.c_uint64, vp, vp]
    lib.so_format_residual_frame_symbols.restype = i64
    lib.so_format_residual_frame_symbols.argtypes = [vp, vp, vp, i32, C.c_char_p, i64]
    lib.so_decode_sequence.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp]
    lib.so_encode_yuv420_file.argtypes = [vp, C.c_char_p, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.so_write_bitstream_files.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, C.c_char_p, C.c_char_p, i32]
    lib.so_set_symbol_output.argtypes = [vp, C.POINTER(so_symbol_out)]
    lib.so_fetch_symbols.argtypes = [vp, C.POINTER(so_symbol_out)]
    lib.so_format_residual_frame_packed.restype = i64
    lib.so_format_residual_frame_packed.argtypes = [vp, vp, i64, i32, i32, C.c_char_p, i64]
    lib.so_symbols_to_levels.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, i32]
    lib.so_write_bitstream_files_symbols.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, C.c_char_p, C.c_char_p, i32]
    lib.so_parse_bitstream_files.argtypes = [C.c_char_p, C.c_char_p, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32]
    lib.so_last_search_timing.argtypes = [vp, C.POINTER(C.c_double)]
    lib.so_last_finish_timing.argtypes = [vp, C.POINTER(C.c_double)]
    lib.so_last_timing.argtypes = [vp, C.POINTER(C.c_double)]
    lib.so_sea_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.so_last_me_launches.argtypes = [vp]
    lib.so_format_mv_frame.restype = i64
    lib.so_format_mv_frame.argtypes = [i32, vp, vp, i32, i32, vp, C.c_char_p, i64]
    lib.so_format_residual_frame.restype = i64
    lib.so_format_residual_frame.argtypes = [vp, vp, i32, i32, i32, C.c_char_p, i64]
    for name in EXPORTS:
        getattr(lib, name)
    _lib = lib
    return lib


def check(ctx, rc):
    if rc != 0:
        msg = load().so_last_error(ctx)
        raise NativeError(rc, msg.decode() if msg else "?")


class Context:
    """Owns one ``so_ctx`` (one CUDA device, not thread-safe)."""

    def __init__(self, *, width, height, block_size, search_range, qp, intra_dur, n_ref_frames=1, fme=False, fast_me=False,
                 vbs=False, rc_flag=0, parallel_mode=0, lam=0.0, intra_thresh=0, max_batch=1, device=0, sea=False):
        self.lib = load()
        p = so_params(width=width, height=height, block_size=block_size, search_range=search_range, qp=qp,
                      intra_dur=intra_dur, n_ref_frames=n_ref_frames,
                      flags=(SO_FLAG_FME if fme else 0) | (SO_FLAG_FAST_ME if fast_me else 0) | (SO_FLAG_VBS if vbs else 0) |
                      (SO_FLAG_SEA if sea else 0) | (SO_FLAG_SEA_AUTO if sea == "auto" else 0),
                      rc_flag=rc_flag or 0, parallel_mode=parallel_mode, lam=float(lam or 0.0),
                      intra_thresh=int(intra_thresh or 0), max_batch=max_batch, reserved=0)
        self.params = p
        self.handle = C.c_void_p()
        rc = self.lib.so_ctx_create(C.byref(self.handle), C.byref(p), device)
        if rc != 0:
            msg = self.lib.so_last_error(None)
            self.handle = None
            raise NativeError(rc, msg.decode() if msg else "?")
        self.width, self.height, self.bs = width, height, block_size
        self.nblk = (width // block_size) * (height // block_size)
        self.rows = height // block_size
        self.max_batch = max_batch
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self.lib.so_ctx_destroy(self.handle)
            self.handle = None

    __del__ = close

    def set_qp(self, qp):
        check(self.handle, self.lib.so_set_qp(self.handle, int(qp)))

    def set_row_qps(self, qps):
        arr = (C.c_int32 * len(qps))(*[int(q) for q in qps])
        check(self.handle, self.lib.so_set_row_qps(self.handle, arr, len(qps)))

    def set_block_qps(self, qp_blocks):
        """ROI extension: per-block QPs [F, nblk] for the next sequence, or None to clear."""
        if qp_blocks is None:
            check(self.handle, self.lib.so_set_block_qps(self.handle, None, 0))
            return
        import numpy as np
        a = np.ascontiguousarray(qp_blocks, np.int32)
        assert a.ndim == 2 and a.shape[1] == self.nblk
        check(self.handle, self.lib.so_set_block_qps(self.handle, a.ctypes.data, a.shape[0]))

    def sea_stats(self):
        """Counters of the pruned exhaustive search (``sea=True``) since the context was created."""
        out = (C.c_uint64 * 4)()
        check(self.handle, self.lib.so_sea_stats(self.handle, out))
        return dict(exact_sads=int(out[0]), p_frames=int(out[1]))

    def last_timing(self):
        out = (C.c_double * 4)()
        check(self.handle, self.lib.so_last_timing(self.handle, out))
        xs = (C.c_double * 4)()
        check(self.handle, self.lib.so_last_search_timing(self.handle, xs))
        # me_ms / tq_ms / search_ms are sums over the frames that carry per-kernel events (timed_frames of frames)
        fi = (C.c_double * 4)()
        check(self.handle, self.lib.so_last_finish_timing(self.handle, fi))
        return dict(device_ms=out[0], me_ms=out[1], tq_ms=out[2], launches=int(out[3]), finish_inter_ms=fi[0], finish_inter_launches=int(fi[1]),
                    me_launches=int(self.lib.so_last_me_launches(self.handle)), search_ms=xs[0], search_launches=int(xs[1]),
                    timed_frames=int(xs[2]), frames=int(xs[3]))
