"""Host-side pieces that need no GPU: the C text formatters and the ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from streamoptima_b200 import _native
from streamoptima_b200.Encoder import Y_Video_codec, EncodedPackage
from tests.golden_util import case_names, load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from streamoptima_b200 import build
    build.build()
    return _native.load()


def test_library_exports_every_declared_symbol():
    lib = _lib()
    hdr = open(os.path.join(ROOT, "include", "streamoptima_b200.h")).read()
    declared = set(re.findall(r"\b(so_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_native.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.so_abi_version() == 2
    assert ctypes.sizeof(_native.so_frame_stats) == 32 and ctypes.sizeof(_native.so_params) == 64


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _lib()
    with pytest.raises(_native.NativeError):
        _native.Context(width=64, height=64, block_size=8, search_range=2, qp=3, intra_dur=4)


@pytest.mark.parametrize("name", case_names())
def test_text_formatters_match_reference(name):
    _lib()
    frames, enc, g = load_case(name)
    F, H, W = frames.shape
    bs = enc["block_size"]
    c = Y_Video_codec(H, W, F, bs, enc["search_range"], enc["Qp"], enc["intra_dur"], 0,
                      **{k: v for k, v in enc.items() if k not in ("block_size", "search_range", "Qp", "intra_dur")})
    mv_lines, res_lines = [], []
    for f in range(F):
        t = int(g["frame_types"][f])
        qp = [int(q) for q in g["qp_rows"][f] if q >= 0]
        mv_lines.append(str(t) + "|" + c.differential_encoder_frame(t, g["split"][f], g["mv"][f], qp))
        res_lines.append(c.entropy_encoder_frame(g["split"][f], g["levels"][f], bs))
    assert "".join(l + "\n" for l in mv_lines) == g["mv_text"]
    assert "".join(l + "\n" for l in res_lines) == g["res_text"]


def test_lazy_package_matches_reference_structure():
    frames, enc, g = load_case("s_vbs")
    F, H, W = frames.shape
    pkg = EncodedPackage({"frame_type_seq": g["frame_types"].tolist()}, g["frame_types"], g["split"], g["mv"], g["levels"],
                         enc["block_size"])
    assert "MVS per Frame" in pkg
    mvs = pkg["MVS per Frame"]
    lev = pkg["approx residual"]
    assert len(mvs) == F and len(mvs[0]) == (H // 8) * (W // 8)
    assert mvs[0][0] == (0, -1) and isinstance(mvs[1][0][1], tuple)
    from oracle.packing import package_to_arrays
    split, mv, levels = package_to_arrays(g["frame_types"].tolist(), mvs, lev, H, W, enc["block_size"])
    np.testing.assert_array_equal(split, g["split"])
    np.testing.assert_array_equal(mv, g["mv"])
    np.testing.assert_array_equal(levels, g["levels"])


@pytest.mark.parametrize("name", case_names())
def test_sequence_bitstream_writer_matches_reference_files(name, tmp_path):
    """so_write_bitstream_files (multi-threaded whole-sequence writer behind transmit_bitstream) against the reference's
    own text streams for every golden case."""
    lib = _lib()
    frames, enc, g = load_case(name)
    F, H, W = frames.shape
    bs = enc["block_size"]
    ft = np.ascontiguousarray(g["frame_types"], np.uint8)
    split = np.ascontiguousarray(g["split"], np.uint8)
    mv = np.ascontiguousarray(g["mv"], np.int16)
    lev = np.ascontiguousarray(g["levels"], np.int16)
    rc_on = (g["qp_rows"] >= 0).any()
    qp = np.ascontiguousarray(g["qp_rows"], np.int32) if rc_on else None
    mvf, rsf = tmp_path / "mv.txt", tmp_path / "res.txt"
    for threads in (1, 3):
        rc = lib.so_write_bitstream_files(ft.ctypes.data, split.ctypes.data, mv.ctypes.data, lev.ctypes.data,
                                          qp.ctypes.data if qp is not None else None, F, W, H, bs, os.fsencode(mvf), os.fsencode(rsf), threads)
        assert rc == 0
        assert open(mvf).read() == g["mv_text"]
        assert open(rsf).read() == g["res_text"]
