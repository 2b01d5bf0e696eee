"""Frame ingest from planar YUV 4:2:0 files (read_yuv Encoder.py:110-126 + pad_hw :140-155) fused with the encode:
``so_encode_yuv420_file`` must give exactly what encoding the extracted (and 128-padded) luma array gives."""
import numpy as np
import pytest

from streamoptima_b200 import synth

pytestmark = pytest.mark.gpu


def _write_yuv420(path, frames):
    F, H, W = frames.shape
    with open(path, "wb") as f:
        for i in range(F):
            f.write(frames[i].tobytes())
            f.write(bytes([(37 * i + 11) % 256]) * (int(H * W / 4) * 2))      # chroma: read and discarded (Encoder.py:124)


def _codec(H, W, F, **kw):
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    return Y_Video_codec(H, W, F, 16, 4, 3, 8, 0, FMEEnable=True, nRefFrames=2, **kw)


def test_yuv_file_equals_array_encode(tmp_path):
    F, H, W = 21, 96, 128                               # 3 chunks of 8: exercises the rotating staging buffers
    frames = synth.translating(F, H, W, seed=81)
    path = tmp_path / "clip.yuv"
    _write_yuv420(path, frames)
    a = _codec(H, W, F, y_only_frame_arr=frames)
    a.encode()
    b = _codec(H, W, F, yuv_file=str(path))
    psnr = b.encode()                                   # goes through so_encode_yuv420_file: the array is never built
    assert b._y_arr is None
    pa, pb = a.encoded_package.packed, b.encoded_package.packed
    for k in ("split", "mv", "levels", "recon", "row_sizes", "frame_types"):
        np.testing.assert_array_equal(pa[k], pb[k], err_msg=k)
    assert psnr == a.encoded_package["PSNR per frame"]
    np.testing.assert_array_equal(b.y_only_f_arr, frames)          # the reference attribute still works (lazy read_yuv)


def test_yuv_file_padding_and_offset(tmp_path):
    F, sh, sw = 10, 90, 120                             # source smaller than the coded 96x128: padded with 128 (pad_hw)
    src = synth.zooming(F, sh, sw, seed=82)
    path = tmp_path / "odd.yuv"
    _write_yuv420(path, src)
    H, W, first, n = 96, 128, 3, 6
    padded = np.full((n, H, W), 128, np.uint8)
    padded[:, :sh, :sw] = src[first:first + n]
    a = _codec(H, W, n)
    oa = {k: np.array(v) for k, v in a.encode_arrays(padded).items() if k in ("split", "mv", "levels", "recon")}
    b = _codec(H, W, n)
    ob = b.encode_yuv_file(str(path), src_height=sh, src_width=sw, first_frame=first, n_frames=n)
    for k in oa:
        np.testing.assert_array_equal(oa[k], ob[k], err_msg=k)


def test_yuv_file_errors(tmp_path):
    from streamoptima_b200._native import NativeError
    c = _codec(96, 128, 4)
    with pytest.raises(NativeError):
        c.encode_yuv_file(str(tmp_path / "missing.yuv"))
    short = tmp_path / "short.yuv"
    _write_yuv420(short, synth.translating(2, 96, 128, seed=83))
    with pytest.raises(NativeError):
        c.encode_yuv_file(str(short), n_frames=4)


def test_reference_main_flow_end_to_end(tmp_path):
    """examples/main.py: the reference's main.py settings (CIF, i = 16, r = 16, half-pel + fast ME + VBS) from a YUV 4:2:0
    file through encode, the two text files, the host parser and the GPU decoder; decoded frames == encoder reconstruction."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("so_example_main", os.path.join(root, "examples", "main.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    psnr, same = mod.main(qp=5, frames=21, workdir=str(tmp_path)).main(debug_prints=False)
    assert same and len(psnr) == 21 and min(psnr) > 25
    assert os.path.getsize(tmp_path / "mvs_per_frame_0.txt") > 0 and os.path.getsize(tmp_path / "res_per_frame_0.txt") > 0
