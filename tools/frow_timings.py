"""Timings of the SURVEY 8(f) rows on the C2 sequence (1080p, 300 frames, i=16, r=16 half-pel, 4 refs, QP 4): text writer,
text parser, GPU decoder, file ingest, device-side symbol packing.  Prints one JSON object (committed under profiles/)."""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import synth_frames_torch
from streamoptima_b200 import _native, decoder as dec
from streamoptima_b200.Encoder import Y_Video_codec

F, H, W, QP = int(os.environ.get("F", 300)), 1088, 1920, 4
dev = torch.device("cuda", 0)
fr = synth_frames_torch(F, H, W, 0, dev)
pin = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True); pin.copy_(fr); frames = pin.numpy(); del fr
Y_Video_codec.write_recon_yuv = False
c = Y_Video_codec(H, W, F, 16, 16, QP, 30, 0, nRefFrames=4, FMEEnable=True, y_only_frame_arr=frames)
out = {}


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    return min(ts), r


# device-side symbol packing: device time of the sequence with and without it
t_plain, _ = best(lambda: c.encode_arrays(frames, want_levels=False, want_recon=False))
d_plain = c.last_timing["device_ms"]
t_sym, r = best(lambda: c.encode_arrays(frames, want_levels=False, want_recon=False, want_symbols=True))
d_sym = c.last_timing["device_ms"]
t_lev, _ = best(lambda: c.encode_arrays(frames, want_levels=True, want_recon=True))
out["symbols"] = {"device_ms_without": d_plain, "device_ms_with": d_sym, "packer_us_per_frame": 1e3 * (d_sym - d_plain) / F,
                  "wall_s_symbols_only": t_sym, "wall_s_levels_and_recon": t_lev, "symbols": int(r["sym_needed"]),
                  "d2h_bytes_symbols": int(r["sym_needed"]) * 2, "d2h_bytes_levels": F * H * W * 2}
del r
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    mvf, rsf = os.path.join(d, "mv.txt"), os.path.join(d, "res.txt")
    t_enc, _ = best(lambda: c.encode(), 2)
    t_wr, _ = best(lambda: c.transmit_bitstream(mv_file=mvf, residual_file=rsf), 2)
    nbytes = os.path.getsize(mvf) + os.path.getsize(rsf)
    p = c.encoded_package.packed
    lev = np.ascontiguousarray(p["levels"])          # rebuilt from the symbols (so_symbols_to_levels)
    t0 = time.perf_counter(); _ = c._last_package.packed.result._levels_from_symbols(); t_inv = time.perf_counter() - t0
    lib = _native.load()
    ft, sp, mv = np.ascontiguousarray(p["frame_types"]), np.ascontiguousarray(p["split"]), np.ascontiguousarray(p["mv"])
    t_wr_lev, _ = best(lambda: lib.so_write_bitstream_files(ft.ctypes.data, sp.ctypes.data, mv.ctypes.data, lev.ctypes.data, None, F, W, H, 16,
                                                           os.fsencode(mvf + "2"), os.fsencode(rsf + "2"), 0), 2)
    same = open(rsf, "rb").read() == open(rsf + "2", "rb").read() and open(mvf, "rb").read() == open(mvf + "2", "rb").read()
    out["writer"] = {"encode_s": t_enc, "transmit_bitstream_s_from_symbols": t_wr, "write_from_levels_s": t_wr_lev, "text_bytes": nbytes,
                     "MB_per_s": nbytes / t_wr / 1e6, "files_identical": bool(same), "symbols_to_levels_s": t_inv,
                     "host_threads": os.cpu_count()}
    d1 = dec.decoder(0, 30, 16, F, H, W, QP, 4, True, None, False)
    t_parse, parsed = best(lambda: d1.parse_bitstream(mvf, rsf), 2)
    out["parser"] = {"seconds": t_parse, "MB_per_s": nbytes / t_parse / 1e6}
    ft2, sp2, mv2, lev2, _ = parsed
    assert np.array_equal(lev2, lev) and np.array_equal(mv2, mv)
    t_dec, rec = best(lambda: d1.decode_arrays(ft2, sp2, mv2, lev2, None, reset_at_intra=False), 2)
    out["gpu_decoder"] = {"seconds": t_dec, "frames_per_s": F / t_dec, "equals_encoder_recon": bool(np.array_equal(rec, p["recon"]))}
    # ingest: a planar 4:2:0 file on tmpfs
    yuv = os.path.join(d, "in.yuv")
    with open(yuv, "wb") as f:
        uv = bytes(H * W // 2)
        for i in range(F):
            f.write(frames[i].tobytes()); f.write(uv)
    c2 = Y_Video_codec(H, W, F, 16, 16, QP, 30, 0, nRefFrames=4, FMEEnable=True, yuv_file=yuv)
    t_file, r2 = best(lambda: c2.encode_yuv_file(yuv, want_levels=False, want_recon=False, want_symbols=True), 3)
    t_read, arr = best(lambda: Y_Video_codec.read_yuv(yuv, H, W, F), 1)
    out["ingest"] = {"encode_yuv_file_s": t_file, "frames_per_s": F / t_file, "python_read_yuv_s": t_read,
                     "encode_arrays_s": t_sym, "file_bytes": os.path.getsize(yuv)}
print(json.dumps(out, indent=1))
