#!/usr/bin/env python
"""The reference's ``main.py`` flow (file extraction -> encoder -> bitstream files -> decoder) on the B200 path.

Same parameters as ``/root/reference/main.py:18-45`` (CIF, i = 16, r = 16, half-pel + fast ME + VBS, lambda 0.015, one
reference frame, 21 frames, I_Period 21); the input is a synthetic YUV 4:2:0 file because the reference's ``video/cif.yuv``
is not part of its repository.  Every step runs through the library: the luma planes are read from the file while earlier
chunks are encoded (``so_encode_yuv420_file``), the two text files are written and parsed on host threads, the decoder
reconstructs on the GPU; the decoded frames must equal the encoder's reconstruction.

    python examples/main.py [--frames 21] [--qp 5] [--keep DIR]
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from streamoptima_b200 import synth                         # noqa: E402
from streamoptima_b200.Encoder import Y_Video_codec         # noqa: E402
from streamoptima_b200 import decoder as dec                # noqa: E402


class main:
    """Mirror of the reference's ``main`` class (main.py:9-73)."""

    def __init__(self, targetBR=None, idx=0, qp=5, RCflag=None, frames=21, workdir=None):
        self.targetBR, self.idx, self.Qp, self.RCflag, self.frames, self.workdir = targetBR, idx, qp, RCflag, frames, workdir

    def main(self, debug_prints=True, qp_tables=None):
        start_time = time.time()
        block_size, search_range, Qp = 16, 16, self.Qp
        intra_dur, intra_mode, frames = 21, 0, self.frames
        h_pixels, w_pixels = 288, 352
        nRefFrames, FMEEnable, fast_me, VBSEnable, lam = 1, True, True, True, 0.015
        workdir = self.workdir or tempfile.mkdtemp(prefix="so_main_")
        os.makedirs(workdir, exist_ok=True)
        yuv = os.path.join(workdir, "cif.yuv")
        mv_file = os.path.join(workdir, f"mvs_per_frame_{self.idx}.txt")
        residual_file = os.path.join(workdir, f"res_per_frame_{self.idx}.txt")
        # stand-in for video/cif.yuv: planar 4:2:0, grey chroma
        luma = synth.translating(frames, h_pixels, w_pixels, seed=0)
        with open(yuv, "wb") as f:
            for i in range(frames):
                f.write(luma[i].tobytes())
                f.write(bytes([128]) * (h_pixels * w_pixels // 2))
        if debug_prints: print("[INFO] YUV 4:2:0 file written. Now running encoder.")
        Y_Video_codec.write_recon_yuv = False
        encoder = Y_Video_codec(h_pixels, w_pixels, frames, block_size, search_range, Qp, intra_dur, intra_mode, lam, VBSEnable,
                                nRefFrames=nRefFrames, yuv_file=yuv, fast_me=fast_me, FMEEnable=FMEEnable, RCFlag=self.RCflag,
                                targetBR=self.targetBR, frame_rate=30, qp_rate_tables=qp_tables, intra_thresh=70000)
        psnr = encoder.encode(block_size=block_size)
        if debug_prints: print("[INFO] Encoded; generating bitstream")
        encoder.transmit_bitstream(block_size=block_size, mv_file=mv_file, residual_file=residual_file)
        t_enc = time.time() - start_time
        decoder = dec.decoder(intra_mode, intra_dur, block_size, frames, h_pixels, w_pixels, Qp, nRefFrames, FMEEnable, lam, VBSEnable,
                              False, RCFlag=self.RCflag, targetBR=self.targetBR, frame_rate=30, qp_rate_tables=qp_tables)
        decoded = decoder.decode_bitstream(mv_file, residual_file, block_size=block_size)
        recon = encoder.encoded_package.packed["recon"]
        same = all(np.array_equal(decoded[i], recon[i]) for i in range(frames))
        if debug_prints:
            print(f"[INFO] encode + bitstream {t_enc:.3f} s, total {time.time() - start_time:.3f} s; mean PSNR {np.mean(psnr):.2f} dB; "
                  f"decoded == encoder reconstruction: {same}; files in {workdir}")
        return psnr, same


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=21)
    ap.add_argument("--qp", type=int, default=5)
    ap.add_argument("--keep", default=None)
    a = ap.parse_args()
    _, ok = main(qp=a.qp, frames=a.frames, workdir=a.keep).main()
    sys.exit(0 if ok else 1)
