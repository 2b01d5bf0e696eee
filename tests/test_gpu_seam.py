"""The per-frame seam of the C ABI, driven exactly as INTEGRATION.md section 2 hands it to a reference maintainer.

``so_encode_intra`` / ``so_encode_inter`` replace ``complete_intra_flow`` (Encoder.py:1582) / ``complete_inter_flow`` (:1644)
and ``so_ref_reset`` / ``so_ref_push`` the ``ref_frames`` list (:1798, :1864-1867); the frame loop stays with the caller.
The tests are TEACHER-FORCED: frame t is encoded against the reference encoder's own reconstructions
(the golden recon is what gets pushed, not ours), so every frame is compared with the golden frame in isolation.
``unit`` selects one of ``max_batch`` independent chains: two units run the same sequence one frame apart and must not
disturb each other.
"""
import ctypes as C

import numpy as np
import pytest

from tests.golden_util import case_names, load_case

pytestmark = pytest.mark.gpu

SO_ALL_UNITS = -1


def _seam_cases():
    from oracle.golden_cases import CASES
    out = []
    for n in case_names():
        e = CASES[n]["enc"]
        if e.get("ParallelMode", 0) != 0 or (e.get("RCFlag") or 0) > 1:
            continue            # ParallelMode 1 resets the list per frame and RCFlag 2 re-encodes: covered by the sequence tests
        if n.startswith("w1920") or n.startswith("cif_fme"):
            continue            # large cases are covered through the sequence API (tests/test_gpu_parity.py)
        out.append(n)
    return out


class Seam:
    """Minimal binding of the per-frame calls -- the same few lines INTEGRATION.md shows."""

    def __init__(self, H, W, enc, max_batch=1, sea=False):
        import torch
        from streamoptima_b200 import _native
        from streamoptima_b200.Encoder import Y_Video_codec
        self.torch, self.nat = torch, _native
        e = dict(enc)
        self.bs = e["block_size"]
        self.ctx = _native.Context(width=W, height=H, block_size=self.bs, search_range=e["search_range"], qp=e["Qp"],
                                   intra_dur=e["intra_dur"], n_ref_frames=e.get("nRefFrames", 1), fme=e.get("FMEEnable", False),
                                   fast_me=e.get("fast_me", False), vbs=e.get("VBSEnable", False), rc_flag=e.get("RCFlag") or 0,
                                   parallel_mode=0, lam=e.get("lam") or 0.0, max_batch=max_batch, sea=sea)
        if (e.get("RCFlag") or 0) > 0:        # data-independent row QPs (quirk Q9), computed by the host class like the reference does
            helper = Y_Video_codec(H, W, 1, self.bs, e["search_range"], e["Qp"], e["intra_dur"], 0, RCFlag=e["RCFlag"],
                                   targetBR=e["targetBR"], qp_rate_tables=e["qp_rate_tables"])
            self.ctx.set_row_qps(helper._rc_row_qps(H // self.bs))
        self.H, self.W = H, W
        self.nblk = (H // self.bs) * (W // self.bs)
        self.lib, self.h = self.ctx.lib, self.ctx.handle

    def outputs(self, n=1):
        t = self.torch
        dev = "cuda:0"
        o = dict(split=t.zeros((n, self.nblk), dtype=t.uint8, device=dev),
                 mv=t.zeros((n, self.nblk, 4, 3), dtype=t.int16, device=dev),
                 levels=t.zeros((n, self.H, self.W), dtype=t.int16, device=dev),
                 recon=t.zeros((n, self.H, self.W), dtype=t.uint8, device=dev),
                 row_sizes=t.zeros((n, self.H // self.bs), dtype=t.int32, device=dev),
                 stats=t.zeros((n, 32), dtype=t.uint8, device=dev))
        fo = self.nat.so_frame_out(*[o[k].data_ptr() for k in ("split", "mv", "levels", "recon", "row_sizes", "stats")])
        return o, fo

    def check(self, rc):
        self.nat.check(self.h, rc)

    def reset(self, unit):
        self.check(self.lib.so_ref_reset(self.h, unit, None))

    def push(self, unit, recon_u8):
        r = self.torch.as_tensor(np.ascontiguousarray(recon_u8), device="cuda:0")
        self.check(self.lib.so_ref_push(self.h, unit, C.c_void_p(r.data_ptr()), None))
        self.torch.cuda.synchronize()

    def encode(self, unit, frame_u8, intra):
        n = frame_u8.shape[0] if frame_u8.ndim == 3 else 1
        cur = self.torch.as_tensor(np.ascontiguousarray(frame_u8), device="cuda:0")
        o, fo = self.outputs(n)
        fn = self.lib.so_encode_intra if intra else self.lib.so_encode_inter
        self.check(fn(self.h, unit, C.c_void_p(cur.data_ptr()), C.byref(fo), None))
        self.torch.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in o.items()}


def _assert_frame(out, g, f, k=0):
    np.testing.assert_array_equal(out["split"][k], g["split"][f])
    np.testing.assert_array_equal(out["mv"][k], g["mv"][f])
    np.testing.assert_array_equal(out["levels"][k], g["levels"][f])
    np.testing.assert_array_equal(out["recon"][k], g["recon"][f])


@pytest.mark.parametrize("name", _seam_cases())
def test_teacher_forced_per_frame_seam(name):
    frames, enc, g = load_case(name)
    F, H, W = frames.shape
    s = Seam(H, W, enc)
    s.reset(0)                                                        # ref_frames = [128 frame]   (Encoder.py:1798)
    for f in range(F):
        intra = int(g["frame_types"][f]) == 0                         # Encoder.py:1839 stays the caller's decision
        out = s.encode(0, frames[f], intra)
        _assert_frame(out, g, f)
        st = out["stats"].view(s.nat.STATS_DTYPE)[0, 0]
        assert int(st["frame_type"]) == (0 if intra else 1)
        mse = float(st["sse"]) / (H * W)
        psnr = float("inf") if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
        assert psnr == pytest.approx(float(g["psnr"][f]), abs=1e-9)
        s.push(0, g["recon"][f])                                      # teacher forcing: the REFERENCE's reconstruction
    s.ctx.close()


@pytest.mark.parametrize("name", ["s_fme_nref3_i16", "s_vbs_fme_nref2", "s_nref3", "s_fast_nref2_i16"])
def test_units_are_independent_chains(name):
    """Unit 1 runs the same sequence one frame behind unit 0: their reference lists differ at every call."""
    frames, enc, g = load_case(name)
    F, H, W = frames.shape
    s = Seam(H, W, enc, max_batch=2)
    s.reset(0)
    s.reset(1)
    for step in range(F + 1):
        for unit, f in ((0, step), (1, step - 1)):
            if not 0 <= f < F:
                continue
            out = s.encode(unit, frames[f], int(g["frame_types"][f]) == 0)
            _assert_frame(out, g, f)
            s.push(unit, g["recon"][f])
    s.ctx.close()


def test_all_units_lock_step_then_fork():
    """SO_ALL_UNITS drives every chain with dense [max_batch] buffers; afterwards single units continue from that state,
    and going back to lock step with diverged chains is refused."""
    name = "s_fme_nref3_i16"
    frames, enc, g = load_case(name)
    F, H, W = frames.shape
    s = Seam(H, W, enc, max_batch=2)
    s.reset(SO_ALL_UNITS)
    both = lambda a: np.stack([a, a])
    for f in range(3):
        out = s.encode(SO_ALL_UNITS, both(frames[f]), int(g["frame_types"][f]) == 0)
        _assert_frame(out, g, f, 0)
        _assert_frame(out, g, f, 1)
        s.push(SO_ALL_UNITS, both(g["recon"][f]))
    out = s.encode(1, frames[3], int(g["frame_types"][3]) == 0)       # fork: unit 1 alone goes on
    _assert_frame(out, g, 3)
    s.push(1, g["recon"][3])
    o, fo = s.outputs(2)
    cur = s.torch.as_tensor(both(frames[4]), device="cuda:0")
    rc = s.lib.so_encode_inter(s.h, SO_ALL_UNITS, C.c_void_p(cur.data_ptr()), C.byref(fo), None)
    assert rc == -4, "diverged chains must be refused with SO_E_STATE"
    out = s.encode(0, frames[3], int(g["frame_types"][3]) == 0)       # unit 0 is still where lock step left it
    _assert_frame(out, g, 3)
    assert s.lib.so_ref_reset(s.h, 2, None) == -1                     # unit out of range
    s.ctx.close()
