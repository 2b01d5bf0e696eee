"""CPU oracle: a NumPy/SciPy restatement of StreamOptima's per-block encode path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import this module; the product (``streamoptima_b200/``) never does and
fails loudly when its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors, so this oracle is pinned against
outputs of the *reference itself*, produced in the build container by ``oracle/gen_golden.py`` and committed
under ``tests/golden/`` (``tests/test_oracle_golden.py`` replays every case bit-for-bit: frame types, MVs, split
flags, levels, reconstruction, both text streams, PSNR, MAE).

Third-party arithmetic: the transform is ``scipy.fftpack.dct/idct`` exactly as the reference calls it
(``Encoder.py:10,781,812``); the reference pins no SciPy version, the goldens were made with SciPy 1.18.1
(``_duccfft`` backend) and the GPU box runs the same image.

Every function cites the reference lines it restates.  The structure is *not* the reference's: full search is
vectorised over all blocks of a frame per candidate (same candidate order and the same replace rule, so the
sequential tie-break semantics are preserved), everything else is per block.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.fftpack import dct, idct

INF = float("inf")


# --------------------------------------------------------------------------------------------------------------
# transform / quantisation                                                               Encoder.py:779-827,938-959
# --------------------------------------------------------------------------------------------------------------
def q_matrix(n: int, qp: int) -> np.ndarray:
    """``generate_Q_matrix`` (Encoder.py:938-945): 2^QP above the anti-diagonal, 2^(QP+1) on it, 2^(QP+2) below."""
    i, j = np.mgrid[0:n, 0:n]
    s = i + j
    return np.where(s < n - 1, 2 ** qp, np.where(s == n - 1, 2 ** (qp + 1), 2 ** (qp + 2))).astype(np.int64)


def qm1_qp(qp: int) -> int:
    """Sub-block QP: QP-1 clamped at 0 (Encoder.py:71-76, 954-959)."""
    return qp - 1 if qp > 0 else qp


def dct2_round(blocks: np.ndarray) -> np.ndarray:
    """``apply_2d_dct`` (Encoder.py:779-784) on ``[..., n, n]``: ortho DCT-II on axis 0 then axis 1, np.round -> int."""
    t = dct(dct(blocks, axis=-2, norm="ortho"), axis=-1, norm="ortho")
    return np.round(t).astype(np.int64)


def idct2_round(coefs: np.ndarray) -> np.ndarray:
    """``apply_2d_idct`` (Encoder.py:810-817)."""
    t = idct(idct(coefs.astype(np.float64), axis=-2, norm="ortho"), axis=-1, norm="ortho")
    return np.round(t).astype(np.int64)


def quantize(tc: np.ndarray, Q: np.ndarray) -> np.ndarray:
    """``quantize_TC`` (Encoder.py:787-789): np.round(TC / Q) -- half-to-even, second rounding."""
    return np.round(tc / Q).astype(np.int64)


def rle_symbols(levels: np.ndarray) -> list:
    """``entropy_encoder_block`` (Encoder.py:1086-1131): anti-diagonal scan + run-level list."""
    n = levels.shape[0]
    result = []
    nz_vals = []
    nz_count = 0
    zero_count = 0
    flag = 1
    for k in range(2 * n - 1):
        if k < n:
            i, j = 0, k
        else:
            i, j = k - n + 1, n - 1
        while i < n and j >= 0:
            v = int(levels[i][j])
            if v != 0:
                if flag == 0:
                    if zero_count:
                        result.append(zero_count)
                        zero_count = 0
                    nz_vals = []
                    nz_count = 0
                    flag = 1
                nz_vals.append(v)
                nz_count += 1
            else:
                if flag == 1:
                    if nz_count:
                        result.append(-nz_count)
                        result.extend(nz_vals)
                        nz_vals = []
                        nz_count = 0
                    zero_count = 0
                    flag = 0
                zero_count += 1
            i += 1
            j -= 1
    if nz_count:
        result.append(-nz_count)
        result.extend(nz_vals)
    if zero_count:
        result.append(0)
    return result


_SCAN_CACHE = {}


def scan_order(n: int) -> np.ndarray:
    """Row-major indices in the anti-diagonal scan order of Encoder.py:1095-1123 (A3: 0,1,4,2,5,8,... for n=4)."""
    if n not in _SCAN_CACHE:
        idx = []
        for k in range(2 * n - 1):
            i, j = (0, k) if k < n else (k - n + 1, n - 1)
            while i < n and j >= 0:
                idx.append(i * n + j)
                i += 1
                j -= 1
        _SCAN_CACHE[n] = np.array(idx, np.int64)
    return _SCAN_CACHE[n]


def rle_length(levels: np.ndarray) -> int:
    """``len(entropy_encoder_block(...))`` in closed form: #non-zeros + #non-zero runs + #zero runs."""
    n = levels.shape[0]
    s = levels.reshape(-1)[scan_order(n)] != 0
    starts = np.empty_like(s)
    starts[0] = True
    starts[1:] = s[1:] != s[:-1]
    return int(s.sum() + starts.sum())


# --------------------------------------------------------------------------------------------------------------
# half-pel reference frames                                                                   Encoder.py:388-406
# --------------------------------------------------------------------------------------------------------------
def halfpel_frames(ref_frames: list) -> list:
    """``frac_me_reference_frame``: (H,W) -> (2H-1, 2W-1) float frames.

    ``np.copy(list)`` promotes to a common dtype: uint8 only if *every* frame is uint8, in which case
    ``row + np.roll(row, -1)`` wraps mod 256 before the division (quirk Q1); the column pass is float.
    """
    arr = np.copy(ref_frames)
    out = []
    for ar in arr:
        H, W = ar.shape
        avg = (ar + np.roll(ar, -1, axis=1)) / 2.0
        rows = np.empty((H, 2 * W), np.float64)
        rows[:, 0::2] = ar
        rows[:, 1::2] = avg
        rows = rows[:, :-1]
        avgc = (rows + np.roll(rows, -1, axis=0)) / 2.0
        up = np.empty((2 * H, 2 * W - 1), np.float64)
        up[0::2] = rows
        up[1::2] = avgc
        out.append(np.ceil(up[:-1]))
    return out


# --------------------------------------------------------------------------------------------------------------
# motion search
# --------------------------------------------------------------------------------------------------------------
def _valid_range(pos: np.ndarray, d: int, size: int, bs: int, second_test: bool) -> np.ndarray:
    """``0 <= p+d < size-bs`` and optionally ``0 <= p+d+2bs < size-bs`` (Encoder.py:695,698,728,730)."""
    p = pos + d
    ok = (p >= 0) & (p < size - bs)
    if second_test:
        ok &= (p + 2 * bs >= 0) & (p + 2 * bs < size - bs)
    return ok


def full_search(cur: np.ndarray, refs: list, bs: int, R: int, fme: bool):
    """``find_best_match`` (Encoder.py:678-717) for every ``bs x bs`` block of ``cur`` at once.

    ``refs`` are the frames the reference passes in (half-pel frames when ``fme``; then block positions are
    doubled by the caller, Encoder.py:558, and ``R`` is already doubled, Encoder.py:1649).
    Returns ``mv [nby,nbx,3]`` and ``sad [nby,nbx]`` (``-1`` where no candidate was valid: MAE = inf, quirk Q3).
    """
    H, W = cur.shape
    nby, nbx = H // bs, W // bs
    mult = 2 if fme else 1
    bx = (np.arange(nbx) * bs * mult)[None, :]
    by = (np.arange(nby) * bs * mult)[:, None]
    BIG = np.iinfo(np.int64).max
    best_sad = np.full((nby, nbx), BIG, np.int64)
    best_mv = np.zeros((nby, nbx, 3), np.int64)
    best_l1 = np.zeros((nby, nbx), np.int64)
    cur_i = cur.astype(np.int64)
    pad = R + 2 * bs + 2
    for ref_idx, ref in enumerate(refs):
        Hr, Wr = ref.shape
        ref_i = np.asarray(ref).astype(np.int64)
        if fme:
            planes = {}
            for py in (0, 1):
                for px in (0, 1):
                    p = ref_i[py::2, px::2]
                    pp = np.zeros((H + 2 * pad, W + 2 * pad), np.int64)
                    pp[pad:pad + p.shape[0], pad:pad + p.shape[1]] = p
                    planes[(px, py)] = pp
        else:
            pp = np.zeros((H + 2 * pad, W + 2 * pad), np.int64)
            pp[pad:pad + Hr, pad:pad + Wr] = ref_i
        for dx in range(-R, R + 1):
            okx = _valid_range(bx, dx, Wr, bs, fme)
            if not okx.any():
                continue
            for dy in range(-R, R + 1):
                ok = okx & _valid_range(by, dy, Hr, bs, fme)
                if not ok.any():
                    continue
                if fme:
                    src = planes[(dx & 1, dy & 1)]
                    ox, oy = dx >> 1, dy >> 1
                else:
                    src = pp
                    ox, oy = dx, dy
                shifted = src[pad + oy:pad + oy + H, pad + ox:pad + ox + W]
                sad = np.abs(cur_i - shifted).reshape(nby, bs, nbx, bs).sum(axis=(1, 3))
                l1 = abs(dx) + abs(dy)
                # mae < best  |  (mae == best and is_better_mv)            Encoder.py:701-706, 771-773
                better = ok & ((sad < best_sad) | ((sad == best_sad) & ((l1 < best_l1) | ((l1 == best_l1) & (ref_idx < best_mv[..., 2])))))
                best_sad[better] = sad[better]
                best_l1[better] = l1
                best_mv[better] = (dx, dy, ref_idx)
    sad_out = np.where(best_sad == BIG, -1, best_sad)
    return best_mv, sad_out


def fast_search_block(cur_block, refs, x, y, bs, mvp, nref, fme: bool):
    """``fast_motion_estimation`` (Encoder.py:719-742): 3x3 around mvp, strict <, both validity tests always.

    Returns ``(mv, best_ref_idx)`` -- the second value is the reference index, not an MAE (quirk Q4).
    """
    best = INF
    best_mv = tuple(int(v) for v in mvp)
    best_ref = 0
    cb = cur_block.astype(np.int64)
    for ref_idx, ref in enumerate(refs[:nref]):
        Hr, Wr = ref.shape
        for dx in range(mvp[0] - 1, mvp[0] + 2):
            for dy in range(mvp[1] - 1, mvp[1] + 2):
                if 0 <= x + dx < Wr - bs and 0 <= y + dy < Hr - bs:
                    if 0 <= x + dx + bs * 2 < Wr - bs and 0 <= y + dy + bs * 2 < Hr - bs:
                        if fme:
                            rb = ref[y + dy:y + dy + bs * 2:2, x + dx:x + dx + bs * 2:2]
                        else:
                            rb = ref[y + dy:y + dy + bs, x + dx:x + dx + bs]
                        sad = int(np.abs(cb - rb.astype(np.int64)).sum())
                        if sad < best:
                            best = sad
                            best_mv = (dx, dy, ref_idx)
                            best_ref = ref_idx
    return best_mv, best_ref


def boundary_block(ref, y, x, bs):
    """``handle_boundary_conditions`` (Encoder.py:750-768): zero-padded *contiguous* crop."""
    Hh, Ww = ref.shape
    out = np.zeros((bs, bs), np.int64)
    ys, ye = max(y, 0), min(y + bs, Hh)
    xs, xe = max(x, 0), min(x + bs, Ww)
    if ye - ys > 0 and xe - xs > 0:
        out[ys - y:ye - y, xs - x:xe - x] = np.asarray(ref[ys:ye, xs:xe]).astype(np.int64)
    return out


def inter_predictor(refs, x, y, mv, bs, fme: bool, second_bs=None):
    """Predictor selection shared by ``calculate_inter_frame_residual`` (Encoder.py:432-460) and
    ``reconstruct_frame`` (Encoder.py:862-873, 907-919).  ``second_bs`` is the block size used in the *second*
    (``+2*bs``) FME test: the residual path uses the (sub-)block size, the reconstruction of split blocks uses
    ``pred + block_size < W - block_size`` with the *parent* size (quirk Q5, Encoder.py:910)."""
    ref = refs[mv[2]]
    Hr, Wr = ref.shape
    px, py = x + mv[0], y + mv[1]
    if 0 <= px < Wr - bs and 0 <= py < Hr - bs:
        if fme:
            if second_bs is None:
                ok = 0 <= px + bs * 2 < Wr - bs and 0 <= py + bs * 2 < Hr - bs
            else:
                ok = 0 <= px + second_bs < Wr - second_bs and 0 <= py + second_bs < Hr - second_bs
            if ok:
                return np.asarray(ref[py:py + bs * 2:2, px:px + bs * 2:2]).astype(np.int64)
            return np.full((bs, bs), 128, np.int64)
        return np.asarray(ref[py:py + bs, px:px + bs]).astype(np.int64)
    return boundary_block(ref, py, px, bs)


def intra_search_block(cur_block, ref_frame, x, y, bs, r):
    """``intra_find_best_match_horizontal`` (Encoder.py:1010-1045): key (SAD, |dx|, -dx); mv=-1 / 128 at x==0."""
    cb = cur_block.astype(np.int64)
    if x == 0:
        pred = np.full((bs, bs), 128, np.int64)
        return -1, int(np.abs(cb - pred).sum()), cb - pred
    best = INF
    best_mv = 0
    res = None
    for dx in range(-r, r + 1):
        if x + dx >= 0 and x + dx + bs <= ref_frame.shape[1]:
            rb = ref_frame[y:y + bs, x + dx:x + dx + bs]
            sad = int(np.abs(cb - rb).sum())
            if sad < best:
                best, best_mv, res = sad, dx, cb - rb
            elif sad == best:
                if abs(dx) <= abs(best_mv):
                    best_mv, res = dx, cb - rb
    return best_mv, best, res


# --------------------------------------------------------------------------------------------------------------
# the codec
# --------------------------------------------------------------------------------------------------------------
class OracleCodec:
    """Restatement of ``Y_Video_codec`` (Encoder.py:17-1898) limited to the hot path, producing packed arrays."""

    def __init__(self, h_pixels, w_pixels, frames, block_size, search_range, Qp, intra_dur, intra_mode=0, lam=None,
                 VBSEnable=False, nRefFrames=1, yuv_file=None, y_only_frame_arr=None, fast_me=False, FMEEnable=False,
                 RCFlag=None, targetBR=None, frame_rate=30, qp_rate_tables=None, intra_thresh=None, ParallelMode=0):
        assert intra_mode == 0, "intra_mode=1 crashes in the reference (Encoder.py:1399-1407); nothing to match"
        assert ParallelMode in (0, 1, 2), "ParallelMode=3 is racy/broken in the reference (SURVEY.md row 22)"
        assert h_pixels % block_size == 0 and w_pixels % block_size == 0, "Encoder.py:1382 requires multiples of i"
        self.H, self.W, self.F = h_pixels, w_pixels, frames
        self.bs, self.sub = block_size, block_size // 2
        self.r, self.qp0 = search_range, Qp
        self.intra_dur = intra_dur
        self.lam, self.vbs, self.nref = lam, VBSEnable, nRefFrames
        self.fast, self.fme = fast_me, FMEEnable
        self.rc = RCFlag if RCFlag is not None else 0
        self.tables, self.intra_thresh = qp_rate_tables, intra_thresh
        self.pm = ParallelMode
        self.frames = y_only_frame_arr
        self.nbx, self.nby = w_pixels // block_size, h_pixels // block_size
        self.bitrate_per_row = None
        if targetBR is not None:                                                   # Encoder.py:78-88
            num, unit = targetBR.split(" ")[:2]
            num = int(num)
            tb = num * 1024 if unit == "kbps" else num * 1048576 if unit == "mbps" else num
            self.bitrate_per_row = (tb // frame_rate) / (h_pixels / block_size)
        self.set_qp(Qp)

    # Encoder.py:948-959
    def set_qp(self, qp):
        self.qp = qp
        self.Q = q_matrix(self.bs, qp)
        self.Qm1 = q_matrix(self.sub, qm1_qp(qp))

    # Encoder.py:1133-1158
    def rd_cost(self, frame_type, split, mae, residuals):
        if split == 0:
            n = rle_length(quantize(dct2_round(residuals.astype(np.float64)), self.Q))
            bits = (8 if frame_type == 0 else 16) + 8 * n
        else:
            bits = 32 if frame_type == 0 else 64
            for r in residuals:
                bits += 8 * rle_length(quantize(dct2_round(r.astype(np.float64)), self.Qm1))
        return self.lam * bits + mae

    # Encoder.py:1576-1580
    def rc_qp(self, frame_type, budget):
        for qp, rate in enumerate(self.tables[frame_type]):
            if rate < budget:
                return qp, rate
        raise TypeError("no QP satisfies the row budget (the reference returns None and crashes here)")

    def rc_rows(self):
        """Per-row QPs (Encoder.py:1599-1609 / 1668-1678).  Data-independent (Q9); *both* flows index table 0."""
        qps = []
        budget = self.bitrate_per_row
        spent = 0
        for row in range(self.nby):
            budget = self.bitrate_per_row if row == 0 else self.bitrate_per_row + (budget - spent)
            qp, spent = self.rc_qp(0, budget)
            qps.append(qp)
        return qps

    # ---- intra ---------------------------------------------------------------- Encoder.py:1238-1347, 1582-1642
    def intra_flow(self, cur):
        bs, sub, r = self.bs, self.sub, self.r
        H, W = self.H, self.W
        ref_frame = np.full((H, W), 128, np.int64)      # Q6: the reference hard-codes (288, 352)
        cur_i = cur.astype(np.int64)
        mvs, residuals = [], []
        total_mae = 0.0
        for y in range(0, H, bs):
            for x in range(0, W, bs):
                use_vbs = self.vbs and x != 0 and y != 0
                if use_vbs:
                    v_mv, v_res, v_mae = [], [], 0
                    for yv in range(y, y + bs, sub):
                        for xv in range(x, x + bs, sub):
                            m, s, res = intra_search_block(cur_i[yv:yv + sub, xv:xv + sub], ref_frame, xv, yv, sub, r)
                            v_mv.append(m)
                            v_res.append(res)
                            v_mae = v_mae + s / (sub * sub)
                    v_mae = v_mae / 4
                m, s, res = intra_search_block(cur_i[y:y + bs, x:x + bs], ref_frame, x, y, bs, r)
                mae = s / (bs * bs)
                if use_vbs:
                    if self.rd_cost(0, 0, mae, res) < self.rd_cost(0, 1, v_mae, v_res):
                        mvs.append((0, m))
                        residuals.append((0, res))
                    else:
                        mvs.append((1, v_mv))
                        residuals.append((1, v_res))
                    mae = v_mae
                else:
                    mvs.append((0, m))
                    residuals.append((0, res))
                # lossless reconstruction into the search frame (Encoder.py:1329-1338): pred + residual == cur
                ref_frame[y:y + bs, x:x + bs] = cur_i[y:y + bs, x:x + bs]
                total_mae += mae
        avg_mae = total_mae / (len(mvs) or 1)
        qp_rows = self.rc_rows() if self.rc > 0 else []
        levels, qsize, row_sizes = self.quantise_blocks(residuals, qp_rows)
        recon = self.intra_reconstruct(mvs, levels, qp_rows)
        return mvs, avg_mae, levels, qp_rows, recon, qsize, row_sizes

    # Encoder.py:1611-1628 / 1680-1697
    def quantise_blocks(self, residuals, qp_rows):
        levels = []
        qsize = 0
        row_sizes = [0]
        for n, (split, res) in enumerate(residuals):
            if self.rc > 0 and n % self.nbx == 0:
                self.set_qp(qp_rows[n // self.nbx])
            if split == 0:
                q = quantize(dct2_round(res.astype(np.float64)), self.Q)
                qsize += rle_length(q)
                levels.append((0, q))
            else:
                qs = []
                for sres in res:
                    q = quantize(dct2_round(sres.astype(np.float64)), self.Qm1)
                    qsize += rle_length(q)
                    qs.append(q)
                levels.append((1, qs))
            if (n + 1) % self.nbx == 0:
                row_sizes.append(qsize)
        return levels, qsize, [row_sizes[i] - row_sizes[i - 1] for i in range(1, len(row_sizes))]

    # Encoder.py:1350-1417
    def intra_reconstruct(self, mvs, levels, qp_rows):
        bs, sub = self.bs, self.sub
        recon = np.full((self.H, self.W), 128, np.int64)       # float, unclipped in the reference
        res_blocks = []
        for n, (split, lv) in enumerate(levels):
            if self.rc > 0 and n % self.nbx == 0:
                self.set_qp(qp_rows[n // self.nbx])
            if split == 0:
                res_blocks.append((0, idct2_round(lv * self.Q)))
            else:
                res_blocks.append((1, [idct2_round(l * self.Qm1) for l in lv]))
        idx = 0
        for y in range(0, self.H, bs):
            for x in range(0, self.W, bs):
                split, rb = res_blocks[idx]
                if x == 0:
                    block = 128 + rb
                elif split == 0:
                    m = mvs[idx][1]
                    block = recon[y:y + bs, x + m:x + m + bs] + rb
                else:
                    block = np.ones((bs, bs), np.int64)
                    j = 0
                    for yv in range(y, y + bs, sub):
                        for xv in range(x, x + bs, sub):
                            m = mvs[idx][1][j]
                            block[yv - y:yv - y + sub, xv - x:xv - x + sub] = recon[yv:yv + sub, xv + m:xv + m + sub] + rb[j]
                            j += 1
                recon[y:y + bs, x:x + bs] = block
                idx += 1
        return (recon & 0xFF).astype(np.uint8)                  # astype(np.uint8) wraps mod 256 (A5)

    # ---- inter ---------------------------------------------------------------- Encoder.py:462-675, 1644-1709
    def inter_flow(self, cur, ref_frames):
        bs, sub = self.bs, self.sub
        fme = self.fme
        mult = 2 if fme else 1
        refs = halfpel_frames(ref_frames) if fme else ref_frames            # Encoder.py:1649
        R = self.r * 2 if fme else self.r
        cur_i = cur.astype(np.int64)
        parallel = self.pm in (1, 2)
        use_fast = self.fast and not (parallel and self.pm == 1)            # Encoder.py:641
        nref_fast = 1 if parallel else self.nref                            # Encoder.py:590
        if not use_fast:
            mv_b, sad_b = full_search(cur_i, refs, bs, R, fme)
            if self.vbs:
                mv_s, sad_s = full_search(cur_i, refs, sub, R, fme)
        if parallel and self.vbs and self.fast:
            raise UnboundLocalError("reference reads mvp before assignment here (Encoder.py:616)")
        mvs, residuals = [], []
        total_mae = 0.0
        mvp = (0, 0, 0)
        for by in range(self.nby):
            for bx in range(self.nbx):
                x, y = bx * bs, by * bs
                use_vbs = self.vbs and x != 0 and y != 0
                if use_vbs:
                    v_mv, v_res, v_mae = [], [], 0
                    for k in range(4):
                        xv, yv = x + (k % 2) * sub, y + (k // 2) * sub
                        cb = cur_i[yv:yv + sub, xv:xv + sub]
                        if use_fast:
                            m, mae_k = fast_search_block(cb, refs, xv * mult, yv * mult, sub, mvp, nref_fast, fme)
                        else:
                            m = tuple(int(v) for v in mv_s[yv // sub, xv // sub])
                            s = sad_s[yv // sub, xv // sub]
                            mae_k = INF if s < 0 else s / (sub * sub)
                        v_mv.append(m)
                        v_res.append(cb - inter_predictor(refs, xv * mult, yv * mult, m, sub, fme))
                        v_mae = v_mae + mae_k
                    v_mae = v_mae / 4
                cb = cur_i[y:y + bs, x:x + bs]
                if use_fast:
                    if parallel:
                        mvp = (0, 0, 0)                                      # Encoder.py:642
                    m, mae = fast_search_block(cb, refs, x * mult, y * mult, bs, mvp, nref_fast, fme)
                else:
                    m = tuple(int(v) for v in mv_b[by, bx])
                    mae = INF if sad_b[by, bx] < 0 else sad_b[by, bx] / (bs * bs)
                res = cb - inter_predictor(refs, x * mult, y * mult, m, bs, fme)
                if use_vbs:
                    if self.rd_cost(1, 0, mae, res) < self.rd_cost(1, 1, v_mae, v_res):    # Encoder.py:565-573
                        mvs.append((0, m))
                        residuals.append((0, res))
                    else:
                        mvs.append((1, v_mv))
                        residuals.append((1, v_res))
                    mae = v_mae
                else:
                    mvs.append((0, m))
                    residuals.append((0, res))
                total_mae += mae
                mvp = m                                                      # Encoder.py:581
        avg_mae = total_mae / (len(mvs) or 1)
        qp_rows = self.rc_rows() if self.rc > 0 else []
        levels, qsize, row_sizes = self.quantise_blocks(residuals, qp_rows)
        recon = self.inter_reconstruct(mvs, ref_frames, levels, qp_rows)
        return mvs, avg_mae, levels, qp_rows, recon, qsize, row_sizes

    # Encoder.py:831-932
    def inter_reconstruct(self, mvs, ref_frames, levels, qp_rows):
        bs, sub, fme = self.bs, self.sub, self.fme
        mult = 2 if fme else 1
        refs = halfpel_frames(ref_frames) if fme else ref_frames            # Encoder.py:834
        recon = np.zeros((self.H, self.W), np.uint8)
        for idx, (split, mv) in enumerate(mvs):
            if self.rc > 0 and idx % self.nbx == 0:
                self.set_qp(qp_rows[idx // self.nbx])
            y, x = (idx // self.nbx) * bs, (idx % self.nbx) * bs
            if split == 0:
                pred = inter_predictor(refs, x * mult, y * mult, mv, bs, fme)
                recon[y:y + bs, x:x + bs] = ((pred + idct2_round(levels[idx][1] * self.Q)) & 0xFF).astype(np.uint8)
            else:
                for k, m in enumerate(mv):
                    xv, yv = x + (k % 2) * sub, y + (k // 2) * sub
                    pred = inter_predictor(refs, xv * mult, yv * mult, m, sub, fme, second_bs=bs)   # quirk Q5
                    recon[yv:yv + sub, xv:xv + sub] = ((pred + idct2_round(levels[idx][1][k] * self.Qm1)) & 0xFF).astype(np.uint8)
        return recon

    # ---- sequence ------------------------------------------------------------------------ Encoder.py:1790-1898
    def encode(self):
        H, W = self.H, self.W
        ref_frames = [np.ones((H, W)) * 128]
        out = dict(frame_types=[], mvs=[], levels=[], qp_rows=[], recon=[], psnr=[], mae=[], qsize=[], row_sizes=[])
        for i in range(self.F):
            cur = self.frames[i]
            if i % self.intra_dur == 0 and self.pm != 1:
                self.set_qp(self.qp0)
                mvs, mae, levels, qp_rows, recon, qsize, row_sizes = self.intra_flow(cur)
                ftype = 0
            else:
                self.set_qp(self.qp0)
                if self.pm == 1:
                    ref_frames = [np.ones((H, W)) * 128]                    # Encoder.py:1846
                mvs, mae, levels, qp_rows, recon, qsize, row_sizes = self.inter_flow(cur, ref_frames)
                ftype = 1
                if self.rc > 1 and qsize > self.intra_thresh:               # Encoder.py:1851-1856 (no set_Qp reset)
                    mvs, mae, levels, qp_rows, recon, qsize, row_sizes = self.intra_flow(cur)
                    ftype = 0
            out["frame_types"].append(ftype)
            out["mvs"].append(mvs)
            out["levels"].append(levels)
            out["qp_rows"].append(qp_rows)
            out["recon"].append(recon)
            out["mae"].append(mae)
            out["qsize"].append(qsize)
            out["row_sizes"].append(row_sizes)
            if i < self.F - 1:                                               # Encoder.py:1864-1867
                if len(ref_frames) >= self.nref:
                    ref_frames.pop(0)
                ref_frames.append(recon)
            out["psnr"].append(psnr(cur, recon))
        out["recon"] = np.stack(out["recon"])
        return out


def psnr(a, b):
    """PSNR as the harness stub computes it (skimage is absent): 10 log10(255^2 / MSE) in float64."""
    mse = np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)
    return float("inf") if mse == 0 else float(10.0 * np.log10(255.0 ** 2 / mse))


# --------------------------------------------------------------------------------------------------------------
# text streams                                                                             Encoder.py:1419-1542
# --------------------------------------------------------------------------------------------------------------
def mv_text_frame(frame_type, mvs, qp_rows, nbx, rc):
    """``"<type>|" + differential_encoder_frame(...)`` (Encoder.py:1419-1520, 1567)."""
    items = []
    rc_on = rc is not None and rc > 0
    ref_qp = 0
    if frame_type == 0:
        ref = 0
        for j, (split, mv) in enumerate(mvs):
            row_start = rc_on and j % nbx == 0
            pre = ""
            if row_start:
                pre = str(qp_rows[j // nbx] - ref_qp) + "@"
                ref_qp = qp_rows[j // nbx]
            if split == 0:
                items.append(pre + "0'(" + str(mv - ref) + ")")
                ref = mv
            else:   # never at a row start (VBS requires x != 0), so the diff_mv@ formatting bug at :1460 is unreachable
                d = []
                for s in mv:
                    d.append(str(s - ref))
                    ref = s
                items.append(pre + "1'(" + ",".join(d) + ")")
    else:
        ref = (0, 0, 0)
        for j, (split, mv) in enumerate(mvs):
            row_start = rc_on and j % nbx == 0
            pre = ""
            if row_start:
                pre = str(qp_rows[j // nbx] - ref_qp) + "@"
                ref_qp = qp_rows[j // nbx]
            if split == 0:
                items.append(pre + "0'" + str((mv[0] - ref[0], mv[1] - ref[1], mv[2] - ref[2])))
                ref = mv
            else:
                d = []
                for s in mv:
                    d.append(str((s[0] - ref[0], s[1] - ref[1], s[2] - ref[2])))
                    ref = s
                items.append(pre + "1'(" + ",".join(d) + ")")
    return str(frame_type) + "|" + ";".join(items)


def res_text_frame(levels):
    """``entropy_encoder_frame`` (Encoder.py:1522-1542), canonical ints (Q10)."""
    items = []
    for split, lv in levels:
        if split == 0:
            items.append("0'(" + str(rle_symbols(lv)) + ")")
        else:
            items.append("1'(" + ",".join(str(rle_symbols(l)) for l in lv) + ")")
    return ";".join(items)
