// Per-frame kernels other than the exhaustive search: half-pel planes, fast ME chain, intra search, the
// residual/transform/quant/RD/reconstruct "finish" kernels and the intra reconstruction chain.
#pragma once
#include "so_common.cuh"
#include "so_transform.cuh"
#include "../../include/streamoptima_b200.h"

// ------------------------------------------------------------------------------------------------------------
// reference ring planes: half-pel phases (frac_me_reference_frame, Encoder.py:388-406; appendix A1) and the four
// byte-shifted copies of every phase that the TMA search windows are loaded from
// ------------------------------------------------------------------------------------------------------------
// slot layout [phase 0..3][shift 0..3][H][pitch]; input = phase 0 / shift 0 (the reconstruction).
// phase 1 = ceil((a[x]+a[x+1])/2) with the uint8 wrap of the sum when `wrap` (quirk Q1), phase 2 = vertical (never
// wraps: the column pass is float), phase 3 = ceil((sh[y]+sh[y+1])/4) on the (possibly wrapped) horizontal sums.
// One thread per 4 output bytes; columns past the frame edge replicate the last column (only invalid candidates and
// never-sampled half-pel positions see them).
// src: the frame the planes are derived from -- the dense reconstruction at push time (write_p0 = 1: it is also stored as
// phase 0 / shift 0), or the slot's own phase-0 plane when only the wrap mode changed (write_p0 = 0).
__global__ void ring_planes_kernel(uint8_t* slot0, size_t unit_stride, size_t plane_bytes, const uint8_t* src, size_t src_unit_stride,
                                   int src_pitch, int W, int H, int pitch, int fme, int wrap, int write_p0) {
    pdl_trigger();
    pdl_wait();
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int x = x4 * 4;
    if (x >= W) return;
    uint8_t* base = slot0 + blockIdx.z * unit_stride;
    const uint8_t* sbase = src + blockIdx.z * src_unit_stride;
    const int yn = min(y + 1, H - 1);
    if (wrap && fme) {
        // uint8-wrap variant (quirk Q1; every frame after the first nRefFrames): all four pixels at once with packed byte
        // arithmetic.  (s + 1) >> 1 = vavgu4(s, 0); (s0 + s1 + 3) >> 2 = ceil(ceil((s0 + s1) / 2) / 2) = vavgu4(vavgu4(s0, s1), 0)
        const uint32_t* r0 = reinterpret_cast<const uint32_t*>(sbase + (size_t)y * src_pitch);
        const uint32_t* r1 = reinterpret_cast<const uint32_t*>(sbase + (size_t)yn * src_pitch);
        const bool in1 = x + 4 < W;
        const uint32_t w00 = r0[x4], w10 = r1[x4];
        const uint32_t w01 = in1 ? r0[x4 + 1] : (w00 >> 24) * 0x01010101u;      // replicate the last column
        const uint32_t w11 = in1 ? r1[x4 + 1] : (w10 >> 24) * 0x01010101u;
        const uint32_t s0l = __vadd4(w00, __funnelshift_r(w00, w01, 8)), s0h = __vadd4(w01, w01 >> 8);
        const uint32_t s1l = __vadd4(w10, __funnelshift_r(w10, w11, 8)), s1h = __vadd4(w11, w11 >> 8);
        uint32_t lo[4], hi[4];
        lo[0] = w00; hi[0] = w01;
        lo[1] = __vavgu4(s0l, 0u); hi[1] = __vavgu4(s0h, 0u);
        lo[2] = __vavgu4(w00, w10); hi[2] = __vavgu4(w01, w11);
        lo[3] = __vavgu4(__vavgu4(s0l, s1l), 0u); hi[3] = __vavgu4(__vavgu4(s0h, s1h), 0u);
        const size_t o = (size_t)y * pitch + x;
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (p == 0 && c == 0 && !write_p0) continue;
                *reinterpret_cast<uint32_t*>(base + (size_t)(p * 4 + c) * plane_bytes + o) = c ? __funnelshift_r(lo[p], hi[p], 8 * c) : lo[p];
            }
        return;
    }
    int a0[9], a1[9];
    {
        const uint32_t* r0 = reinterpret_cast<const uint32_t*>(sbase + (size_t)y * src_pitch);
        const uint32_t* r1 = reinterpret_cast<const uint32_t*>(sbase + (size_t)yn * src_pitch);
        uint32_t w0[3], w1[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const bool in = x + 4 * q < W;
            w0[q] = in ? r0[x4 + q] : 0u;
            w1[q] = in ? r1[x4 + q] : 0u;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int col = min(x + i, W - 1) - x;            // replicate the last column
            a0[i] = (w0[col >> 2] >> (8 * (col & 3))) & 255;
            a1[i] = (w1[col >> 2] >> (8 * (col & 3))) & 255;
        }
    }
    int ph[4][7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        int sh0 = a0[i] + a0[i + 1], sh1 = a1[i] + a1[i + 1];
        if (wrap) { sh0 &= 255; sh1 &= 255; }
        ph[0][i] = a0[i];
        ph[1][i] = (sh0 + 1) >> 1;
        ph[2][i] = (a0[i] + a1[i] + 1) >> 1;
        ph[3][i] = (sh0 + sh1 + 3) >> 2;
    }
    const size_t o = (size_t)y * pitch + x;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        if (p > 0 && !fme) break;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (p == 0 && c == 0 && !write_p0) continue;      // the input plane itself
            const uint32_t v = (uint32_t)ph[p][c] | ((uint32_t)ph[p][c + 1] << 8) | ((uint32_t)ph[p][c + 2] << 16) | ((uint32_t)ph[p][c + 3] << 24);
            *reinterpret_cast<uint32_t*>(base + (size_t)(p * 4 + c) * plane_bytes + o) = v;
        }
    }
}

// Integer search only (no half-pel phases): the frame and its three byte-shifted copies, 16 pixels per thread with 128-bit
// stores (the generic kernel above computes all four phases per pixel before it knows that only phase 0 is stored).  W % 16 == 0.
__global__ void __launch_bounds__(128) ring_shift_kernel(uint8_t* slot0, size_t unit_stride, size_t plane_bytes, const uint8_t* src,
                                                         size_t src_unit_stride, int src_pitch, int W, int pitch) {
    pdl_trigger();
    pdl_wait();
    const int x16 = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, x = x16 * 16;
    if (x >= W) return;
    const uint8_t* srow = src + blockIdx.z * src_unit_stride + (size_t)y * src_pitch + x;
    const uint4 v = *reinterpret_cast<const uint4*>(srow);
    // word past the 16 pixels; past the frame edge the last column is replicated (only invalid candidates see it)
    const uint32_t nx = x + 16 < W ? *reinterpret_cast<const uint32_t*>(srow + 16) : (v.w >> 24) * 0x01010101u;
    uint8_t* base = slot0 + blockIdx.z * unit_stride + (size_t)y * pitch + x;
    *reinterpret_cast<uint4*>(base) = v;
#pragma unroll
    for (int c = 1; c < 4; ++c)
        *reinterpret_cast<uint4*>(base + (size_t)c * plane_bytes) =
            make_uint4(__funnelshift_r(v.x, v.y, 8 * c), __funnelshift_r(v.y, v.z, 8 * c), __funnelshift_r(v.z, v.w, 8 * c), __funnelshift_r(v.w, nx, 8 * c));
}

__global__ void ring_fill_kernel(uint8_t* dst, size_t dst_unit_stride, size_t bytes16, uint32_t v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < bytes16) reinterpret_cast<uint4*>(dst + blockIdx.y * dst_unit_stride)[i] = make_uint4(v, v, v, v);
}

// ------------------------------------------------------------------------------------------------------------
// launch arguments shared by the finish / intra / fast kernels
// ------------------------------------------------------------------------------------------------------------
struct FlowArgs {
    FrameGeom g;                 // g.bs = parent block size
    RefRing ring;
    const uint8_t* cur;          // [unit][H][W] dense
    size_t cur_unit_stride;
    int unit0;                   // first unit processed by this launch (grid.y / grid.x index is added)
    int vbs, fast, chain;        // chain: fast ME carries mvp across blocks (ParallelMode 0)
    uint32_t mae_den, frame_type; // written into the frame's statistics by block 0 of the finish kernels
    int me_packed;               // me_parent / me_sub hold packed search keys (1: 64-bit keys, 2: compact keys of so_me_ring2.cuh); the finish kernel decodes and resets them
    int nref_fast;               // refs[:nRefFrames] of fast ME (1 in ParallelMode 2, Encoder.py:590)
    int qp_final;                // QP when no rate control
    int qp_rd;                   // QP of self.Q during prediction (RD cost)
    const int* qp_rows;          // device, per block row, or nullptr
    const int* qp_blocks;        // device, per block (ROI extension, overrides qp_rows / qp_final), or nullptr
    double lam;
    MeResult* me_parent;         // [unit][nblk]
    MeResult* me_sub;            // [unit][4*nblk], sub grid (2nby x 2nbx)
    size_t me_parent_stride, me_sub_stride;
    int16_t* res_frame;          // intra: dequantised residual, [unit][H][W]
    int32_t* band;               // intra: unclipped reconstruction, [unit][H][W]
    // outputs (dense per unit)
    uint8_t* split; int16_t* mv; int16_t* levels; uint8_t* recon; uint32_t* row_sizes; so_frame_stats* stats;
    size_t split_stride, mv_stride, frame_stride, rows_stride, stats_stride;   // per-unit strides in elements
    size_t scratch_stride;       // per-unit stride of res_frame / band (one frame)
    const short4* mvp_in;        // fast ME, table-driven chain: predictor (x, y, ref) of every block, [unit][nblk]; else nullptr
    size_t mvp_in_stride;
    uint32_t* blk_len;           // optional: RLE symbols of every block (all its sub-blocks), [unit][nblk] -- the count pass of the
    size_t blk_len_stride;       // symbol packer comes for free from the size statistics
};

// ME results are either plain MeResult records (fast ME, intra search) or the packed 64-bit keys the exhaustive search
// merges with atomicMin.
__device__ __forceinline__ MeResult me_get(const MeResult* p, int packed, int R) {
    if (!packed) return *p;
    const unsigned long long key = *reinterpret_cast<const unsigned long long*>(p);
    MeResult r;
    if (key == ~0ull) { r.dx = 0; r.dy = 0; r.ref = 0; r.none = 1; r.sad = 0; }          // (0,0,0), MAE = inf (Encoder.py:684-685)
    else if (packed == 2) {
        // compact key of the item-ring search kernel (so_me_ring2.cuh): SAD (16) | |dx|+|dy| (8) | ref (8) | dx + R (7) | dy > 0 (1)
        const uint32_t lo = (uint32_t)key, hi = (uint32_t)(key >> 32);
        const int l1 = (int)((lo >> 16) & 0xFFu), dx = (int)((lo >> 1) & 0x7Fu) - R, ady = l1 - abs(dx);
        r.sad = ((hi & 0xFFu) << 8) | (lo >> 24); r.ref = (int16_t)((lo >> 8) & 0xFFu);
        r.dx = (int16_t)dx; r.dy = (int16_t)((lo & 1u) ? ady : -ady); r.none = 0;
    } else {
        r.sad = (uint32_t)(key >> 40); r.ref = (int16_t)((key >> 16) & 0xFF);
        r.dx = (int16_t)((int)((key >> 8) & 0xFF) - R); r.dy = (int16_t)((int)(key & 0xFF) - R); r.none = 0;
    }
    return r;
}

__device__ __forceinline__ double me_mae(const MeResult& m, int n, int fast) {
    if (fast) return (double)m.sad;                         // quirk Q4: the "MAE" is the reference index
    if (m.none) return __longlong_as_double(0x7FF0000000000000LL);
    return (double)m.sad / (double)(n * n);
}

// number of RLE symbols of the block(s) whose quantised values are spread one per thread:
// #non-zeros + #runs (oracle/codec_oracle.py:rle_length).  nzbuf: BS*BS ints of shared memory.
template <int BS, int N>
__device__ __forceinline__ int rle_len_cta(int level, int u, int v, int sub_index, int* nzbuf, bool active) {
    const int p = c_scanpos[tbl_off(N) + u * N + v];
    const int base = sub_index * N * N;
    __syncthreads();
    if (active) nzbuf[base + p] = (level != 0);
    __syncthreads();
    const bool nz = active && level != 0;
    const bool start = active && (p == 0 || nzbuf[base + p] != nzbuf[base + p - 1]);
    return __syncthreads_count(nz) + __syncthreads_count(start);
}

__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* sbuf) {
    // sbuf: 32 entries
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sbuf[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long s = 0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) s += sbuf[i];
    return s;
}

// ------------------------------------------------------------------------------------------------------------
// Plain exhaustive search, one warp per block, any block size >= 2 (find_best_match, Encoder.py:678-717).  Lanes stride
// over the candidates (ref, dx, dy); the packed 64-bit key (SAD, |dx|+|dy|, ref, dx, dy) carries the reference's replace
// rule (appendix A4), so the visiting order is free.  It serves the 2x2 sub-blocks of VBS with block_size 4, which are
// too narrow for the word-packed kernels, and (SO_ME_SIMPLE=1) as an independent cross-check of those kernels in tests.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) me_simple_kernel(const FlowArgs a, int bs, MeResult* out, size_t out_unit_stride) {
    const FrameGeom& g = a.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nbx = g.W / bs, nby = g.H / bs;
    const int blk = blockIdx.x * 4 + warp, unit = a.unit0 + blockIdx.y;
    if (blk >= nbx * nby) return;
    const int bx = blk % nbx, by = blk / nbx;
    const int x = bx * bs, y = by * bs;
    const int mult = g.fme ? 2 : 1;
    int xlo, xhi, ylo, yhi;
    valid_range(x, g.W, bs, g.fme, g.fme, xlo, xhi);
    valid_range(y, g.H, bs, g.fme, g.fme, ylo, yhi);
    xlo = max(xlo, -g.R); xhi = min(xhi, g.R);
    ylo = max(ylo, -g.R); yhi = min(yhi, g.R);
    const uint8_t* cur = a.cur + unit * a.cur_unit_stride + (size_t)y * g.W + x;
    unsigned long long best = ~0ull;
    const int nx = xhi - xlo + 1, ny = yhi - ylo + 1;
    if (nx > 0 && ny > 0) {
        const int per_ref = nx * ny, total = per_ref * g.nref;
        for (int c = lane; c < total; c += 32) {
            const int ref = c / per_ref, rem = c - ref * per_ref;
            const int dx = xlo + rem / ny, dy = ylo + rem % ny;
            const int Xh = x * mult + dx, Yh = y * mult + dy;
            const int ph = g.fme ? (((Yh & 1) << 1) | (Xh & 1)) : 0;
            const int X0 = g.fme ? (Xh >> 1) : Xh, Y0 = g.fme ? (Yh >> 1) : Yh;
            const uint8_t* pl = a.ring.plane(unit, ref, ph) + (size_t)Y0 * g.pitch + X0;
            unsigned sad = 0;
            for (int j = 0; j < bs; ++j)
                for (int i = 0; i < bs; ++i) sad += (unsigned)abs((int)cur[(size_t)j * g.W + i] - (int)pl[(size_t)j * g.pitch + i]);
            const unsigned long long key = ((unsigned long long)sad << 40) | ((unsigned long long)(abs(dx) + abs(dy)) << 24) |
                                           ((unsigned long long)ref << 16) | ((unsigned long long)(dx + g.R) << 8) | (unsigned long long)(dy + g.R);
            best = key < best ? key : best;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        best = other < best ? other : best;
    }
    if (lane == 0) *reinterpret_cast<unsigned long long*>(out + unit * out_unit_stride + blk) = best;
}

// ------------------------------------------------------------------------------------------------------------
// inter finish: residual -> DCT -> [VBS RD decision] -> quant -> RLE size -> dequant -> IDCT -> reconstruct
// (inter_prediction tail Encoder.py:564-581, complete_inter_flow :1680-1697, reconstruct_frame :831-932)
// one CTA per block, one thread per pixel (blockDim = max(BS*BS, 32); extra threads are inactive pixels)
// ------------------------------------------------------------------------------------------------------------
template <int BS>
__global__ void inter_finish_kernel(const FlowArgs a) {
    constexpr int S = BS / 2;
    constexpr int P = BS + 1;
    __shared__ double ws[BS * P];
    __shared__ int nzbuf[BS * BS];
    __shared__ unsigned long long sbuf[32];
    const FrameGeom& g = a.g;
    const int blk = blockIdx.x, unit = a.unit0 + blockIdx.y;
    const int bx = blk % g.nbx, by = blk / g.nbx;
    const int t = threadIdx.x;
    const bool active = t < BS * BS;
    const int i = active ? t % BS : 0, j = active ? t / BS : 0;
    const int x = bx * BS, y = by * BS;
    const int mult = g.fme ? 2 : 1;
    RefList rl;
    for (int r = 0; r < g.nref; ++r)
        for (int ph = 0; ph < 4; ++ph) rl.plane[r][ph] = a.ring.plane(unit, r, ph);

    const int c = a.cur[unit * a.cur_unit_stride + (size_t)(y + j) * g.W + x + i];
    MeResult* pme = a.me_parent + unit * a.me_parent_stride + blk;
    const MeResult mp = me_get(pme, a.me_packed, g.R);
    MeResult msub[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        msub[kk] = mp;
        if (a.vbs) msub[kk] = me_get(a.me_sub + unit * a.me_sub_stride + (by * 2 + (kk >> 1)) * (g.nbx * 2) + bx * 2 + (kk & 1), a.me_packed, g.R);
    }
    if (a.me_packed) {          // every thread holds its copy: reset the keys for the next frame's atomicMin merge
        __syncthreads();
        if (t == 0) *reinterpret_cast<unsigned long long*>(pme) = ~0ull;
        if (a.vbs && t < 4)
            *reinterpret_cast<unsigned long long*>(a.me_sub + unit * a.me_sub_stride + (by * 2 + (t >> 1)) * (g.nbx * 2) + bx * 2 + (t & 1)) = ~0ull;
    }
    const PredSel selp = pred_select(g, x * mult, y * mult, mp.dx, mp.dy, BS, -1);
    const int predp = pred_sample(g, rl, selp, mp.ref, i, j);
    const int resp = c - predp;

    if (active) ws[j * P + i] = so_i2d(resp);
    transform2d<BS, BS, false>(ws, t);
    const int tcp = active ? so_d2i_rint(ws[j * P + i]) : 0;

    const bool eligible = a.vbs && bx != 0 && by != 0;
    const int qrow = a.qp_blocks ? a.qp_blocks[blk] : (a.qp_rows ? a.qp_rows[by] : a.qp_final);
    int split = 0;
    double mae_blk = me_mae(mp, BS, a.fast);
    // sub-block data (computed only when eligible; eligibility is CTA-uniform)
    const int k = (j >= S ? 2 : 0) + (i >= S ? 1 : 0);
    const int si = i % S, sj = j % S;
    int tcs = 0, preds_q5 = 0;
    MeResult ms = mp;
    if (eligible) {
        ms = k == 0 ? msub[0] : (k == 1 ? msub[1] : (k == 2 ? msub[2] : msub[3]));
        const int xs = x + (k & 1) * S, ys = y + (k >> 1) * S;
        const PredSel sels = pred_select(g, xs * mult, ys * mult, ms.dx, ms.dy, S, -1);
        const int preds = pred_sample(g, rl, sels, ms.ref, si, sj);
        const PredSel selq = pred_select(g, xs * mult, ys * mult, ms.dx, ms.dy, S, BS);     // quirk Q5
        preds_q5 = pred_sample(g, rl, selq, ms.ref, si, sj);
        if (active) ws[j * P + i] = so_i2d(c - preds);
        transform2d<BS, S, false>(ws, t);
        tcs = active ? so_d2i_rint(ws[j * P + i]) : 0;
        // RD costs with the prediction-time QP (calculate_RD_cost, Encoder.py:1133-1158)
        const int lenp = rle_len_cta<BS, BS>(quant_rhe(tcp, q_shift(j, i, BS, a.qp_rd)), j, i, 0, nzbuf, active);
        const int qs = a.qp_rd > 0 ? a.qp_rd - 1 : a.qp_rd;
        const int lens = rle_len_cta<BS, S>(quant_rhe(tcs, q_shift(sj, si, S, qs)), sj, si, k, nzbuf, active);
        // vbs_mae = (sum of the four sub MAEs) / 4, accumulated in Z order
        double vm = 0.0;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) vm = __dadd_rn(vm, me_mae(msub[kk], S, a.fast));
        vm = vm / 4.0;
        const double rd_bs = __dadd_rn(__dmul_rn(a.lam, (double)(16 + 8 * lenp)), mae_blk);
        const double rd_vbs = __dadd_rn(__dmul_rn(a.lam, (double)(64 + 8 * lens)), vm);
        split = (rd_bs < rd_vbs) ? 0 : 1;
        mae_blk = vm;
    }

    // final quantisation with the row QP (self.Q / self.Qm1 after set_Qp, Encoder.py:1668-1694)
    int level, shift, pred_rec;
    if (!split) {
        shift = q_shift(j, i, BS, qrow);
        level = quant_rhe(tcp, shift);
        pred_rec = predp;
    } else {
        const int qs = qrow > 0 ? qrow - 1 : qrow;
        shift = q_shift(sj, si, S, qs);
        level = quant_rhe(tcs, shift);
        pred_rec = preds_q5;
    }
    const int len = split ? rle_len_cta<BS, S>(level, sj, si, k, nzbuf, active)
                          : rle_len_cta<BS, BS>(level, j, i, 0, nzbuf, active);
    if (active) {
        a.levels[unit * a.frame_stride + (size_t)(y + j) * g.W + x + i] = (int16_t)level;
        ws[j * P + i] = so_i2d(level * (1 << shift));          // rescale_QTC, Encoder.py:820
    }
    if (split) transform2d<BS, S, true>(ws, t); else transform2d<BS, BS, true>(ws, t);
    unsigned long long se = 0;
    if (active) {
        const int rec = (pred_rec + so_d2i_rint(ws[j * P + i])) & 0xFF;     // astype(np.uint8) wraps (A5)
        a.recon[unit * a.frame_stride + (size_t)(y + j) * g.W + x + i] = (uint8_t)rec;
        const int d = rec - c;
        se = (unsigned long long)(d * d);
    }
    se = block_sum_u64(se, sbuf);
    if (t == 0) {
        a.split[unit * a.split_stride + blk] = (uint8_t)split;
        int16_t* mvo = a.mv + unit * a.mv_stride + (size_t)blk * 12;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const MeResult m = split ? msub[kk] : mp;
            const bool on = split || kk == 0;
            mvo[kk * 3 + 0] = on ? m.dx : 0; mvo[kk * 3 + 1] = on ? m.dy : 0; mvo[kk * 3 + 2] = on ? m.ref : 0;
        }
        so_frame_stats* st = a.stats + unit * a.stats_stride;
        if (blk == 0) { st->mae_den = a.mae_den; st->frame_type = a.frame_type; }
        atomicAdd(reinterpret_cast<unsigned long long*>(&st->sse), se);
        atomicAdd(&st->qsize, (unsigned)len);
        if (a.blk_len) a.blk_len[unit * a.blk_len_stride + blk] = (uint32_t)len;
        atomicAdd(a.row_sizes + unit * a.rows_stride + by, (unsigned)len);
        // MAE numerator in units of 1/mae_den: full search 1/BS^2 (sub MAEs: sum sad_k/S^2/4 = sum sad_k/BS^2),
        // fast ME 1/4 (values are reference indices; VBS averages four of them)
        if (a.fast) {
            unsigned long long n = (unsigned long long)mp.sad * 4ull;
            if (eligible) {
                n = 0;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) n += msub[kk].sad;
            }
            atomicAdd(reinterpret_cast<unsigned long long*>(&st->mae_num), n);
        } else {
            unsigned long long n = mp.sad;
            bool inf = mp.none;
            if (eligible) {
                n = 0; inf = false;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) { n += msub[kk].sad; inf = inf || msub[kk].none; }
            }
            if (inf) atomicOr(&st->mae_inf, 1u); else atomicAdd(reinterpret_cast<unsigned long long*>(&st->mae_num), n);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// inter finish for 16x16 blocks, warp-synchronous: one half-warp per block (lane = block row), two blocks per warp, no
// block-wide barriers.  Same arithmetic as inter_finish_kernel<16> (which stays the generic version for 8x8 / 4x4).
//   * a lane owns one row of 16 pixels: 128-bit loads/stores of cur, levels and recon
//   * 2-D transforms through a per-warp shared tile: column pass (lane = column), row pass (lane = row)
//   * RLE symbol count from 16-bit non-zero row masks: the scan predecessor of (i, j) is (i-1, j+1) inside the block,
//     (k-1, 0) for the first-row element (0, k) and (n-1, i-1) for the last-column element (i, n-1); a run starts where
//     the non-zero flag differs from the predecessor's, so  len = #nonzeros + #starts  needs only shifts, XORs and
//     popcounts of row masks exchanged by shuffles / ballots (see the closed form in oracle/codec_oracle.py:rle_length)
// ------------------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ int rle_rows_partial(uint32_t m, int rr, uint32_t m_prev, uint32_t c0, uint32_t cl, uint32_t m_last) {
    constexpr uint32_t MASK = (1u << (N - 1)) - 1u;
    int v = __popc(m);
    if (rr > 0) v += __popc((m ^ (m_prev >> 1)) & MASK);
    else v += 1 + __popc(((m >> 1) ^ c0) & MASK) + __popc(((cl >> 1) ^ m_last) & MASK);
    return v;
}

#ifndef FIN16_MINB
#define FIN16_MINB 1                  // resident CTAs per SM the non-VBS finish kernel is compiled for (register cap)
#endif
template <bool VBS>
__global__ void __launch_bounds__(128, VBS ? 1 : FIN16_MINB) inter_finish16_kernel(const FlowArgs a) {
    constexpr int BS = 16, S = 8, P = 17;
    __shared__ double tiles[4][2][BS * P];
    pdl_trigger();
    pdl_wait();
    const FrameGeom& g = a.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = lane >> 4, r = lane & 15;
    const int nblk = g.nbx * g.nby;
    const int unit = a.unit0 + blockIdx.y;
    int blk = (blockIdx.x * 4 + warp) * 2 + h;
    const bool live = blk < nblk;
    if (!live) blk = nblk - 1;                        // keep the lanes in the warp-synchronous flow; results are not stored
    const int bx = blk % g.nbx, by = blk / g.nbx;
    const int x = bx * BS, y = by * BS;
    const int mult = g.fme ? 2 : 1;
    double* ws = tiles[warp][h];
    const unsigned FULL = 0xFFFFFFFFu;
    RefList rl;
    for (int q = 0; q < g.nref; ++q)
        for (int ph = 0; ph < 4; ++ph) rl.plane[q][ph] = a.ring.plane(unit, q, ph);

    // ---- current row, motion results (decoded from the packed keys, which are reset for the next frame)
    int c[BS];
    {
        const uint4 v = *reinterpret_cast<const uint4*>(a.cur + unit * a.cur_unit_stride + (size_t)(y + r) * g.W + x);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < BS; ++i) c[i] = (w[i >> 2] >> (8 * (i & 3))) & 255;
    }
    MeResult* pme = a.me_parent + unit * a.me_parent_stride + blk;
    const MeResult mp = me_get(pme, a.me_packed, g.R);
    MeResult msub[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        msub[kk] = mp;
        if (VBS && a.vbs) msub[kk] = me_get(a.me_sub + unit * a.me_sub_stride + (by * 2 + (kk >> 1)) * (g.nbx * 2) + bx * 2 + (kk & 1), a.me_packed, g.R);
    }
    __syncwarp();
    if (a.me_packed && live) {
        if (r == 0) *reinterpret_cast<unsigned long long*>(pme) = ~0ull;
        if (VBS && a.vbs && r < 4)
            *reinterpret_cast<unsigned long long*>(a.me_sub + unit * a.me_sub_stride + (by * 2 + (r >> 1)) * (g.nbx * 2) + bx * 2 + (r & 1)) = ~0ull;
    }

    // ---- whole-block predictor row and residual -> forward transform
    int predp[BS];
    {
        const PredSel sel = pred_select(g, x * mult, y * mult, mp.dx, mp.dy, BS, -1);
        if (sel.mode == 0) {
            // in-bounds predictor = 16 contiguous bytes of one phase plane (appendix A1); the copy shifted by X & 3 bytes
            // holds them at a word-aligned address: four 32-bit loads instead of sixteen byte gathers
            const int X = g.fme ? (sel.PX >> 1) : sel.PX, Y = (g.fme ? (sel.PY >> 1) : sel.PY) + r;
            const int ph = g.fme ? (((sel.PY & 1) << 1) | (sel.PX & 1)) : 0;
            const int cs = X & 3;
            const uint32_t* pw = reinterpret_cast<const uint32_t*>(a.ring.plane(unit, mp.ref, ph) + (size_t)cs * (a.ring.plane_stride >> 2) +
                                                                   (size_t)Y * g.pitch + (X - cs));
            const uint32_t w[4] = {__ldg(pw), __ldg(pw + 1), __ldg(pw + 2), __ldg(pw + 3)};
#pragma unroll
            for (int i = 0; i < BS; ++i) predp[i] = (w[i >> 2] >> (8 * (i & 3))) & 255;
        } else {
#pragma unroll
            for (int i = 0; i < BS; ++i) predp[i] = pred_sample(g, rl, sel, mp.ref, i, r);
        }
    }
#pragma unroll
    for (int i = 0; i < BS; ++i) ws[r * P + i] = so_i2d(c[i] - predp[i]);
    // column r, then row r: one copy of the straight-line transform in the instruction stream (the kernel is latency-bound
    // and its code does not fit the instruction cache when every pass is inlined separately)
    if constexpr (VBS) {                       // the VBS variant is register-bound: run-time strides cost it 29 registers
        __syncwarp();
        dct1d<BS>(ws + r, P);
        __syncwarp();
        dct1d<BS>(ws + r * P, 1);
    } else {
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            __syncwarp();
            dct1d<BS>(pass ? ws + r * P : ws + r, pass ? 1 : P);
        }
    }
    int tcp[BS];
#pragma unroll
    for (int i = 0; i < BS; ++i) tcp[i] = so_d2i_rint(ws[r * P + i]);
    __syncwarp();

    const bool eligible = VBS && a.vbs && bx != 0 && by != 0;
    const int qrow = a.qp_blocks ? a.qp_blocks[blk] : (a.qp_rows ? a.qp_rows[by] : a.qp_final);
    const int ky = r >> 3, sr = r & 7;        // sub-block row group of this lane
    int split = 0;
    int tcs[BS], predq5[BS];
#pragma unroll
    for (int i = 0; i < BS; ++i) { tcs[i] = 0; predq5[i] = 0; }
    const bool any_elig = VBS ? __any_sync(FULL, eligible) : false;
    if (any_elig) {
        // sub-block residuals: columns 0..7 belong to sub-block (ky, 0), columns 8..15 to (ky, 1)
        if (eligible) {
#pragma unroll
            for (int kx = 0; kx < 2; ++kx) {
                const MeResult ms = ky ? (kx ? msub[3] : msub[2]) : (kx ? msub[1] : msub[0]);
                const int xs = x + kx * S, ys = y + ky * S;
                const PredSel sels = pred_select(g, xs * mult, ys * mult, ms.dx, ms.dy, S, -1);
                const PredSel selq = pred_select(g, xs * mult, ys * mult, ms.dx, ms.dy, S, BS);      // quirk Q5
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    ws[r * P + kx * S + i] = so_i2d(c[kx * S + i] - pred_sample(g, rl, sels, ms.ref, i, sr));
                    predq5[kx * S + i] = pred_sample(g, rl, selq, ms.ref, i, sr);
                }
            }
        }
        __syncwarp();
        if (eligible) { dct1d<S>(ws + r, P); dct1d<S>(ws + S * P + r, P); }          // column r of the top and bottom sub-blocks
        __syncwarp();
        if (eligible) { dct1d<S>(ws + r * P, 1); dct1d<S>(ws + r * P + S, 1); }      // row r of the left and right sub-blocks
        if (eligible) {
#pragma unroll
            for (int i = 0; i < BS; ++i) tcs[i] = so_d2i_rint(ws[r * P + i]);
        }
        __syncwarp();
        // RD costs at the prediction-time QP (calculate_RD_cost, Encoder.py:1133-1158)
        uint32_t mP = 0, mL = 0, mR = 0;
        {
            const int qs = a.qp_rd > 0 ? a.qp_rd - 1 : a.qp_rd;
#pragma unroll
            for (int i = 0; i < BS; ++i) {
                if (quant_rhe(tcp[i], q_shift(r, i, BS, a.qp_rd)) != 0) mP |= 1u << i;
                if (quant_rhe(tcs[i], q_shift(sr, i & 7, S, qs)) != 0) { if (i < S) mL |= 1u << i; else mR |= 1u << (i - S); }
            }
        }
        // whole block: group = half-warp
        int lenp, lens;
        {
            const uint32_t prev = __shfl_up_sync(FULL, mP, 1);
            const uint32_t c0 = (__ballot_sync(FULL, mP & 1u) >> (16 * h)) & 0xFFFFu;
            const uint32_t cl = (__ballot_sync(FULL, (mP >> 15) & 1u) >> (16 * h)) & 0xFFFFu;
            const uint32_t last = __shfl_sync(FULL, mP, 16 * h + 15);
            int v = rle_rows_partial<BS>(mP, r, prev, c0, cl, last);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
            lenp = v;
        }
        {   // four sub-blocks: groups of 8 lanes, two masks per lane
            const int q8 = lane >> 3;
            const uint32_t prevL = __shfl_up_sync(FULL, mL, 1), prevR = __shfl_up_sync(FULL, mR, 1);
            const uint32_t c0L = (__ballot_sync(FULL, mL & 1u) >> (8 * q8)) & 0xFFu, clL = (__ballot_sync(FULL, (mL >> 7) & 1u) >> (8 * q8)) & 0xFFu;
            const uint32_t c0R = (__ballot_sync(FULL, mR & 1u) >> (8 * q8)) & 0xFFu, clR = (__ballot_sync(FULL, (mR >> 7) & 1u) >> (8 * q8)) & 0xFFu;
            const uint32_t lastL = __shfl_sync(FULL, mL, 8 * q8 + 7), lastR = __shfl_sync(FULL, mR, 8 * q8 + 7);
            int v = rle_rows_partial<S>(mL, sr, prevL, c0L, clL, lastL) + rle_rows_partial<S>(mR, sr, prevR, c0R, clR, lastR);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
            lens = v;
        }
        if (eligible) {
            double vm = 0.0;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) vm = __dadd_rn(vm, me_mae(msub[kk], S, a.fast));
            vm = vm / 4.0;
            const double rd_bs = __dadd_rn(__dmul_rn(a.lam, (double)(16 + 8 * lenp)), me_mae(mp, BS, a.fast));
            const double rd_vbs = __dadd_rn(__dmul_rn(a.lam, (double)(64 + 8 * lens)), vm);
            split = (rd_bs < rd_vbs) ? 0 : 1;
        }
    }

    // ---- final quantisation with the row QP, RLE size, dequantisation, inverse transform, reconstruction
    int level[BS];
    uint32_t m0 = 0, m1 = 0;                  // non-zero masks: whole row (no split) or left / right sub-block rows (split)
    {
        const int qs = qrow > 0 ? qrow - 1 : qrow;
#pragma unroll
        for (int i = 0; i < BS; ++i) {
            const int shift = split ? q_shift(sr, i & 7, S, qs) : q_shift(r, i, BS, qrow);
            level[i] = quant_rhe(split ? tcs[i] : tcp[i], shift);
            ws[r * P + i] = so_i2d(level[i] * (1 << shift));
            if (level[i] != 0) { if (!split) m0 |= 1u << i; else if (i < S) m0 |= 1u << i; else m1 |= 1u << (i - S); }
        }
    }
    int len;
    {
        const int q8 = lane >> 3;
        const uint32_t prev0 = __shfl_up_sync(FULL, m0, 1), prev1 = __shfl_up_sync(FULL, m1, 1);
        const unsigned b0lo = __ballot_sync(FULL, m0 & 1u), b0hi16 = __ballot_sync(FULL, (m0 >> 15) & 1u), b0hi8 = __ballot_sync(FULL, (m0 >> 7) & 1u);
        const unsigned b1lo = __ballot_sync(FULL, m1 & 1u), b1hi8 = __ballot_sync(FULL, (m1 >> 7) & 1u);
        const uint32_t last16 = __shfl_sync(FULL, m0, 16 * h + 15);
        const uint32_t last0 = __shfl_sync(FULL, m0, 8 * q8 + 7), last1 = __shfl_sync(FULL, m1, 8 * q8 + 7);
        int v;
        if (!split) v = rle_rows_partial<BS>(m0, r, prev0, (b0lo >> (16 * h)) & 0xFFFFu, (b0hi16 >> (16 * h)) & 0xFFFFu, last16);
        else v = rle_rows_partial<S>(m0, sr, prev0, (b0lo >> (8 * q8)) & 0xFFu, (b0hi8 >> (8 * q8)) & 0xFFu, last0) +
                 rle_rows_partial<S>(m1, sr, prev1, (b1lo >> (8 * q8)) & 0xFFu, (b1hi8 >> (8 * q8)) & 0xFFu, last1);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        len = v;
    }
    if (live) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = ((uint32_t)(uint16_t)(int16_t)level[2 * i]) | ((uint32_t)(uint16_t)(int16_t)level[2 * i + 1] << 16);
        uint4* lp = reinterpret_cast<uint4*>(a.levels + unit * a.frame_stride + (size_t)(y + r) * g.W + x);
        lp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        lp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    if constexpr (VBS) {
        __syncwarp();
        if (split) { idct1d<S>(ws + r, P); idct1d<S>(ws + S * P + r, P); } else idct1d<BS>(ws + r, P);
        __syncwarp();
        if (split) { idct1d<S>(ws + r * P, 1); idct1d<S>(ws + r * P + S, 1); } else idct1d<BS>(ws + r * P, 1);
    } else {
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            __syncwarp();
            idct1d<BS>(pass ? ws + r * P : ws + r, pass ? 1 : P);
        }
    }
    unsigned long long se = 0;
    {
        uint32_t pk[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < BS; ++i) {
            const int rec = ((split ? predq5[i] : predp[i]) + so_d2i_rint(ws[r * P + i])) & 0xFF;      // astype(np.uint8) wraps (A5)
            pk[i >> 2] |= (uint32_t)rec << (8 * (i & 3));
            const int d = rec - c[i];
            se += (unsigned long long)(d * d);
        }
        if (live) *reinterpret_cast<uint4*>(a.recon + unit * a.frame_stride + (size_t)(y + r) * g.W + x) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) se += __shfl_xor_sync(FULL, se, o);
    if (r == 0 && live) {
        a.split[unit * a.split_stride + blk] = (uint8_t)split;
        int16_t* mvo = a.mv + unit * a.mv_stride + (size_t)blk * 12;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const MeResult m = split ? msub[kk] : mp;
            const bool on = split || kk == 0;
            mvo[kk * 3 + 0] = on ? m.dx : 0; mvo[kk * 3 + 1] = on ? m.dy : 0; mvo[kk * 3 + 2] = on ? m.ref : 0;
        }
        if (a.blk_len) a.blk_len[unit * a.blk_len_stride + blk] = (uint32_t)len;
    }
    // ---- frame statistics: the eight blocks of the CTA are summed in shared memory and leave as ONE set of atomics (the
    //      counters of a frame live in one 128-byte line: 3 same-line atomics per block serialise in L2 -- they, not the
    //      arithmetic, bounded this kernel)
    __shared__ unsigned long long red_se[8], red_n[8];
    __shared__ unsigned int red_len[8];
    __shared__ int red_by[8];
    if (r == 0) {
        const int slot = warp * 2 + h;
        unsigned long long n = 0;
        bool inf = false;
        if (a.fast) {
            n = (unsigned long long)mp.sad * 4ull;
            if (eligible) { n = 0; for (int kk = 0; kk < 4; ++kk) n += msub[kk].sad; }
        } else {
            n = mp.sad;
            inf = mp.none;
            if (eligible) { n = 0; inf = false; for (int kk = 0; kk < 4; ++kk) { n += msub[kk].sad; inf = inf || msub[kk].none; } }
        }
        red_se[slot] = live ? se : 0ull;
        red_len[slot] = live ? (unsigned)len : 0u;
        red_n[slot] = (live && !inf) ? n : 0ull;
        red_by[slot] = live ? (inf ? (by | 0x40000000) : by) : -1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        so_frame_stats* st = a.stats + unit * a.stats_stride;
        if (blockIdx.x == 0) { st->mae_den = a.mae_den; st->frame_type = a.frame_type; }
        unsigned long long sse = 0, num = 0;
        unsigned int tl = 0, rowl = 0;
        bool inf = false;
        int row = -1;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int b = red_by[k];
            if (b < 0) continue;
            sse += red_se[k]; num += red_n[k]; tl += red_len[k];
            inf = inf || (b & 0x40000000);
            const int rb = b & 0x3FFFFFFF;
            if (rb != row) {
                if (row >= 0) atomicAdd(a.row_sizes + unit * a.rows_stride + row, rowl);
                row = rb; rowl = 0;
            }
            rowl += red_len[k];
        }
        if (row >= 0) atomicAdd(a.row_sizes + unit * a.rows_stride + row, rowl);
        atomicAdd(reinterpret_cast<unsigned long long*>(&st->sse), sse);
        atomicAdd(&st->qsize, tl);
        if (inf) atomicOr(&st->mae_inf, 1u);
        if (num) atomicAdd(reinterpret_cast<unsigned long long*>(&st->mae_num), num);
    }
}

// ------------------------------------------------------------------------------------------------------------
// fast motion estimation (fast_motion_estimation, Encoder.py:719-742; chained by inter_prediction :581)
// grid.x = 1 (chain over all blocks) or nblk (ParallelMode 2: mvp = (0,0,0) for every block), grid.y = units
// The four sub-blocks search the same nine offsets around the same predictor, so their SADs are the quadrant
// partial sums of the parent's candidates; only the validity tests differ per (sub-)block.
// ------------------------------------------------------------------------------------------------------------
template <int BS>
__global__ void fast_me_kernel(const FlowArgs a) {
    constexpr int S = BS / 2;
    __shared__ unsigned int sadq[SO_MAX_REF * 9][4];
    __shared__ int s_mvp[3];
    const FrameGeom& g = a.g;
    const int unit = a.unit0 + blockIdx.y;
    const int t = threadIdx.x;
    const bool active = t < BS * BS;
    const int i = active ? t % BS : 0, j = active ? t / BS : 0;
    const int mult = g.fme ? 2 : 1;
    const int nref = min(a.nref_fast, g.nref);
    const int ncand = nref * 9;
    const int nblk = g.nbx * g.nby;
    const int b0 = a.chain ? 0 : blockIdx.x, b1 = a.chain ? nblk : blockIdx.x + 1;
    const uint8_t* planes[SO_MAX_REF][4];
    for (int r = 0; r < nref; ++r)
        for (int ph = 0; ph < 4; ++ph) planes[r][ph] = a.ring.plane(unit, r, ph);
    if (t < 3) s_mvp[t] = 0;
    const int Wr = g.fme ? 2 * g.W - 1 : g.W, Hr = g.fme ? 2 * g.H - 1 : g.H;
    const int q = (j >= S ? 2 : 0) + (i >= S ? 1 : 0);

    for (int blk = b0; blk < b1; ++blk) {
        const int bx = blk % g.nbx, by = blk / g.nbx;
        const int x = bx * BS, y = by * BS;
        for (int e = t; e < ncand * 4; e += blockDim.x) sadq[e >> 2][e & 3] = 0;
        __syncthreads();
        const int mvx = s_mvp[0], mvy = s_mvp[1], mvr = s_mvp[2];
        const int c = a.cur[unit * a.cur_unit_stride + (size_t)(y + j) * g.W + x + i];
        for (int cand = 0; cand < ncand; ++cand) {
            const int ref = cand / 9, dx = mvx - 1 + (cand % 9) / 3, dy = mvy - 1 + (cand % 3);
            // sample position of pixel (i,j) for this candidate; out-of-frame samples belong to invalid candidates only
            const int X = x * mult + dx + mult * i, Y = y * mult + dy + mult * j;
            int d = 0;
            if (active && X >= 0 && Y >= 0 && X < Wr && Y < Hr) {
                const int v = g.fme ? planes[ref][((Y & 1) << 1) | (X & 1)][(size_t)(Y >> 1) * g.pitch + (X >> 1)]
                                    : planes[ref][0][(size_t)Y * g.pitch + X];
                d = abs(c - v);
            }
            // quadrant sums: reduce within the warp per quadrant, then one shared atomic per warp and quadrant
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const unsigned s = __reduce_add_sync(0xFFFFFFFFu, (unsigned)((active && q == qq) ? d : 0));
                if ((t & 31) == 0 && s) atomicAdd(&sadq[cand][qq], s);
            }
        }
        __syncthreads();
        // entity e: 0 = whole block, 1..4 = sub-blocks; scan candidates in (ref, dx, dy) order with strict <
        const bool eligible = a.vbs && bx != 0 && by != 0;
        if (t < 5 && (t == 0 || eligible)) {
            const int e = t;
            const int n = e == 0 ? BS : S;
            const int ex = (e == 0 ? x : x + ((e - 1) & 1) * S) * mult, ey = (e == 0 ? y : y + ((e - 1) >> 1) * S) * mult;
            unsigned best = 0xFFFFFFFFu;
            int bdx = mvx, bdy = mvy, bref = mvr, best_ref_idx = 0;
            for (int cand = 0; cand < ncand; ++cand) {
                const int ref = cand / 9, dx = mvx - 1 + (cand % 9) / 3, dy = mvy - 1 + (cand % 3);
                const int px = ex + dx, py = ey + dy;
                const bool ok = px >= 0 && px < Wr - n && py >= 0 && py < Hr - n &&
                                px + 2 * n >= 0 && px + 2 * n < Wr - n && py + 2 * n >= 0 && py + 2 * n < Hr - n;
                if (!ok) continue;
                const unsigned s = e == 0 ? sadq[cand][0] + sadq[cand][1] + sadq[cand][2] + sadq[cand][3] : sadq[cand][e - 1];
                if (s < best) { best = s; bdx = dx; bdy = dy; bref = ref; best_ref_idx = ref; }
            }
            MeResult r;
            r.dx = (int16_t)bdx; r.dy = (int16_t)bdy; r.ref = (int16_t)bref; r.none = 0; r.sad = (uint32_t)best_ref_idx;
            if (e == 0) a.me_parent[unit * a.me_parent_stride + blk] = r;
            else {
                const int kk = e - 1;
                a.me_sub[unit * a.me_sub_stride + (by * 2 + (kk >> 1)) * (g.nbx * 2) + bx * 2 + (kk & 1)] = r;
            }
            if (e == 0 && a.chain) { s_mvp[0] = bdx; s_mvp[1] = bdy; s_mvp[2] = bref; }
        }
        __syncthreads();
    }
}

// 16x16 variant of the fast search.  The chain is serial (block b needs the vector of block b - 1), so what counts is
// the latency of one step.  Thread = (candidate, block row): 36 candidates x 16 rows = 576 threads issue all loads of a
// step at once -- a candidate row is 16 contiguous bytes of ONE phase plane (appendix A1), read as four aligned words
// from the copy shifted by (column & 3) bytes -- then 4 VABSDIFF4 per thread, three shuffles to the quadrant sums, and
// one warp per entity (whole block + four sub-blocks) picks the winner with a REDUX on (SAD, scan index): strict '<' in
// scan order (ref, dx, dy) (Encoder.py:726-740) == lexicographic minimum.  The current block of the next step is
// fetched during this one.
// blocks [b0, b1) of one unit by the 576 threads of a CTA; chain: the predictor is carried from block to block starting at
// mv0 (otherwise it is a.mvp_in[blk] or zero); state_out (optional) receives the predictor every block used, mv_out
// (optional, shared memory) the whole-block vector of the last block
template <int BS>
__device__ __forceinline__ void fast_me16_run(const FlowArgs& a, int unit, int b0, int b1, bool chain, int mvx0, int mvy0, int mvr0,
                                              short4* state_out, int* mv_out) {
    static_assert(BS == 16 || BS == 8, "word-packed rows: 16x16 or 8x8 blocks");
    constexpr int S = BS / 2, WPR = BS / 4, CPP = 576 / BS;           // words per row, candidates per pass
    __shared__ unsigned int sadq[SO_MAX_REF * 9][4];
    __shared__ __align__(16) uint32_t s_cur[BS][WPR];
    __shared__ int s_mvp[3];
    const FrameGeom& g = a.g;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int cl = t / BS, row = t % BS;
    const int mult = g.fme ? 2 : 1;
    const int nref = min(a.nref_fast, g.nref);
    const int ncand = nref * 9;
    const int Wr = g.fme ? 2 * g.W - 1 : g.W, Hr = g.fme ? 2 * g.H - 1 : g.H;
    const uint8_t* cur = a.cur + unit * a.cur_unit_stride;
    const size_t shift_stride = a.ring.plane_stride >> 2;
    if (t == 0) { s_mvp[0] = mvx0; s_mvp[1] = mvy0; s_mvp[2] = mvr0; }
    if (t < BS * WPR) s_cur[t / WPR][t % WPR] = *reinterpret_cast<const uint32_t*>(cur + (size_t)((b0 / g.nbx) * BS + t / WPR) * g.W + (b0 % g.nbx) * BS + (t % WPR) * 4);
    __syncthreads();
    for (int blk = b0; blk < b1; ++blk) {
        const int bx = blk % g.nbx, by = blk / g.nbx;
        const int x = bx * BS, y = by * BS;
        int mvx = s_mvp[0], mvy = s_mvp[1], mvr = s_mvp[2];
        if (!chain && a.mvp_in) { const short4 m = a.mvp_in[unit * a.mvp_in_stride + blk]; mvx = m.x; mvy = m.y; mvr = m.z; }
        if (state_out && t == 0) state_out[blk] = make_short4((short)mvx, (short)mvy, (short)mvr, 0);
        uint32_t cnext = 0;
        if (t < BS * WPR && blk + 1 < b1) {
            const int nb = blk + 1;
            cnext = *reinterpret_cast<const uint32_t*>(cur + (size_t)((nb / g.nbx) * BS + t / WPR) * g.W + (nb % g.nbx) * BS + (t % WPR) * 4);
        }
        uint32_t cwv[WPR];
#pragma unroll
        for (int w = 0; w < WPR; ++w) cwv[w] = s_cur[row][w];
        for (int c0 = 0; c0 < ncand; c0 += CPP) {
            const int cand = c0 + cl;
            uint32_t sl = 0, sr = 0;
            if (cand < ncand) {
                const int ref = cand / 9, dx = mvx - 1 + (cand % 9) / 3, dy = mvy - 1 + (cand % 3);
                const int Xh = x * mult + dx, Yh = y * mult + dy;
                const int ph = g.fme ? (((Yh & 1) << 1) | (Xh & 1)) : 0;
                const int col0 = g.fme ? (Xh >> 1) : Xh, Y = (g.fme ? (Yh >> 1) : Yh) + row;
                uint32_t pw[WPR];
#pragma unroll
                for (int w = 0; w < WPR; ++w) pw[w] = 0u;
                if (Y >= 0 && Y < g.H) {
                    const uint8_t* pl = a.ring.plane(unit, ref, ph);
                    const int cs = col0 & 3, colA = col0 - cs;
                    const uint8_t* rowp = pl + (size_t)Y * g.pitch;
#pragma unroll
                    for (int w = 0; w < WPR; ++w) {
                        const int cA = colA + 4 * w;
                        if (cA >= 0 && col0 + 4 * w + 3 < g.W) {
                            pw[w] = __ldg(reinterpret_cast<const uint32_t*>(rowp + (size_t)cs * shift_stride + cA));
                        } else {                        // straddles the frame edge: only invalid (sub-)candidates see these bytes
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                const int col = col0 + 4 * w + b;
                                if (col >= 0 && col < g.W) pw[w] |= (uint32_t)rowp[col] << (8 * b);
                            }
                        }
                    }
                }
#pragma unroll
                for (int w = 0; w < WPR; ++w) {
                    if (w < WPR / 2) sl = sad4_acc(cwv[w], pw[w], sl); else sr = sad4_acc(cwv[w], pw[w], sr);
                }
            }
#pragma unroll
            for (int o = 1; o < S; o <<= 1) {
                sl += __shfl_xor_sync(0xFFFFFFFFu, sl, o);
                sr += __shfl_xor_sync(0xFFFFFFFFu, sr, o);
            }
            if (cand < ncand && (row % S) == 0) {
                sadq[cand][(row / S) * 2] = sl;
                sadq[cand][(row / S) * 2 + 1] = sr;
            }
        }
        __syncthreads();
        if (t < BS * WPR && blk + 1 < b1) s_cur[t / WPR][t % WPR] = cnext;
        const bool eligible = a.vbs && bx != 0 && by != 0;
        if (warp < 5 && (warp == 0 || eligible)) {
            const int e = warp;
            const int n = e == 0 ? BS : S;
            const int ex = (e == 0 ? x : x + ((e - 1) & 1) * S) * mult, ey = (e == 0 ? y : y + ((e - 1) >> 1) * S) * mult;
            uint32_t best = 0xFFFFFFFFu;
            for (int cand = lane; cand < ncand; cand += 32) {
                const int dx = mvx - 1 + (cand % 9) / 3, dy = mvy - 1 + (cand % 3);
                const int px = ex + dx, py = ey + dy;
                const bool ok = px >= 0 && px < Wr - n && py >= 0 && py < Hr - n &&
                                px + 2 * n >= 0 && px + 2 * n < Wr - n && py + 2 * n >= 0 && py + 2 * n < Hr - n;
                if (ok) {
                    const unsigned sv = e == 0 ? sadq[cand][0] + sadq[cand][1] + sadq[cand][2] + sadq[cand][3] : sadq[cand][e - 1];
                    best = min(best, (sv << 7) | (uint32_t)cand);
                }
            }
            best = __reduce_min_sync(0xFFFFFFFFu, best);
            if (lane == 0) {
                int bdx = mvx, bdy = mvy, bref = mvr, best_ref_idx = 0;        // no valid candidate: best_mv = mvp (Encoder.py:722)
                if (best != 0xFFFFFFFFu) {
                    const int cand = (int)(best & 127u);
                    bref = cand / 9; bdx = mvx - 1 + (cand % 9) / 3; bdy = mvy - 1 + (cand % 3); best_ref_idx = bref;
                }
                MeResult r;
                r.dx = (int16_t)bdx; r.dy = (int16_t)bdy; r.ref = (int16_t)bref; r.none = 0; r.sad = (uint32_t)best_ref_idx;   // quirk Q4
                if (e == 0) a.me_parent[unit * a.me_parent_stride + blk] = r;
                else {
                    const int kk = e - 1;
                    a.me_sub[unit * a.me_sub_stride + (by * 2 + (kk >> 1)) * (g.nbx * 2) + bx * 2 + (kk & 1)] = r;
                }
                if (e == 0 && chain) { s_mvp[0] = bdx; s_mvp[1] = bdy; s_mvp[2] = bref; }
                if (e == 0 && mv_out) { mv_out[0] = bdx; mv_out[1] = bdy; mv_out[2] = bref; }
            }
        }
        __syncthreads();
    }
}

template <int BS>
__global__ void __launch_bounds__(576) fast_me16_kernel(const FlowArgs a) {
    const int nblk = a.g.nbx * a.g.nby;
    fast_me16_run<BS>(a, a.unit0 + blockIdx.y, a.chain ? 0 : (int)blockIdx.x, a.chain ? nblk : (int)blockIdx.x + 1, a.chain != 0, 0, 0, 0, nullptr, nullptr);
}

// ------------------------------------------------------------------------------------------------------------
// Table-driven fast-ME chain for 16x16 blocks.  The chain mvp(b+1) = argmin around mvp(b) is serial over all blocks of a
// frame (Encoder.py:581), but only the WHOLE-block decision feeds it, each step moves the predictor by at most one unit
// per axis, and motion repeats from frame to frame.  So:
//   1. fast_table16_kernel (parallel over blocks): whole-block SADs of every reference at the (2K+3)^2 offsets around the
//      predictor the same block used in the previous P frame of the stream (`state`, zero at start), kept in shared
//      memory, and from them the block's TRANSITION TABLE: for each of the (2K+1)^2 predictors within K of that centre,
//      which candidate fast_motion_estimation picks (one byte);
//   2. fast_chain16_kernel (one warp walks the chain): next predictor = T_b[predictor - centre_b] -- one dependent
//      shared-memory byte load and a few integer instructions per block (tables staged ahead with cp.async).  A block
//      whose predictor lies outside its window (cold start, scene change) is decided by the whole CTA with the cooperative
//      step of fast_me16_run.  The predictor of every block is recorded in `state`;
//   3. fast_me16_kernel with mvp_in = state (parallel over blocks): whole-block and sub-block results exactly as the
//      chained kernel would produce them.
// Results do not depend on the table centres; only the speed does.
// ------------------------------------------------------------------------------------------------------------
constexpr int FT_K = 4;
constexpr int FT_N = 2 * FT_K + 3;          // table offsets per axis: centre +- (K + 1)

// One CTA per block.  The pixels any of the 11 x 11 offsets can touch form, per (reference, phase plane), a region of
// 26 rows x 32 bytes: it is copied to shared memory once (aligned words, zero outside the frame), then each thread takes
// candidates and reads its 16 rows as five words + funnel shifts -- no scattered global loads.
constexpr int FTR_H = 16 + FT_N - 1, FTR_W = 32;              // region rows (integer search: 16 + 10; half-pel needs 16 + 5) / bytes per row

constexpr int FT_S = 2 * FT_K + 1;          // predictor states per axis served by a table
constexpr int FT_TRANS = 96;                // bytes of transition table per block (81 used, 16-byte granules)

template <int BS>
__global__ void __launch_bounds__(128) fast_table16_kernel(const FlowArgs a, uint8_t* trans, size_t trans_unit_stride, const short4* state,
                                                           size_t state_unit_stride) {
    constexpr int WPR = BS / 4;
    extern __shared__ __align__(16) unsigned char ft_smem[];   // [nref * nph][FTR_H][FTR_W] regions, then the current block
    const FrameGeom& g = a.g;
    const int blk = blockIdx.x, unit = a.unit0 + blockIdx.y;
    const int bx = blk % g.nbx, by = blk / g.nbx;
    const int x = bx * BS, y = by * BS;
    const int nref = min(a.nref_fast, g.nref), nph = g.fme ? 4 : 1;
    const short4 c = state[unit * state_unit_stride + blk];
    const int mult = g.fme ? 2 : 1;
    const int Wr = g.fme ? 2 * g.W - 1 : g.W, Hr = g.fme ? 2 * g.H - 1 : g.H;
    // first offset of the window in search units and the region origin in plane pixels (floor), columns aligned to 4
    const int dx0 = c.x - (FT_K + 1), dy0 = c.y - (FT_K + 1);
    const int X0 = g.fme ? ((x * 2 + dx0) >> 1) : x + dx0, Y0 = g.fme ? ((y * 2 + dy0) >> 1) : y + dy0;
    const int XA = X0 & ~3;
    uint32_t* s_cur = reinterpret_cast<uint32_t*>(ft_smem + (size_t)nref * nph * FTR_H * FTR_W);
    {   // staging.  Thread = (word of the row, row): no divisions in the loop, and the two loads a thread makes per plane
        // (rows r and r + 16) are independent -- the kernel is bound by the latency of these L2 reads, not by their volume
        const int w = threadIdx.x & 7, r0 = threadIdx.x >> 3;                 // 128 threads: 8 words x 16 rows
        const int X = XA + 4 * w;
        const bool xin = X >= 0 && X + 3 < g.W;
        for (int ref = 0; ref < nref; ++ref) {
#pragma unroll 4
            for (int ph = 0; ph < nph; ++ph) {
                const uint8_t* plane = a.ring.plane(unit, ref, ph);
                uint32_t* dst = reinterpret_cast<uint32_t*>(ft_smem) + (size_t)(ref * nph + ph) * FTR_H * (FTR_W / 4);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int row = r0 + 16 * k, Y = Y0 + row;
                    if (row < FTR_H) {
                        uint32_t v = 0u;
                        if (Y >= 0 && Y < g.H) {
                            const uint8_t* pl = plane + (size_t)Y * g.pitch;
                            if (xin) v = __ldg(reinterpret_cast<const uint32_t*>(pl + X));
                            else {
#pragma unroll
                                for (int b = 0; b < 4; ++b) if (X + b >= 0 && X + b < g.W) v |= (uint32_t)pl[X + b] << (8 * b);
                            }
                        }
                        dst[row * (FTR_W / 4) + w] = v;
                    }
                }
            }
        }
    }
    if (threadIdx.x < BS * WPR)
        s_cur[threadIdx.x] = *reinterpret_cast<const uint32_t*>(a.cur + unit * a.cur_unit_stride + (size_t)(y + threadIdx.x / WPR) * g.W + x + (threadIdx.x % WPR) * 4);
    __syncthreads();
    uint16_t* s_sad = reinterpret_cast<uint16_t*>(s_cur + BS * WPR);         // [nref][FT_N][FT_N]
    // Work item = (reference, horizontal offset, vertical set): the up-to-six vertical offsets of a set lie in ONE phase plane
    // one row apart (half-pel: the even / the odd offsets; integer: the first six / the last five), so a thread walks the
    // region rows once -- five aligned words per row, shifted into place -- and feeds every row to all the offsets it belongs
    // to, with the current block held in registers: 1/7 of the shared-memory loads of one-candidate-per-thread.  The
    // horizontal offset is the fastest index: the lanes of a warp read the same few rows (broadcast, disjoint banks).
    {
        constexpr int VT = 6;
        for (int it = threadIdx.x; it < nref * FT_N * 2; it += blockDim.x) {
            const int ref = it / (FT_N * 2), rem = it - ref * (FT_N * 2);
            const int vs = rem / FT_N, ix = rem - vs * FT_N;
            const int iy0 = g.fme ? vs : vs * VT, stepi = g.fme ? 2 : 1, n = vs ? FT_N - VT : VT;
            const int dx = dx0 + ix, px = x * mult + dx;
            const int py0 = y * mult + dy0 + iy0;
            const bool xok = px >= 0 && px <= Wr - 3 * BS - 1;       // Encoder.py:728-730 on the x axis
            const int ph = g.fme ? (((py0 & 1) << 1) | (px & 1)) : 0;
            const int co = (g.fme ? (px >> 1) : px) - XA, ro0 = (g.fme ? (py0 >> 1) : py0) - Y0;
            const unsigned char* rg = ft_smem + ((size_t)(ref * nph + ph) * FTR_H + ro0) * FTR_W + (co & ~3);
            const int sh = (co & 3) * 8;
            uint32_t acc[VT];
#pragma unroll
            for (int t = 0; t < VT; ++t) acc[t] = 0u;
            uint32_t cw[BS][WPR];                   // current rows are loaded when first needed: at most VT of them are live
#pragma unroll
            for (int r = 0; r < BS + VT - 1; ++r) {
                if (r < BS) {
#pragma unroll
                    for (int w = 0; w < WPR; ++w) cw[r][w] = s_cur[r * WPR + w];
                }
                if (r < BS + n - 1) {
                    const uint32_t* rw = reinterpret_cast<const uint32_t*>(rg + r * FTR_W);
                    uint32_t q[WPR + 1], wv[WPR];
#pragma unroll
                    for (int w = 0; w <= WPR; ++w) q[w] = rw[w];
#pragma unroll
                    for (int w = 0; w < WPR; ++w) wv[w] = __funnelshift_r(q[w], q[w + 1], sh);
#pragma unroll
                    for (int t = 0; t < VT; ++t) {
                        const int cr = r - t;                       // current-block row this region row meets at vertical offset t
                        if (cr >= 0 && cr < BS && t < n) {
#pragma unroll
                            for (int w = 0; w < WPR; ++w) acc[t] = sad4_acc(cw[cr][w], wv[w], acc[t]);
                        }
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < VT; ++t) {
                if (t < n) {
                    const int py = py0 + t * stepi;
                    const bool ok = xok && py >= 0 && py <= Hr - 3 * BS - 1;       // invalid offsets get 0xFFFF (> any SAD)
                    s_sad[ref * (FT_N * FT_N) + ix * FT_N + iy0 + t * stepi] = ok ? (uint16_t)acc[t] : (uint16_t)0xFFFFu;
                }
            }
        }
    }
    __syncthreads();
    // transition table: for every predictor within K of the centre, the winner of fast_motion_estimation's scan (ref, dx,
    // dy ascending, strict '<': the minimum of (SAD, scan index)) packed as ref << 4 | (dx - mvp.x + 1) << 2 | (dy - mvp.y + 1);
    // 0xFF = no valid candidate (the predictor is kept, Encoder.py:722)
    uint8_t* tr = trans + unit * trans_unit_stride + (size_t)blk * FT_TRANS;
    for (int sI = threadIdx.x; sI < FT_S * FT_S; sI += blockDim.x) {
        const int sx = sI / FT_S, sy = sI - sx * FT_S;                      // predictor = centre - K + (sx, sy)
        uint32_t best = 0xFFFFFFFFu;
        for (int ref = 0; ref < nref; ++ref)
#pragma unroll
            for (int o = 0; o < 9; ++o) {
                const uint32_t sad = s_sad[ref * (FT_N * FT_N) + (sx + o / 3) * FT_N + sy + o % 3];
                best = min(best, (sad << 7) | (uint32_t)((ref << 4) | ((o / 3) << 2) | (o % 3)));
            }
        tr[sI] = best < (0xFFFFu << 7) ? (uint8_t)(best & 127u) : (uint8_t)0xFF;
    }
}

template <int BS>
__device__ __forceinline__ void fast_me16_run(const FlowArgs& a, int unit, int b0, int b1, bool chain, int mvx0, int mvy0, int mvr0,
                                              short4* state_out, int* mv_out);

template <int BS>
__global__ void __launch_bounds__(576) fast_chain16_kernel(const FlowArgs a, const uint8_t* trans, size_t trans_unit_stride, short4* state,
                                                           size_t state_unit_stride) {
    // Warp 0 walks the chain.  With the transition tables a step is one dependent shared-memory byte load plus a few integer
    // instructions: next predictor = T_b[predictor - centre_b].  Tables and centres are staged in groups of four blocks,
    // FT_DG groups ahead, with cp.async.  When the predictor leaves the window of a block (cold start, scene change) the
    // block is decided by the whole CTA with the cooperative step of fast_me16_run (the other 17 warps sleep at the barrier
    // until then).
    constexpr int GB = 4, DG = 6, GSLOTS = DG + 1;                   // blocks per group, groups in flight
    constexpr int GBYTES = GB * FT_TRANS + GB * 8;                   // tables + state records of a group (416 B)
    __shared__ __align__(16) unsigned char ring[GSLOTS * GBYTES];
    __shared__ int s_req[4];                    // {block to decide cooperatively (nblk: done), predictor x, y, ref}
    __shared__ int s_mvout[3];                  // its result
    const FrameGeom& g = a.g;
    const int unit = a.unit0 + blockIdx.y;
    const int lane = threadIdx.x & 31;
    const bool walker = threadIdx.x < 32;
    const int nblk = g.nbx * g.nby, ngroups = (nblk + GB - 1) / GB;
    const char* tsrc = reinterpret_cast<const char*>(trans + unit * trans_unit_stride);      // buffers are padded to whole groups
    short4* st = state + unit * state_unit_stride;
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(ring);
    int sg = 0, sg_slot = 0;                     // next group to stage / its slot
    auto stage = [&]() {
        if (sg < ngroups) {
            const uint32_t dst = smem0 + sg_slot * GBYTES;
            if (lane < GB * FT_TRANS / 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + lane * 16), "l"(tsrc + (size_t)sg * (GB * FT_TRANS) + lane * 16) : "memory");
            else if (lane < GB * FT_TRANS / 16 + GB / 2)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + lane * 16),
                             "l"(reinterpret_cast<const char*>(st + (size_t)sg * GB) + (lane - GB * FT_TRANS / 16) * 16) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        ++sg;
        if (++sg_slot == GSLOTS) sg_slot = 0;
    };
    int mvx = 0, mvy = 0, mvr = 0, blk = 0, gslot = 0, craw = 0;
    bool pending = false;                        // block `blk` was handed to the cooperative step
    // centre (x | y << 16) of block b, whose group sits in slot gs
    auto centre = [&](int b, int gs) { return *reinterpret_cast<const int*>(ring + gs * GBYTES + GB * FT_TRANS + (b & (GB - 1)) * 8); };
    // block blk is done: move on; at a group boundary stage one more group and make sure the group after the next has landed
    auto next_block = [&]() {
        if ((++blk & (GB - 1)) == 0) {
            __syncwarp(); stage();
            asm volatile("cp.async.wait_group %0;" ::"n"(DG - 1) : "memory");
            __syncwarp();
            if (++gslot == GSLOTS) gslot = 0;
        }
    };
    if (walker) {
        for (int p = 0; p < DG; ++p) stage();
        stage();
        asm volatile("cp.async.wait_group %0;" ::"n"(DG - 1) : "memory");     // groups 0 and 1 have landed
        __syncwarp();
        craw = centre(0, 0);
    }
    while (true) {
        if (walker) {
            if (pending) {
                mvx = s_mvout[0]; mvy = s_mvout[1]; mvr = s_mvout[2]; pending = false;
                next_block();
                craw = centre(blk, gslot);
            }
            while (blk < nblk) {
                const int sx = mvx - (int)(short)(craw & 0xFFFF) + FT_K, sy = mvy - (craw >> 16) + FT_K;
                if (!((unsigned)sx < (unsigned)FT_S && (unsigned)sy < (unsigned)FT_S)) { pending = true; break; }
                const int t = ring[gslot * GBYTES + (blk & (GB - 1)) * FT_TRANS + sx * FT_S + sy];
                // while the table byte is on its way: the next block's centre (its group has landed already) and the record
                const int gnext = ((blk & (GB - 1)) == GB - 1) ? (gslot + 1 == GSLOTS ? 0 : gslot + 1) : gslot;
                const int craw_next = centre(blk + 1, gnext);
                if (lane == 0) st[blk] = make_short4((short)mvx, (short)mvy, (short)mvr, 0);      // the predictor this block used
                if (t != 0xFF) { mvr = t >> 4; mvx += ((t >> 2) & 3) - 1; mvy += (t & 3) - 1; }
                next_block();
                craw = craw_next;
            }
            if (lane == 0) { s_req[0] = pending ? blk : nblk; s_req[1] = mvx; s_req[2] = mvy; s_req[3] = mvr; }
        }
        __syncthreads();
        const int rb = s_req[0];
        if (rb >= nblk) break;
        fast_me16_run<BS>(a, unit, rb, rb + 1, true, s_req[1], s_req[2], s_req[3], st, s_mvout);     // records st[rb], ends with a barrier
    }
    if (walker) asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------------------
// The chain as a scan.  A block's transition table is a FUNCTION predictor -> predictor, and function composition is
// associative: the chain over all blocks of a frame (Encoder.py:581) does not have to be walked block by block.
//   A. fast_scan_chunk_kernel (parallel over chunks of L consecutive blocks): the composition F_c of the chunk's tables for
//      every predictor of the first block's window (81 inputs, one thread each, L dependent shared-memory lookups): where
//      the chain leaves the chunk, or "escaped" when it runs out of some block's window on the way;
//   B. fast_scan_walk_kernel (one CTA per unit): walks the <= 256 CHUNKS -- one lookup in F_c per chunk (all F_c sit in
//      shared memory) -- and records the predictor every chunk is entered with.  A chunk whose F_c has no answer for the
//      incoming predictor (cold start, scene change) is walked block by block right here, with the cooperative step of
//      fast_me16_run for blocks whose window the predictor has left, and marked as done;
//   C. fast_scan_fill_kernel (parallel over chunks): replays each chunk from its entry predictor and records the
//      predictor of every block in `state` -- what fast_me16_kernel (mvp_in = state) then turns into results.
// Serial depth: L + nchunks + L table lookups (1080p: 32 + 255 + 32) instead of 8160 dependent steps.
// F_c entry: .x = px | py << 16 (predictor after the chunk), .y = ref (0xFF: unchanged) | escaped << 8; entry 81 of a
// chunk holds the centre of its first block.
// ------------------------------------------------------------------------------------------------------------
constexpr int FS_STATES = FT_S * FT_S;          // 81
constexpr int FS_ROW = FS_STATES + 1;           // + the first centre

__device__ __forceinline__ int fs_pack(int x, int y) { return (x & 0xFFFF) | (y << 16); }
__device__ __forceinline__ int fs_x(int v) { return (int)(short)(v & 0xFFFF); }
__device__ __forceinline__ int fs_y(int v) { return v >> 16; }

// stage the tables and centres of blocks [b0, b0 + n) of one unit in shared memory: sm = [n][FT_TRANS] tables, then n centres
__device__ __forceinline__ int* fs_stage(unsigned char* sm, int L, const uint8_t* trans_u, const short4* state_u, int b0, int n) {
    const uint4* src = reinterpret_cast<const uint4*>(trans_u + (size_t)b0 * FT_TRANS);
    for (int e = threadIdx.x; e < n * (FT_TRANS / 16); e += blockDim.x) reinterpret_cast<uint4*>(sm)[e] = src[e];
    int* cen = reinterpret_cast<int*>(sm + (size_t)L * FT_TRANS);
    for (int e = threadIdx.x; e < n; e += blockDim.x) { const short4 c = state_u[b0 + e]; cen[e] = fs_pack(c.x, c.y); }
    return cen;
}

__global__ void __launch_bounds__(96) fast_scan_chunk_kernel(const uint8_t* trans, size_t trans_unit_stride, const short4* state,
                                                             size_t state_unit_stride, int unit0, int nblk, int L, int2* F, size_t F_unit_stride) {
    extern __shared__ __align__(16) unsigned char fs_smem[];
    const int chunk = blockIdx.x, unit = unit0 + blockIdx.y;
    const int b0 = chunk * L, n = min(L, nblk - b0);
    const int* cen = fs_stage(fs_smem, L, trans + unit * trans_unit_stride, state + unit * state_unit_stride, b0, n);
    __syncthreads();
    int2* out = F + unit * F_unit_stride + (size_t)chunk * FS_ROW;
    const int t = threadIdx.x;
    if (t == FS_STATES) out[t] = make_int2(cen[0], 0);
    if (t >= FS_STATES) return;
    int px = fs_x(cen[0]) - FT_K + t / FT_S, py = fs_y(cen[0]) - FT_K + t % FT_S, ref = 0xFF, esc = 0;
    for (int i = 0; i < n; ++i) {
        const int c = cen[i];
        const int rx = px - fs_x(c) + FT_K, ry = py - fs_y(c) + FT_K;
        if ((unsigned)rx >= (unsigned)FT_S || (unsigned)ry >= (unsigned)FT_S) { esc = 1; break; }
        const int tr = fs_smem[i * FT_TRANS + rx * FT_S + ry];
        if (tr != 0xFF) { ref = tr >> 4; px += ((tr >> 2) & 3) - 1; py += (tr & 3) - 1; }
    }
    out[t] = make_int2(fs_pack(px, py), ref | (esc << 8));
}

// whole-block result of fast_motion_estimation for a block entered with predictor (px, py, ref) whose table says `tr`
// (Encoder.py:722-742; the second return value is the reference index, quirk Q4)
__device__ __forceinline__ MeResult fs_result(int px, int py, int ref, int tr) {
    MeResult r;
    r.none = 0;
    if (tr != 0xFF) { r.ref = (int16_t)(tr >> 4); r.dx = (int16_t)(px + ((tr >> 2) & 3) - 1); r.dy = (int16_t)(py + (tr & 3) - 1); r.sad = (uint32_t)(tr >> 4); }
    else { r.dx = (int16_t)px; r.dy = (int16_t)py; r.ref = (int16_t)ref; r.sad = 0u; }          // no valid candidate: best_mv = mvp, best_ref_idx = 0
    return r;
}

template <int BS>
__global__ void __launch_bounds__(576) fast_scan_walk_kernel(const FlowArgs a, const uint8_t* trans, size_t trans_unit_stride, short4* state,
                                                             size_t state_unit_stride, int L, int nchunks, const int2* F, size_t F_unit_stride,
                                                             int2* entry, size_t entry_unit_stride, int write_results) {
    extern __shared__ __align__(16) unsigned char fs_smem[];        // [nchunks][FS_ROW] int2, then the centres of one chunk
    __shared__ int s_mvout[3];
    __shared__ int s_req[4];                                        // {chunk that needs the whole CTA (nchunks: done), predictor x, y, ref}
    const FrameGeom& g = a.g;
    const int unit = a.unit0 + blockIdx.y;
    const int nblk = g.nbx * g.nby;
    int2* sF = reinterpret_cast<int2*>(fs_smem);
    int* cen = reinterpret_cast<int*>(sF + (size_t)nchunks * FS_ROW);
    {
        const int4* src = reinterpret_cast<const int4*>(F + unit * F_unit_stride);          // FS_ROW is even: whole int4s
        const int n4 = nchunks * FS_ROW / 2;
        for (int e0 = threadIdx.x; e0 < n4; e0 += blockDim.x * 4) {                         // four loads in flight per thread
            int4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int e = e0 + u * blockDim.x; if (e < n4) v[u] = __ldg(src + e); }
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int e = e0 + u * blockDim.x; if (e < n4) reinterpret_cast<int4*>(sF)[e] = v[u]; }
        }
    }
    __syncthreads();
    const uint8_t* tr_u = trans + unit * trans_unit_stride;
    short4* st = state + unit * state_unit_stride;
    int2* en = entry + unit * entry_unit_stride;
    MeResult* res = a.me_parent + unit * a.me_parent_stride;
    int px = 0, py = 0, ref = 0, c = 0;
    while (true) {
        if (threadIdx.x < 32) {
            // warp 0 walks the chunks: one dependent lookup in the chunk's composed table per step (the centre of the next
            // chunk does not depend on the walk and is fetched alongside)
            int c0 = c < nchunks ? sF[(size_t)c * FS_ROW + FS_STATES].x : 0;
            while (c < nchunks) {
                const int2* row = sF + (size_t)c * FS_ROW;
                const int c0n = c + 1 < nchunks ? row[FS_ROW + FS_STATES].x : 0;
                const int sx = px - fs_x(c0) + FT_K, sy = py - fs_y(c0) + FT_K;
                if (!((unsigned)sx < (unsigned)FT_S && (unsigned)sy < (unsigned)FT_S)) break;
                const int2 f = row[sx * FT_S + sy];
                if ((f.y >> 8) & 1) break;
                if (threadIdx.x == 0) en[c] = make_int2(fs_pack(px, py), ref);
                px = fs_x(f.x); py = fs_y(f.x);
                if ((f.y & 0xFF) != 0xFF) ref = f.y & 0xFF;
                ++c;
                c0 = c0n;
            }
            if (threadIdx.x == 0) { s_req[0] = c; s_req[1] = px; s_req[2] = py; s_req[3] = ref; }
        }
        __syncthreads();
        c = s_req[0]; px = s_req[1]; py = s_req[2]; ref = s_req[3];
        if (c >= nchunks) break;
        // this chunk has no answer for the incoming predictor (cold start, scene change): block by block, all threads in step
        // (its centres are copied first -- `state` is overwritten with the predictors as we go)
        if (threadIdx.x == 0) en[c] = make_int2(0, -1);
        const int b0 = c * L, n = min(L, nblk - b0);
        for (int e = threadIdx.x; e < n; e += blockDim.x) { const short4 cc = st[b0 + e]; cen[e] = fs_pack(cc.x, cc.y); }
        __syncthreads();
        for (int i = 0; i < n; ++i) {
            const int cb = cen[i];
            const int rx = px - fs_x(cb) + FT_K, ry = py - fs_y(cb) + FT_K;
            if ((unsigned)rx < (unsigned)FT_S && (unsigned)ry < (unsigned)FT_S) {
                const int tr = __ldg(tr_u + (size_t)(b0 + i) * FT_TRANS + rx * FT_S + ry);
                if (threadIdx.x == 0) {
                    st[b0 + i] = make_short4((short)px, (short)py, (short)ref, 0);
                    if (write_results) res[b0 + i] = fs_result(px, py, ref, tr);
                }
                if (tr != 0xFF) { ref = tr >> 4; px += ((tr >> 2) & 3) - 1; py += (tr & 3) - 1; }
            } else {
                fast_me16_run<BS>(a, unit, b0 + i, b0 + i + 1, true, px, py, ref, st, s_mvout);     // records st[] and the result, ends with a barrier
                px = s_mvout[0]; py = s_mvout[1]; ref = s_mvout[2];
                __syncthreads();                                                                    // s_mvout is rewritten by the next step
            }
        }
        ++c;
        __syncthreads();                                                                            // s_req is rewritten by warp 0
    }
}

// res != nullptr: the whole-block results are written here as well (no VBS: nothing else is needed from a search kernel)
__global__ void __launch_bounds__(32) fast_scan_fill_kernel(const uint8_t* trans, size_t trans_unit_stride, short4* state, size_t state_unit_stride,
                                                            int unit0, int nblk, int L, const int2* entry, size_t entry_unit_stride,
                                                            MeResult* res, size_t res_unit_stride) {
    extern __shared__ __align__(16) unsigned char fs_smem[];
    const int chunk = blockIdx.x, unit = unit0 + blockIdx.y;
    const int2 e = entry[unit * entry_unit_stride + chunk];
    if (e.y == -1) return;                                  // the walker went through this chunk block by block
    const int b0 = chunk * L, n = min(L, nblk - b0);
    short4* st = state + unit * state_unit_stride;
    const int* cen = fs_stage(fs_smem, L, trans + unit * trans_unit_stride, st, b0, n);
    __syncwarp();
    if (threadIdx.x != 0) return;
    int px = fs_x(e.x), py = fs_y(e.x), ref = e.y;
    for (int i = 0; i < n; ++i) {
        const int c = cen[i];
        const int rx = px - fs_x(c) + FT_K, ry = py - fs_y(c) + FT_K;         // inside the window: the walker checked this path
        st[b0 + i] = make_short4((short)px, (short)py, (short)ref, 0);
        const int tr = fs_smem[i * FT_TRANS + rx * FT_S + ry];
        if (res) res[unit * res_unit_stride + b0 + i] = fs_result(px, py, ref, tr);
        if (tr != 0xFF) { ref = tr >> 4; px += ((tr >> 2) & 3) - 1; py += (tr & 3) - 1; }
    }
}

// ------------------------------------------------------------------------------------------------------------
// intra search (intra_find_best_match_horizontal, Encoder.py:1010-1045; intra_prediction :1272-1338)
// The search frame holds ORIGINAL pixels left of the current parent block and 128 elsewhere, so all blocks are
// independent.  Sub-block SADs are the quadrant sums of the parent's candidates (same dx, same pixels).
// ------------------------------------------------------------------------------------------------------------
template <int BS>
__global__ void intra_search_kernel(const FlowArgs a) {
    constexpr int S = BS / 2;
    extern __shared__ unsigned int isad[];          // [(2r+1)][4]
    const FrameGeom& g = a.g;
    const int blk = blockIdx.x, unit = a.unit0 + blockIdx.y;
    const int bx = blk % g.nbx, by = blk / g.nbx;
    const int x = bx * BS, y = by * BS;
    const int ncand = 2 * g.r + 1;
    const uint8_t* cur = a.cur + unit * a.cur_unit_stride;
    for (int e = threadIdx.x; e < ncand * 4; e += blockDim.x) isad[e] = 0;
    __syncthreads();
    // task = (candidate, row): left and right half-row SADs against pixels at columns x+dx+i (128 at/after column x)
    for (int task = threadIdx.x; task < ncand * BS; task += blockDim.x) {
        const int cand = task / BS, j = task % BS;
        const int dx = bx == 0 ? 0 : cand - g.r;
        const uint8_t* row = cur + (size_t)(y + j) * g.W;
        unsigned sl = 0, sr = 0;
        for (int i = 0; i < BS; ++i) {
            const int col = x + dx + i;
            const int pv = (bx == 0 || col >= x || col < 0) ? 128 : row[col];
            const int d = abs((int)row[x + i] - pv);
            if (i < S) sl += d; else sr += d;
        }
        const int qrow = j >= S ? 2 : 0;
        atomicAdd(&isad[cand * 4 + qrow], sl);
        atomicAdd(&isad[cand * 4 + qrow + 1], sr);
    }
    __syncthreads();
    const bool eligible = a.vbs && bx != 0 && by != 0;
    if (threadIdx.x < 5 && (threadIdx.x == 0 || eligible)) {
        const int e = threadIdx.x;
        const int n = e == 0 ? BS : S;
        const int ex = e == 0 ? x : x + ((e - 1) & 1) * S;
        MeResult r;
        r.dy = 0; r.ref = 0; r.none = 0;
        if (ex == 0) {          // x == 0: mode -1, predictor 128 (only the whole block can be here: VBS needs x != 0)
            r.dx = -1;
            r.sad = isad[0] + isad[1] + isad[2] + isad[3];
        } else {
            unsigned best = 0xFFFFFFFFu; int bmv = 0; bool any = false;
            for (int cand = 0; cand < ncand; ++cand) {
                const int dx = cand - g.r;
                if (!(ex + dx >= 0 && ex + dx + n <= g.W)) continue;
                const unsigned s = e == 0 ? isad[cand * 4] + isad[cand * 4 + 1] + isad[cand * 4 + 2] + isad[cand * 4 + 3]
                                          : isad[cand * 4 + (e - 1)];
                if (!any || s < best) { best = s; bmv = dx; any = true; }
                else if (s == best && abs(dx) <= abs(bmv)) bmv = dx;
            }
            r.dx = (int16_t)bmv; r.sad = any ? best : 0; r.none = any ? 0 : 1;
        }
        if (e == 0) a.me_parent[unit * a.me_parent_stride + blk] = r;
        else {
            const int kk = e - 1;
            a.me_sub[unit * a.me_sub_stride + (by * 2 + (kk >> 1)) * (g.nbx * 2) + bx * 2 + (kk & 1)] = r;
        }
    }
}

// 16x16 blocks, one warp per block, lane = candidate.  The search frame holds original pixels LEFT of the parent block and
// 128 from the parent's first column on (Encoder.py:1248, :1329-1338), so every offset dx >= 0 predicts a flat 128 block:
// their SADs all equal SAD(dx = 0) and the replace rule (SAD, |dx|, -dx) (appendix A4) keeps dx = 0.  Only dx = -t,
// t = 0..r, has to be evaluated: predictor byte i = row[x - t + i] for i < t, else 128.  Each lane builds its 16 predictor
// bytes per row from aligned 32-bit loads + funnel shifts + a byte mask and accumulates the four quadrant SADs with
// VABSDIFF4; the five argmins (parent + four 8x8 sub-blocks = quadrant sums, same pixels) are REDUX reductions.
__global__ void __launch_bounds__(128) intra_search16_kernel(const FlowArgs a) {
    constexpr int BS = 16;
    const FrameGeom& g = a.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = blockIdx.x * 4 + warp, unit = a.unit0 + blockIdx.y;
    if (blk >= g.nbx * g.nby) return;
    const int bx = blk % g.nbx, by = blk / g.nbx;
    const int x = bx * BS, y = by * BS;
    const uint8_t* cur = a.cur + unit * a.cur_unit_stride + (size_t)y * g.W;
    const bool eligible = a.vbs && bx != 0 && by != 0;
    uint32_t bestk[5] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};     // SAD << 8 | t: parent, TL, TR, BL, BR
    const int tmax = bx == 0 ? 0 : g.r;
    for (int t0 = 0; t0 <= tmax; t0 += 32) {
        const int t = t0 + lane;
        const bool cand = t <= tmax;
        const int start = x - t;                                  // first predictor column
        // start may be negative by up to 8 columns: the candidate is then invalid for the parent and the left sub-blocks but
        // still valid for the right ones, whose eight predictor columns start at x + 8 - t >= 0 (words wholly left of the
        // frame are skipped; they only feed bytes of the invalid half)
        const bool readable = cand && t > 0;
        const int b0 = start & ~3, sh = (start & 3) * 8;
        uint32_t m[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int cnt = min(max(t - 4 * k, 0), 4);            // real pixels in word k
            m[k] = cnt == 4 ? 0xFFFFFFFFu : ((1u << (8 * cnt)) - 1u);
        }
        uint32_t q4[4] = {0u, 0u, 0u, 0u};                        // TL, TR, BL, BR
#pragma unroll 4
        for (int j = 0; j < BS; ++j) {
            const uint8_t* row = cur + (size_t)j * g.W;
            const uint4 cw = *reinterpret_cast<const uint4*>(row + x);       // same address for all lanes: one broadcast load
            uint32_t q[5];
#pragma unroll
            for (int k = 0; k < 5; ++k)
                q[k] = (readable && b0 + 4 * k >= 0 && b0 + 4 * k < x) ? __ldg(reinterpret_cast<const uint32_t*>(row + b0 + 4 * k)) : 0u;
            const uint32_t cwv[4] = {cw.x, cw.y, cw.z, cw.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t pw = (__funnelshift_r(q[k], q[k + 1], sh) & m[k]) | (0x80808080u & ~m[k]);
                const int qi = (j >= 8 ? 2 : 0) + (k >= 2 ? 1 : 0);
                q4[qi] = sad4_acc(cwv[k], pw, q4[qi]);
            }
        }
        // validity (Encoder.py:1026: x + dx >= 0 and x + dx + bs <= W; the second holds for dx <= 0)
        const uint32_t par = q4[0] + q4[1] + q4[2] + q4[3];
        if (cand && x - t >= 0) bestk[0] = min(bestk[0], (par << 8) | (uint32_t)t);
        if (eligible) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ex = x + (e & 1) * 8;
                if (cand && ex - t >= 0) bestk[1 + e] = min(bestk[1 + e], (q4[e] << 8) | (uint32_t)t);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 5; ++e) {
        if (e > 0 && !eligible) break;
        const uint32_t k = __reduce_min_sync(0xFFFFFFFFu, bestk[e]);
        if (lane == 0) {
            MeResult r;
            r.dy = 0; r.ref = 0; r.none = 0;
            r.sad = k >> 8;
            r.dx = (int16_t)(-(int)(k & 0xFFu));
            if (e == 0 && bx == 0) r.dx = -1;                     // x == 0: mode -1, predictor 128 (Encoder.py:1016-1019)
            if (e == 0) a.me_parent[unit * a.me_parent_stride + blk] = r;
            else a.me_sub[unit * a.me_sub_stride + (by * 2 + ((e - 1) >> 1)) * (g.nbx * 2) + bx * 2 + ((e - 1) & 1)] = r;
        }
    }
}

// intra predictor sample for a (sub-)block at column ex with offset mv, parent block starting at column xpar
__device__ __forceinline__ int intra_pred(const uint8_t* row, int ex, int mv, int i, int xpar, bool first_col) {
    if (first_col) return 128;
    const int col = ex + mv + i;
    return (col >= xpar || col < 0) ? 128 : row[col];
}

// intra finish: residual -> DCT -> [RD] -> quant -> levels, RLE size; dequant -> IDCT -> res_frame
// (intra_prediction :1313-1327, complete_intra_flow :1611-1628, reconstruct_frame_intra :1358-1376)
template <int BS>
__global__ void intra_finish_kernel(const FlowArgs a) {
    constexpr int S = BS / 2;
    constexpr int P = BS + 1;
    __shared__ double ws[BS * P];
    __shared__ int nzbuf[BS * BS];
    const FrameGeom& g = a.g;
    const int blk = blockIdx.x, unit = a.unit0 + blockIdx.y;
    const int bx = blk % g.nbx, by = blk / g.nbx;
    const int t = threadIdx.x;
    const bool active = t < BS * BS;
    const int i = active ? t % BS : 0, j = active ? t / BS : 0;
    const int x = bx * BS, y = by * BS;
    const uint8_t* row = a.cur + unit * a.cur_unit_stride + (size_t)(y + j) * g.W;
    const int c = row[x + i];
    const MeResult mp = a.me_parent[unit * a.me_parent_stride + blk];
    const int resp = c - intra_pred(row, x, mp.dx, i, x, bx == 0);
    if (active) ws[j * P + i] = so_i2d(resp);
    transform2d<BS, BS, false>(ws, t);
    const int tcp = active ? so_d2i_rint(ws[j * P + i]) : 0;

    const bool eligible = a.vbs && bx != 0 && by != 0;
    const int qrow = a.qp_blocks ? a.qp_blocks[blk] : (a.qp_rows ? a.qp_rows[by] : a.qp_final);
    int split = 0;
    const int k = (j >= S ? 2 : 0) + (i >= S ? 1 : 0);
    const int si = i % S, sj = j % S;
    int tcs = 0;
    unsigned long long mae_n = mp.sad;      // units of 1/BS^2
    bool mae_inf = mp.none;
    if (eligible) {
        const int sb = (by * 2 + (k >> 1)) * (g.nbx * 2) + bx * 2 + (k & 1);
        const MeResult ms = a.me_sub[unit * a.me_sub_stride + sb];
        const int xs = x + (k & 1) * S;
        const int ress = c - intra_pred(row, xs, ms.dx, si, x, false);
        if (active) ws[j * P + i] = so_i2d(ress);
        transform2d<BS, S, false>(ws, t);
        tcs = active ? so_d2i_rint(ws[j * P + i]) : 0;
        const int lenp = rle_len_cta<BS, BS>(quant_rhe(tcp, q_shift(j, i, BS, a.qp_rd)), j, i, 0, nzbuf, active);
        const int qs = a.qp_rd > 0 ? a.qp_rd - 1 : a.qp_rd;
        const int lens = rle_len_cta<BS, S>(quant_rhe(tcs, q_shift(sj, si, S, qs)), sj, si, k, nzbuf, active);
        double vm = 0.0;
        unsigned long long vn = 0; bool vinf = false;
        for (int kk = 0; kk < 4; ++kk) {
            const MeResult m = a.me_sub[unit * a.me_sub_stride + (by * 2 + (kk >> 1)) * (g.nbx * 2) + bx * 2 + (kk & 1)];
            vm = __dadd_rn(vm, me_mae(m, S, 0));
            vn += m.sad; vinf = vinf || m.none;
        }
        vm = vm / 4.0;
        const double rd_bs = __dadd_rn(__dmul_rn(a.lam, (double)(8 + 8 * lenp)), me_mae(mp, BS, 0));
        const double rd_vbs = __dadd_rn(__dmul_rn(a.lam, (double)(32 + 8 * lens)), vm);
        split = (rd_bs < rd_vbs) ? 0 : 1;
        mae_n = vn; mae_inf = vinf;
    }
    int level, shift;
    if (!split) { shift = q_shift(j, i, BS, qrow); level = quant_rhe(tcp, shift); }
    else { const int qs = qrow > 0 ? qrow - 1 : qrow; shift = q_shift(sj, si, S, qs); level = quant_rhe(tcs, shift); }
    const int len = split ? rle_len_cta<BS, S>(level, sj, si, k, nzbuf, active)
                          : rle_len_cta<BS, BS>(level, j, i, 0, nzbuf, active);
    if (active) {
        a.levels[unit * a.frame_stride + (size_t)(y + j) * g.W + x + i] = (int16_t)level;
        ws[j * P + i] = so_i2d(level * (1 << shift));
    }
    if (split) transform2d<BS, S, true>(ws, t); else transform2d<BS, BS, true>(ws, t);
    if (active) a.res_frame[unit * a.scratch_stride + (size_t)(y + j) * g.W + x + i] = (int16_t)so_d2i_rint(ws[j * P + i]);
    if (t == 0) {
        a.split[unit * a.split_stride + blk] = (uint8_t)split;
        int16_t* mvo = a.mv + unit * a.mv_stride + (size_t)blk * 12;
        for (int kk = 0; kk < 4; ++kk) {
            MeResult m = mp;
            if (split) m = a.me_sub[unit * a.me_sub_stride + (by * 2 + (kk >> 1)) * (g.nbx * 2) + bx * 2 + (kk & 1)];
            const bool on = split || kk == 0;
            mvo[kk * 3 + 0] = on ? m.dx : 0; mvo[kk * 3 + 1] = 0; mvo[kk * 3 + 2] = 0;
        }
        so_frame_stats* st = a.stats + unit * a.stats_stride;
        if (blk == 0) { st->mae_den = a.mae_den; st->frame_type = a.frame_type; }
        atomicAdd(&st->qsize, (unsigned)len);
        if (a.blk_len) a.blk_len[unit * a.blk_len_stride + blk] = (uint32_t)len;
        atomicAdd(a.row_sizes + unit * a.rows_stride + by, (unsigned)len);
        if (mae_inf) atomicOr(&st->mae_inf, 1u); else atomicAdd(reinterpret_cast<unsigned long long*>(&st->mae_num), mae_n);
    }
}

// intra reconstruction chain (reconstruct_frame_intra, Encoder.py:1378-1414): one CTA per block row, blocks in
// sequence; the frame being built is int32 and unclipped, the block under construction still reads as 128.
template <int BS>
__global__ void intra_recon_kernel(const FlowArgs a) {
    constexpr int S = BS / 2;
    __shared__ unsigned long long sbuf[32];
    const FrameGeom& g = a.g;
    const int by = blockIdx.x, unit = a.unit0 + blockIdx.y;
    const int t = threadIdx.x;
    const bool active = t < BS * BS;
    const int i = active ? t % BS : 0, j = active ? t / BS : 0;
    const int y = by * BS;
    int32_t* band = a.band + unit * a.scratch_stride + (size_t)(y + j) * g.W;
    const int16_t* res = a.res_frame + unit * a.scratch_stride + (size_t)(y + j) * g.W;
    const uint8_t* cur = a.cur + unit * a.cur_unit_stride + (size_t)(y + j) * g.W;
    uint8_t* rec = a.recon + unit * a.frame_stride + (size_t)(y + j) * g.W;
    if (active) for (int xx = i; xx < g.W; xx += BS) band[xx] = 128;
    __syncthreads();
    unsigned long long se = 0;
    const int k = (j >= S ? 2 : 0) + (i >= S ? 1 : 0);
    for (int bx = 0; bx < g.nbx; ++bx) {
        const int blk = by * g.nbx + bx, x = bx * BS;
        const int split = a.split[unit * a.split_stride + blk];
        const int16_t* mvo = a.mv + unit * a.mv_stride + (size_t)blk * 12;
        int val = 0;
        if (active) {
            const int r = res[x + i];
            if (bx == 0) val = 128 + r;
            else if (!split) val = band[x + mvo[0] + i] + r;
            else val = band[x + (k & 1) * S + mvo[k * 3] + (i % S)] + r;
        }
        __syncthreads();
        if (active) {
            band[x + i] = val;
            const int rv = val & 0xFF;
            rec[x + i] = (uint8_t)rv;
            const int d = rv - (int)cur[x + i];
            se += (unsigned long long)(d * d);
        }
        __syncthreads();
    }
    se = block_sum_u64(se, sbuf);
    if (t == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&(a.stats + unit * a.stats_stride)->sse), se);
}

// 16x16 variant of the row chain: a pixel row only ever reads its own row (the predictor is horizontal), so the 16 lanes
// of a half warp carry one row on their own -- no CTA barrier, no global read-after-write.  The reconstructed row lives in
// a 128-column shared-memory ring (a predictor reaches at most r <= 63 columns back); columns at or right of the block
// being built read as 128 like the unwritten frame does.  Residuals, split flags and vectors of the next block are
// fetched while the current one is computed.
__global__ void __launch_bounds__(256) intra_recon16_kernel(const FlowArgs a) {
    constexpr int BS = 16, S = 8, RING = 128, PF = 8;
    __shared__ int ring[BS][RING];
    __shared__ unsigned long long sbuf[32];
    extern __shared__ int16_t s_mv[];                 // [nbx][4] horizontal offsets (slot 0 only when not split), then [nbx] split flags
    const FrameGeom& g = a.g;
    const int by = blockIdx.x, unit = a.unit0 + blockIdx.y;
    const int t = threadIdx.x, i = t & 15, j = t >> 4;
    const int y = by * BS;
    const int16_t* res = a.res_frame + unit * a.scratch_stride + (size_t)(y + j) * g.W;
    const uint8_t* cur = a.cur + unit * a.cur_unit_stride + (size_t)(y + j) * g.W;
    uint8_t* rec = a.recon + unit * a.frame_stride + (size_t)(y + j) * g.W;
    uint8_t* s_split = reinterpret_cast<uint8_t*>(s_mv + (size_t)g.nbx * 4);
    {   // vectors and split flags of the whole block row: one parallel load, the chain below never waits on them
        const uint8_t* splitp = a.split + unit * a.split_stride + (size_t)by * g.nbx;
        const int16_t* mvp = a.mv + unit * a.mv_stride + (size_t)by * g.nbx * 12;
        for (int e = t; e < g.nbx * 4; e += blockDim.x) s_mv[e] = mvp[(size_t)(e >> 2) * 12 + (e & 3) * 3];
        for (int e = t; e < g.nbx; e += blockDim.x) s_split[e] = splitp[e];
    }
    const int k = (j >= S ? 2 : 0) + (i >= S ? 1 : 0);
    unsigned long long se = 0;
    int rq[PF], cq[PF];                               // residual / current pixel of the next PF blocks
#pragma unroll
    for (int p = 0; p < PF; ++p) {
        rq[p] = p < g.nbx ? res[p * BS + i] : 0;
        cq[p] = p < g.nbx ? cur[p * BS + i] : 0;
    }
    __syncthreads();
    for (int bx0 = 0; bx0 < g.nbx; bx0 += PF) {
        int rn[PF], cn[PF];                           // the whole next group is requested before this one is worked through
#pragma unroll
        for (int p = 0; p < PF; ++p) {
            const int bn = bx0 + PF + p;
            rn[p] = bn < g.nbx ? res[bn * BS + i] : 0;
            cn[p] = bn < g.nbx ? cur[bn * BS + i] : 0;
        }
#pragma unroll
        for (int p = 0; p < PF; ++p) {
            const int bx = bx0 + p;
            if (bx < g.nbx) {
                const int x = bx * BS;
                const int r = rq[p], c = cq[p];
                const int split = s_split[bx];
                const int mv = s_mv[bx * 4 + (split ? k : 0)];
                int val;
                if (bx == 0) val = 128 + r;
                else {
                    const int col = split ? x + (k & 1) * S + mv + (i & (S - 1)) : x + mv + i;
                    const int pv = (col >= x || col < 0) ? 128 : ring[j][col & (RING - 1)];
                    val = pv + r;
                }
                __syncwarp();                       // everybody has read the ring before anybody overwrites it
                ring[j][(x + i) & (RING - 1)] = val;
                const int rv = val & 0xFF;          // astype(np.uint8) of the whole frame wraps (appendix A5)
                rec[x + i] = (uint8_t)rv;
                const int d = rv - c;
                se += (unsigned long long)(d * d);
                __syncwarp();
            }
        }
#pragma unroll
        for (int p = 0; p < PF; ++p) { rq[p] = rn[p]; cq[p] = cn[p]; }
    }
    se = block_sum_u64(se, sbuf);
    if (t == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&(a.stats + unit * a.stats_stride)->sse), se);
}

// ------------------------------------------------------------------------------------------------------------
// decoder side (decoder.py:97-211 decode_frame_inter, :330-432 decode_frame_intra): levels + vectors -> frame.
// Same arithmetic as the encoder's reconstruction (including the split-block bounds quirk Q5, decoder.py:185).
// intra != 0: only the dequantised residual is produced (res_frame); intra_recon_kernel then runs the row chain.
// ------------------------------------------------------------------------------------------------------------
template <int BS>
__global__ void decode_block_kernel(const FlowArgs a, int intra) {
    constexpr int S = BS / 2;
    constexpr int P = BS + 1;
    __shared__ double ws[BS * P];
    const FrameGeom& g = a.g;
    const int blk = blockIdx.x, unit = a.unit0 + blockIdx.y;
    const int bx = blk % g.nbx, by = blk / g.nbx;
    const int t = threadIdx.x;
    const bool active = t < BS * BS;
    const int i = active ? t % BS : 0, j = active ? t / BS : 0;
    const int x = bx * BS, y = by * BS;
    const int split = a.split[unit * a.split_stride + blk];
    const int16_t* mvo = a.mv + unit * a.mv_stride + (size_t)blk * 12;
    const int qrow = a.qp_blocks ? a.qp_blocks[blk] : (a.qp_rows ? a.qp_rows[by] : a.qp_final);
    const int k = (j >= S ? 2 : 0) + (i >= S ? 1 : 0);
    const int si = i % S, sj = j % S;
    const int level = a.levels[unit * a.frame_stride + (size_t)(y + j) * g.W + x + i];
    int shift;
    if (!split) shift = q_shift(j, i, BS, qrow);
    else { const int qs = qrow > 0 ? qrow - 1 : qrow; shift = q_shift(sj, si, S, qs); }
    if (active) ws[j * P + i] = so_i2d(level * (1 << shift));
    if (split) transform2d<BS, S, true>(ws, t); else transform2d<BS, BS, true>(ws, t);
    if (!active) return;
    const int r = so_d2i_rint(ws[j * P + i]);
    if (intra) {
        a.res_frame[unit * a.scratch_stride + (size_t)(y + j) * g.W + x + i] = (int16_t)r;
        return;
    }
    RefList rl;
    for (int q = 0; q < g.nref; ++q)
        for (int ph = 0; ph < 4; ++ph) rl.plane[q][ph] = a.ring.plane(unit, q, ph);
    const int mult = g.fme ? 2 : 1;
    int pred;
    if (!split) {
        const PredSel sel = pred_select(g, x * mult, y * mult, mvo[0], mvo[1], BS, -1);
        pred = pred_sample(g, rl, sel, mvo[2], i, j);
    } else {
        const int xs = x + (k & 1) * S, ys = y + (k >> 1) * S;
        const PredSel sel = pred_select(g, xs * mult, ys * mult, mvo[k * 3], mvo[k * 3 + 1], S, BS);
        pred = pred_sample(g, rl, sel, mvo[k * 3 + 2], si, sj);
    }
    a.recon[unit * a.frame_stride + (size_t)(y + j) * g.W + x + i] = (uint8_t)((pred + r) & 0xFF);
}

// ------------------------------------------------------------------------------------------------------------
// run-level symbol generation (entropy_encoder_block, Encoder.py:1086-1131) with a device-side prefix scan
//   pass 1 (emit == 0): per (sub-)block symbol counts -> lens[frame][blk][4]
//   scan              : exclusive prefix sum of lens per frame -> offs[frame][4*nblk + 1]
//   pass 2 (emit == 1): symbols written at their final position of the packed per-frame stream (int16)
// One CTA per block, one thread per coefficient in scan order.  A symbol stream is: for every maximal run, a header
// (-count for non-zero runs followed by the values, +count for zero runs, a single 0 for a trailing zero run).
// Every non-zero coefficient owns one slot, every run start owns one slot:
//   slot(header of run starting at p) = #nonzeros before p + #run starts before p
//   slot(value at p)                  = #nonzeros before p + #run starts up to and including p's run
// ------------------------------------------------------------------------------------------------------------
// The launch covers frames [f0, f0 + nf) of every unit of a [unit][F] sequence: blockIdx.y = unit * nf + i -> frame
// unit * F + f0 + i (whole sequence: f0 = 0, nf = F).
__device__ __forceinline__ int seq_frame(int y, int f0, int nf, int F) { return (y / nf) * F + f0 + (y % nf); }

template <int BS>
__global__ void rle_symbols_kernel(const int16_t* levels, const uint8_t* split, uint32_t* lens, const uint32_t* offs, int16_t* syms,
                                   size_t sym_frame_stride, int W, int nbx, int nblk, int emit, int f0, int nf, int F) {
    constexpr int S = BS / 2;
    __shared__ int16_t sv[BS * BS];         // coefficient at (segment, scan position)
    __shared__ int snz[BS * BS + 1], sst[BS * BS + 1];     // inclusive prefix sums of non-zero / run-start flags
    __shared__ int wsum[2][32];
    const int blk = blockIdx.x, frame = seq_frame(blockIdx.y, f0, nf, F);
    const int t = threadIdx.x;
    const bool active = t < BS * BS;
    const int bx = blk % nbx, by = blk / nbx;
    const int sp = split[(size_t)frame * nblk + blk];
    // element handled by this thread: pixel (i, j) -> (segment, scan position)
    const int i = active ? t % BS : 0, j = active ? t / BS : 0;
    const int n = sp ? S : BS;
    const int seg = sp ? ((j >= S ? 2 : 0) + (i >= S ? 1 : 0)) : 0;
    const int u = sp ? j % S : j, v = sp ? i % S : i;
    const int p = c_scanpos[tbl_off(n) + u * n + v];
    const int e = seg * n * n + p;                       // index in segment-major scan order
    if (active) sv[e] = levels[(size_t)frame * W * (nblk / nbx) * BS + (size_t)(by * BS + j) * W + bx * BS + i];
    __syncthreads();
    // thread t now owns element t of the segment-major order
    const int nseg_elems = n * n;
    const int myseg = active ? t / nseg_elems : 0, myp = active ? t % nseg_elems : 0;
    const int val = active ? sv[t] : 0;
    const int nz = active && val != 0;
    const int prev_nz = (active && myp > 0) ? (sv[t - 1] != 0) : 0;
    const int st = active && (myp == 0 || nz != prev_nz);
    // block-wide inclusive scans of nz and st
    int a = nz, b = st;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xFFFFFFFFu, a, o), y = __shfl_up_sync(0xFFFFFFFFu, b, o);
        if ((t & 31) >= o) { a += x; b += y; }
    }
    if ((t & 31) == 31) { wsum[0][t >> 5] = a; wsum[1][t >> 5] = b; }
    __syncthreads();
    int ba = 0, bb = 0;
    for (int w = 0; w < (t >> 5); ++w) { ba += wsum[0][w]; bb += wsum[1][w]; }
    a += ba; b += bb;
    if (active) { snz[t + 1] = a; sst[t + 1] = b; }
    if (t == 0) { snz[0] = 0; sst[0] = 0; }
    __syncthreads();
    const int nsegs = sp ? 4 : 1;
    if (!emit) {
        if (t < 4) {
            uint32_t len = 0;
            if (t < nsegs) {
                const int lo = t * nseg_elems, hi = lo + nseg_elems;
                len = (uint32_t)((snz[hi] - snz[lo]) + (sst[hi] - sst[lo]));
            }
            lens[((size_t)frame * nblk + blk) * 4 + t] = len;
        }
        return;
    }
    if (!active) return;
    const int lo = myseg * nseg_elems;
    const uint32_t base = offs[(size_t)frame * (4 * nblk + 1) + (size_t)blk * 4 + myseg];
    int16_t* out = syms + (size_t)frame * sym_frame_stride + base;
    const int nzb = snz[t] - snz[lo], stb = sst[t] - sst[lo];           // counts strictly before this element, within the segment
    if (st) {
        int len = 1;
        while (myp + len < nseg_elems && ((sv[t + len] != 0) == nz)) ++len;
        const bool trailing = (myp + len == nseg_elems);
        out[nzb + stb] = (int16_t)(nz ? -len : (trailing ? 0 : len));
    }
    if (nz) out[nzb + (sst[t + 1] - sst[lo])] = (int16_t)val;
}

// Emit pass of the symbol packer used on the end-to-end path: one WARP per block.  The per-block symbol counts come from the
// finish kernels (blk_len), their exclusive scan per frame (scan_lens_kernel) gives every block its place in the frame's
// packed stream.  Lane l owns scan positions l, l + 32, ...: a ballot per 32 positions yields the block's non-zero bitmap
// in scan order, replicated in every lane, and everything else is bit arithmetic on it --
//   run starts   S = nz ^ (nz << 1)  |  segment starts (a split block is four consecutive segments)
//   slot(header of the run starting at p) = #nz before p + #starts before p
//   slot(value at p)                      = #nz before p + #starts up to and including p
//   run length = distance to the next start (or the block end); a zero run that reaches its segment's end is the symbol 0.
template <int BS>
__global__ void __launch_bounds__(256) rle_emit_warp_kernel(const int16_t* levels, const uint8_t* split, const uint32_t* boffs, int16_t* syms,
                                                            size_t sym_frame_stride, int W, int H, int nbx, int nblk, int f0, int nf, int F) {
    constexpr int NN = BS * BS, NW = (NN + 31) / 32, S = BS / 2;
    __shared__ uint8_t inv_full[NN], inv_sub[NN / 4];          // scan position -> pixel index inside the (sub-)block, row-major
    for (int t = threadIdx.x; t < NN; t += blockDim.x) inv_full[c_scanpos[tbl_off(BS) + t]] = (uint8_t)t;
    for (int t = threadIdx.x; t < NN / 4; t += blockDim.x) inv_sub[c_scanpos[tbl_off(S) + t]] = (uint8_t)t;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = blockIdx.x * 8 + warp;
    if (blk >= nblk) return;
    const int frame = seq_frame(blockIdx.y, f0, nf, F);
    const int sp = split[(size_t)frame * nblk + blk];
    const int bx = blk % nbx, by = blk / nbx;
    const int16_t* lv = levels + (size_t)frame * W * H + (size_t)(by * BS) * W + bx * BS;
    const int seglen = sp ? NN / 4 : NN;
    int val[NW];
    uint32_t nzw[NW], stw[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const int p = lane + 32 * k;
        int v = 0;
        if (p < NN) {
            int i, j;
            if (sp) {
                const int seg = p / (NN / 4), pix = inv_sub[p % (NN / 4)];
                i = pix / S + (seg >> 1) * S; j = pix % S + (seg & 1) * S;
            } else {
                const int pix = inv_full[p];
                i = pix / BS; j = pix % BS;
            }
            v = lv[(size_t)i * W + j];
        }
        val[k] = v;
        nzw[k] = __ballot_sync(0xFFFFFFFFu, v != 0);
    }
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const uint32_t prev = (nzw[k] << 1) | (k ? (nzw[k > 0 ? k - 1 : 0] >> 31) : 0u);
        uint32_t segm;
        if (seglen >= 32) segm = ((32 * k) % seglen == 0) ? 1u : 0u;
        else segm = seglen == 16 ? 0x00010001u : (seglen == 8 ? 0x01010101u : 0x11111111u);
        const uint32_t act = (NN - 32 * k >= 32) ? 0xFFFFFFFFu : ((1u << (NN - 32 * k > 0 ? NN - 32 * k : 0)) - 1u);
        stw[k] = ((nzw[k] ^ prev) | segm) & act;
    }
    int16_t* out = syms + (size_t)frame * sym_frame_stride + boffs[(size_t)frame * (nblk + 1) + blk];
    const uint32_t below = (1u << lane) - 1u;
    int nzpre = 0, stpre = 0;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const int p = lane + 32 * k;
        const int nz = (nzw[k] >> lane) & 1, st = (stw[k] >> lane) & 1;
        const int nzb = nzpre + __popc(nzw[k] & below), stb = stpre + __popc(stw[k] & below);
        if (st) {
            const uint32_t m = lane == 31 ? 0u : (stw[k] >> (lane + 1));
            int q = NN;                                  // next run start (or the end of the block)
            if (m) q = p + __ffs(m);
            else {
#pragma unroll
                for (int w = NW - 1; w > k; --w) if (stw[w]) q = 32 * w + __ffs(stw[w]) - 1;
            }
            const int len = q - p;
            out[nzb + stb] = (int16_t)(nz ? -len : ((q % seglen == 0) ? 0 : len));
        }
        if (nz) out[nzb + stb + st] = (int16_t)val[k];
        nzpre += __popc(nzw[k]); stpre += __popc(stw[k]);
    }
}

// exclusive prefix sum of n entries per frame (one CTA per frame): offs[frame][0..n], offs[frame][n] = totals[frame] = total
__global__ void scan_lens_kernel(const uint32_t* lens, uint32_t* offs, int n, uint32_t* totals, int f0, int nf, int F) {
    __shared__ uint32_t wtot[32];
    __shared__ uint32_t carry;
    const int frame = seq_frame(blockIdx.x, f0, nf, F), t = threadIdx.x;
    const uint32_t* in = lens + (size_t)frame * n;
    uint32_t* out = offs + (size_t)frame * (n + 1);
    if (t == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + t;
        const uint32_t v = i < n ? in[i] : 0u;
        uint32_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, s, o);
            if ((t & 31) >= o) s += x;
        }
        if ((t & 31) == 31) wtot[t >> 5] = s;
        __syncthreads();
        uint32_t wb = 0;
        for (int w = 0; w < (t >> 5); ++w) wb += wtot[w];
        const uint32_t c0 = carry;
        if (i < n) out[i] = c0 + wb + s - v;
        __syncthreads();
        if (t == blockDim.x - 1) carry = c0 + wb + s;
        __syncthreads();
    }
    if (t == 0) { out[n] = carry; totals[frame] = carry; }
}
