// Exhaustive motion search with successive elimination (16x16 blocks, r = 16, no VBS): the same result as
// find_best_match (Encoder.py:678-717) -- the lexicographic minimum of (SAD, |dx|+|dy|, ref, dx, dy) over every valid
// candidate -- without evaluating the SAD of candidates that provably cannot reach it.
//
// Bound.  For a candidate window w and the current block c, split both into their four 8x8 quadrants:
//     SAD(c, w) = sum |c_ij - w_ij|  >=  sum_q |S_q(c) - S_q(w)|            (triangle inequality per quadrant)
// With quadrant sums quantised to bytes, s_q = S_q >> 6 (S_q <= 64 * 255), |S_q(c) - S_q(w)| >= 64 |s_q(c) - s_q(w)| - 63, so
//     SAD >= 64 * D - 252,   D = sum_q |s_q(c) - s_q(w)|  = ONE VABSDIFF4 on two packed words.
// Let U be the exact SAD of ANY valid candidate of the block (so U >= the final minimum).  A candidate with D > T,
// T = (U + 252) >> 6, has 64 * D - 252 > U, hence SAD > U: it cannot be the minimum key, ties included (a tie needs
// SAD == minimum <= U), and is dropped.  Every candidate that passes gets its exact SAD, is merged into the block's 64-bit key
// like in the plain search, and tightens T.  The first U comes from predictor candidates (the vector the block chose in the
// previous P frame, the same motion scaled to the newest reference, and (0, 0)); all of them are ordinary valid candidates.
// The filter costs one 128-bit load and four VABSDIFF4 per FOUR candidates instead of 256 VABSDIFF4.
//
// Kernels: sea_qplane_kernel (per new reference: packed quadrant bytes of every window position of every phase plane),
// sea_search_kernel (per P frame: predictors, filter, exact SADs of the survivors, next frame's predictors).
#pragma once
#include "so_common.cuh"

constexpr int SEA_NB = 4;                        // horizontally adjacent blocks per CTA (window columns re-read from L1)
constexpr int SEA_LCAP = 4096;                   // survivor list of a CTA
#ifndef SEA_CHUNK
#define SEA_CHUNK 4                              // window rows a lane loads before it looks at them
#endif
#ifndef SEA_MINB
#define SEA_MINB 4                               // resident CTAs per SM the search kernel is compiled for
#endif

struct SeaArgs {
    FrameGeom g;                       // g.bs == 16; g.nref = length of the reference list
    const uint8_t* ring;               // reference ring of unit 0: [unit][slot][phase 4][shift 4][H][pitch]
    size_t unit_stride, slot_stride, plane_bytes;
    const uint32_t* pq;                // packed quadrant bytes [unit][slot][phase][H][W]
    size_t pq_unit_stride, pq_slot_stride, pq_plane_stride;      // in words
    unsigned int slot_packed;          // list index -> ring slot, 4 bits each
    const uint8_t* cur;                // current frames, first unit of the launch
    size_t cur_unit_stride;
    unsigned long long* out;           // packed keys of the first unit of the launch, stride of MeResult (16 B) per block
    size_t out_unit_stride;            // in MeResult elements
    uint32_t* prev;                    // [unit][nblk]: winner of the last P frame (ref << 16 | dx + R << 8 | dy + R), all ones = none
    unsigned int* ctr;                 // [0..1] 64-bit count of exact SADs, [2] launches, [6] finished CTAs / [7] exact SADs of this launch
    unsigned int* host_stat;           // mapped host memory or null: {exact SADs of the last finished launch, its sequence number}
    unsigned int seq;                  // sequence number of this launch
    int unit0, units, nph, nblk;
};

__device__ __forceinline__ unsigned long long sea_key64(uint32_t sad, int ref, int dx, int dy, int R) {
    return ((unsigned long long)sad << 40) | ((unsigned long long)(abs(dx) + abs(dy)) << 24) | ((unsigned long long)ref << 16) |
           ((unsigned long long)(dx + R) << 8) | (unsigned long long)(dy + R);
}

// ---- packed quadrant bytes ------------------------------------------------------------------------------------------------
// pq[y][x] = {s(y, x), s(y, x + 8), s(y + 8, x), s(y + 8, x + 8)} (bytes 0..3), s(y, x) = (sum of the 8x8 pixels at (y, x)) >> 6.
// Row sums of 8 pixels come from two aligned words of the copy of the plane shifted by x & 3 bytes (DP4A with ones).
// Positions whose 16x16 window leaves the plane hold partial sums: only invalid candidates sit there (valid_range).
constexpr int SQ_TX = 64, SQ_TY = 32, SQ_THREADS = 4 * (SQ_TX + 8);       // tile of outputs per CTA; 72 columns x 4 row groups of threads
__global__ void __launch_bounds__(SQ_THREADS) sea_qplane_kernel(const uint8_t* slot0, size_t unit_stride, size_t plane_bytes, uint32_t* pq0,
                                                                size_t pq_unit_stride, size_t pq_plane_stride, int W, int H, int pitch, int nph) {
    pdl_trigger();
    pdl_wait();
    __shared__ uint16_t hs[SQ_TY + 15][SQ_TX + 8];                  // sums of 8 pixels of a row
    __shared__ __align__(16) uint8_t qs[SQ_TY + 8][SQ_TX + 8];      // quantised 8x8 sums
    const int unit = blockIdx.z / nph, ph = blockIdx.z - unit * nph;
    const int x0 = blockIdx.x * SQ_TX, y0 = blockIdx.y * SQ_TY;
    const int cx = threadIdx.x % (SQ_TX + 8), rg = threadIdx.x / (SQ_TX + 8);
    {
        // rows rg, rg + 4, ...: all loads of the thread first (they are independent)
        const int x = x0 + cx, m = x >> 2, p4 = pitch >> 2;
        const bool in_x = x < W, has1 = 4 * (m + 1) < pitch;
        const uint32_t* col = reinterpret_cast<const uint32_t*>(slot0 + unit * unit_stride + (size_t)(ph * 4 + (x & 3)) * plane_bytes) + m;
        uint32_t w0[12], w1[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int r = rg + 4 * i, y = y0 + r;
            w0[i] = 0u; w1[i] = 0u;
            if (r < SQ_TY + 15 && in_x && y < H) {
                w0[i] = __ldg(col + y * p4);
                if (has1) w1[i] = __ldg(col + y * p4 + 1);
            }
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int r = rg + 4 * i;
            if (r < SQ_TY + 15) hs[r][cx] = (uint16_t)__dp4a(w1[i], 0x01010101u, __dp4a(w0[i], 0x01010101u, 0u));
        }
    }
    __syncthreads();
    {
        // vertical sums of 8 rows, sliding: thread (cx, rg) owns rows 10 rg .. 10 rg + 9 of the 40
        const int r0 = rg * ((SQ_TY + 8) / 4);
        uint32_t sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += hs[r0 + i][cx];
        qs[r0][cx] = (uint8_t)(sum >> 6);
#pragma unroll
        for (int j = 1; j < (SQ_TY + 8) / 4; ++j) {
            sum += hs[r0 + j + 7][cx];
            sum -= hs[r0 + j - 1][cx];
            qs[r0 + j][cx] = (uint8_t)(sum >> 6);
        }
    }
    __syncthreads();
    uint32_t* out = pq0 + unit * pq_unit_stride + (size_t)ph * pq_plane_stride;
    for (int idx = threadIdx.x; idx < SQ_TY * (SQ_TX / 4); idx += SQ_THREADS) {
        // four horizontally adjacent outputs: a 4x4 byte transpose of {s(y,x..), s(y,x+8..), s(y+8,x..), s(y+8,x+8..)}
        const int r = idx / (SQ_TX / 4), c4 = (idx - r * (SQ_TX / 4)) * 4;
        const int y = y0 + r, x = x0 + c4;
        if (y >= H || x >= W) continue;
        const uint32_t a = *reinterpret_cast<const uint32_t*>(&qs[r][c4]), b = *reinterpret_cast<const uint32_t*>(&qs[r][c4 + 8]);
        const uint32_t c = *reinterpret_cast<const uint32_t*>(&qs[r + 8][c4]), d = *reinterpret_cast<const uint32_t*>(&qs[r + 8][c4 + 8]);
        const uint32_t t0 = __byte_perm(a, b, 0x5140), t1 = __byte_perm(a, b, 0x7362);
        const uint32_t u0 = __byte_perm(c, d, 0x5140), u1 = __byte_perm(c, d, 0x7362);
        *reinterpret_cast<uint4*>(out + (size_t)y * W + x) =
            make_uint4(__byte_perm(t0, u0, 0x5410), __byte_perm(t0, u0, 0x7632), __byte_perm(t1, u1, 0x5410), __byte_perm(t1, u1, 0x7632));
    }
}

// candidate (ref, dx, dy) of block (bx, by) -> address of its first window row in the shift copy that makes it word aligned
__device__ __forceinline__ const uint8_t* sea_window(const SeaArgs& a, const uint8_t* ring_u, int bx, int by, int ref, int dx, int dy) {
    const FrameGeom& g = a.g;
    const int px = g.fme ? (dx & 1) : 0, py = g.fme ? (dy & 1) : 0;
    const int ox = g.fme ? ((dx - px) >> 1) : dx, oy = g.fme ? ((dy - py) >> 1) : dy;
    const int X = bx * 16 + ox, Y = by * 16 + oy, c = X & 3;
    const int slot = (int)((a.slot_packed >> (4 * ref)) & 15u);
    return ring_u + slot * a.slot_stride + (size_t)(((py << 1) | px) * 4 + c) * a.plane_bytes + (size_t)(Y * g.pitch + X - c);
}

// Exact SADs of listed candidates, two per step and warp (half a warp each: lane -> block row, the current row from shared
// memory, four aligned words of the window row), merged into the block's key; every result tightens the block's threshold.
// List entry: dx + R | (dy + R) << 8 | ref << 16 | block in CTA << 20 | D << 22.  Entries first, first + stride, ... < total;
// called by all 32 lanes of the warp.
__device__ __forceinline__ unsigned int sea_eval(const SeaArgs& a, const uint8_t* ring_u, int bx0, int by, const uint8_t (*cur)[256], uint32_t* thr,
                                                 unsigned long long* key, const uint32_t* list, int first, int stride, int total, int lane) {
    const FrameGeom& g = a.g;
    const int l = lane & 15, half = lane >> 4;
    unsigned int evals = 0;
    for (int j = first; j < total; j += stride) {
        const int idx = j + half;
        const bool act = idx < total;
        const uint32_t e = list[act ? idx : j];
        const int b = (int)((e >> 20) & 3u);
        const bool go = act && (e >> 22) <= *reinterpret_cast<volatile uint32_t*>(thr + b);      // the threshold may have moved since the filter
        const int ref = (int)((e >> 16) & 15u), dx = (int)(e & 0xFFu) - g.R, dy = (int)((e >> 8) & 0xFFu) - g.R;
        uint32_t s = 0;
        if (go) {
            const uint4 c = *reinterpret_cast<const uint4*>(&cur[b][l * 16]);
            const uint32_t* w = reinterpret_cast<const uint32_t*>(sea_window(a, ring_u, bx0 + b, by, ref, dx, dy) + (size_t)(l * g.pitch));
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
            s = sad4_acc(w0, c.x, 0u);
            s = sad4_acc(w1, c.y, s);
            s = sad4_acc(w2, c.z, s);
            s = sad4_acc(w3, c.w, s);
        }
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        if (go && l == 0) {
            atomicMin(key + b, sea_key64(s, ref, dx, dy, g.R));
            atomicMin(thr + b, (s + 252u) >> 6);
            ++evals;
        }
    }
    return evals;
}

// Exact SADs of listed candidates, one per THREAD: entries first, first + stride, ... < total.  The survivors of a block
// cluster around a few positions, so the 16 x 4 word loads of neighbouring lanes mostly hit the same L1 lines.
__device__ __forceinline__ unsigned int sea_eval_thread(const SeaArgs& a, const uint8_t* ring_u, int bx0, int by, const uint8_t (*cur)[256], uint32_t* thr,
                                                        unsigned long long* key, const uint32_t* list, int first, int stride, int total) {
    const FrameGeom& g = a.g;
    unsigned int evals = 0;
    for (int j = first; j < total; j += stride) {
        const uint32_t e = list[j];
        const int b = (int)((e >> 20) & 3u);
        if ((e >> 22) > *reinterpret_cast<volatile uint32_t*>(thr + b)) continue;      // the threshold may have moved since the filter
        const int ref = (int)((e >> 16) & 15u), dx = (int)(e & 0xFFu) - g.R, dy = (int)((e >> 8) & 0xFFu) - g.R;
        const uint32_t* w = reinterpret_cast<const uint32_t*>(sea_window(a, ring_u, bx0 + b, by, ref, dx, dy));
        const int p4 = g.pitch >> 2;
        uint32_t s0 = 0, s1 = 0;
        bool hopeless = false;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int r = 8 * half; r < 8 * half + 8; r += 2) {
                const uint4 c0 = *reinterpret_cast<const uint4*>(&cur[b][r * 16]), c1 = *reinterpret_cast<const uint4*>(&cur[b][r * 16 + 16]);
                const uint32_t* w0 = w + r * p4;
                const uint32_t* w1 = w0 + p4;
                s0 = sad4_acc(__ldg(w0), c0.x, s0); s0 = sad4_acc(__ldg(w0 + 1), c0.y, s0);
                s0 = sad4_acc(__ldg(w0 + 2), c0.z, s0); s0 = sad4_acc(__ldg(w0 + 3), c0.w, s0);
                s1 = sad4_acc(__ldg(w1), c1.x, s1); s1 = sad4_acc(__ldg(w1 + 1), c1.y, s1);
                s1 = sad4_acc(__ldg(w1 + 2), c1.z, s1); s1 = sad4_acc(__ldg(w1 + 3), c1.w, s1);
            }
            // the SAD of the upper half alone already exceeds the block's best SAD so far (top 24 bits of its key): this candidate
            // cannot win, not even a tie -- skip its lower half
            if (half == 0 && s0 + s1 > (reinterpret_cast<volatile uint32_t*>(key + b)[1] >> 8)) { hopeless = true; break; }
        }
        ++evals;
        if (hopeless) continue;
        const uint32_t sad = s0 + s1;
        atomicMin(key + b, sea_key64(sad, ref, dx, dy, g.R));
        atomicMin(thr + b, (sad + 252u) >> 6);
    }
    return evals;
}

// the same for a single candidate that found the CTA's list full (kept out of line: it is rare and register hungry)
__device__ __noinline__ unsigned int sea_eval_one(const SeaArgs& a, const uint8_t* ring_u, int bx0, int by, const uint8_t (*cur)[256], uint32_t* thr,
                                                  unsigned long long* key, uint32_t e) {
    return sea_eval_thread(a, ring_u, bx0, by, cur, thr, key, &e, 0, 1, 1);
}

// CTA = SEA_NB horizontally adjacent blocks.  Phase 0: predictors -> first thresholds.  Phase A: every warp filters (reference,
// phase plane, block) triples -- lane -> (quad of four horizontally adjacent offsets 0..8, row 0..2 of a group of three window
// rows): 27 lanes, 11 steps whose loads are independent -- and appends the survivors to the CTA's list.  Phase B: all sixteen
// half warps take exact SADs of listed candidates side by side.  A list that runs full (no usable bound: cold start at a
// frame edge, scene cut) is handled on the spot by the warp that found the candidates, which also tightens the threshold.
__global__ void __launch_bounds__(256, SEA_MINB) sea_search_kernel(const SeaArgs a) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) uint8_t s_cur[SEA_NB][256];
    __shared__ unsigned long long s_key[SEA_NB];
    __shared__ uint32_t s_thr[SEA_NB], s_cq[SEA_NB];
    __shared__ int s_xlo[SEA_NB][2], s_xhi[SEA_NB][2], s_ylo[2], s_yhi[2];
    __shared__ uint32_t s_list[SEA_LCAP];
    __shared__ uint32_t s_q[SEA_NB][8];                      // predictor candidates
    __shared__ unsigned int s_n, s_evals;
    const FrameGeom& g = a.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gpr = (g.nbx + SEA_NB - 1) / SEA_NB;           // block groups per row
    const int by = blockIdx.x / gpr, bx0 = (blockIdx.x - by * gpr) * SEA_NB, u = blockIdx.y;
    const uint8_t* cur = a.cur + u * a.cur_unit_stride;
    const uint8_t* ring_u = a.ring + (size_t)(a.unit0 + u) * a.unit_stride;
    if (tid < SEA_NB * 16) {
        const int b = tid >> 4, row = tid & 15, bx = bx0 + b;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (bx < g.nbx) v = *reinterpret_cast<const uint4*>(cur + (size_t)(by * 16 + row) * g.W + bx * 16);
        *reinterpret_cast<uint4*>(&s_cur[b][row * 16]) = v;
    } else if (tid < SEA_NB * 16 + SEA_NB * 2) {
        // valid offsets of phase px in plane units: mul * ox + px in [max(lo, -R), min(hi, R)]  (Encoder.py:695-698)
        const int b = (tid - SEA_NB * 16) >> 1, px = tid & 1, bx = bx0 + b;
        int lo = 1, hi = 0;
        if (bx < g.nbx) {
            valid_range(bx * 16, g.W, 16, g.fme, g.fme, lo, hi);
            lo = max(lo, -g.R); hi = min(hi, g.R);
            if (g.fme) { lo = (lo - px + 1) >> 1; hi = (hi - px) >> 1; }
        }
        s_xlo[b][px] = lo; s_xhi[b][px] = hi;
    } else if (tid < SEA_NB * 16 + SEA_NB * 2 + 2) {
        const int py = tid & 1;
        int lo, hi;
        valid_range(by * 16, g.H, 16, g.fme, g.fme, lo, hi);
        lo = max(lo, -g.R); hi = min(hi, g.R);
        if (g.fme) { lo = (lo - py + 1) >> 1; hi = (hi - py) >> 1; }
        s_ylo[py] = lo; s_yhi[py] = hi;
    }
    if (tid >= 96 && tid < 96 + SEA_NB) { s_key[tid - 96] = ~0ull; s_thr[tid - 96] = 1023u; }       // D <= 1020: everything passes
    if (tid == 127) { s_evals = 0u; s_n = 0u; }
    __syncthreads();
    unsigned int evals = 0;
    if (warp < SEA_NB && bx0 + warp < g.nbx) {
        // ---- quadrant bytes of the current block: lane -> row (lane >> 1), 8-pixel half (lane & 1)
        const int b = warp, bx = bx0 + b;
        const uint2 w = *reinterpret_cast<const uint2*>(&s_cur[b][(lane >> 1) * 16 + (lane & 1) * 8]);
        uint32_t s = __dp4a(w.x, 0x01010101u, 0u);
        s = __dp4a(w.y, 0x01010101u, s);
        s += __shfl_xor_sync(0xFFFFFFFFu, s, 2);
        s += __shfl_xor_sync(0xFFFFFFFFu, s, 4);
        s += __shfl_xor_sync(0xFFFFFFFFu, s, 8);             // lanes 0 / 1: top left / right, lanes 16 / 17: bottom left / right
        const uint32_t q0 = __shfl_sync(0xFFFFFFFFu, s, 0) >> 6, q1 = __shfl_sync(0xFFFFFFFFu, s, 1) >> 6;
        const uint32_t q2 = __shfl_sync(0xFFFFFFFFu, s, 16) >> 6, q3 = __shfl_sync(0xFFFFFFFFu, s, 17) >> 6;
        if (lane == 0) s_cq[b] = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
        // ---- predictors -> first threshold
        int n = 0;
        if (lane == 0) {
            int xlo, xhi, ylo, yhi;
            valid_range(bx * 16, g.W, 16, g.fme, g.fme, xlo, xhi);
            valid_range(by * 16, g.H, 16, g.fme, g.fme, ylo, yhi);
            xlo = max(xlo, -g.R); xhi = min(xhi, g.R);
            ylo = max(ylo, -g.R); yhi = min(yhi, g.R);
            auto push = [&](int ref, int dx, int dy) {
                if (ref < g.nref && dx >= xlo && dx <= xhi && dy >= ylo && dy <= yhi)
                    s_q[warp][n++] = (uint32_t)(dx + g.R) | ((uint32_t)(dy + g.R) << 8) | ((uint32_t)ref << 16) | ((uint32_t)b << 20);
            };
            push(0, 0, 0);
            const uint32_t pv = a.prev[(size_t)(a.unit0 + u) * a.nblk + by * g.nbx + bx];
            if (pv != 0xFFFFFFFFu) {
                const int ref = (int)((pv >> 16) & 0xFFu), dx = (int)((pv >> 8) & 0xFFu) - g.R, dy = (int)(pv & 0xFFu) - g.R;
                push(ref, dx, dy);
                if (ref > 0) push(0, dx / (ref + 1), dy / (ref + 1));           // the same motion seen from the newest reference
                else if (g.nref > 1) push(1, max(-g.R, min(g.R, 2 * dx)), max(-g.R, min(g.R, 2 * dy)));
            }
        }
        n = __shfl_sync(0xFFFFFFFFu, n, 0);
        __syncwarp();
        evals += sea_eval(a, ring_u, bx0, by, s_cur, s_thr, s_key, s_q[warp], 0, 2, n, lane);
    }
    __syncthreads();
    // ================================= phase A: filter =================================
    const int mul = g.fme ? 2 : 1;
    const int npairs = g.nref * a.nph * SEA_NB;
    const int quad = lane % 9, rsub = lane / 9;              // lanes 27..31 idle
    const int ox0 = 4 * quad - 16;
    const uint32_t* pq_u = a.pq + (size_t)(a.unit0 + u) * a.pq_unit_stride;
    const int off0 = (rsub - 16) * g.W + ox0, rowstep = 3 * g.W;
    const int nwarps = (int)(blockDim.x >> 5);             // 8, or 4 when there is one (reference, phase plane) only
    for (int pair = warp; pair < npairs; pair += nwarps) {
        const int item = pair / SEA_NB, b = pair - item * SEA_NB, bx = bx0 + b;
        if (bx >= g.nbx) continue;
        const int ref = item / a.nph, ph = item - ref * a.nph;
        const int px = ph & 1, py = ph >> 1;                    // nph == 1: ph = 0
        const int lo = s_xlo[b][px], hi = s_xhi[b][px], ylo = s_ylo[py], yhi = s_yhi[py];
        if (lo > hi || ylo > yhi) continue;                     // no valid candidate on this plane
        const int slot = (int)((a.slot_packed >> (4 * ref)) & 15u);
        // window origin of the triple (offset (0, 0): always inside the plane); lane offsets are 32-bit word offsets from it
        const uint32_t* pb = pq_u + slot * a.pq_slot_stride + (size_t)ph * a.pq_plane_stride + (size_t)(by * 16) * g.W + bx * 16;
        const bool lane_ok = lane < 27 && ox0 + 3 >= lo && ox0 <= hi;
        // steps k whose row 3k + rsub - 16 is valid: k in [klo, khi]
        const int nlo = ylo + 16 - rsub, nhi = yhi + 16 - rsub;
        const int klo = nlo <= 0 ? 0 : (nlo + 2) / 3, khi = nhi < 0 ? -1 : min(10, nhi / 3);
        const uint32_t kmask = (lane_ok && klo <= khi) ? ((2u << khi) - (1u << klo)) : 0u;
        const uint32_t cq = s_cq[b];
        const uint32_t tag = ((uint32_t)ref << 16) | ((uint32_t)b << 20);
#pragma unroll
        for (int c0 = 0; c0 < 11; c0 += SEA_CHUNK) {
            // the loads of the steps are independent of each other: issue them first.  Step kk covers rows 3k .. 3k + 2 with
            // k = 5, 4, 6, 3, 7, ... (centre rows first); lanes without a valid row read the window origin and are masked
            uint4 v[SEA_CHUNK];
#pragma unroll
            for (int t = 0; t < SEA_CHUNK; ++t) {
                const int kk = c0 + t;
                if (kk < 11) {
                    const int k = 5 + ((kk & 1) ? -((kk + 1) >> 1) : (kk >> 1));
                    const int off = (kmask >> k) & 1u ? off0 + k * rowstep : 0;
                    v[t] = __ldg(reinterpret_cast<const uint4*>(pb + off));
                }
            }
            const uint32_t T = *reinterpret_cast<volatile uint32_t*>(&s_thr[b]);
#pragma unroll
            for (int t = 0; t < SEA_CHUNK; ++t) {
                const int kk = c0 + t;
                if (kk >= 11) continue;
                const int k = 5 + ((kk & 1) ? -((kk + 1) >> 1) : (kk >> 1));
                const uint32_t d0 = sad4_acc(v[t].x, cq, 0u), d1 = sad4_acc(v[t].y, cq, 0u), d2 = sad4_acc(v[t].z, cq, 0u), d3 = sad4_acc(v[t].w, cq, 0u);
                const bool rowok = (kmask >> k) & 1u;
                if (!__any_sync(0xFFFFFFFFu, rowok && min(min(d0, d1), min(d2, d3)) <= T)) continue;
                // ---- some lane has a candidate within the bound
                const int oy = 3 * k + rsub - 16;
                uint32_t m = 0;
                if (rowok)
                    m = (d0 <= T && ox0 >= lo && ox0 <= hi ? 1u : 0u) | (d1 <= T && ox0 + 1 >= lo && ox0 + 1 <= hi ? 2u : 0u) |
                        (d2 <= T && ox0 + 2 >= lo && ox0 + 2 <= hi ? 4u : 0u) | (d3 <= T && ox0 + 3 >= lo && ox0 + 3 <= hi ? 8u : 0u);
                if (m) {
                    const uint32_t base_e = (uint32_t)(mul * ox0 + px + g.R) | ((uint32_t)(mul * oy + py + g.R) << 8) | tag;
                    const uint32_t dd[4] = {d0, d1, d2, d3};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if ((m >> i) & 1u) {
                            const uint32_t e = (base_e + (uint32_t)(mul * i)) | (dd[i] << 22);
                            const unsigned int pos = atomicAdd(&s_n, 1u);
                            if (pos < (unsigned)SEA_LCAP) s_list[pos] = e;
                            else evals += sea_eval_one(a, ring_u, bx0, by, s_cur, s_thr, s_key, e);      // list full (no usable bound yet)
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    // ================================= phase B: exact SADs of the listed survivors =================================
    {
        const int n = (int)min(s_n, (unsigned)SEA_LCAP);
        evals += sea_eval_thread(a, ring_u, bx0, by, s_cur, s_thr, s_key, s_list, tid, (int)blockDim.x, n);
    }
    evals = __reduce_add_sync(0xFFFFFFFFu, evals);
    if (lane == 0 && evals) atomicAdd(&s_evals, evals);
    __syncthreads();
    if (tid < SEA_NB && bx0 + tid < g.nbx) {
        const unsigned long long key = s_key[tid];
        const int blk = by * g.nbx + bx0 + tid;
        atomicMin(reinterpret_cast<unsigned long long*>(reinterpret_cast<MeResult*>(a.out) + u * a.out_unit_stride + blk), key);
        a.prev[(size_t)(a.unit0 + u) * a.nblk + blk] = key == ~0ull ? 0xFFFFFFFFu : (uint32_t)(key & 0xFFFFFFull);
    }
    if (tid == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(a.ctr), (unsigned long long)s_evals);
        if (blockIdx.x == 0 && blockIdx.y == 0) atomicAdd(&a.ctr[2], 1u);
        if (a.host_stat) {
            // the last CTA of the launch reports how many exact SADs the launch took: the host reads it without synchronising
            // (a few frames late) and decides whether pruning pays on this content (SO_FLAG_SEA_AUTO)
            atomicAdd(&a.ctr[7], s_evals);
            __threadfence();
            if (atomicAdd(&a.ctr[6], 1u) == gridDim.x * gridDim.y - 1u) {
                const unsigned int v = atomicExch(&a.ctr[7], 0u);
                a.ctr[6] = 0u;
                *reinterpret_cast<volatile unsigned int*>(a.host_stat) = v;
                __threadfence_system();
                *reinterpret_cast<volatile unsigned int*>(a.host_stat + 1) = a.seq;
            }
        }
    }
}

// SO_FLAG_SEA_AUTO while the plain kernel runs: keep the pruned search's predictors fresh (winners of the frame, read from the
// compact keys of me_ring2_kernel -- me_get format 2) so that the next probe does not start cold.
__global__ void __launch_bounds__(256) sea_save_kernel(const SeaArgs a) {
    pdl_trigger();
    pdl_wait();
    const int blk = blockIdx.x * 256 + threadIdx.x, u = blockIdx.y;
    if (blk >= a.nblk) return;
    const unsigned long long key = *reinterpret_cast<const unsigned long long*>(reinterpret_cast<const MeResult*>(a.out) + u * a.out_unit_stride + blk);
    uint32_t pv = 0xFFFFFFFFu;
    if (key != ~0ull) {
        const uint32_t lo = (uint32_t)key;
        const int l1 = (int)((lo >> 16) & 0xFFu), dxr = (int)((lo >> 1) & 0x7Fu), ady = l1 - abs(dxr - a.g.R);
        pv = (((lo >> 8) & 0xFFu) << 16) | ((uint32_t)dxr << 8) | (uint32_t)(((lo & 1u) ? ady : -ady) + a.g.R);
    }
    a.prev[(size_t)(a.unit0 + u) * a.nblk + blk] = pv;
}
