// Integer / half-pel exhaustive motion search (find_best_match, Encoder.py:678-717) for sm_100a.
//
// Work decomposition
//   item  = (block, reference frame, phase plane)         one search window of (bs+2r)^2 bytes
//   task  = (item, byte shift c in 0..3, group of G consecutive vertical offsets)
//   a task accumulates NDX x G candidates: horizontal offsets ox = -r + c + 4k (k < NDX), vertical oy0 + g.
// The window of an item is held in shared memory as FOUR copies shifted by 0..3 bytes so that every candidate of a
// task reads 32-bit-aligned words: one 128-bit LDS brings 4 words that feed up to 4 candidates x 4 words of
// VABSDIFF4.U8.ACC (4 pixels per instruction, measured 64 lanes/clk/SM on B200 -- tools/int_peak.cu).
// Candidates are compared with the reference's replace rule through a packed key
//   (SAD, |dx|+|dy|, ref, dx, dy)   -- lexicographic minimum == sequential scan of Encoder.py:688-715 (appendix A4).
#pragma once
#include "so_common.cuh"

struct RefRing {
    const uint8_t* base;
    size_t unit_stride, slot_stride, plane_stride;
    int slot[SO_MAX_REF];      // list index -> ring slot
    __device__ __host__ const uint8_t* plane(int unit, int idx, int ph) const {
        return base + unit * unit_stride + slot[idx] * slot_stride + ph * plane_stride;
    }
};

struct MeFullArgs {
    FrameGeom g;               // g.bs is the block size searched by this launch (sub-block size for VBS passes)
    RefRing ring;
    const uint8_t* cur;        // current frame of unit 0, dense [H][W]
    size_t cur_unit_stride;
    MeResult* out;             // [unit][nby*nbx] for this launch's block grid
    size_t out_unit_stride;    // in elements
    int nph;                   // phase planes searched per reference: 4 (fme) or 1
    int items_per_unit;        // nblk * nref * nph
    int WI;                    // items per CTA
    int NG;                    // vertical groups per (item, shift): ceil((2r+1)/G)
    int rows;                  // window rows = bs + 2r
    int wpitch;                // window row pitch in bytes (16 * odd)
    int copy_stride;           // bytes between the 4 shifted copies of an item
    int item_stride;           // bytes between items
    int tma;                   // 1: windows are fetched with TMA (cp.async.bulk.tensor), 0: SIMT loader
};

// shared memory carve-up (dynamic):  [keys: WI x u64][cur: WI x bs*bs bytes][windows]
template <int BS, int NDX, int G>
__global__ void __launch_bounds__(384, 1) me_full_kernel(const MeFullArgs a) {
    constexpr int WPR = BS / 4;                 // 32-bit words per block row
    constexpr int NW = NDX + WPR - 1;           // words of a window row a task touches
    constexpr int NV = (NW + 3) / 4;            // 128-bit loads per window row
    extern __shared__ __align__(128) unsigned char smem[];
    const FrameGeom& g = a.g;
    const int unit = blockIdx.y;
    const int item0 = blockIdx.x * a.WI;
    const int nitems = min(a.WI, a.items_per_unit - item0);
    const int per_blk = g.nref * a.nph;
    const int blk0 = item0 / per_blk;
    const int blk_last = (item0 + nitems - 1) / per_blk;
    const int nblk_local = blk_last - blk0 + 1;

    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);
    uint32_t* curs = reinterpret_cast<uint32_t*>(smem + a.WI * 8);
    unsigned char* wins = smem + a.WI * 8 + a.WI * BS * BS;

    const uint8_t* cur = a.cur + unit * a.cur_unit_stride;

    // ---- stage current blocks and reset keys
    for (int i = threadIdx.x; i < nblk_local; i += blockDim.x) keys[i] = ~0ull;
    for (int i = threadIdx.x; i < nblk_local * BS * WPR; i += blockDim.x) {
        const int lb = i / (BS * WPR), rem = i % (BS * WPR), row = rem / WPR, w = rem % WPR;
        const int blk = blk0 + lb;
        const int bx = blk % g.nbx, by = blk / g.nbx;
        curs[i] = *reinterpret_cast<const uint32_t*>(cur + (size_t)(by * BS + row) * g.W + bx * BS + w * 4);
    }
    // ---- stage windows: 4 byte-shifted copies per item (SIMT loader: aligned words + funnel shift, zero fill outside)
    {
        const int wpr = a.wpitch / 4;
        const int words_per_item = 4 * a.rows * wpr;
        for (int i = threadIdx.x; i < nitems * words_per_item; i += blockDim.x) {
            const int li = i / words_per_item;
            int rem = i % words_per_item;
            const int c = rem / (a.rows * wpr);
            rem %= a.rows * wpr;
            const int row = rem / wpr, w = rem % wpr;
            const int item = item0 + li;
            const int blk = item / per_blk, rp = item % per_blk;
            const int ref = rp / a.nph, ph = rp % a.nph;
            const int bx = blk % g.nbx, by = blk / g.nbx;
            const uint8_t* plane = a.ring.plane(unit, ref, ph);
            const int Y = by * BS - g.r + row;
            const int X = bx * BS - g.r + c + 4 * w;          // byte address of this shifted word
            uint32_t v = 0;
            if (Y >= 0 && Y < g.H) {
                const int xa = (X >= 0) ? (X & ~3) : -((-X + 3) & ~3);      // floor to a multiple of 4
                const int sh = X - xa;
                const uint32_t* rowp = reinterpret_cast<const uint32_t*>(plane + (size_t)Y * g.pitch);
                const uint32_t lo = (xa >= 0 && xa < g.W) ? __ldg(rowp + (xa >> 2)) : 0u;
                const uint32_t hi = (sh && xa + 4 >= 0 && xa + 4 < g.W) ? __ldg(rowp + (xa >> 2) + 1) : 0u;
                v = __funnelshift_r(lo, hi, sh * 8);
            }
            *reinterpret_cast<uint32_t*>(wins + li * a.item_stride + c * a.copy_stride + row * a.wpitch + w * 4) = v;
        }
    }
    __syncthreads();

    // ---- SAD tasks
    const int tasks_per_item = 4 * a.NG;
    for (int task = threadIdx.x; task < nitems * tasks_per_item; task += blockDim.x) {
        const int li = task / tasks_per_item;
        const int rem = task % tasks_per_item;
        const int c = rem / a.NG, grp = rem % a.NG;
        const int item = item0 + li;
        const int blk = item / per_blk, rp = item % per_blk;
        const int ref = rp / a.nph, ph = rp % a.nph;
        const int px = ph & 1, py = ph >> 1;
        const int lb = blk - blk0;
        const int oy0 = grp * G;                                   // window row of the first vertical offset (oy = -r + oy0)

        const unsigned char* win = wins + li * a.item_stride + c * a.copy_stride + oy0 * a.wpitch;
        const uint32_t* cb = curs + lb * BS * WPR;

        uint32_t acc[G][NDX];
#pragma unroll
        for (int gg = 0; gg < G; ++gg)
#pragma unroll
            for (int k = 0; k < NDX; ++k) acc[gg][k] = 0;

        uint32_t curq[G][WPR];
#pragma unroll
        for (int rho = 0; rho < BS + G - 1; ++rho) {
            uint32_t refw[NV * 4];
            if (oy0 + rho < a.rows) {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const uint4 q = *reinterpret_cast<const uint4*>(win + rho * a.wpitch + v * 16);
                    refw[4 * v] = q.x; refw[4 * v + 1] = q.y; refw[4 * v + 2] = q.z; refw[4 * v + 3] = q.w;
                }
            } else {
#pragma unroll
                for (int v = 0; v < NV * 4; ++v) refw[v] = 0;
            }
#pragma unroll
            for (int gg = G - 1; gg > 0; --gg)
#pragma unroll
                for (int w = 0; w < WPR; ++w) curq[gg][w] = curq[gg - 1][w];
            if (rho < BS) {
#pragma unroll
                for (int w = 0; w < WPR; ++w) curq[0][w] = cb[rho * WPR + w];
            }
#pragma unroll
            for (int gg = 0; gg < G; ++gg) {
                const int j = rho - gg;
                if (j >= 0 && j < BS) {
#pragma unroll
                    for (int k = 0; k < NDX; ++k)
#pragma unroll
                        for (int w = 0; w < WPR; ++w) acc[gg][k] = sad4_acc(refw[k + w], curq[gg][w], acc[gg][k]);
                }
            }
        }

        // ---- candidate validity + thread-local argmin on (SAD, L1, [ref], dx, dy)
        const int bx = blk % g.nbx, by = blk / g.nbx;
        int xlo, xhi, ylo, yhi;
        valid_range(bx * BS, g.W, BS, g.fme, g.fme, xlo, xhi);
        valid_range(by * BS, g.H, BS, g.fme, g.fme, ylo, yhi);
        xlo = max(xlo, -g.R); xhi = min(xhi, g.R);
        ylo = max(ylo, -g.R); yhi = min(yhi, g.R);
        uint32_t best = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < NDX; ++k) {
            const int ox = -g.r + c + 4 * k;
            const int dx = g.fme ? 2 * ox + px : ox;
            const bool vx = (ox <= g.r) && dx >= xlo && dx <= xhi;
#pragma unroll
            for (int gg = 0; gg < G; ++gg) {
                const int oy = -g.r + oy0 + gg;
                const int dy = g.fme ? 2 * oy + py : oy;
                const bool v = vx && (oy <= g.r) && dy >= ylo && dy <= yhi;
                const uint32_t key = (acc[gg][k] << 16) | (uint32_t)((abs(dx) + abs(dy)) << 8) | (uint32_t)(k * G + gg);
                best = v ? min(best, key) : best;
            }
        }
        if (best != 0xFFFFFFFFu) {
            const int idx = best & 0xFF, k = idx / G, gg = idx % G;
            const int ox = -g.r + c + 4 * k, oy = -g.r + oy0 + gg;
            const int dx = g.fme ? 2 * ox + px : ox;
            const int dy = g.fme ? 2 * oy + py : oy;
            const unsigned long long key = ((unsigned long long)(best >> 16) << 40) |
                                           ((unsigned long long)((best >> 8) & 0xFF) << 24) |
                                           ((unsigned long long)ref << 16) |
                                           ((unsigned long long)(dx + g.R) << 8) | (unsigned long long)(dy + g.R);
            if (key < keys[lb]) atomicMin(&keys[lb], key);
        }
    }
    __syncthreads();

    // ---- write results.  A block whose items straddle two CTAs is merged with a global atomicMin on the packed key.
    for (int i = threadIdx.x; i < nblk_local; i += blockDim.x) {
        const int blk = blk0 + i;
        unsigned long long* o = reinterpret_cast<unsigned long long*>(a.out + unit * a.out_unit_stride + blk);
        const unsigned long long key = keys[i];
        // packed key -> MeResult is done by me_unpack_kernel; here the raw key is min-merged
        atomicMin(o, key);
    }
}

// out[] holds packed keys after me_full_kernel; convert in place to MeResult.
__global__ void me_unpack_kernel(MeResult* out, int n, int R) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = *reinterpret_cast<unsigned long long*>(out + i);
    MeResult r;
    if (key == ~0ull) {
        r.dx = 0; r.dy = 0; r.ref = 0; r.none = 1; r.sad = 0;      // best_mv = (0,0,0), MAE = inf (Encoder.py:684-685)
    } else {
        r.sad = (uint32_t)(key >> 40);
        r.ref = (int16_t)((key >> 16) & 0xFF);
        r.dx = (int16_t)((int)((key >> 8) & 0xFF) - R);
        r.dy = (int16_t)((int)(key & 0xFF) - R);
        r.none = 0;
    }
    out[i] = r;
}
