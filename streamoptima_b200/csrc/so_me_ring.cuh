// Exhaustive motion search for 16x16 blocks at r = 16 (BASELINE configs 2-5), item-ring version (sm_100a).
//
// Same arithmetic and the same reference ring as so_me_tma.cuh (four byte-shifted copies of every phase plane, TMA box
// loads straight into shared memory); what changes is the pipeline around the VABSDIFF4 loop:
//   * every CTA owns a CONTIGUOUS range of items (item = (block, reference, phase plane)) and keeps a ring of MR_NS = 22
//     item slots in shared memory.  A slot is refilled as soon as the 44 tasks of its item have finished (per-item
//     mbarriers), so up to 22 items are in flight instead of two 8-item stages;
//   * tasks are ordered [item][vertical group][byte shift]; a bundle = 32 consecutive tasks (it spans at most two items),
//     fetched from a CTA-wide counter by 15 search warps (one producer warp issues the TMA loads);
//   * all four shifts of an (item, group) sit in adjacent lanes.  The 33rd horizontal offset (ox = +16, shift 0 only) is
//     no longer a separate warp-divergent pass: its 16 rows are split over the four lanes (4 rows each, partial SADs
//     combined with two shuffles), so every lane of every bundle runs the same code;
//   * per-bundle bookkeeping is cut down: no divisions by run-time values, REDUX-based argmin merge per item, one
//     global atomicMin per (bundle, item).
// Bank conflicts: the four shift planes of an item are 2400 B apart (6 bank groups of 16 B mod 8), consecutive vertical
// groups are 144 B apart (1 mod 8) and the box of an item in an odd slot is loaded one row above the window (3 mod 8):
// the eight lanes of a quarter warp (two (item, group) pairs x four shifts) hit eight distinct 16-byte bank groups.
#pragma once
#include "so_me_tma.cuh"

constexpr int MR_NS = 22;                       // item slots (even: the slot parity alternates across the wrap)
constexpr int MR_NG = 11;                       // vertical groups of 3 offsets: 33 = 2r + 1
constexpr int MR_TPI = 4 * MR_NG;               // tasks per item
constexpr int MR_WP = 48;                       // window row pitch (bytes): 12 words = offsets -16..16 plus 15 block pixels
constexpr int MR_BOXROWS = 50;                  // window rows + row offset (0 | 1); 50 * 48 B = 150 x 16 B = 6 (mod 8) bank groups
constexpr int MR_PLANE = MR_BOXROWS * MR_WP;    // the four shift planes of an item come in ONE box of depth 4 (consecutive z)
constexpr int MR_SLOT = 4 * MR_PLANE;           // 9600 B: a multiple of 128 (TMA destination alignment)
static_assert(MR_SLOT % 128 == 0 && (MR_PLANE / 16) % 8 == 6, "slot alignment / bank-group stride of the shift planes");
constexpr int MR_CUR = 256;
constexpr int MR_CHUNK = 8;                     // items per work chunk handed out by the device-wide counter
constexpr int MR_SMEM = MR_NS * (MR_SLOT + MR_CUR) + 1024;

struct MeRingArgs {
    FrameGeom g;                 // g.bs == 16, g.r == 16
    unsigned long long* out;     // packed keys, stride of MeResult (16 B) per block
    size_t out_unit_stride;      // in MeResult elements
    unsigned long long* out_sub; // QUAD: keys of the four 8x8 sub-blocks, grid (2nby x 2nbx)
    size_t out_sub_unit_stride;
    int units;
    int nph;                     // 4 (fme) or 1
    int items_per_unit;          // blocks * nref * nph
    int z_per_unit;              // planes per unit in the ring tensor = nslots * 16
    int z_unit0;                 // plane offset of unit 0 of this launch
    unsigned int slot_packed;    // list index -> ring slot, 4 bits each
    unsigned int* work;          // [0] device-wide chunk counter, [1] finished CTAs; both zero at launch, reset by the last CTA
};

__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, uint32_t cnt) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(cnt) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

template <bool QUAD>
__global__ void __launch_bounds__(QUAD ? 384 : 512, 1) me_ring_kernel(const __grid_constant__ CUtensorMap ring_map,
                                                                      const __grid_constant__ CUtensorMap cur_map, const MeRingArgs a) {
    constexpr int BS = 16, WPR = 4, G = 3;
    extern __shared__ __align__(1024) unsigned char smem_r[];
    unsigned char* const wins = smem_r;                                        // [NS][4][MR_PLANE]
    unsigned char* const curs = wins + MR_NS * MR_SLOT;                        // [NS][256]
    uint64_t* const ready = reinterpret_cast<uint64_t*>(curs + MR_NS * MR_CUR);
    uint64_t* const empty = ready + MR_NS;
    int4* const meta = reinterpret_cast<int4*>(empty + MR_NS);                  // [NS]: {out index, bx, by, ref | ph << 8 | interior << 16}
    unsigned int* const counter = reinterpret_cast<unsigned int*>(meta + MR_NS);
    volatile int* const issued = reinterpret_cast<volatile int*>(counter + 1);  // items whose loads have been issued
    volatile int* const final_cnt = issued + 1;                                 // number of items of this CTA, once known

    const FrameGeom& g = a.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = (int)(blockDim.x >> 5);
    if (tid == 0) {
        for (int s = 0; s < MR_NS; ++s) { mbar_init(&ready[s], 1); mbar_init(&empty[s], MR_TPI); }
        *counter = 0;
        *issued = 0;
        *final_cnt = 0x7FFFFFFF;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_trigger();
    __syncthreads();
    pdl_wait();          // everything below reads what earlier kernels of the stream wrote (ring planes, keys, the work counter)

    const long long total = (long long)a.units * a.items_per_unit;
    const int nchunks = (int)((total + MR_CHUNK - 1) / MR_CHUNK);
    const int per_blk = g.nref * a.nph;
    const int mul = g.fme ? 2 : 1;

    if (warp == nwarps - 1) {
        // ================================= producer =================================
        // work is handed out in chunks of MR_CHUNK consecutive items from a device-wide counter: all CTAs stay in the same
        // neighbourhood of the frame (window rows are re-read from L2, not DRAM) and the tail balances itself
        int q = 0, qn = 0;
        if (lane == 0) q = (int)atomicAdd(a.work, 1u);
        q = __shfl_sync(0xFFFFFFFFu, q, 0);
        int slot = 0, n = 0;
        uint32_t par = 1;                       // parity of (use - 1) for the wait on `empty`
        bool first_round = true;
        while (q < nchunks) {
            if (lane == 0) qn = (int)atomicAdd(a.work, 1u);             // next chunk: the latency hides behind this one
            const long long it0 = (long long)q * MR_CHUNK;
            const int nit = (int)min((long long)MR_CHUNK, total - it0);
            int unit = (int)(it0 / a.items_per_unit);
            int rem = (int)(it0 - (long long)unit * a.items_per_unit);
            int blk = rem / per_blk;
            int rp = rem - blk * per_blk;
            int ref = rp / a.nph, ph = rp - ref * a.nph;
            int by = blk / g.nbx, bx = blk - by * g.nbx;
            for (int k = 0; k < nit; ++k) {
                if (!first_round) {
                    while (!mbar_try(&empty[slot], par)) __nanosleep(40);
                }
                // phase planes are visited in the order 0, 2, 1, 3: two even-px items, then two odd-px ones, so that many bundles
                // hold odd horizontal phases only and skip the 33rd-offset pass (it is invalid there: dx = 33 > R)
                const int phz = a.nph == 4 ? (((ph & 1) << 1) | (ph >> 1)) : ph;
                if (lane == 0) {
                    int l0, h0, l1, h1;
                    valid_range(bx * BS, g.W, BS, g.fme, g.fme, l0, h0);
                    valid_range(by * BS, g.H, BS, g.fme, g.fme, l1, h1);
                    const int interior = (l0 <= -g.R && h0 >= g.R && l1 <= -g.R && h1 >= g.R) ? 1 : 0;   // every offset of the range is valid
                    meta[slot] = make_int4((int)(unit * a.out_unit_stride) + blk, bx, by, ref | (phz << 8) | (interior << 16) | (unit << 17));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the slot was read through the generic proxy
                __syncwarp();
                if (lane == 0) mbar_arrive_expect_tx(&ready[slot], (uint32_t)(4 * MR_BOXROWS * MR_WP + BS * BS));
                __syncwarp();
                if (lane == 0) {
                    const int z = a.z_unit0 + unit * a.z_per_unit + (int)((a.slot_packed >> (4 * ref)) & 15u) * 16 + phz * 4;
                    tma_load_3d(wins + slot * MR_SLOT, &ring_map, &ready[slot], bx * BS - 16, by * BS - 16 - (slot & 1), z);
                } else if (lane == 1) {
                    tma_load_3d(curs + slot * MR_CUR, &cur_map, &ready[slot], bx * BS, by * BS, unit);
                }
                ++n;
                if (lane == 0) *issued = n;
                // next item
                if (++ph == a.nph) {
                    ph = 0;
                    if (++ref == g.nref) {
                        ref = 0; ++blk;
                        if (++bx == g.nbx) { bx = 0; if (++by == g.nby) { by = 0; blk = 0; ++unit; } }
                    }
                }
                if (++slot == MR_NS) { slot = 0; par ^= 1; first_round = false; }
            }
            q = __shfl_sync(0xFFFFFFFFu, qn, 0);
        }
        if (lane == 0) *final_cnt = n;          // no item with local index >= n will ever exist
    } else {
    // ================================= search warps =================================
    unsigned b = 0;
    if (lane == 0) b = atomicAdd(counter, 1u);
    b = __shfl_sync(0xFFFFFFFFu, b, 0);
    while (true) {
        const unsigned t0 = b * 32u;
        const unsigned first = t0 / (unsigned)MR_TPI;
        const unsigned r0 = t0 - first * MR_TPI;
        unsigned rem = r0 + lane;
        const unsigned second = rem >= (unsigned)MR_TPI ? 1u : 0u;
        rem -= second * MR_TPI;
        // does item `first` (and `first + 1` when the bundle runs into it) exist?  `issued` counts the items whose loads
        // have been issued; `final_cnt` is set once the producer has run out of chunks.  Waiting for `issued` also makes
        // the parity test below unambiguous: ready[slot] is then in phase `use` (pending or complete), never an older one.
        bool two = r0 + 31u >= (unsigned)MR_TPI;
        {
            bool exists = true;
            SpinWait sw;
            while (true) {
                const int iss = *issued, fin = *final_cnt;
                if (iss > (int)first) break;
                if (fin <= (int)first) { exists = false; break; }
                sw.pause();
            }
            if (!exists) break;
            while (two) {
                const int iss = *issued, fin = *final_cnt;
                if (iss > (int)first + 1) break;
                if (fin <= (int)first + 1) two = false;
                else sw.pause();
            }
        }
        const bool has = second == 0u || two;
        const int grp = (int)(rem >> 2), c = (int)(rem & 3u);
        unsigned nb = 0;
        if (lane == 0) nb = atomicAdd(counter, 1u);         // next bundle index: consumed at the end of this iteration
        const unsigned use0 = __umulhi(first, 0xBA2E8BA3u) >> 4, slot0 = first - use0 * MR_NS;       // first / 22
        const unsigned slot1 = slot0 + 1 == MR_NS ? 0u : slot0 + 1, use1 = slot1 == 0 ? use0 + 1 : use0;
        mbar_wait(&ready[slot0], use0 & 1u);
        if (two) mbar_wait(&ready[slot1], use1 & 1u);
        const unsigned slot = (second && two) ? slot1 : slot0;
        const int4 mt = meta[slot];
        const int bx = mt.y, by = mt.z;
        const int ph = (mt.w >> 8) & 255;
        const int px = g.fme ? (ph & 1) : 0, py = g.fme ? (ph >> 1) : 0;
        const int p = (int)(slot & 1u);
        const unsigned char* wslot = wins + slot * MR_SLOT;
        const unsigned char* win = wslot + c * MR_PLANE + (p + G * grp) * MR_WP;
        const uint32_t* cb = reinterpret_cast<const uint32_t*>(curs + slot * MR_CUR);
        const int oy0 = -16 + G * grp;

        if constexpr (QUAD) {
            // ---- VBS: the four 8x8 sub-blocks search the same offsets, so their SADs are the quadrant sums of the parent's
            // candidates (Encoder.py:517-536 vs :558).  Two half passes (top / bottom 8 rows), left and right words in
            // separate accumulators; quadrant minima are folded after each half, the parent sums are kept.
            uint32_t par[3][8];
#pragma unroll
            for (int gg = 0; gg < 3; ++gg)
#pragma unroll
                for (int k = 0; k < 8; ++k) par[gg][k] = 0;
            // 33rd horizontal offset first: lane c takes block rows 4c..4c+3 (c = 0, 1: top half; 2, 3: bottom half)
            uint32_t exq[5][3];                      // parent, TL, TR, BL, BR sums of candidate k = 8, complete on every lane
#pragma unroll
            for (int e = 0; e < 5; ++e)
#pragma unroll
                for (int gg = 0; gg < 3; ++gg) exq[e][gg] = 0u;
            if (__any_sync(0xFFFFFFFFu, px == 0)) {  // odd horizontal phases have no 33rd offset (dx = 33 > R): masked below
                uint32_t eL[3] = {0u, 0u, 0u}, eR[3] = {0u, 0u, 0u};
                const unsigned char* w0 = wslot + (p + G * grp + 4 * c) * MR_WP + 32;
                uint4 cr[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) cr[i] = reinterpret_cast<const uint4*>(cb)[4 * c + i];
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const uint4 w = *reinterpret_cast<const uint4*>(w0 + i * MR_WP);
#pragma unroll
                    for (int gg = 0; gg < 3; ++gg) {
                        const int r = i - gg;
                        if (r >= 0 && r < 4) {
                            eL[gg] = sad4_acc(w.x, cr[r].x, eL[gg]); eL[gg] = sad4_acc(w.y, cr[r].y, eL[gg]);
                            eR[gg] = sad4_acc(w.z, cr[r].z, eR[gg]); eR[gg] = sad4_acc(w.w, cr[r].w, eR[gg]);
                        }
                    }
                }
#pragma unroll
                for (int gg = 0; gg < 3; ++gg) {
                    const uint32_t hl = eL[gg] + __shfl_xor_sync(0xFFFFFFFFu, eL[gg], 1);      // my half (top for c < 2)
                    const uint32_t hr = eR[gg] + __shfl_xor_sync(0xFFFFFFFFu, eR[gg], 1);
                    const uint32_t ol = __shfl_xor_sync(0xFFFFFFFFu, hl, 2), orr = __shfl_xor_sync(0xFFFFFFFFu, hr, 2);   // the other half
                    const bool top = c < 2;
                    exq[1][gg] = top ? hl : ol; exq[2][gg] = top ? hr : orr;
                    exq[3][gg] = top ? ol : hl; exq[4][gg] = top ? orr : hr;
                    exq[0][gg] = hl + hr + ol + orr;
                }
            }
            uint32_t bq[5] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};   // parent, TL, TR, BL, BR
            const bool fast_valid = __all_sync(0xFFFFFFFFu, (mt.w >> 16) & 1);
            // distance parts of the keys
            uint32_t ly8[3], lx8[9];
#pragma unroll
            for (int gg = 0; gg < 3; ++gg) ly8[gg] = (uint32_t)(abs(mul * (oy0 + gg) + py) << 8) + gg;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int dx = mul * (-16 + c + 4 * k) + px;                 // k < 4: negative, k >= 4: non-negative
                lx8[k] = ((uint32_t)(k < 4 ? -dx : dx) << 8) + k * 3;
            }
            // interior block: parent and sub-blocks share the two special cases (ox = 16 on odd horizontal phases, oy = 16
            // on odd vertical ones); keys are folded rows-first like in the plain search (3 IMAD + min3 + add per column)
            const uint32_t xbl = (c == 0 && px == 0) ? 0u : 0xFFFFFFFFu;
            const uint32_t ybl = (grp == MR_NG - 1 && py) ? 0xFFFFFFFFu : 0u;
            auto fold = [&](uint32_t s0, uint32_t s1, uint32_t s2, int k) {
                const uint32_t t0 = s0 * 65536u + ly8[0], t1 = s1 * 65536u + ly8[1], t2 = (s2 * 65536u + ly8[2]) | ybl;
                uint32_t v = min(min(t0, t1), t2) + lx8[k];
                if (k == 8) v |= xbl;
                return v;
            };
            // edge block: parent and sub-blocks have their own rectangles (Encoder.py:695-698 with their own size / position)
            auto xbad = [&](int k, int pos, int n) {
                int l, h;
                valid_range(pos, g.W, n, g.fme, g.fme, l, h);
                const int dx = mul * (-16 + c + 4 * k) + px;
                return ((k < 8 || c == 0) && dx >= -g.R && dx <= g.R && dx >= l && dx <= h) ? 0u : 0xFFFFFFFFu;
            };
            auto ybad = [&](int gg, int pos, int n) {
                int l, h;
                valid_range(pos, g.H, n, g.fme, g.fme, l, h);
                const int dy = mul * (oy0 + gg) + py;
                return (dy >= -g.R && dy <= g.R && dy >= l && dy <= h) ? 0u : 0xFFFFFFFFu;
            };
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {            // rolled: one copy of the SAD pass in the instruction stream
                uint32_t aL[3][8], aR[3][8];
#pragma unroll
                for (int gg = 0; gg < 3; ++gg)
#pragma unroll
                    for (int k = 0; k < 8; ++k) { aL[gg][k] = 0; aR[gg][k] = 0; }
                sad_pass_g3<WPR, 8, 8, BS / 2, MR_WP, true>(win + half * (BS / 2) * MR_WP, cb + half * (BS / 2) * WPR, aL, aR);
                uint32_t eL[3], eR[3];
#pragma unroll
                for (int gg = 0; gg < 3; ++gg) { eL[gg] = half ? exq[3][gg] : exq[1][gg]; eR[gg] = half ? exq[4][gg] : exq[2][gg]; }
                uint32_t bl = 0xFFFFFFFFu, br = 0xFFFFFFFFu;
                if (fast_valid) {
                    bl = fold(eL[0], eL[1], eL[2], 8);
                    br = fold(eR[0], eR[1], eR[2], 8);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        bl = min(bl, fold(aL[0][k], aL[1][k], aL[2][k], k));
                        br = min(br, fold(aR[0][k], aR[1][k], aR[2][k], k));
                    }
                } else {
                    uint32_t yb[3];
#pragma unroll
                    for (int gg = 0; gg < 3; ++gg) yb[gg] = ybad(gg, by * BS + half * (BS / 2), BS / 2);
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        const uint32_t xl = xbad(k, bx * BS, BS / 2), xr = xbad(k, bx * BS + BS / 2, BS / 2);
#pragma unroll
                        for (int gg = 0; gg < 3; ++gg) {
                            const uint32_t l1v = lx8[k] + ly8[gg];
                            const uint32_t sl = k < 8 ? aL[gg][k < 8 ? k : 0] : eL[gg], sr = k < 8 ? aR[gg][k < 8 ? k : 0] : eR[gg];
                            bl = min(bl, (sl * 65536u + l1v) | xl | yb[gg]);
                            br = min(br, (sr * 65536u + l1v) | xr | yb[gg]);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k)
#pragma unroll
                    for (int gg = 0; gg < 3; ++gg) par[gg][k] += aL[gg][k] + aR[gg][k];
                if (half == 0) { bq[1] = bl; bq[2] = br; } else { bq[3] = bl; bq[4] = br; }
            }
            if (fast_valid) {
                uint32_t bp = fold(exq[0][0], exq[0][1], exq[0][2], 8);
#pragma unroll
                for (int k = 0; k < 8; ++k) bp = min(bp, fold(par[0][k], par[1][k], par[2][k], k));
                bq[0] = bp;
            } else {
                uint32_t yb[3];
#pragma unroll
                for (int gg = 0; gg < 3; ++gg) yb[gg] = ybad(gg, by * BS, BS);
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const uint32_t xp = xbad(k, bx * BS, BS);
#pragma unroll
                    for (int gg = 0; gg < 3; ++gg) {
                        const uint32_t sp = k < 8 ? par[gg][k < 8 ? k : 0] : exq[0][gg];
                        bq[0] = min(bq[0], (sp * 65536u + (lx8[k] + ly8[gg])) | xp | yb[gg]);
                    }
                }
            }
            // ---- merge: five keys per segment
#pragma unroll
            for (int e = 0; e < 5; ++e) {
                const uint32_t best = bq[e];
                const int idx = (int)(best & 0xFFu), k = (idx * 11) >> 5, gg = idx - 3 * k;
                const int dx = mul * (-16 + c + 4 * k) + px, dy = mul * (oy0 + gg) + py;
                const uint32_t xy = ((uint32_t)(dx + g.R) << 8) | (uint32_t)(dy + g.R);
                const uint32_t v1 = (has && best != 0xFFFFFFFFu) ? (best >> 8) : 0xFFFFFFFFu;
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    if (s2 == 1 && !two) break;
                    const bool mine = second == (unsigned)s2;
                    const uint32_t m1 = __reduce_min_sync(0xFFFFFFFFu, mine ? v1 : 0xFFFFFFFFu);
                    const uint32_t m2 = __reduce_min_sync(0xFFFFFFFFu, (mine && v1 == m1) ? xy : 0xFFFFFFFFu);
                    if (lane == 0 && m1 != 0xFFFFFFFFu) {
                        const int4 ms = meta[s2 ? slot1 : slot0];
                        const unsigned long long key = ((unsigned long long)(m1 >> 8) << 40) | ((unsigned long long)(m1 & 0xFFu) << 24) |
                                                       ((unsigned long long)(ms.w & 255) << 16) | (unsigned long long)m2;
                        unsigned long long* okey;
                        if (e == 0) okey = reinterpret_cast<unsigned long long*>(reinterpret_cast<MeResult*>(a.out) + ms.x);
                        else {
                            const int kx = (e - 1) & 1, ky = (e - 1) >> 1, un = ms.w >> 17;
                            okey = reinterpret_cast<unsigned long long*>(reinterpret_cast<MeResult*>(a.out_sub) + un * a.out_sub_unit_stride +
                                                                         (size_t)(ms.z * 2 + ky) * (g.nbx * 2) + ms.y * 2 + kx);
                        }
                        atomicMin(okey, key);
                    }
                }
            }
        } else {
        // ---- main pass: 8 horizontal x 3 vertical offsets
        uint32_t acc[3][8];
#pragma unroll
        for (int gg = 0; gg < 3; ++gg)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[gg][k] = 0;
        {
            uint32_t unused[3][8];
            sad_pass_g3<WPR, 8, 8, BS, MR_WP, false>(win, cb, acc, unused);
        }
        // ---- 33rd horizontal offset (ox = +16: words 8..11 of the shift-0 copy): lane c of the four takes block rows 4c..4c+3
        uint32_t ex[3] = {0u, 0u, 0u};
        if (__any_sync(0xFFFFFFFFu, px == 0)) {
            const unsigned char* w0 = wslot + (p + G * grp + 4 * c) * MR_WP + 32;
            uint4 cr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) cr[i] = reinterpret_cast<const uint4*>(cb)[4 * c + i];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const uint4 w = *reinterpret_cast<const uint4*>(w0 + i * MR_WP);
#pragma unroll
                for (int gg = 0; gg < 3; ++gg) {
                    const int r = i - gg;
                    if (r >= 0 && r < 4) {
                        ex[gg] = sad4_acc(w.x, cr[r].x, ex[gg]); ex[gg] = sad4_acc(w.y, cr[r].y, ex[gg]);
                        ex[gg] = sad4_acc(w.z, cr[r].z, ex[gg]); ex[gg] = sad4_acc(w.w, cr[r].w, ex[gg]);
                    }
                }
            }
#pragma unroll
            for (int gg = 0; gg < 3; ++gg) {
                ex[gg] += __shfl_xor_sync(0xFFFFFFFFu, ex[gg], 1);
                ex[gg] += __shfl_xor_sync(0xFFFFFFFFu, ex[gg], 2);
            }
        }

        // ---- thread-local argmin.  key32 = sad << 16 | (|dx| + |dy|) << 8 | (k * 3 + g); invalid candidates are OR-ed to all ones.
        uint32_t best = 0xFFFFFFFFu;
        const bool fast_valid = __all_sync(0xFFFFFFFFu, (mt.w >> 16) & 1);
        if (fast_valid) {
            // interior block: the only invalid candidates are ox = 16 on odd horizontal phases and oy = 16 on odd vertical ones
            const uint32_t xbl = (c == 0 && px == 0) ? 0u : 0xFFFFFFFFu;          // candidate k = 8
            const uint32_t ybl = (grp == MR_NG - 1 && py) ? 0xFFFFFFFFu : 0u;     // candidate g = 2 of the last group
            uint32_t ly8[3];
#pragma unroll
            for (int gg = 0; gg < 3; ++gg) ly8[gg] = (uint32_t)(abs(mul * (oy0 + gg) + py) << 8) + gg;
            // min over the three vertical offsets first (their keys differ by SAD and ly8 only), then add the horizontal part:
            // 3 IMAD (FMA pipe) + one 3-input min + one add per column instead of 3 adds + 3 mins on the ALU pipe
            uint32_t kk[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int dx = mul * (-16 + c + 4 * k) + px;                 // k < 4: negative, k >= 4: non-negative (c <= 3)
                const uint32_t lx8 = ((uint32_t)(k < 4 ? -dx : dx) << 8) + k * 3;
                const uint32_t t0 = (k < 8 ? acc[0][k < 8 ? k : 0] : ex[0]) * 65536u + ly8[0];
                const uint32_t t1 = (k < 8 ? acc[1][k < 8 ? k : 0] : ex[1]) * 65536u + ly8[1];
                const uint32_t t2 = ((k < 8 ? acc[2][k < 8 ? k : 0] : ex[2]) * 65536u + ly8[2]) | ybl;
                kk[k] = min(min(t0, t1), t2) + lx8;
            }
            kk[8] |= xbl;
            best = min(min(min(kk[0], kk[1]), min(kk[2], kk[3])), min(min(kk[4], kk[5]), min(kk[6], min(kk[7], kk[8]))));
        } else {
            int xlo, xhi, ylo, yhi;
            valid_range(bx * BS, g.W, BS, g.fme, g.fme, xlo, xhi);
            valid_range(by * BS, g.H, BS, g.fme, g.fme, ylo, yhi);
            xlo = max(xlo, -g.R); xhi = min(xhi, g.R);
            ylo = max(ylo, -g.R); yhi = min(yhi, g.R);
            uint32_t ly8[3], ybad[3];
#pragma unroll
            for (int gg = 0; gg < 3; ++gg) {
                const int dy = mul * (oy0 + gg) + py;
                ly8[gg] = (uint32_t)(abs(dy) << 8) + gg;
                ybad[gg] = (dy >= ylo && dy <= yhi) ? 0u : 0xFFFFFFFFu;
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int dx = mul * (-16 + c + 4 * k) + px;
                const uint32_t lx8 = (uint32_t)(abs(dx) << 8) + k * 3;
                const uint32_t xbad = ((k < 8 || c == 0) && dx >= xlo && dx <= xhi) ? 0u : 0xFFFFFFFFu;
#pragma unroll
                for (int gg = 0; gg < 3; ++gg) {
                    const uint32_t key = ((k < 8 ? acc[gg][k < 8 ? k : 0] : ex[gg]) * 65536u + (lx8 + ly8[gg])) | xbad | ybad[gg];
                    best = min(best, key);
                }
            }
        }
        // ---- merge per item: order (SAD, |dx|+|dy|, ref, dx, dy); all lanes of a segment share ref, so two REDUX steps
        //      ((SAD, L1), then (dx, dy) among the lanes that tie) give the winner of the segment
        uint32_t xy;
        {
            const int idx = (int)(best & 0xFFu), k = (idx * 11) >> 5, gg = idx - 3 * k;
            const int dx = mul * (-16 + c + 4 * k) + px, dy = mul * (oy0 + gg) + py;
            xy = ((uint32_t)(dx + g.R) << 8) | (uint32_t)(dy + g.R);
        }
        const uint32_t v1 = (has && best != 0xFFFFFFFFu) ? (best >> 8) : 0xFFFFFFFFu;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (s == 1 && !two) break;
            const bool mine = second == (unsigned)s;
            const uint32_t m1 = __reduce_min_sync(0xFFFFFFFFu, mine ? v1 : 0xFFFFFFFFu);
            const uint32_t m2 = __reduce_min_sync(0xFFFFFFFFu, (mine && v1 == m1) ? xy : 0xFFFFFFFFu);
            if (lane == 0 && m1 != 0xFFFFFFFFu) {
                const int4 ms = meta[s ? slot1 : slot0];
                const unsigned long long key = ((unsigned long long)(m1 >> 8) << 40) | ((unsigned long long)(m1 & 0xFFu) << 24) |
                                               ((unsigned long long)(ms.w & 255) << 16) | (unsigned long long)m2;
                atomicMin(reinterpret_cast<unsigned long long*>(reinterpret_cast<MeResult*>(a.out) + ms.x), key);
            }
        }
        }   // !QUAD
        __syncwarp();
        if (lane == 0) {
            const uint32_t c0 = min(32u, (unsigned)MR_TPI - r0);
            mbar_arrive_cnt(&empty[slot0], c0);
            if (two) mbar_arrive_cnt(&empty[slot1], 32u - c0);
        }
        b = __shfl_sync(0xFFFFFFFFu, nb, 0);
    }
    }   // search warps
    // the last CTA to get here resets the device-wide counters for the next launch (every producer has stopped fetching)
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(a.work + 1, 1u);
        if (prev == gridDim.x - 1) { a.work[0] = 0u; a.work[1] = 0u; }
    }
}
