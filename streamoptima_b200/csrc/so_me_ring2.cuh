// Exhaustive motion search for 16x16 blocks at r = 16 (BASELINE configs 2-5), item-ring kernel, second version (sm_100a).
//
// Same arithmetic, ring planes, TMA boxes and shared-memory layout as so_me_ring.cuh; what changes is how work is cut:
//   * work is handed out in HOMOGENEOUS CHUNKS of 8 items: four consecutive (block, reference) pairs x the two phase planes
//     of ONE horizontal parity.  A chunk of interior blocks is 8 x 44 tasks = exactly 11 bundles, so no bundle mixes even and
//     odd horizontal phases any more: the 33rd-offset pass (dx = +32, even phases only) runs exactly where it is needed;
//   * a chunk only carries the vertical groups that hold a valid offset for at least one of its items (edge block rows:
//     Encoder.py:695-698): tasks per item = 4 x groups, tasks ordered [item][group][shift] as before.  The producer writes,
//     for every bundle, one DESCRIPTOR WORD per lane quad {slot, phase parity of the slot's use, group, valid} into a ring
//     indexed by the bundle number and publishes the bundle count once the chunk's loads are issued: a search warp's
//     header is one shared-memory load -- no cursor, no division, no "issued" check;
//   * the producer issues the (up to) eight items of a chunk lane-parallel: lane k waits for its own slot, writes the item
//     meta and launches both TMA boxes;
//   * a bundle may now cover several items (edge chunks have few tasks per item): the per-item merge uses
//     __match_any_sync + REDUX over the lanes of an item, and each item's lanes arrive on its `empty` barrier together
//     (the producer pre-arrives for the groups a chunk does not carry, so the barrier count stays 44).
#pragma once
#include "so_me_ring.cuh"

#ifndef MR2_PRODUCER_SLEEP
#define MR2_PRODUCER_SLEEP 200          // ns between two polls of a slot's `empty` barrier by the producer
#endif
constexpr int MR2_BD = 64;                      // bundles whose descriptors are kept (ring; at most 55 are live, see the producer)
constexpr int MR2_CTRL = 1024 + MR2_BD * 8 * 16;                 // barriers, item meta, counters | bundle descriptors
constexpr int MR2_SMEM = MR_NS * (MR_SLOT + MR_CUR) + MR2_CTRL + 512;   // + vertical addends

struct MeRing2Args {
    MeRingArgs b;                // geometry, outputs, units, nph, z_*, slot_packed, work counters (items_per_unit unused)
    int npairs;                  // (block, reference) pairs per unit = blocks * nref
    int chunks_per_unit;         // fme: ceil(npairs / 4) * 2 (pair group x horizontal parity); else ceil(npairs / 8)
};

// One 32-bit key per candidate, thread-local argmin and per-item warp merge alike: SAD (16) | |dx|+|dy| (8) | dx + R (7) | dy > 0 (1).
// Given the L1 distance and dx, dy is known up to its sign, so the lexicographic order (SAD, L1, dx, dy) of Encoder.py:771
// survives; the low byte is additive -- 2 dx + (2 R + (dy > 0)) -- so it rides along in the addends of the key fold and the
// winner needs no index decode.
// The 64-bit key merged with atomicMin: the 32-bit key with the reference between its L1 byte and its low byte -- the order
// (SAD, L1, ref, dx, dy) of Encoder.py:771 -- put together with two byte permutes; me_get (so_kernels.cuh, format 2) decodes it
// once per block.  refw: a word whose byte 0 is the reference index.
__device__ __forceinline__ unsigned long long mr2_key64(uint32_t m, uint32_t refw) {
    const uint32_t lo = __byte_perm(m, refw, 0x2140), hi = __byte_perm(m, 0u, 0x4443);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ uint4 mr2_lds_v4(const volatile void* p) {       // one volatile 128-bit shared load (descriptor ring)
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(const_cast<const void*>(p))) : "memory");
    return v;
}
__device__ __forceinline__ void mr2_sts_v4(volatile void* p, const uint4& v) {
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(const_cast<const void*>(p))), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned int mr2_next(unsigned int* counter) {       // one lane: no warp-aggregation code around it
    unsigned int v;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"(smem_u32(counter)) : "memory");
    return v;
}

template <bool QUAD>
__global__ void __launch_bounds__(QUAD ? 384 : 512, 1) me_ring2_kernel(const __grid_constant__ CUtensorMap ring_map,
                                                                       const __grid_constant__ CUtensorMap cur_map, const MeRing2Args a2) {
    const MeRingArgs& a = a2.b;
    constexpr int BS = 16, WPR = 4, G = 3;
    extern __shared__ __align__(1024) unsigned char smem_r[];
    unsigned char* const wins = smem_r;                                        // [NS][4][MR_PLANE]
    unsigned char* const curs = wins + MR_NS * MR_SLOT;                        // [NS][256]
    uint64_t* const ready = reinterpret_cast<uint64_t*>(curs + MR_NS * MR_CUR);
    uint64_t* const empty = ready + MR_NS;
    int4* const meta = reinterpret_cast<int4*>(empty + MR_NS);                  // [NS]: {out index, bx, by, ref | ph << 8 | interior << 16}
    unsigned int* const counter = reinterpret_cast<unsigned int*>(meta + MR_NS);
    volatile int* const bundles_pub = reinterpret_cast<volatile int*>(counter + 1);   // bundles whose descriptors are written AND whose items are issued
    volatile int* const final_bundles = bundles_pub + 1;                        // number of bundles of this CTA, once known
    // [MR2_BD][8] descriptors, one per lane quad of a bundle, everything a search warp needs before the SAD loop ready-made:
    //   x: slot | use parity of the slot << 8 | flags << 16 (1 valid, 2 every offset of the item is valid, 4 odd horizontal phase,
    //      8 odd vertical phase) | vertical group << 24
    //   y, z, w: byte offsets from smem_r of the window row (slot parity + 3 * group) of shift plane 0, of the group's vertical
    //      addends (lytab) and of the current block
    volatile uint4* const bdesc = reinterpret_cast<volatile uint4*>(smem_r + MR_NS * (MR_SLOT + MR_CUR) + 1024);
    // vertical parts of the keys per (group, vertical phase): {|dy| << 8 | 2R + (dy > 0)} of the three offsets of the group and the
    // multiplier of the third SAD (0 with an all-ones addend where oy = 16 does not exist: odd vertical phases, last group)
    constexpr int LYTAB_OFF = MR_NS * (MR_SLOT + MR_CUR) + MR2_CTRL;
    uint4* const lytab = reinterpret_cast<uint4*>(smem_r + LYTAB_OFF);      // [MR_NG][2]

    const FrameGeom& g = a.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = (int)(blockDim.x >> 5);
    if (tid >= 32 && tid < 32 + 2 * MR_NG) {
        const int grp = (tid - 32) >> 1, py = (tid - 32) & 1, m = g.fme ? 2 : 1;
        uint32_t v[3];
#pragma unroll
        for (int gg = 0; gg < 3; ++gg) {
            const int dy = m * (-16 + 3 * grp + gg) + (g.fme ? py : 0);
            v[gg] = ((uint32_t)abs(dy) << 8) + (uint32_t)(2 * g.R) + (dy > 0 ? 1u : 0u);
        }
        const bool ybl = grp == MR_NG - 1 && py && g.fme;
        lytab[tid - 32] = make_uint4(v[0], v[1], ybl ? 0xFFFFFF00u : v[2], ybl ? 0u : 65536u);
    }
    if (tid == 0) {
        for (int s = 0; s < MR_NS; ++s) { mbar_init(&ready[s], 1); mbar_init(&empty[s], MR_TPI); }
        *counter = 0;
        *bundles_pub = 0;
        *final_bundles = 0x7FFFFFFF;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_trigger();
    __syncthreads();
    pdl_wait();          // everything below reads what earlier kernels of the stream wrote (ring planes, keys, the work counter)

    const int cpu = a2.chunks_per_unit;
    const int nchunks = a.units * cpu;
    const int mul = g.fme ? 2 : 1;

    if (warp == nwarps - 1) {
        // ================================= producer =================================
        // chunks are handed out by a device-wide counter: all CTAs stay in the same neighbourhood of the frame (window rows
        // are re-read from L2, not DRAM) and the tail balances itself
        int q = 0, qn = 0;
        if (lane == 0) q = (int)atomicAdd(a.work, 1u);
        q = __shfl_sync(0xFFFFFFFFu, q, 0);
        int slot = 0, n = 0, bundle_base = 0;
        uint32_t upar = 0;                      // parity of the use (n / MR_NS) of `slot`
        while (q < nchunks) {
            if (lane == 0) qn = (int)atomicAdd(a.work, 1u);             // next chunk: the latency hides behind this one
            const int unit = q / cpu, qq = q - unit * cpu;
            int pair0, nit, hpar = 0;
            if (a.nph == 4) { pair0 = (qq >> 1) * 4; hpar = qq & 1; nit = 2 * min(4, a2.npairs - pair0); }
            else { pair0 = qq * 8; nit = min(8, a2.npairs - pair0); }
            // item k of the chunk belongs to lane k < nit
            int i_blk = 0, i_ref = 0, i_ph = 0, i_bx = 0, i_by = 0, i_int = 0, glo = 99, ghi = -1;
            if (lane < nit) {
                const int pair = pair0 + (a.nph == 4 ? (lane >> 1) : lane);
                i_ph = a.nph == 4 ? hpar + 2 * (lane & 1) : 0;          // phase plane = (py << 1) | px
                i_blk = pair / g.nref; i_ref = pair - i_blk * g.nref;
                i_by = i_blk / g.nbx; i_bx = i_blk - i_by * g.nbx;
                int l0, h0, l1, h1;
                valid_range(i_bx * BS, g.W, BS, g.fme, g.fme, l0, h0);
                valid_range(i_by * BS, g.H, BS, g.fme, g.fme, l1, h1);
                i_int = (l0 <= -g.R && h0 >= g.R && l1 <= -g.R && h1 >= g.R) ? 1 : 0;       // every offset of the range is valid
                // vertical groups with a valid offset: dy = mul * oy + py in [max(l1, -R), min(h1, R)], group = (oy + 16) / 3
                const int py = g.fme ? (i_ph >> 1) : 0;
                if constexpr (QUAD) {           // the 8x8 sub-blocks searched in the same pass have wider valid ranges: union of all three
                    int ls, hs;
                    valid_range(i_by * BS, g.H, BS / 2, g.fme, g.fme, ls, hs);
                    l1 = min(l1, ls); h1 = max(h1, hs);
                    valid_range(i_by * BS + BS / 2, g.H, BS / 2, g.fme, g.fme, ls, hs);
                    l1 = min(l1, ls); h1 = max(h1, hs);
                }
                const int dlo = max(l1, -g.R) - py, dhi = min(h1, g.R) - py;
                const int olo = max(-16, dlo >= 0 ? (dlo + mul - 1) / mul : -((-dlo) / mul));     // ceil(dlo / mul)
                const int ohi = min(16, dhi >= 0 ? dhi / mul : -((-dhi + mul - 1) / mul));       // floor(dhi / mul)
                if (olo <= ohi) { glo = (olo + 16) / 3; ghi = (ohi + 16) / 3; }
            }
            glo = __reduce_min_sync(0xFFFFFFFFu, glo);
            ghi = __reduce_max_sync(0xFFFFFFFFu, ghi);
            if (ghi >= glo) {                   // else: no item of the chunk has a valid candidate -- the keys stay all ones
                const int ng = ghi - glo + 1, tpi = 4 * ng, nb = (nit * tpi + 31) >> 5;
                // ---- bundle descriptors: entry e = 8 * (bundle in chunk) + lane quad = the (item, group) pair index of the quad.
                // Ring safety: when these are written every item up to n - 1 has been issued, so every item up to n - 1 - MR_NS has
                // been consumed; the live bundles are those of the last MR_NS items (they lie in at most 4 chunks = 44 bundles) plus
                // the 11 of this chunk: 55 <= MR2_BD.
                const int mg = (1024 + ng - 1) / ng;        // exact division by ng of any e < 96 (ng <= 11): (e * mg) >> 10
                const int i_info = i_ph | (i_int << 8);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int e = lane + 32 * j;
                    int ki = (e * mg) >> 10, gl = e - ki * ng;
                    const bool has = ki < nit;              // quads past the chunk's last task shadow its last item
                    if (!has) { ki = nit - 1; gl = 0; }
                    const int info = __shfl_sync(0xFFFFFFFFu, i_info, ki);       // phase plane and interior flag of the quad's item
                    if (e < nb * 8) {
                        int s = slot + ki;
                        const bool wrap = s >= MR_NS;
                        if (wrap) s -= MR_NS;
                        const int grp = glo + gl, ph = info & 255;
                        const uint32_t flags = (has ? 1u : 0u) | ((info >> 8) ? 2u : 0u) | ((g.fme && (ph & 1)) ? 4u : 0u) | ((g.fme && (ph >> 1)) ? 8u : 0u);
                        uint4 d;
                        d.x = (uint32_t)s | ((upar ^ (wrap ? 1u : 0u)) << 8) | (flags << 16) | ((uint32_t)grp << 24);
                        d.y = (uint32_t)(s * MR_SLOT + ((s & 1) + G * grp) * MR_WP);
                        d.z = (uint32_t)(LYTAB_OFF + (grp * 2 + ((flags >> 3) & 1u)) * 16);
                        d.w = (uint32_t)(MR_NS * MR_SLOT + s * MR_CUR);
                        mr2_sts_v4(&bdesc[(bundle_base * 8 + e) & (MR2_BD * 8 - 1)], d);
                    }
                }
                // ---- the chunk's items, one lane each
                if (lane < nit) {
                    int s = slot + lane;
                    const bool wrap = s >= MR_NS;
                    if (wrap) s -= MR_NS;
                    if (n + lane >= MR_NS) {            // not the first use of the slot: wait until its previous item is consumed
                        const uint32_t par = upar ^ (wrap ? 1u : 0u) ^ 1u;
                        while (!mbar_try(&empty[s], par)) __nanosleep(MR2_PRODUCER_SLEEP);      // 22 items ahead: a slot frees up every ~0.6 us
                    }
                    meta[s] = make_int4((int)(unit * a.out_unit_stride) + i_blk, i_bx, i_by, i_ref | (i_ph << 8) | (i_int << 16) | (unit << 17));
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the slot was read through the generic proxy
                    mbar_arrive_expect_tx(&ready[s], (uint32_t)(4 * MR_BOXROWS * MR_WP + BS * BS));
                    if (tpi < MR_TPI) mbar_arrive_cnt(&empty[s], (uint32_t)(MR_TPI - tpi));     // the groups this chunk does not carry
                    const int z = a.z_unit0 + unit * a.z_per_unit + (int)((a.slot_packed >> (4 * i_ref)) & 15u) * 16 + i_ph * 4;
                    tma_load_3d(wins + s * MR_SLOT, &ring_map, &ready[s], i_bx * BS - 16, i_by * BS - 16 - (s & 1), z);
                    tma_load_3d(curs + s * MR_CUR, &cur_map, &ready[s], i_bx * BS, i_by * BS, unit);
                }
                __syncwarp();
                n += nit;
                slot += nit;
                if (slot >= MR_NS) { slot -= MR_NS; upar ^= 1u; }
                bundle_base += nb;
                if (lane == 0) *bundles_pub = bundle_base;          // descriptors written, meta written, loads issued
            }
            q = __shfl_sync(0xFFFFFFFFu, qn, 0);
        }
        if (lane == 0) *final_bundles = bundle_base;      // no bundle with index >= bundle_base will ever exist
    } else {
    // ================================= search warps =================================
    unsigned b = 0;
    if (lane == 0) b = mr2_next(counter);
    b = __shfl_sync(0xFFFFFFFFu, b, 0);
    while (true) {
        // ---- wait until bundle b is published (its descriptors are written and its items' loads issued)
        {
            SpinWait sw;
            bool none = false;
            while ((int)b >= *bundles_pub) {
                if ((int)b >= *final_bundles) { none = true; break; }      // written after the last publication
                sw.pause();
            }
            if (none) break;
        }
        const uint4 dsc = mr2_lds_v4(&bdesc[((b & (MR2_BD - 1)) << 3) | (lane >> 2)]);
        const unsigned slot = dsc.x & 0xFFu;
        const int grp = (int)(dsc.x >> 24), c = (int)(lane & 3u);
        const bool has = dsc.x & 0x10000u;
        unsigned nb = 0;
        if (lane == 0) nb = mr2_next(counter);              // next bundle index: consumed at the end of this iteration
        mbar_wait(&ready[slot], (dsc.x >> 8) & 1u);
        __syncwarp();
        const unsigned seg = __match_any_sync(0xFFFFFFFFu, has ? slot : 0xFFu);    // the lanes of my item (one slot each)
        const bool leader = has && (int)lane == __ffs(seg) - 1;
        const int4 mt = meta[slot];
        const int bx = mt.y, by = mt.z;
        const int px = (dsc.x >> 18) & 1u, py = (dsc.x >> 19) & 1u;          // 0 without half-pel search
        const unsigned char* win = smem_r + (dsc.y + (unsigned)(c * MR_PLANE));
        const uint32_t* cb = reinterpret_cast<const uint32_t*>(smem_r + dsc.w);
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
                const unsigned char* w0 = smem_r + (dsc.y + (unsigned)(4 * c * MR_WP + 32));
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
            const bool fast_valid = __all_sync(0xFFFFFFFFu, dsc.x & 0x20000u);
            // distance parts of the keys
            uint32_t ly8[3], lx8[9];
#pragma unroll
            for (int gg = 0; gg < 3; ++gg) {
                const int dy = mul * (oy0 + gg) + py;
                ly8[gg] = ((uint32_t)abs(dy) << 8) + (uint32_t)(2 * g.R) + (dy > 0 ? 1u : 0u);
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int dx = mul * (-16 + c + 4 * k) + px;                 // k < 4: negative, k >= 4: non-negative
                lx8[k] = (uint32_t)(dx * (k < 4 ? -254 : 258));              // |dx| << 8 | 2 dx
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
                const uint32_t v1 = has ? bq[e] : 0xFFFFFFFFu;
                const uint32_t m1 = __reduce_min_sync(seg, v1);
                if (leader && m1 != 0xFFFFFFFFu) {
                    const int4 ms = mt;
                    const unsigned long long key = mr2_key64(m1, (uint32_t)ms.w);
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
        const bool any_px0 = __any_sync(0xFFFFFFFFu, px == 0);
        if (any_px0) {
            const unsigned char* w0 = smem_r + (dsc.y + (unsigned)(4 * c * MR_WP + 32));
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
        const bool fast_valid = __all_sync(0xFFFFFFFFu, dsc.x & 0x20000u);
        if (fast_valid) {
            // interior block: the only invalid candidates are ox = 16 on odd horizontal phases and oy = 16 on odd vertical ones
            const uint32_t xbl = (c == 0 && px == 0) ? 0u : 0xFFFFFFFFu;          // candidate k = 8
            // min over the three vertical offsets first (their keys differ by SAD and the vertical addend only), then the horizontal
            // part: 3 IMAD (FMA pipe) + ONE 3-input min on the ALU pipe + one IMAD (dx * -254 = |dx| << 8 | 2 dx for dx < 0, dx * 258
            // for dx >= 0) per column.  The invalid third row of the last group on odd vertical phases is taken out through its
            // multiplier (0) and addend (all ones) in the table: no predicated 2-input mins
            const uint4 ly = *reinterpret_cast<const uint4*>(smem_r + dsc.z);
            const int dx0 = mul * (c - 16) + px, dxs = 4 * mul;
            uint32_t kk[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int dx = dx0 + k * dxs;                                // k < 4: negative, k >= 4: non-negative (c <= 3)
                const uint32_t t0 = acc[0][k] * 65536u + ly.x, t1 = acc[1][k] * 65536u + ly.y, t2 = acc[2][k] * ly.w + ly.z;
                kk[k] = (uint32_t)(dx * (k < 4 ? -254 : 258)) + min(min(t0, t1), t2);
            }
            best = min(min(min(kk[0], kk[1]), min(kk[2], kk[3])), min(min(kk[4], kk[5]), min(kk[6], kk[7])));
            if (any_px0) {                      // the 33rd horizontal offset exists on even horizontal phases only (uniform per bundle)
                const int dx = dx0 + 8 * dxs;
                const uint32_t t0 = ex[0] * 65536u + ly.x, t1 = ex[1] * 65536u + ly.y, t2 = ex[2] * ly.w + ly.z;
                best = min(best, ((uint32_t)(dx * 258) + min(min(t0, t1), t2)) | xbl);
            }
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
                ly8[gg] = ((uint32_t)abs(dy) << 8) + (uint32_t)(2 * g.R) + (dy > 0 ? 1u : 0u);
                ybad[gg] = (dy >= ylo && dy <= yhi) ? 0u : 0xFFFFFFFFu;
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int dx = mul * (-16 + c + 4 * k) + px;
                const uint32_t lx8 = ((uint32_t)abs(dx) << 8) + (uint32_t)(2 * dx);
                const uint32_t xbad = ((k < 8 || c == 0) && dx >= xlo && dx <= xhi) ? 0u : 0xFFFFFFFFu;
#pragma unroll
                for (int gg = 0; gg < 3; ++gg) {
                    const uint32_t key = ((k < 8 ? acc[gg][k < 8 ? k : 0] : ex[gg]) * 65536u + (lx8 + ly8[gg])) | xbad | ybad[gg];
                    best = min(best, key);
                }
            }
        }
        // ---- merge per item: order (SAD, |dx|+|dy|, ref, dx, dy); all lanes of a segment share ref, so ONE REDUX over the
        //      32-bit key (SAD, L1, dx, sign of dy) gives the winner of the segment
        const uint32_t v1 = has ? best : 0xFFFFFFFFu;
        {
            const uint32_t m1 = __reduce_min_sync(seg, v1);
            if (leader && m1 != 0xFFFFFFFFu)
                atomicMin(reinterpret_cast<unsigned long long*>(reinterpret_cast<MeResult*>(a.out) + mt.x), mr2_key64(m1, (uint32_t)mt.w));
        }
        }   // !QUAD
        __syncwarp();
        if (leader) mbar_arrive_cnt(&empty[slot], (uint32_t)__popc(seg));        // the tasks of this item done by this bundle
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
