// Exhaustive motion search, persistent + TMA + warp-specialised version (sm_100a).
//
//   grid  = one CTA per SM (persistent); stage = SI consecutive items (item = (block, reference, phase plane))
//   load    : one TMA box load per item (cp.async.bulk.tensor.3d, SASS UTMALDG) brings the raw (bs+2r)-row search
//             window, 16-byte aligned in x (TMA faults on unaligned inner coordinates -- tools/tma_probe.cu), with the
//             out-of-frame part zero-filled by the TMA unit (those candidates are invalid anyway, Encoder.py:695-698).
//             Raw windows are double buffered; the load of stage n+2 is issued as soon as stage n has been expanded.
//   expand  : all threads turn the raw window of stage n+1 into FOUR copies shifted by 0..3 bytes (funnel shifts, ~3 %
//             of the integer-pipe work) so that every candidate reads 32-bit aligned words, and stage the current blocks
//   search  : one task per thread = (item, byte shift c, group of G vertical offsets), NDX x G candidates accumulated
//             with VABSDIFF4.U8.ACC from 128-bit shared loads; per-thread argmin on a packed 32-bit key, per-stage
//             merge through warp shuffles / shared atomics, per-block merge through a global atomicMin on the 64-bit
//             key (SAD, |dx|+|dy|, ref, dx, dy) that encodes the reference's replace rule (appendix A4).
//   one bar.sync per stage separates expand(n+1)/search(n) from expand(n+2)/search(n+1) (copies are double buffered).
// The whole reference ring is ONE 3-D tensor map {W, H, units*slots*4 planes}: plane index = z coordinate.
#pragma once
#include <cuda.h>

#include "so_common.cuh"
#include "so_me_full.cuh"

struct MeTmaArgs {
    FrameGeom g;                 // g.bs = block size searched by this launch
    const uint8_t* cur;          // unit 0, dense [H][W]
    size_t cur_unit_stride;
    unsigned long long* out;     // packed keys, stride of MeResult (16 B) per block
    size_t out_unit_stride;      // in MeResult elements
    int units;
    int nph;                     // 4 (fme) or 1
    int items_per_unit;
    int stages_per_unit;
    int SI;                      // items per stage
    int NG;                      // vertical groups per (item, shift)
    int rows;                    // window rows = bs + 2r
    int wpitch;                  // row pitch of the shifted copies (bytes, 16 * odd)
    int copy_stride;             // bytes between shifted copies (multiple of 128)
    int item_stride;             // 4 * copy_stride
    int stage_bytes;             // SI * item_stride
    int raw_w;                   // TMA box width (bytes, multiple of 16; a power of two times 16 when aligned16)
    int raw_item_stride;         // bytes between raw windows (multiple of 128)
    int raw_stage_bytes;         // SI * raw_item_stride
    int aligned16;               // bx*bs - r is a multiple of 16 for every block
    int z_per_unit;              // planes per unit in the ring tensor = nslots * 4
    int slot[SO_MAX_REF];        // list index -> ring slot
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    unsigned long long spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > (1ull << 24)) __trap();        // a lost arrival must fault, never hang the GPU
    }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
                 : "memory");
}

constexpr int ME_TMA_STAGES = 2;

template <int BS, int NDX, int G>
__global__ void __launch_bounds__(352, 1) me_tma_kernel(const __grid_constant__ CUtensorMap ring_map, const MeTmaArgs a) {
    constexpr int WPR = BS / 4;
    constexpr int NW = NDX + WPR - 1;
    constexpr int NV = (NW + 3) / 4;
    constexpr int NCH = ((NW * 4 + 15) / 16) | 1;            // 16-byte chunks per copy row (wpitch / 16)
    extern __shared__ __align__(1024) unsigned char smem_t[];
    unsigned char* const smem = smem_t;
    const FrameGeom& g = a.g;
    // carve-up: [raw: 2 stages (TMA destinations, 128-B aligned)][copies: 2 stages][cur tiles: 2 x SI x BS*BS][keys: 2 x SI][rawfull[2]]
    unsigned char* raws = smem;
    unsigned char* wins = smem + ME_TMA_STAGES * a.raw_stage_bytes;
    uint32_t* curs = reinterpret_cast<uint32_t*>(wins + ME_TMA_STAGES * a.stage_bytes);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(curs) + ME_TMA_STAGES * a.SI * BS * BS);
    uint64_t* rawfull = reinterpret_cast<uint64_t*>(keys + ME_TMA_STAGES * a.SI);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthr = blockDim.x;
    if (tid == 0) {
        for (int s = 0; s < ME_TMA_STAGES; ++s) mbar_init(&rawfull[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < ME_TMA_STAGES * a.SI; i += nthr) keys[i] = ~0ull;
    __syncthreads();

    const int per_blk = g.nref * a.nph;
    const int total_stages = a.units * a.stages_per_unit;
    const int nloc = (total_stages - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // stages of this CTA
    auto stage_of = [&](int j) { return (int)blockIdx.x + j * (int)gridDim.x; };

    // ---- TMA issue of local stage j (one thread)
    auto issue_raw = [&](int j) {
        const int sg = stage_of(j), rb = j & 1;
        const int unit = sg / a.stages_per_unit, sidx = sg % a.stages_per_unit;
        const int item0 = sidx * a.SI;
        const int nitems = min(a.SI, a.items_per_unit - item0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the buffer was last read through the generic proxy
        mbar_arrive_expect_tx(&rawfull[rb], (uint32_t)(nitems * a.rows * a.raw_w));
        for (int li = 0; li < nitems; ++li) {
            const int item = item0 + li;
            const int blk = item / per_blk, rp = item % per_blk;
            const int ref = rp / a.nph, ph = rp % a.nph;
            const int bx = blk % g.nbx, by = blk / g.nbx;
            const int z = unit * a.z_per_unit + a.slot[ref] * 4 + ph;
            const int X0 = bx * BS - g.r;
            tma_load_3d(raws + rb * a.raw_stage_bytes + li * a.raw_item_stride, &ring_map, &rawfull[rb], X0 & ~15, by * BS - g.r, z);
        }
    };

    // ---- expand local stage j: raw window -> four byte-shifted copies, current blocks -> shared
    auto expand = [&](int j) {
        const int sg = stage_of(j), sb = j & 1;
        const int unit = sg / a.stages_per_unit, sidx = sg % a.stages_per_unit;
        const int item0 = sidx * a.SI;
        const int nitems = min(a.SI, a.items_per_unit - item0);
        const uint8_t* cur = a.cur + unit * a.cur_unit_stride;
        uint32_t* cdst = curs + sb * a.SI * BS * WPR;
        for (int i = tid; i < nitems * BS * WPR; i += nthr) {
            const int li = i / (BS * WPR), rem = i % (BS * WPR), row = rem / WPR, w = rem % WPR;
            const int blk = (item0 + li) / per_blk;
            const int bx = blk % g.nbx, by = blk / g.nbx;
            cdst[i] = __ldg(reinterpret_cast<const uint32_t*>(cur + (size_t)(by * BS + row) * g.W + bx * BS + w * 4));
        }
        mbar_wait(&rawfull[sb], (uint32_t)((j >> 1) & 1));
        const int lpr = a.raw_w >> 4;                         // 16-byte chunks per raw row
        if (a.aligned16) {
            // unit = (item, row, chunk m < NCH): one LDS.128 + the next word feed all four copies
            const int per_item = a.rows * NCH;
            for (int u = tid; u < nitems * per_item; u += nthr) {
                const int li = u / per_item, rem = u % per_item, row = rem / NCH, m = rem % NCH;
                const unsigned char* rsrc = raws + sb * a.raw_stage_bytes + li * a.raw_item_stride + row * a.raw_w + m * 16;
                const uint4 v = *reinterpret_cast<const uint4*>(rsrc);
                const uint32_t nx = (m + 1 < lpr) ? *reinterpret_cast<const uint32_t*>(rsrc + 16) : 0u;
                unsigned char* o = wins + sb * a.stage_bytes + li * a.item_stride + row * a.wpitch + m * 16;
                *reinterpret_cast<uint4*>(o) = v;
#pragma unroll
                for (int c = 1; c < 4; ++c) {
                    uint4 w4;
                    w4.x = __funnelshift_r(v.x, v.y, 8 * c); w4.y = __funnelshift_r(v.y, v.z, 8 * c);
                    w4.z = __funnelshift_r(v.z, v.w, 8 * c); w4.w = __funnelshift_r(v.w, nx, 8 * c);
                    *reinterpret_cast<uint4*>(o + c * a.copy_stride) = w4;
                }
            }
        } else {
            // generic alignment: output chunk q of copy c = raw bytes [off + c + 16q, +16)
            const int per_item = 4 * a.rows * NCH;
            for (int u = tid; u < nitems * per_item; u += nthr) {
                const int li = u / per_item;
                int rem = u % per_item;
                const int c = rem / (a.rows * NCH);
                rem %= a.rows * NCH;
                const int row = rem / NCH, q = rem % NCH;
                const int blk = (item0 + li) / per_blk;
                const int X0 = (blk % g.nbx) * BS - g.r;
                const int sbyte = X0 - (X0 & ~15) + c, cq = sbyte >> 4, wo = (sbyte & 15) >> 2, bits = (sbyte & 3) * 8;
                const unsigned char* rsrc = raws + sb * a.raw_stage_bytes + li * a.raw_item_stride + row * a.raw_w;
                const int m0 = q + cq;
                const uint4 A = m0 < lpr ? *reinterpret_cast<const uint4*>(rsrc + m0 * 16) : make_uint4(0, 0, 0, 0);
                const uint4 B = m0 + 1 < lpr ? *reinterpret_cast<const uint4*>(rsrc + (m0 + 1) * 16) : make_uint4(0, 0, 0, 0);
                uint4 w4;
                switch (wo) {
                    case 0: w4 = make_uint4(__funnelshift_r(A.x, A.y, bits), __funnelshift_r(A.y, A.z, bits),
                                            __funnelshift_r(A.z, A.w, bits), __funnelshift_r(A.w, B.x, bits)); break;
                    case 1: w4 = make_uint4(__funnelshift_r(A.y, A.z, bits), __funnelshift_r(A.z, A.w, bits),
                                            __funnelshift_r(A.w, B.x, bits), __funnelshift_r(B.x, B.y, bits)); break;
                    case 2: w4 = make_uint4(__funnelshift_r(A.z, A.w, bits), __funnelshift_r(A.w, B.x, bits),
                                            __funnelshift_r(B.x, B.y, bits), __funnelshift_r(B.y, B.z, bits)); break;
                    default: w4 = make_uint4(__funnelshift_r(A.w, B.x, bits), __funnelshift_r(B.x, B.y, bits),
                                             __funnelshift_r(B.y, B.z, bits), __funnelshift_r(B.z, B.w, bits)); break;
                }
                *reinterpret_cast<uint4*>(wins + sb * a.stage_bytes + li * a.item_stride + c * a.copy_stride + row * a.wpitch + q * 16) = w4;
            }
        }
    };

    if (nloc <= 0) return;
    if (tid == 0) { issue_raw(0); if (nloc > 1) issue_raw(1); }
    expand(0);
    __syncthreads();

    const int tasks_per_item = 4 * a.NG;
    for (int j = 0; j < nloc; ++j) {
        const int sg = stage_of(j), sb = j & 1;
        const int unit = sg / a.stages_per_unit, sidx = sg % a.stages_per_unit;
        const int item0 = sidx * a.SI;
        const int nitems = min(a.SI, a.items_per_unit - item0);
        const int blk0 = item0 / per_blk;
        unsigned long long* skeys = keys + sb * a.SI;
        // raw[sb] has been expanded (barrier at the end of the previous iteration): refill it with stage j+2
        if (tid == 0 && j + 2 < nloc) issue_raw(j + 2);
        if (j + 1 < nloc) expand(j + 1);

        for (int task = tid; task < nitems * tasks_per_item; task += nthr) {
            const int li = task / tasks_per_item;
            const int rem = task % tasks_per_item;
            const int c = rem / a.NG, grp = rem % a.NG;
            const int item = item0 + li;
            const int blk = item / per_blk, rp = item % per_blk;
            const int ref = rp / a.nph, ph = rp % a.nph;
            const int px = ph & 1, py = ph >> 1;
            const int lb = blk - blk0;
            const int oy0 = grp * G;
            const unsigned char* win = wins + sb * a.stage_bytes + li * a.item_stride + c * a.copy_stride + oy0 * a.wpitch;
            const uint4* cb4 = reinterpret_cast<const uint4*>(curs + (sb * a.SI + li) * BS * WPR);
            const uint32_t* cb = curs + (sb * a.SI + li) * BS * WPR;

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
                    if constexpr (WPR == 4) {
                        const uint4 q = cb4[rho];
                        curq[0][0] = q.x; curq[0][1] = q.y; curq[0][2] = q.z; curq[0][3] = q.w;
                    } else if constexpr (WPR == 2) {
                        const uint2 q = reinterpret_cast<const uint2*>(cb)[rho];
                        curq[0][0] = q.x; curq[0][1] = q.y;
                    } else {
                        curq[0][0] = cb[rho];
                    }
                }
#pragma unroll
                for (int gg = 0; gg < G; ++gg) {
                    const int jr = rho - gg;
                    if (jr >= 0 && jr < BS) {
#pragma unroll
                        for (int k = 0; k < NDX; ++k)
#pragma unroll
                            for (int w = 0; w < WPR; ++w) acc[gg][k] = sad4_acc(refw[k + w], curq[gg][w], acc[gg][k]);
                    }
                }
            }

            // ---- thread-local argmin.  key32 = sad<<16 | (|dx|+|dy|)<<8 | (k*G+g); invalid candidates are OR-ed to all ones.
            const int bx = blk % g.nbx, by = blk / g.nbx;
            int xlo, xhi, ylo, yhi;
            valid_range(bx * BS, g.W, BS, g.fme, g.fme, xlo, xhi);
            valid_range(by * BS, g.H, BS, g.fme, g.fme, ylo, yhi);
            xlo = max(xlo, -g.R); xhi = min(xhi, g.R);
            ylo = max(ylo, -g.R); yhi = min(yhi, g.R);
            const int mul = g.fme ? 2 : 1;
            uint32_t ly8[G], ybad[G];
#pragma unroll
            for (int gg = 0; gg < G; ++gg) {
                const int oy = -g.r + oy0 + gg, dy = mul * oy + (g.fme ? py : 0);
                ly8[gg] = (uint32_t)(abs(dy) << 8) + gg;
                ybad[gg] = (oy <= g.r && dy >= ylo && dy <= yhi) ? 0u : 0xFFFFFFFFu;
            }
            uint32_t best = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < NDX; ++k) {
                const int ox = -g.r + c + 4 * k, dx = mul * ox + (g.fme ? px : 0);
                const uint32_t lx8 = (uint32_t)(abs(dx) << 8) + k * G;
                const uint32_t xbad = (ox <= g.r && dx >= xlo && dx <= xhi) ? 0u : 0xFFFFFFFFu;
#pragma unroll
                for (int gg = 0; gg < G; ++gg) {
                    const uint32_t key = ((acc[gg][k] << 16) + (lx8 + ly8[gg])) | xbad | ybad[gg];
                    best = min(best, key);
                }
            }
            unsigned long long key = ~0ull;
            if (best != 0xFFFFFFFFu) {
                const int idx = best & 0xFF, k = idx / G, gg = idx % G;
                const int ox = -g.r + c + 4 * k, oy = -g.r + oy0 + gg;
                const int dx = mul * ox + (g.fme ? px : 0);
                const int dy = mul * oy + (g.fme ? py : 0);
                key = ((unsigned long long)(best >> 16) << 40) | ((unsigned long long)((best >> 8) & 0xFF) << 24) |
                      ((unsigned long long)ref << 16) | ((unsigned long long)(dx + g.R) << 8) | (unsigned long long)(dy + g.R);
            }
            // merge: warp shuffle when the whole warp works on the same block, shared atomics otherwise
            const unsigned act = __activemask();
            const int lb0 = __shfl_sync(act, lb, __ffs(act) - 1);
            if (act == 0xFFFFFFFFu && __all_sync(act, lb == lb0)) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
                    key = other < key ? other : key;
                }
                if (lane == 0 && key < skeys[lb]) atomicMin(&skeys[lb], key);
            } else if (key < skeys[lb]) {
                atomicMin(&skeys[lb], key);
            }
        }
        __syncthreads();        // search(j) and expand(j+1) complete
        if (warp == 0) {        // flush the per-block keys of stage j (next written by search(j+2), after the next barrier)
            const int blk_last = (item0 + nitems - 1) / per_blk;
            for (int i = lane; i <= blk_last - blk0; i += 32) {
                const unsigned long long k = skeys[i];
                skeys[i] = ~0ull;
                if (k != ~0ull)
                    atomicMin(reinterpret_cast<unsigned long long*>(reinterpret_cast<MeResult*>(a.out) + unit * a.out_unit_stride + blk0 + i), k);
            }
        }
    }
}
