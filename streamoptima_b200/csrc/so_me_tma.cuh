// Exhaustive motion search, persistent + TMA + warp-specialised version (sm_100a).
//
//   grid  = one CTA per SM (persistent), block = 1 producer warp + NCW consumer warps
//   stage = SI consecutive items (item = (block, reference, phase plane)); two stages are resident in shared memory
//   producer: one TMA box load per item (cp.async.bulk.tensor.3d, SASS UTMALDG) brings the raw (bs+2r)-row search
//             window, 16-byte aligned in x (TMA faults on unaligned inner coordinates -- tools/tma_probe.cu), with the
//             out-of-frame part zero-filled by the TMA unit (those candidates are invalid anyway, Encoder.py:695-698).
//             While the consumers work on stage n the producer warp expands the raw window of stage n+1 into FOUR
//             copies shifted by 0..3 bytes (funnel shifts), so that every candidate reads 32-bit aligned words, and
//             stages the current blocks.  The raw load of stage n+2 is already in flight meanwhile.
//   consumer: one task per thread = (item, byte shift c, group of G vertical offsets), NDX x G candidates accumulated
//             with VABSDIFF4.U8.ACC from 128-bit shared loads; per-thread argmin on a packed 32-bit key, per-stage
//             merge through warp shuffles / shared atomics, per-block merge through a global atomicMin on the 64-bit
//             key (SAD, |dx|+|dy|, ref, dx, dy) that encodes the reference's replace rule (appendix A4).
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
__global__ void __launch_bounds__(384, 1) me_tma_kernel(const __grid_constant__ CUtensorMap ring_map, const MeTmaArgs a) {
    constexpr int WPR = BS / 4;
    constexpr int NW = NDX + WPR - 1;
    constexpr int NV = (NW + 3) / 4;
    constexpr int NCH = ((NW * 4 + 15) / 16) | 1;            // 16-byte chunks per copy row (wpitch / 16)
    extern __shared__ __align__(1024) unsigned char smem_t[];
    unsigned char* const smem = smem_t;
    const FrameGeom& g = a.g;
    // carve-up: [copies: 2 stages][raw: 2 stages][cur tiles: 2 x SI x BS*BS][keys: 2 x SI][mbarriers: rawfull, ready, empty]
    unsigned char* wins = smem;
    unsigned char* raws = smem + ME_TMA_STAGES * a.stage_bytes;
    uint32_t* curs = reinterpret_cast<uint32_t*>(raws + ME_TMA_STAGES * a.raw_stage_bytes);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(curs) + ME_TMA_STAGES * a.SI * BS * BS);
    uint64_t* rawfull = reinterpret_cast<uint64_t*>(keys + ME_TMA_STAGES * a.SI);
    uint64_t* ready = rawfull + ME_TMA_STAGES;
    uint64_t* empty = ready + ME_TMA_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = (blockDim.x >> 5) - 1;                    // consumer warps
    if (threadIdx.x == 0) {
        for (int s = 0; s < ME_TMA_STAGES; ++s) { mbar_init(&rawfull[s], 1); mbar_init(&ready[s], 1); mbar_init(&empty[s], ncw); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < ME_TMA_STAGES * a.SI; i += blockDim.x) keys[i] = ~0ull;
    __syncthreads();

    const int per_blk = g.nref * a.nph;
    const int total_stages = a.units * a.stages_per_unit;

    if (warp == 0) {
        // ============================ producer ============================
        auto issue_raw = [&](int itx, int sgx) {
            const int rb = itx % ME_TMA_STAGES;
            const int unit = sgx / a.stages_per_unit, sidx = sgx % a.stages_per_unit;
            const int item0 = sidx * a.SI;
            const int nitems = min(a.SI, a.items_per_unit - item0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // raw buffer was read through the generic proxy
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(&rawfull[rb], (uint32_t)(nitems * a.rows * a.raw_w));
            __syncwarp();
            for (int li = lane; li < nitems; li += 32) {
                const int item = item0 + li;
                const int blk = item / per_blk, rp = item % per_blk;
                const int ref = rp / a.nph, ph = rp % a.nph;
                const int bx = blk % g.nbx, by = blk / g.nbx;
                const int z = unit * a.z_per_unit + a.slot[ref] * 4 + ph;
                const int X0 = bx * BS - g.r;
                tma_load_3d(raws + rb * a.raw_stage_bytes + li * a.raw_item_stride, &ring_map, &rawfull[rb], X0 & ~15, by * BS - g.r, z);
            }
        };
        int it = 0;
        int sg = blockIdx.x;
        if (sg < total_stages) issue_raw(0, sg);
        for (; sg < total_stages; sg += gridDim.x, ++it) {
            const int sb = it % ME_TMA_STAGES;
            const uint32_t par = (it / ME_TMA_STAGES) & 1;
            if (sg + (int)gridDim.x < total_stages) issue_raw(it + 1, sg + gridDim.x);   // raw[(it+1)%2] was consumed in iteration it-1
            mbar_wait(&rawfull[sb], par);
            mbar_wait(&empty[sb], par ^ 1);
            const int unit = sg / a.stages_per_unit, sidx = sg % a.stages_per_unit;
            const int item0 = sidx * a.SI;
            const int nitems = min(a.SI, a.items_per_unit - item0);
            // ---- current blocks
            const uint8_t* cur = a.cur + unit * a.cur_unit_stride;
            uint32_t* cdst = curs + sb * a.SI * BS * WPR;
            for (int i = lane; i < nitems * BS * WPR; i += 32) {
                const int li = i / (BS * WPR), rem = i % (BS * WPR), row = rem / WPR, w = rem % WPR;
                const int blk = (item0 + li) / per_blk;
                const int bx = blk % g.nbx, by = blk / g.nbx;
                cdst[i] = __ldg(reinterpret_cast<const uint32_t*>(cur + (size_t)(by * BS + row) * g.W + bx * BS + w * 4));
            }
            // ---- raw window -> four byte-shifted copies
            if (a.aligned16) {
                // lanes = (row, 16-byte chunk): one LDS.128 + one shuffle feed all four copies
                const int lpr = a.raw_w >> 4;                 // chunks per raw row (power of two)
                const int rpi = 32 / lpr;                     // rows per iteration
                const int m = lane & (lpr - 1), rr = lane / lpr;
                for (int li = 0; li < nitems; ++li) {
                    const unsigned char* rsrc = raws + sb * a.raw_stage_bytes + li * a.raw_item_stride;
                    unsigned char* wdst = wins + sb * a.stage_bytes + li * a.item_stride;
                    for (int row0 = 0; row0 < a.rows; row0 += rpi) {
                        const int row = row0 + rr;
                        uint4 v = make_uint4(0, 0, 0, 0);
                        if (row < a.rows) v = *reinterpret_cast<const uint4*>(rsrc + row * a.raw_w + m * 16);
                        uint32_t nx = __shfl_down_sync(0xFFFFFFFFu, v.x, 1);
                        if (m == lpr - 1) nx = 0;
                        if (row < a.rows && m < NCH) {
                            unsigned char* o = wdst + row * a.wpitch + m * 16;
                            *reinterpret_cast<uint4*>(o) = v;
#pragma unroll
                            for (int c = 1; c < 4; ++c) {
                                uint4 w4;
                                w4.x = __funnelshift_r(v.x, v.y, 8 * c); w4.y = __funnelshift_r(v.y, v.z, 8 * c);
                                w4.z = __funnelshift_r(v.z, v.w, 8 * c); w4.w = __funnelshift_r(v.w, nx, 8 * c);
                                *reinterpret_cast<uint4*>(o + c * a.copy_stride) = w4;
                            }
                        }
                    }
                }
            } else {
                // generic alignment: output chunk q of copy c = raw bytes [off + c + 16q, +16)
                const int lpr = a.raw_w >> 4;
                for (int li = 0; li < nitems; ++li) {
                    const int blk = (item0 + li) / per_blk;
                    const int X0 = (blk % g.nbx) * BS - g.r;
                    const int off = X0 - (X0 & ~15);
                    const unsigned char* rsrc = raws + sb * a.raw_stage_bytes + li * a.raw_item_stride;
                    unsigned char* wdst = wins + sb * a.stage_bytes + li * a.item_stride;
                    for (int c = 0; c < 4; ++c) {
                        const int sbyte = off + c, cq = sbyte >> 4, wo = (sbyte & 15) >> 2, bits = (sbyte & 3) * 8;
                        for (int u = lane; u < a.rows * NCH; u += 32) {
                            const int row = u / NCH, q = u % NCH;
                            const int m0 = q + cq;
                            uint32_t W8[9];
                            const uint4 A = m0 < lpr ? *reinterpret_cast<const uint4*>(rsrc + row * a.raw_w + m0 * 16) : make_uint4(0, 0, 0, 0);
                            const uint4 B = m0 + 1 < lpr ? *reinterpret_cast<const uint4*>(rsrc + row * a.raw_w + (m0 + 1) * 16) : make_uint4(0, 0, 0, 0);
                            W8[0] = A.x; W8[1] = A.y; W8[2] = A.z; W8[3] = A.w; W8[4] = B.x; W8[5] = B.y; W8[6] = B.z; W8[7] = B.w; W8[8] = 0;
                            uint4 w4;
                            switch (wo) {
                                case 0: w4 = make_uint4(__funnelshift_r(W8[0], W8[1], bits), __funnelshift_r(W8[1], W8[2], bits),
                                                        __funnelshift_r(W8[2], W8[3], bits), __funnelshift_r(W8[3], W8[4], bits)); break;
                                case 1: w4 = make_uint4(__funnelshift_r(W8[1], W8[2], bits), __funnelshift_r(W8[2], W8[3], bits),
                                                        __funnelshift_r(W8[3], W8[4], bits), __funnelshift_r(W8[4], W8[5], bits)); break;
                                case 2: w4 = make_uint4(__funnelshift_r(W8[2], W8[3], bits), __funnelshift_r(W8[3], W8[4], bits),
                                                        __funnelshift_r(W8[4], W8[5], bits), __funnelshift_r(W8[5], W8[6], bits)); break;
                                default: w4 = make_uint4(__funnelshift_r(W8[3], W8[4], bits), __funnelshift_r(W8[4], W8[5], bits),
                                                         __funnelshift_r(W8[5], W8[6], bits), __funnelshift_r(W8[6], W8[7], bits)); break;
                            }
                            *reinterpret_cast<uint4*>(wdst + c * a.copy_stride + row * a.wpitch + q * 16) = w4;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[sb]);
        }
    } else {
        // ============================ consumers ============================
        const int ctid = threadIdx.x - 32;
        const int tasks_per_item = 4 * a.NG;
        int it = 0;
        for (int sg = blockIdx.x; sg < total_stages; sg += gridDim.x, ++it) {
            const int sb = it % ME_TMA_STAGES;
            const uint32_t par = (it / ME_TMA_STAGES) & 1;
            const int unit = sg / a.stages_per_unit, sidx = sg % a.stages_per_unit;
            const int item0 = sidx * a.SI;
            const int nitems = min(a.SI, a.items_per_unit - item0);
            const int blk0 = item0 / per_blk;
            unsigned long long* skeys = keys + sb * a.SI;
            mbar_wait(&ready[sb], par);

            for (int task = ctid; task < nitems * tasks_per_item; task += ncw * 32) {
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
                unsigned long long key = ~0ull;
                if (best != 0xFFFFFFFFu) {
                    const int idx = best & 0xFF, k = idx / G, gg = idx % G;
                    const int ox = -g.r + c + 4 * k, oy = -g.r + oy0 + gg;
                    const int dx = g.fme ? 2 * ox + px : ox;
                    const int dy = g.fme ? 2 * oy + py : oy;
                    key = ((unsigned long long)(best >> 16) << 40) | ((unsigned long long)((best >> 8) & 0xFF) << 24) |
                          ((unsigned long long)ref << 16) | ((unsigned long long)(dx + g.R) << 8) | (unsigned long long)(dy + g.R);
                }
                // merge: warp shuffle when the whole warp works on the same block, shared atomics otherwise
                const unsigned act = __activemask();
                const int lb0 = __shfl_sync(act, lb, __ffs(act) - 1);
                if (__all_sync(act, lb == lb0) && act == 0xFFFFFFFFu) {
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
            // all consumer warps done with this stage: flush the per-block keys, release the buffer
            asm volatile("bar.sync 1, %0;" ::"r"(ncw * 32) : "memory");
            if (warp == 1) {
                const int blk_last = (item0 + nitems - 1) / per_blk;
                for (int i = lane; i <= blk_last - blk0; i += 32) {
                    const unsigned long long k = skeys[i];
                    skeys[i] = ~0ull;
                    if (k != ~0ull)
                        atomicMin(reinterpret_cast<unsigned long long*>(reinterpret_cast<MeResult*>(a.out) + unit * a.out_unit_stride + blk0 + i), k);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[sb]);
        }
    }
}
