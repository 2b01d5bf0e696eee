// Probe: 3-D u8 tensor map, 48x48x1 box at unaligned / negative x, mbarrier completion.  Development aid.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, uint8_t* out, int x, int y, int z, int bw, int bh, int mode) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bw * bh) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(sm)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
    }
    uint32_t ok = 0; long spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
        if (++spins > (1 << 22)) { if (threadIdx.x == 0) printf("timeout\n"); return; }
    }
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = sm[i];
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int W = 352, H = 288, Z = 8, pitch = 352;
    std::vector<uint8_t> h((size_t)pitch * H * Z);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *o;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&o, 65536);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    PFN fn = (PFN)p;
    const int boxes[3][2] = {{64, 48}, {16, 12}, {48, 24}};
    for (int b = 0; b < 3; ++b) {
        const int bw = boxes[b][0], bh = boxes[b][1];
        CUtensorMap map;
        cuuint64_t gdim[3] = {W, H, Z}; cuuint64_t gstr[2] = {pitch, (cuuint64_t)pitch * H};
        cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box %dx%d encode -> %d\n", bw, bh, (int)r);
        if (r) continue;
        const int tests[5][3] = {{0, 0, 0}, {16, 5, 3}, {-16, -16, 1}, {W - 32, H - 10, 7}, {-32, 100, 2}};
        for (auto& t : tests) {
            cudaMemset(o, 0xAB, 65536);
            probe<<<1, 128, 16384 + 64>>>(map, o, t[0], t[1], t[2], bw, bh, 0);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("  (%d,%d,%d): CUDA error %s\n", t[0], t[1], t[2], cudaGetErrorString(e)); return 1; }
            std::vector<uint8_t> res(bw * bh);
            cudaMemcpy(res.data(), o, bw * bh, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int yy = 0; yy < bh; ++yy) for (int xx = 0; xx < bw; ++xx) {
                const int X = t[0] + xx, Y = t[1] + yy;
                const uint8_t want = (X < 0 || Y < 0 || X >= W || Y >= H) ? 0 : h[(size_t)t[2] * pitch * H + (size_t)Y * pitch + X];
                if (res[yy * bw + xx] != want) ++bad;
            }
            printf("  (%d,%d,%d): %d mismatches\n", t[0], t[1], t[2], bad);
        }
    }
    return 0;
}
