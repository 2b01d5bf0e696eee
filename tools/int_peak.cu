// Microbenchmark: measured integer-ALU instruction peaks on B200 (sm_100a).
// Produces the denominator for the motion-estimation roofline (SURVEY.md §8d says the INT32
// lane-instruction peak must be measured; MEASURED_PEAKS.json only holds HBM and bf16 figures).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak int_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 8

template <int OP>
__global__ void __launch_bounds__(1024, 2) k(uint32_t* out, uint32_t seed) {
    uint32_t a[NACC], b = seed * 0x9E3779B9u + threadIdx.x, c = seed ^ 0x01020304u;
#pragma unroll
    for (int i = 0; i < NACC; i++) a[i] = threadIdx.x * 0x01010101u + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (OP == 0) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 4) {  // 1:1 mix vabsdiff4 + imad
                if (i & 1) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            }
            if (OP == 5) {  // 1:1 mix vabsdiff4 + add
                if (i & 1) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
                else asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            }
            if (OP == 6) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 7) {  // 3:1 vabsdiff4 : imad
                if ((i & 3) != 3) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            }
            if (OP == 8) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 9) {  // fp32 fma for reference
                float f = __uint_as_float(a[i]);
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c)));
                a[i] = __float_as_uint(f);
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory LDS.128 bandwidth (conflict-free)
__global__ void __launch_bounds__(1024, 2) lds128(uint32_t* out) {
    __shared__ uint4 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    uint4 acc = make_uint4(0, 0, 0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint4 v = buf[(idx + i * 32) & 2047];
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
        idx = (idx + acc.x) & 2047;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x ^ acc.y ^ acc.z ^ acc.w;
}

template <int OP>
double run(const char* name, uint32_t* d, int sms) {
    int grid = sms * 2;
    k<OP><<<grid, 1024>>>(d, 1);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k<OP><<<grid, 1024>>>(d, r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double lane_ops = (double)grid * 1024 * ITERS * NACC;
    double tops = lane_ops / (best * 1e-3) / 1e12;
    printf("{\"op\": \"%s\", \"ms\": %.4f, \"lane_Tops\": %.3f, \"warp_Ginstr\": %.2f}\n", name, best, tops, tops * 1e3 / 32);
    return tops;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d}\n", p.name, p.multiProcessorCount, clk);
    uint32_t* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 2 * 1024 * 4);
    int sms = p.multiProcessorCount;
    run<0>("vabsdiff4.add", d, sms);
    run<6>("vabsdiff4", d, sms);
    run<1>("iadd", d, sms);
    run<2>("lop3", d, sms);
    run<3>("imad", d, sms);
    run<8>("dp4a", d, sms);
    run<9>("ffma", d, sms);
    run<4>("vabsdiff4.add+imad 1:1", d, sms);
    run<5>("vabsdiff4.add+iadd 1:1", d, sms);
    run<7>("vabsdiff4.add+imad 3:1", d, sms);
    {
        lds128<<<sms * 2, 1024>>>(d); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int r = 0; r < 5; r++) {
            cudaEventRecord(e0); lds128<<<sms * 2, 1024>>>(d); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double bytes = (double)sms * 2 * 1024 * ITERS * 8 * 16;
        printf("{\"op\": \"lds128\", \"ms\": %.4f, \"TBps\": %.2f}\n", best, bytes / (best * 1e-3) / 1e12);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
