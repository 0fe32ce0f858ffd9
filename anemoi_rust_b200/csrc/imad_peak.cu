// Integer-multiply issue-rate microbenchmark: the measured denominator of the IMAD roofline.
//
// MAC32 = one 32x32->64 multiply-accumulate. The Anemoi kernels spend > 95 % of their issue slots on
// IMAD.WIDE.U32(.X) (PTX mad.lo.cc.u32 + madc.hi.cc.u32 pairs that ptxas fuses), so the peak we report
// against is the chip-wide rate of exactly that instruction, measured with independent chains.
// Variants:
//   0  mad.lo.u32           (IMAD, 32-bit low product)
//   1  mad.hi.u32           (IMAD.HI)
//   2  mad.wide.u32         (IMAD.WIDE.U32, 64-bit accumulate, no carry)
//   3  carry chains         (IMAD.WIDE.U32 ..P0 / IMAD.WIDE.U32.X: 6-long chains as in a 12-limb row)
//   4  carry chains + 1 IADD3 per 2 IMAD.WIDE (checks that ALU-pipe work co-issues for free)
//   5  DFMA                 (context only)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "anemoi_b200.h"

namespace {

constexpr int kIters = 4096;
constexpr int kThreads = 256;

template <int V>
__global__ void __launch_bounds__(kThreads) imad_kernel(uint32_t* out, const uint32_t* in, long long* cycles) {
    uint32_t a0 = in[threadIdx.x & 31], a1 = in[32 + (threadIdx.x & 31)];
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = in[64 + i] + threadIdx.x;
    uint32_t s0 = a0 ^ a1, s1 = a0 + 7, s2 = a1 + 3;
    double d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = (double)r[i];
    double da = (double)a0 * 1e-9, db = (double)a1 * 1e-9;
    long long t0 = clock64();
    unsigned long long g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
        if (V == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a0), "r"(a1));
        } else if (V == 1) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a0), "r"(a1));
        } else if (V == 2) {
#pragma unroll
            for (int i = 0; i < 16; i += 2)
                asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %0, %3, t; mov.b64 {%0,%1}, t;}"
                             : "+r"(r[i]), "+r"(r[i + 1]) : "r"(a0), "r"(a1));
#pragma unroll
            for (int i = 0; i < 16; i += 2)
                asm volatile("{.reg .u64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %1, %3, t; mov.b64 {%0,%1}, t;}"
                             : "+r"(r[i]), "+r"(r[i + 1]) : "r"(a1), "r"(a0));
        } else if (V == 3 || V == 4 || V == 6 || V == 7) {
            // two interleaved chains of 4 wide MACs each, twice = 16 MAC32 per iteration
#pragma unroll
            for (int rep = 0; rep < 2; rep++) {
                asm volatile(
                    "mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                    "madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
                    "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
                    "madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
                    : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                    : "r"(a0), "r"(a1));
                asm volatile(
                    "mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                    "madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
                    "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
                    "madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
                    : "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                    : "r"(a1), "r"(a0));
                if (V == 6) {
#pragma unroll
                    for (int i = 0; i < 8; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(da), "d"(db));
                }
                if (V == 7) {
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(s0) : "r"(a0), "r"(a1));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(s1) : "r"(a0), "r"(a1));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(s2) : "r"(a0), "r"(a1));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(s0) : "r"(a1), "r"(a1));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(s1) : "r"(a1), "r"(a1));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(s2) : "r"(a1), "r"(a1));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(s0) : "r"(a0), "r"(a0));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(s1) : "r"(a0), "r"(a0));
                }
                if (V == 4) {
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(s0) : "r"(s1));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(s1) : "r"(s2));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(s2) : "r"(s0));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(s0) : "r"(s2));
                }
            }
        } else {
#pragma unroll
            for (int rep = 0; rep < 2; rep++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(da), "d"(db));
        }
    }
    long long t1 = clock64();
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    uint32_t acc = s0 ^ s1 ^ s2;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= r[i];
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= (uint32_t)d[i];
    if (acc == 0x12345678u) out[threadIdx.x] = acc;  // keeps the work alive; practically never taken
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cycles[0] = t1 - t0;
        cycles[1] = (long long)(g1 - g0);
    }
}

template <int V>
int run_variant(int blocks, double* ops_per_s, double* sm_mhz, cudaStream_t st) {
    static uint32_t* d_in = nullptr;
    static uint32_t* d_out = nullptr;
    static long long* d_cyc = nullptr;
    if (!d_in) {
        if (cudaMalloc(&d_in, 4096) != cudaSuccess) return ANEMOI_B200_ERR_CUDA;
        if (cudaMalloc(&d_out, 4096) != cudaSuccess) return ANEMOI_B200_ERR_CUDA;
        if (cudaMalloc(&d_cyc, 64) != cudaSuccess) return ANEMOI_B200_ERR_CUDA;
        uint32_t h[1024];
        for (int i = 0; i < 1024; i++) h[i] = 0x9E3779B9u * (i + 1) | 1u;
        cudaMemcpy(d_in, h, 4096, cudaMemcpyHostToDevice);
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) imad_kernel<V><<<blocks, kThreads, 0, st>>>(d_out, d_in, d_cyc);
    const int reps = 5;
    cudaEventRecord(e0, st);
    for (int w = 0; w < reps; w++) imad_kernel<V><<<blocks, kThreads, 0, st>>>(d_out, d_in, d_cyc);
    cudaEventRecord(e1, st);
    if (cudaEventSynchronize(e1) != cudaSuccess) return ANEMOI_B200_ERR_CUDA;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cyc[2] = {0, 1};
    cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double ops = (double)reps * blocks * kThreads * (double)kIters * 16.0;
    *ops_per_s = ops / (ms * 1e-3);
    // SM clock: cycles (clock64) over nanoseconds (globaltimer) of one block's loop
    *sm_mhz = (double)cyc[0] / (double)cyc[1] * 1e3;
    return ANEMOI_B200_OK;
}

}  // namespace

extern "C" int anemoi_b200_imad_peak(int variant, double* ops_per_s, double* sm_mhz) {
    if (!ops_per_s || !sm_mhz) return ANEMOI_B200_ERR_ARG;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return ANEMOI_B200_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int blocks = sms * 8 * 4;  // 8 resident blocks of 256 threads per SM (32 regs), 4 waves
    switch (variant) {
        case 0: return run_variant<0>(blocks, ops_per_s, sm_mhz, 0);
        case 1: return run_variant<1>(blocks, ops_per_s, sm_mhz, 0);
        case 2: return run_variant<2>(blocks, ops_per_s, sm_mhz, 0);
        case 3: return run_variant<3>(blocks, ops_per_s, sm_mhz, 0);
        case 4: return run_variant<4>(blocks, ops_per_s, sm_mhz, 0);
        case 5: return run_variant<5>(blocks, ops_per_s, sm_mhz, 0);
        case 6: return run_variant<6>(blocks, ops_per_s, sm_mhz, 0);
        case 7: return run_variant<7>(blocks, ops_per_s, sm_mhz, 0);
        default: return ANEMOI_B200_ERR_ARG;
    }
}
