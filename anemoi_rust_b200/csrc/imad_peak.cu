// Integer-multiply issue-rate microbenchmark: the measured denominator of the IMAD roofline.
//
// MAC32 = one 32x32->64 multiply-accumulate. The Anemoi kernels spend their FMA-pipe issue slots on
// IMAD.WIDE.U32(.X) (PTX mad.lo.cc.u32 + madc.hi.cc.u32 pairs that ptxas fuses), so the peak they are
// reported against is the chip-wide rate of IMAD.WIDE.U32 with a 64-bit accumulate, measured with sixteen
// independent accumulators per thread at full occupancy (variant 2). The other variants document what the
// integer pipe does with the neighbouring instruction flavours and whether other pipes overlap with it.
//   0  IMAD          (32-bit low product + 32-bit add)
//   1  IMAD.HI.U32
//   2  IMAD.WIDE.U32 (64-bit accumulate in place)                        <- roofline peak
//   3  IMAD.WIDE.U32 / IMAD.WIDE.U32.X carry chains (inline PTX, 4-link chains)
//   4  variant 2 + one IADD3 per IMAD.WIDE   (does ALU-pipe work co-issue for free?)
//   5  DFMA
//   6  variant 2 + one DFMA per IMAD.WIDE    (MAC32 counted; do FP64 and integer multiply overlap?)
//   7  variant 2 + one IMAD per IMAD.WIDE    (MAC32 counted)
//   8  heterogeneous warps: half the warps IMAD.WIDE only, half DFMA only (ops of both kinds counted together)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "anemoi_b200.h"

namespace {

constexpr int kIters = 4096;
constexpr int kThreads = 256;
constexpr int kOpsPerIter = 16;

struct Probe {
    long long t0;
    unsigned long long g0;
    __device__ __forceinline__ void start() {
        t0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    }
    __device__ __forceinline__ void stop(long long* cycles) {
        long long t1 = clock64();
        unsigned long long g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            cycles[0] = t1 - t0;
            cycles[1] = (long long)(g1 - g0);
        }
    }
};

#define LOAD_INPUTS()                                               \
    uint32_t a = in[threadIdx.x & 31];                              \
    uint32_t b[16];                                                 \
    _Pragma("unroll") for (int i = 0; i < 16; i++) b[i] = in[32 + i] ^ threadIdx.x;

__global__ void __launch_bounds__(kThreads) k_imad_lo(uint32_t* out, const uint32_t* in, long long* cycles) {
    LOAD_INPUTS();
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = in[64 + i];
    Probe p;
    p.start();
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) r[i] = a * b[i] + r[i];
        a += 0x9e3779b9u;
    }
    p.stop(cycles);
    uint32_t acc = a;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= r[i];
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}

__global__ void __launch_bounds__(kThreads) k_imad_hi(uint32_t* out, const uint32_t* in, long long* cycles) {
    LOAD_INPUTS();
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = in[64 + i];
    Probe p;
    p.start();
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a), "r"(b[i]));
        a += 0x9e3779b9u;
    }
    p.stop(cycles);
    uint32_t acc = a;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= r[i];
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}

// MIX: 0 none, 1 IADD3, 2 DFMA, 3 IMAD lo -- one extra instruction per IMAD.WIDE
template <int MIX>
__global__ void __launch_bounds__(kThreads) k_imad_wide(uint32_t* out, const uint32_t* in, long long* cycles) {
    LOAD_INPUTS();
    unsigned long long q[16];
    uint32_t s[16];
    double d[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        q[i] = in[64 + i];
        s[i] = in[80 + i];
        d[i] = (double)in[96 + i];
    }
    const double da = (double)a * 1e-9, db = (double)b[0] * 1e-9;
    Probe p;
    p.start();
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            q[i] = (unsigned long long)a * b[i] + q[i];
            if (MIX == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(s[i]) : "r"(b[i]));
            if (MIX == 2) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(da), "d"(db));
            if (MIX == 3) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(s[i]) : "r"(a), "r"(b[i]));
        }
        a += 0x9e3779b9u;
    }
    p.stop(cycles);
    unsigned long long acc = a;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= q[i] ^ s[i] ^ (unsigned long long)d[i];
    if (acc == 0x12345678u) out[threadIdx.x] = (uint32_t)acc;
}

__global__ void __launch_bounds__(kThreads) k_imad_chain(uint32_t* out, const uint32_t* in, long long* cycles) {
    LOAD_INPUTS();
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = in[64 + i];
    Probe p;
    p.start();
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t* t = r + 8 * c;
            asm volatile(
                "mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                "madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;"
                "madc.lo.cc.u32 %4, %8, %11, %4; madc.hi.cc.u32 %5, %8, %11, %5;"
                "madc.lo.cc.u32 %6, %8, %12, %6; madc.hi.u32 %7, %8, %12, %7;"
                : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7])
                : "r"(a), "r"(b[4 * c]), "r"(b[4 * c + 1]), "r"(b[4 * c + 2]), "r"(b[4 * c + 3]));
        }
        a += 0x9e3779b9u;
    }
    p.stop(cycles);
    uint32_t acc = a;
#pragma unroll
    for (int i = 0; i < 32; i++) acc ^= r[i];
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}

__global__ void __launch_bounds__(kThreads) k_dfma(uint32_t* out, const uint32_t* in, long long* cycles) {
    LOAD_INPUTS();
    double d[16];
#pragma unroll
    for (int i = 0; i < 16; i++) d[i] = (double)in[96 + i];
    const double da = (double)a * 1e-9, db = (double)b[0] * 1e-9;
    Probe p;
    p.start();
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(da), "d"(db));
    }
    p.stop(cycles);
    double acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc += d[i];
    if (acc == 0.12345678) out[threadIdx.x] = (uint32_t)acc;
}

// Heterogeneous warps: even warps run only IMAD.WIDE, odd warps only DFMA (16 independent accumulators each), so that
// neither instruction stream waits on the other inside a warp. If the FP64 pipe and the integer-multiply (FMA-heavy)
// pipe were independent, the two halves would each run at their stand-alone rate; if they share an issue port the
// SUM of port-cycles stays at 100 %. Counted: 16 ops per iteration per thread, whatever the flavour.
__global__ void __launch_bounds__(kThreads) k_hetero(uint32_t* out, const uint32_t* in, long long* cycles) {
    LOAD_INPUTS();
    unsigned long long q[16];
    double d[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        q[i] = in[64 + i];
        d[i] = (double)in[96 + i];
    }
    const double da = (double)a * 1e-9, db = (double)b[0] * 1e-9;
    const bool fp = (threadIdx.x >> 5) & 1;
    Probe p;
    p.start();
    if (fp) {
#pragma unroll 1
        for (int it = 0; it < kIters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(da), "d"(db));
        }
    } else {
#pragma unroll 1
        for (int it = 0; it < kIters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) q[i] = (unsigned long long)a * b[i] + q[i];
            a += 0x9e3779b9u;
        }
    }
    p.stop(cycles);
    unsigned long long acc = a;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= q[i] ^ (unsigned long long)d[i];
    if (acc == 0x12345678u) out[threadIdx.x] = (uint32_t)acc;
}

typedef void (*kernel_t)(uint32_t*, const uint32_t*, long long*);

int run_variant(kernel_t kernel, double* ops_per_s, double* sm_mhz) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return ANEMOI_B200_ERR_NO_DEVICE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, 0);
    if (occ < 1) occ = 1;
    const int blocks = sms * occ * 2;  // two full waves
    uint32_t *d_in = nullptr, *d_out = nullptr;
    long long* d_cyc = nullptr;
    if (cudaMalloc(&d_in, 4096) != cudaSuccess || cudaMalloc(&d_out, 4096) != cudaSuccess ||
        cudaMalloc(&d_cyc, 64) != cudaSuccess)
        return ANEMOI_B200_ERR_NOMEM;
    uint32_t h[1024];
    for (int i = 0; i < 1024; i++) h[i] = (0x9E3779B9u * (i + 1)) | 1u;
    cudaMemcpy(d_in, h, 4096, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) kernel<<<blocks, kThreads>>>(d_out, d_in, d_cyc);
    const int reps = 5;
    cudaEventRecord(e0);
    for (int w = 0; w < reps; w++) kernel<<<blocks, kThreads>>>(d_out, d_in, d_cyc);
    cudaEventRecord(e1);
    int rc = ANEMOI_B200_OK;
    if (cudaEventSynchronize(e1) != cudaSuccess) rc = ANEMOI_B200_ERR_CUDA;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cyc[2] = {0, 1};
    cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_in);
    cudaFree(d_out);
    cudaFree(d_cyc);
    *ops_per_s = (double)reps * blocks * kThreads * (double)kIters * kOpsPerIter / (ms * 1e-3);
    *sm_mhz = (double)cyc[0] / (double)(cyc[1] > 0 ? cyc[1] : 1) * 1e3;  // SM cycles per ns of one block's loop
    return rc;
}

}  // namespace

extern "C" int anemoi_b200_imad_peak(int variant, double* ops_per_s, double* sm_mhz) {
    if (!ops_per_s || !sm_mhz) return ANEMOI_B200_ERR_ARG;
    switch (variant) {
        case 0: return run_variant(k_imad_lo, ops_per_s, sm_mhz);
        case 1: return run_variant(k_imad_hi, ops_per_s, sm_mhz);
        case 2: return run_variant(k_imad_wide<0>, ops_per_s, sm_mhz);
        case 3: return run_variant(k_imad_chain, ops_per_s, sm_mhz);
        case 4: return run_variant(k_imad_wide<1>, ops_per_s, sm_mhz);
        case 5: return run_variant(k_dfma, ops_per_s, sm_mhz);
        case 6: return run_variant(k_imad_wide<2>, ops_per_s, sm_mhz);
        case 7: return run_variant(k_imad_wide<3>, ops_per_s, sm_mhz);
        case 8: return run_variant(k_hetero, ops_per_s, sm_mhz);
        default: return ANEMOI_B200_ERR_ARG;
    }
}
