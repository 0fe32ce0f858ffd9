// Body of one per-field translation unit: instantiates anemoi_kernel<F, 1> (Anemoi-2-1) and
// anemoi_kernel<F, 2> (Anemoi-4-3) for the field selected by ANEMOI_TU_FIELD and exports one launcher.
// Each field is its own TU so that nvcc compiles the seven of them in parallel and each gets its own
// constant bank for the ARK tables.
#pragma once
#include <atomic>

#include "anemoi_kernels.cuh"
#include "launch.h"

namespace anemoi {

// SM count of the current device, looked up once per device (the attribute query costs microseconds, which is what
// the tiny upper Merkle levels consist of).
static inline int cached_sm_count() {
    static std::atomic<int> cache[64];
    int device = 0;
    cudaGetDevice(&device);
    if (device < 0 || device >= 64) return 148;
    int sms = cache[device].load(std::memory_order_relaxed);
    if (sms == 0) {
        sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        cache[device].store(sms, std::memory_order_relaxed);
    }
    return sms;
}

template <class F, int COLS>
static cudaError_t launch_one(const KernelArgs& a, cudaStream_t stream) {
    if (a.n == 0) return cudaSuccess;
    const unsigned long long threads = a.n * COLS;
    // Big batches: F::BLOCK-thread blocks, F::MIN_BLOCKS resident per SM. Small batches (upper Merkle levels, KATs): one
    // warp per block so the few warps spread over all SMs; at <= 2 warps per SM sub-partition the latency form runs.
    const int sms = cached_sm_count();
    const int block = (threads >= (unsigned long long)sms * F::BLOCK * 2) ? F::BLOCK : 32;
    const unsigned long long blocks = (threads + block - 1) / block;
    if (blocks > 0x7fffffffULL) return cudaErrorInvalidValue;
    constexpr bool kHasLatencyForm = (COLS == 1 ? F::CARRY_CHAIN_2_1 : F::CARRY_CHAIN_4_3);
    if (a.mode >= MODE_LAYER_ARK) {
        anemoi_layer_kernel<F, COLS><<<(unsigned)blocks, block, 0, stream>>>(a);
    } else if (kHasLatencyForm && block == 32 && blocks <= (unsigned long long)sms * 8) {
        if constexpr (kHasLatencyForm) anemoi_kernel_lat<F, COLS><<<(unsigned)blocks, block, 0, stream>>>(a);
    } else {
        anemoi_kernel<F, COLS><<<(unsigned)blocks, block, 0, stream>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace anemoi

#define ANEMOI_DEFINE_LAUNCHER(FIELD)                                                                 \
    extern "C" cudaError_t anemoi_launch_##FIELD(int cols, const anemoi::KernelArgs* a, cudaStream_t s) { \
        return cols == 1 ? anemoi::launch_one<anemoi::F_##FIELD, 1>(*a, s)                               \
                         : anemoi::launch_one<anemoi::F_##FIELD, 2>(*a, s);                              \
    }
