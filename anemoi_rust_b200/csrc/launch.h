// Internal: per-field kernel launchers (one TU per field), called by the C ABI in api.cu.
#pragma once
#include <cuda_runtime.h>

#include "kernel_args.h"

extern "C" {
cudaError_t anemoi_launch_bls12_377(int cols, const anemoi::KernelArgs* a, cudaStream_t s);
cudaError_t anemoi_launch_bls12_381(int cols, const anemoi::KernelArgs* a, cudaStream_t s);
cudaError_t anemoi_launch_bn_254(int cols, const anemoi::KernelArgs* a, cudaStream_t s);
cudaError_t anemoi_launch_ed_on_bls12_377(int cols, const anemoi::KernelArgs* a, cudaStream_t s);
cudaError_t anemoi_launch_jubjub(int cols, const anemoi::KernelArgs* a, cudaStream_t s);
cudaError_t anemoi_launch_pallas(int cols, const anemoi::KernelArgs* a, cudaStream_t s);
cudaError_t anemoi_launch_vesta(int cols, const anemoi::KernelArgs* a, cudaStream_t s);
}
