// Internal: device gathers used by the Merkle opening / verification entry points (merkle_aux.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

cudaError_t anemoi_aux_gather_paths(const uint64_t* leaves, const uint64_t* tree, unsigned long long n_leaves, int arity,
                                    int height, int n64, const uint64_t* indices, unsigned long long n_idx,
                                    uint64_t* paths, cudaStream_t stream);
cudaError_t anemoi_aux_assemble_level(const uint64_t* cur, const uint64_t* paths, const uint64_t* indices, int level,
                                      int arity, int height, int n64, unsigned long long n_idx, uint64_t* states,
                                      cudaStream_t stream);
// *d_count = number of elements of elems[0..n) that are >= the modulus (n64 little-endian u64 limbs)
cudaError_t anemoi_aux_count_noncanonical(const uint64_t* elems, unsigned long long n, int n64, const uint64_t* modulus,
                                          unsigned long long* d_count, cudaStream_t stream);
