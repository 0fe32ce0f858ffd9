// Merkle authentication paths on the device: gather the openings of a batch of leaf indices from a
// retained tree, and re-assemble one tree level of a batch of paths for verification. These are plain
// HBM gathers (a few hundred bytes per query); all hashing stays in the fused Anemoi kernels.
//
// Not in the reference (it has no tree code); the node function is its Jive compression, so a path is the
// list of the (arity - 1) siblings of the queried node at every level, leaf level first, each group in
// left-to-right order with the node's own slot skipped.
#include <cuda_runtime.h>
#include <cstdint>

#include "merkle_aux.h"

namespace {

// one thread per (query, level, sibling): copies one field element (n64 words)
__global__ void gather_paths_kernel(const uint64_t* __restrict__ leaves, const uint64_t* __restrict__ tree,
                                    unsigned long long n_leaves, int arity, int height, int n64,
                                    const uint64_t* __restrict__ indices, unsigned long long n_idx,
                                    uint64_t* __restrict__ paths) {
    const unsigned long long per_query = (unsigned long long)height * (arity - 1);
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_idx * per_query) return;
    const unsigned long long q = t / per_query;
    const int r = (int)(t % per_query);
    const int level = r / (arity - 1);
    const int s = r % (arity - 1);
    unsigned long long idx = indices[q];
    unsigned long long level_nodes = n_leaves, offset = 0;  // offset of `level` inside `tree` (levels >= 1)
    for (int l = 0; l < level; l++) {
        idx /= arity;
        if (l >= 1) offset += level_nodes;
        level_nodes /= arity;
    }
    const int pos = (int)(idx % arity);
    const int k = s < pos ? s : s + 1;
    const unsigned long long sib = idx - pos + k;
    const uint64_t* src = (level == 0) ? leaves + sib * n64 : tree + (offset + sib) * n64;
    uint64_t* dst = paths + t * n64;
    for (int w = 0; w < n64; w++) dst[w] = src[w];
}

// one thread per (query, child slot): states[q][k] = own value (k == pos) or the matching sibling
__global__ void assemble_level_kernel(const uint64_t* __restrict__ cur, const uint64_t* __restrict__ paths,
                                      const uint64_t* __restrict__ indices, int level, int arity, int height, int n64,
                                      unsigned long long n_idx, uint64_t* __restrict__ states) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_idx * arity) return;
    const unsigned long long q = t / arity;
    const int k = (int)(t % arity);
    unsigned long long idx = indices[q];
    for (int l = 0; l < level; l++) idx /= arity;
    const int pos = (int)(idx % arity);
    const uint64_t* src;
    if (k == pos) {
        src = cur + q * n64;
    } else {
        const int s = k < pos ? k : k - 1;
        src = paths + ((q * height + level) * (arity - 1) + s) * n64;
    }
    uint64_t* dst = states + t * n64;
    for (int w = 0; w < n64; w++) dst[w] = src[w];
}

// one thread per element: counts the elements that are not canonical (>= p); p as n64 little-endian u64 limbs
struct Modulus { uint64_t w[6]; };
__global__ void count_noncanonical_kernel(const uint64_t* __restrict__ elems, unsigned long long n, int n64, Modulus p,
                                          unsigned long long* __restrict__ count) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint64_t* e = elems + t * n64;
    bool lt = false, eq = true;  // lexicographic compare from the top limb
    for (int w = n64 - 1; w >= 0; w--) {
        const uint64_t v = e[w];
        if (eq && v < p.w[w]) lt = true;
        if (v != p.w[w]) eq = false;
    }
    if (!lt) atomicAdd(count, 1ULL);
}

}  // namespace

cudaError_t anemoi_aux_count_noncanonical(const uint64_t* elems, unsigned long long n, int n64, const uint64_t* modulus,
                                          unsigned long long* d_count, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess || n == 0) return e;
    Modulus p;
    for (int w = 0; w < 6; w++) p.w[w] = w < n64 ? modulus[w] : 0;
    const int block = 256;
    const unsigned long long blocks = (n + block - 1) / block;
    if (blocks > 0x7fffffffULL) return cudaErrorInvalidValue;
    count_noncanonical_kernel<<<(unsigned)blocks, block, 0, stream>>>(elems, n, n64, p, d_count);
    return cudaGetLastError();
}

cudaError_t anemoi_aux_gather_paths(const uint64_t* leaves, const uint64_t* tree, unsigned long long n_leaves, int arity,
                                    int height, int n64, const uint64_t* indices, unsigned long long n_idx,
                                    uint64_t* paths, cudaStream_t stream) {
    const unsigned long long total = n_idx * height * (arity - 1);
    if (total == 0) return cudaSuccess;
    const int block = 256;
    const unsigned long long blocks = (total + block - 1) / block;
    if (blocks > 0x7fffffffULL) return cudaErrorInvalidValue;
    gather_paths_kernel<<<(unsigned)blocks, block, 0, stream>>>(leaves, tree, n_leaves, arity, height, n64, indices, n_idx,
                                                              paths);
    return cudaGetLastError();
}

cudaError_t anemoi_aux_assemble_level(const uint64_t* cur, const uint64_t* paths, const uint64_t* indices, int level,
                                      int arity, int height, int n64, unsigned long long n_idx, uint64_t* states,
                                      cudaStream_t stream) {
    const unsigned long long total = n_idx * arity;
    if (total == 0) return cudaSuccess;
    const int block = 256;
    const unsigned long long blocks = (total + block - 1) / block;
    if (blocks > 0x7fffffffULL) return cudaErrorInvalidValue;
    assemble_level_kernel<<<(unsigned)blocks, block, 0, stream>>>(cur, paths, indices, level, arity, height, n64, n_idx,
                                                                states);
    return cudaGetLastError();
}
