/*
 * anemoi_b200.h -- C ABI of libanemoi_b200.so, the B200 (sm_100a) batched Anemoi engine.
 *
 * This is the drop-in boundary for the batched hot path of anemoi-hash/anemoi-rust. The reference has
 * no FFI of its own: its API is per-item trait functions on zero-sized types. Each entry point below
 * is the batched form of one of those functions and cites the reference item it replaces (paths are
 * relative to the reference crate root). INTEGRATION.md shows the Rust `extern "C"` block and the
 * extension traits a maintainer adds to bind them.
 *
 * DATA LAYOUT (identical to a Rust `&[Felt]` slice of arkworks `Fp<MontBackend<_, N>, N>`):
 *   field element  = N64 little-endian u64 limbs (N64 = 6 for bls12_377/bls12_381, else 4),
 *                    holding a * 2^(64*N64) mod p (Montgomery form), canonical (< p);
 *   state          = W consecutive elements [x_0..x_{c-1}, y_0..y_{c-1}]  (W = 2: Anemoi-2-1, W = 4: Anemoi-4-3);
 *   batch          = n consecutive states / digests, no padding.
 * Inputs >= p are not checked (an arkworks Fp cannot hold them): no UB, result unspecified.
 *
 * OWNERSHIP / THREADING: the caller owns every buffer; the library borrows it for the duration of the
 * call. Host-pointer calls are synchronous (H2D, kernels, D2H inside the call). `_dev` calls take
 * device pointers on the CURRENT device and are ordered on `stream` (a cudaStream_t passed as void*;
 * NULL = legacy default stream); they do not synchronize. All calls are re-entrant and thread-safe.
 *
 * ERRORS: 0 on success, a negative ANEMOI_B200_ERR_* otherwise. The reference panics (assert!) on bad
 * lengths / arities; the ABI returns ERR_LENGTH / ERR_ARITY for the same conditions and never aborts.
 * There is NO CPU fallback: without a usable CUDA device every compute entry returns ERR_NO_DEVICE.
 */
#ifndef ANEMOI_B200_H
#define ANEMOI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANEMOI_B200_VERSION 100 /* 0.1.0 */

/* field ids: module order of src/lib.rs:27-64 */
#define ANEMOI_FIELD_BLS12_377 0       /* src/bls12_377/mod.rs:1        ark_bls12_377::Fq, 6 limbs */
#define ANEMOI_FIELD_BLS12_381 1       /* src/bls12_381/mod.rs:1        ark_bls12_381::Fq, 6 limbs */
#define ANEMOI_FIELD_BN_254 2          /* src/bn_254/mod.rs:1           ark_bn254::Fq,     4 limbs */
#define ANEMOI_FIELD_ED_ON_BLS12_377 3 /* src/ed_on_bls12_377/mod.rs:1  ark_bls12_377::Fr, 4 limbs */
#define ANEMOI_FIELD_JUBJUB 4          /* src/jubjub/mod.rs:1           ark_bls12_381::Fr, 4 limbs */
#define ANEMOI_FIELD_PALLAS 5          /* src/pallas/mod.rs:3           ark_pallas::Fq,    4 limbs */
#define ANEMOI_FIELD_VESTA 6           /* src/vesta/mod.rs:3            ark_pallas::Fr,    4 limbs */
#define ANEMOI_NUM_FIELDS 7

/* instantiation ids */
#define ANEMOI_INST_2_1 0 /* src/<field>/anemoi_2_1/mod.rs:20-38: width 2, rate 1, 1 column  */
#define ANEMOI_INST_4_3 1 /* src/<field>/anemoi_4_3/mod.rs:20-38: width 4, rate 3, 2 columns */

#define ANEMOI_B200_OK 0
#define ANEMOI_B200_ERR_ARG (-1)       /* null pointer, bad device index, bad variant */
#define ANEMOI_B200_ERR_FIELD (-2)     /* field id out of range */
#define ANEMOI_B200_ERR_INST (-3)      /* instantiation id out of range */
#define ANEMOI_B200_ERR_ARITY (-4)     /* k / arity the reference asserts against (hasher.rs:107, 4-3 :163-165) */
#define ANEMOI_B200_ERR_LENGTH (-5)    /* length not a whole number of states / not a power of the arity */
#define ANEMOI_B200_ERR_CUDA (-6)      /* a CUDA runtime call failed; see anemoi_b200_last_cuda_error() */
#define ANEMOI_B200_ERR_NO_DEVICE (-7) /* no CUDA device: there is no CPU fallback */
#define ANEMOI_B200_ERR_NOMEM (-8)     /* device or pinned-host allocation failed */
#define ANEMOI_B200_ERR_NCCL (-9)      /* an NCCL call failed, or libnccl.so.2 could not be loaded at run time */

/* ---- introspection ------------------------------------------------------------------------- */
int anemoi_b200_version(void);
const char* anemoi_b200_strerror(int code);
const char* anemoi_b200_last_cuda_error(void); /* text of the calling thread's last CUDA failure */
int anemoi_b200_device_count(void);            /* 0 when there is no usable device */
int anemoi_b200_field_limbs(int field);        /* N64 (4 or 6), or ERR_FIELD */
int anemoi_b200_state_width(int inst);         /* STATE_WIDTH: 2 or 4 (src/<field>/anemoi_x/mod.rs:20) */
int anemoi_b200_rate_width(int inst);          /* RATE_WIDTH: 1 or 3 (mod.rs:22) */
int anemoi_b200_num_rounds(int field, int inst); /* NUM_HASH_ROUNDS (mod.rs:31-32) */
const char* anemoi_b200_field_name(int field); /* "bls12_377", ... (src/lib.rs module names) */
/* Device buffers of the host-pointer calls come from a memory pool owned by this library (one per device, kept
 * between calls). This returns all but keep_bytes of the cached memory of `device` to the driver. */
int anemoi_b200_pool_trim(int device, size_t keep_bytes);
/* Grows the pool of `device` to at least `bytes` ahead of time, so that the first big host-pointer call of a process
 * does not pay for it (mapping 3 GiB of fresh HBM took 0.1 - 1.3 s on the boxes measured). Optional. */
int anemoi_b200_pool_reserve(int device, size_t bytes);

/* ---- host-pointer entry points (synchronous) ----------------------------------------------- */

/* Anemoi::permutation (src/traits.rs:370-378) on n states, in place. */
int anemoi_b200_permute(int field, int inst, uint64_t* states, size_t n, int device);

/* Anemoi::sbox_layer (src/traits.rs:328-358) on n states, in place. Diagnostic entry: it is what the
 * reference's test_sbox vectors (src/<field>/anemoi_x/mod.rs test_sbox) pin. */
int anemoi_b200_sbox_layer(int field, int inst, uint64_t* states, size_t n, int device);

/* One layer of the round function on n states, in place (diagnostic; the reference exposes these as trait
 * methods): layer 0 = Anemoi::ark_layer(state, round) (src/traits.rs:113-125), 1 = Anemoi::mds_layer
 * (:129-157), 2 = Anemoi::sbox_layer (:328-358), 3 = Anemoi::round(state, round) (:361-367).
 * round >= NUM_ROUNDS -> ERR_LENGTH (the reference asserts, traits.rs:115). */
int anemoi_b200_layer(int field, int inst, int layer, int round, uint64_t* states, size_t n, int device);

/* Jive::compress / Jive::compress_k on n states.
 *   inst 2-1: k must be 2 (src/<field>/anemoi_2_1/hasher.rs:96-110); out = n elements.
 *   inst 4-3: k in {2, 4} (src/<field>/anemoi_4_3/hasher.rs:148-179); out = n * (4/k) elements.
 * k = 2 is Jive::compress. Any other k -> ERR_ARITY (the reference panics). */
int anemoi_b200_compress(int field, int inst, int k, const uint64_t* in, uint64_t* out, size_t n, int device);

/* Same, with the batch split into n_gpus contiguous slices, one per device 0..n_gpus-1, processed concurrently
 * (one host thread + stream per device; independent states need no collective). */
int anemoi_b200_compress_multi(int field, int inst, int k, const uint64_t* in, uint64_t* out, size_t n, int n_gpus);

/* Sponge::hash_field (2-1: src/<field>/anemoi_2_1/hasher.rs:68-85; 4-3: anemoi_4_3/hasher.rs:93-129)
 * on n_msgs messages of felts_per_msg elements each (message i at elems[i*felts_per_msg ...]).
 * digests = n_msgs elements. felts_per_msg may be 0. */
int anemoi_b200_hash_field(int field, int inst, const uint64_t* elems, size_t n_msgs, size_t felts_per_msg,
                           uint64_t* digests, int device);

/* Ragged form: message i is elems[offsets[i] .. offsets[i+1]) (element indices; n_msgs + 1 offsets,
 * non-decreasing, offsets[0] may be non-zero). */
int anemoi_b200_hash_field_ragged(int field, int inst, const uint64_t* elems, const uint64_t* offsets, size_t n_msgs,
                                  uint64_t* digests, int device);

/* Sponge::hash (2-1: anemoi_2_1/hasher.rs:18-66; 4-3: anemoi_4_3/hasher.rs:18-91) on n_msgs byte
 * strings of bytes_per_msg bytes each: 31- (47-) byte little-endian chunks, the 0x01 pad rule, and the
 * canonical -> Montgomery conversion all happen on the device. */
int anemoi_b200_hash_bytes(int field, int inst, const uint8_t* bytes, size_t n_msgs, size_t bytes_per_msg,
                           uint64_t* digests, int device);

/* Ragged form of anemoi_b200_hash_bytes: message i is bytes[offsets[i] .. offsets[i+1]) (n_msgs + 1 byte offsets). */
int anemoi_b200_hash_bytes_ragged(int field, int inst, const uint8_t* bytes, const uint64_t* offsets, size_t n_msgs,
                                  uint64_t* digests, int device);

/* Sponge::merge on n digest pairs (in = 2n elements, out = n elements).
 *   2-1: Jive compress (anemoi_2_1/hasher.rs:87-92).
 *   4-3: the reference's sponge merge, which copies digests[0] twice and never reads digests[1]
 *        (anemoi_4_3/hasher.rs:131-144) -- reproduced literally. */
int anemoi_b200_merge(int field, int inst, const uint64_t* digest_pairs, uint64_t* out, size_t n, int device);

/* Jive Merkle root (the reference has no tree code; node function = its Jive compress):
 *   arity 2 on inst 2-1: node = compress([l, r]);  arity 4 on inst 4-3: node = compress_k([a,b,c,d], 4).
 * n_leaves must be arity^h, h >= 0 (h = 0: root = the leaf). Built level by level on the device;
 * n_gpus > 1 splits the leaves into n_gpus contiguous slices (n_gpus must be a power of two <=
 * device_count and divide n_leaves into whole sub-trees or whole groups of sub-trees), reduces each
 * slice on its own GPU concurrently (one host thread + stream per device), all-gathers the partial roots with
 * NCCL (single-process communicator, ncclCommInitAll, created once) and finishes the top levels on every device. */
int anemoi_b200_merkle_root(int field, int inst, int arity, const uint64_t* leaves, size_t n_leaves, uint64_t* root,
                            int n_gpus);

/* Opt-in input check: *count = how many of the n elements are NOT canonical (>= p). The compute entries do not check
 * their inputs (an arkworks Fp cannot hold a non-canonical value, and the lazy-reduction bounds of the kernels assume
 * canonical input); a host that builds limb arrays by other means can run this first. */
int anemoi_b200_count_noncanonical(int field, const uint64_t* elems, size_t n, uint64_t* count, int device);

/* AnemoiDigest::to_bytes (src/<field>/anemoi_x/digest.rs:42-46): n Montgomery elements -> n canonical
 * little-endian byte strings of 8*N64 bytes each (de-Montgomery on the device). */
int anemoi_b200_digest_to_bytes(int field, const uint64_t* digests, uint8_t* bytes, size_t n, int device);

/* ---- device-pointer, stream-ordered entry points (current device) -------------------------- */
int anemoi_b200_permute_dev(int field, int inst, uint64_t* d_states, size_t n, void* stream);
int anemoi_b200_sbox_layer_dev(int field, int inst, uint64_t* d_states, size_t n, void* stream);
int anemoi_b200_layer_dev(int field, int inst, int layer, int round, uint64_t* d_states, size_t n, void* stream);
int anemoi_b200_compress_dev(int field, int inst, int k, const uint64_t* d_in, uint64_t* d_out, size_t n, void* stream);
int anemoi_b200_hash_field_dev(int field, int inst, const uint64_t* d_elems, size_t n_msgs, size_t felts_per_msg,
                               uint64_t* d_digests, void* stream);
int anemoi_b200_hash_field_ragged_dev(int field, int inst, const uint64_t* d_elems, const uint64_t* d_offsets,
                                      size_t n_msgs, uint64_t* d_digests, void* stream);
int anemoi_b200_hash_bytes_dev(int field, int inst, const uint8_t* d_bytes, size_t n_msgs, size_t bytes_per_msg,
                               uint64_t* d_digests, void* stream);
int anemoi_b200_hash_bytes_ragged_dev(int field, int inst, const uint8_t* d_bytes, const uint64_t* d_offsets,
                                      size_t n_msgs, uint64_t* d_digests, void* stream);
int anemoi_b200_merge_dev(int field, int inst, const uint64_t* d_pairs, uint64_t* d_out, size_t n, void* stream);
int anemoi_b200_digest_to_bytes_dev(int field, const uint64_t* d_digests, uint8_t* d_bytes, size_t n, void* stream);
int anemoi_b200_count_noncanonical_dev(int field, const uint64_t* d_elems, size_t n, uint64_t* d_count, void* stream);

/* Reduce `levels` levels of a Jive Merkle tree on the device: d_leaves (n_leaves elements, not
 * modified) -> d_out (n_leaves / arity^levels elements). d_scratch must hold at least
 * anemoi_b200_merkle_scratch_felts(arity, n_leaves) elements (it may be NULL when levels <= 1).
 * n_leaves must be a multiple of arity^levels. This is the per-GPU leg of a sharded tree: each rank
 * reduces its slice to its partial roots, the ranks all-gather them (NCCL), and every rank finishes the
 * top levels with one more call. */
int anemoi_b200_merkle_reduce_dev(int field, int inst, int arity, const uint64_t* d_leaves, size_t n_leaves, int levels,
                                  uint64_t* d_scratch, uint64_t* d_out, void* stream);
size_t anemoi_b200_merkle_scratch_felts(int arity, size_t n_leaves);

/* ---- sharded Merkle root over NCCL (north_star: "only the subtree roots are gathered over NVLink,
 * a single tiny NCCL allgather") ---------------------------------------------------------------------------------
 * One rank per GPU (one process or one host thread each). Rank g holds the contiguous slice
 * [g * n_local, (g + 1) * n_local) of the leaves on its device, reduces it while it consists of whole sub-trees,
 * all-gathers the <= 2 partial roots per rank (<= 96 bytes) with ONE ncclAllGather on `stream`, and finishes the
 * <= 2 top levels redundantly, so every rank ends with the same root in d_root (1 element). Collective: every rank
 * of the communicator must call it with the same field/inst/arity/n_local. nccl_comm is an ncclComm_t (from the
 * caller's own NCCL, or from anemoi_b200_comm_init_rank); NULL = single rank (whole tree on this device).
 * d_scratch: anemoi_b200_merkle_sharded_scratch_felts(arity, n_local, nranks) elements, or NULL to let the library
 * take it from its stream-ordered pool. ERR_LENGTH when nranks * n_local is not a power of the arity or the ranks do
 * not split the tree into whole sub-trees. NCCL is bound at run time (libnccl.so.2; a copy already loaded by the
 * process, e.g. PyTorch's, is reused): ERR_NCCL if it cannot be loaded. */
int anemoi_b200_merkle_root_sharded_dev(int field, int inst, int arity, const uint64_t* d_local_leaves, size_t n_local,
                                        void* nccl_comm, uint64_t* d_scratch, uint64_t* d_root, void* stream);
size_t anemoi_b200_merkle_sharded_scratch_felts(int arity, size_t n_local, int nranks);
/* Communicator helpers so that a host without its own NCCL binding (the Rust shim) can drive the sharded build:
 * rank 0 calls _unique_id and ships the 128 bytes to the other ranks by any means; every rank then calls
 * _init_rank with its CUDA device current (collective). _comm_info reports (nranks, rank). */
int anemoi_b200_nccl_version(void); /* e.g. 22809; 0 when libnccl.so.2 is not loadable */
int anemoi_b200_comm_unique_id(uint8_t* id128);
int anemoi_b200_comm_init_rank(const uint8_t* id128, int nranks, int rank, void** comm);
int anemoi_b200_comm_info(void* comm, int* nranks, int* rank);
int anemoi_b200_comm_destroy(void* comm);

/* ---- Merkle trees with retained levels, authentication paths, batches of trees -------
 * Not in the reference (no tree code); node function as in anemoi_b200_merkle_root. Layouts:
 *   tree  = every level above the leaves, level 1 first (n/arity nodes), then level 2, ..., the root last:
 *           anemoi_b200_merkle_tree_felts(arity, n_leaves) = (n_leaves - 1) / (arity - 1) elements;
 *   path  = for one leaf index, height * (arity - 1) elements: the siblings of the queried node at each
 *           level, leaf level first, left to right with the node's own slot skipped.
 * A batch of equal-size trees stored back to back is reduced with anemoi_b200_merkle_reduce_dev
 * (levels = height): it returns one root per tree. */
size_t anemoi_b200_merkle_tree_felts(int arity, size_t n_leaves);
/* Build the whole tree over d_leaves (n_leaves = arity^h) into d_tree; the root is its last element. */
int anemoi_b200_merkle_tree_dev(int field, int inst, int arity, const uint64_t* d_leaves, size_t n_leaves,
                                uint64_t* d_tree, void* stream);
/* Openings of n_idx leaf indices (d_indices: u64, each < n_leaves) -> d_paths (n_idx paths). */
int anemoi_b200_merkle_open_dev(int field, int inst, int arity, const uint64_t* d_leaves, const uint64_t* d_tree,
                                size_t n_leaves, const uint64_t* d_indices, size_t n_idx, uint64_t* d_paths, void* stream);
/* Recompute the root implied by each (leaf value, index, path): d_roots (n_idx elements). d_scratch holds
 * n_idx * (arity + 1) elements. The caller compares the results with the committed root. */
int anemoi_b200_merkle_verify_dev(int field, int inst, int arity, const uint64_t* d_leaf_values, const uint64_t* d_indices,
                                  const uint64_t* d_paths, int height, size_t n_idx, uint64_t* d_scratch, uint64_t* d_roots,
                                  void* stream);
/* Host-pointer forms: build the tree on `device`, return the root and the openings of `indices`. */
int anemoi_b200_merkle_open(int field, int inst, int arity, const uint64_t* leaves, size_t n_leaves,
                            const uint64_t* indices, size_t n_idx, uint64_t* root, uint64_t* paths, int device);
int anemoi_b200_merkle_verify(int field, int inst, int arity, const uint64_t* leaf_values, const uint64_t* indices,
                              const uint64_t* paths, int height, size_t n_idx, uint64_t* roots, int device);

/* ---- roofline denominator ------------------------------------------------------------------ */
/* Chip-wide issue rate of one integer-multiply flavour on the current device, measured with
 * independent chains. variant: 0 mad.lo.u32, 1 mad.hi.u32, 2 mad.wide.u32 (the roofline peak), 3 carry-chained
 * mad.lo.cc/madc.hi.cc pairs (= IMAD.WIDE.U32.X, what the kernels issue), 4 = 2 + one add.u32 per multiply,
 * 5 fma.rn.f64, 6 = 2 + one DFMA per multiply, 7 = 2 + one IMAD per multiply, 8 = half the warps IMAD.WIDE only and
 * half DFMA only (6-8: do the FP64 and integer-multiply pipes overlap? -- DESIGN.md 3.3).
 * ops_per_s = instructions (MAC32 for 2-4, 6, 7) per second; sm_mhz = SM clock seen. */
int anemoi_b200_imad_peak(int variant, double* ops_per_s, double* sm_mhz);

#ifdef __cplusplus
}
#endif
#endif /* ANEMOI_B200_H */
