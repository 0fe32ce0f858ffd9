//! Rust side of the drop-in boundary: `extern "C"` bindings of libanemoi_b200.so (include/anemoi_b200.h) and
//! batched extension traits implemented for the reference crate's own marker types.
//!
//! SOURCE ONLY: this image has no cargo/rustc, so this file has not been compiled here. It is what a maintainer adds
//! to anemoi-rust as `src/b200.rs` (+ `pub mod b200;` in src/lib.rs behind a `b200` feature; the module carries
//! `#![allow(unsafe_code)]` because src/lib.rs:13 denies unsafe crate-wide). INTEGRATION.md walks through it.
//! tests/test_rust_shim_abi.py keeps the `extern "C"` block below in lock-step with the header.
//!
//! Surface -- the batched forms of the reference's traits, same types, same panics:
//!   Jive::compress / compress_k        (src/traits.rs:23-33)  -> compress_batch(&[F]) / compress_k_batch(&[F], k)
//!   Sponge::hash_field                 (src/traits.rs:16)     -> hash_field_batch(&[&[F]]) -> Vec<Digest>  (ragged)
//!                                                                hash_field_batch_fixed(&[F], len) -> Vec<Digest>
//!   Sponge::hash                       (src/traits.rs:13)     -> hash_batch(&[&[u8]]) -> Vec<Digest>
//!   Sponge::merge                      (src/traits.rs:19)     -> merge_batch(&[[Digest; 2]]) -> Vec<Digest>
//!   Anemoi::permutation                (src/traits.rs:370)    -> permutation_batch(&mut [F])
//!   AnemoiDigest::to_bytes             (digest.rs:42-46)      -> digests_to_bytes(&[Digest]) -> Vec<u8>
//!   (new) Jive Merkle trees            -> merkle_root(&[Digest], n_gpus) / merkle_open / merkle_verify
//!   (new) device-resident, stream-ordered forms (`*_dev`) and the NCCL-sharded root -> `dev` sub-module
//!
//! No conversion happens at the boundary: an arkworks `Fp<MontBackend<_, N>, N>` is `BigInt<N>([u64; N])` plus a
//! zero-sized marker, i.e. N little-endian u64 Montgomery limbs -- exactly what the C ABI reads and writes. That
//! layout is asserted at compile time below (size) and pinned by `tests::layout_is_montgomery_limbs` under `cargo test`.
#![allow(unsafe_code)]

use ark_ff::PrimeField;
use core::ffi::{c_char, c_int, c_void};

#[link(name = "anemoi_b200")]
extern "C" {
    // BEGIN GENERATED FFI (tools/gen_rust_ffi.py)
    pub fn anemoi_b200_version() -> c_int;
    pub fn anemoi_b200_strerror(code: c_int) -> *const c_char;
    pub fn anemoi_b200_last_cuda_error() -> *const c_char;
    pub fn anemoi_b200_device_count() -> c_int;
    pub fn anemoi_b200_field_limbs(field: c_int) -> c_int;
    pub fn anemoi_b200_state_width(inst: c_int) -> c_int;
    pub fn anemoi_b200_rate_width(inst: c_int) -> c_int;
    pub fn anemoi_b200_num_rounds(field: c_int, inst: c_int) -> c_int;
    pub fn anemoi_b200_field_name(field: c_int) -> *const c_char;
    pub fn anemoi_b200_pool_trim(device: c_int, keep_bytes: usize) -> c_int;
    pub fn anemoi_b200_pool_reserve(device: c_int, bytes: usize) -> c_int;
    pub fn anemoi_b200_permute(field: c_int, inst: c_int, states: *mut u64, n: usize, device: c_int) -> c_int;
    pub fn anemoi_b200_sbox_layer(field: c_int, inst: c_int, states: *mut u64, n: usize, device: c_int) -> c_int;
    pub fn anemoi_b200_layer(field: c_int, inst: c_int, layer: c_int, round: c_int, states: *mut u64, n: usize, device: c_int) -> c_int;
    pub fn anemoi_b200_compress(field: c_int, inst: c_int, k: c_int, input: *const u64, out: *mut u64, n: usize, device: c_int) -> c_int;
    pub fn anemoi_b200_compress_multi(field: c_int, inst: c_int, k: c_int, input: *const u64, out: *mut u64, n: usize, n_gpus: c_int) -> c_int;
    pub fn anemoi_b200_hash_field(field: c_int, inst: c_int, elems: *const u64, n_msgs: usize, felts_per_msg: usize, digests: *mut u64, device: c_int) -> c_int;
    pub fn anemoi_b200_hash_field_ragged(field: c_int, inst: c_int, elems: *const u64, offsets: *const u64, n_msgs: usize, digests: *mut u64, device: c_int) -> c_int;
    pub fn anemoi_b200_hash_bytes(field: c_int, inst: c_int, bytes: *const u8, n_msgs: usize, bytes_per_msg: usize, digests: *mut u64, device: c_int) -> c_int;
    pub fn anemoi_b200_hash_bytes_ragged(field: c_int, inst: c_int, bytes: *const u8, offsets: *const u64, n_msgs: usize, digests: *mut u64, device: c_int) -> c_int;
    pub fn anemoi_b200_merge(field: c_int, inst: c_int, digest_pairs: *const u64, out: *mut u64, n: usize, device: c_int) -> c_int;
    pub fn anemoi_b200_merkle_root(field: c_int, inst: c_int, arity: c_int, leaves: *const u64, n_leaves: usize, root: *mut u64, n_gpus: c_int) -> c_int;
    pub fn anemoi_b200_count_noncanonical(field: c_int, elems: *const u64, n: usize, count: *mut u64, device: c_int) -> c_int;
    pub fn anemoi_b200_digest_to_bytes(field: c_int, digests: *const u64, bytes: *mut u8, n: usize, device: c_int) -> c_int;
    pub fn anemoi_b200_permute_dev(field: c_int, inst: c_int, d_states: *mut u64, n: usize, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_sbox_layer_dev(field: c_int, inst: c_int, d_states: *mut u64, n: usize, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_layer_dev(field: c_int, inst: c_int, layer: c_int, round: c_int, d_states: *mut u64, n: usize, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_compress_dev(field: c_int, inst: c_int, k: c_int, d_in: *const u64, d_out: *mut u64, n: usize, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_hash_field_dev(field: c_int, inst: c_int, d_elems: *const u64, n_msgs: usize, felts_per_msg: usize, d_digests: *mut u64, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_hash_field_ragged_dev(field: c_int, inst: c_int, d_elems: *const u64, d_offsets: *const u64, n_msgs: usize, d_digests: *mut u64, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_hash_bytes_dev(field: c_int, inst: c_int, d_bytes: *const u8, n_msgs: usize, bytes_per_msg: usize, d_digests: *mut u64, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_hash_bytes_ragged_dev(field: c_int, inst: c_int, d_bytes: *const u8, d_offsets: *const u64, n_msgs: usize, d_digests: *mut u64, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_merge_dev(field: c_int, inst: c_int, d_pairs: *const u64, d_out: *mut u64, n: usize, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_digest_to_bytes_dev(field: c_int, d_digests: *const u64, d_bytes: *mut u8, n: usize, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_count_noncanonical_dev(field: c_int, d_elems: *const u64, n: usize, d_count: *mut u64, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_merkle_reduce_dev(field: c_int, inst: c_int, arity: c_int, d_leaves: *const u64, n_leaves: usize, levels: c_int, d_scratch: *mut u64, d_out: *mut u64, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_merkle_scratch_felts(arity: c_int, n_leaves: usize) -> usize;
    pub fn anemoi_b200_merkle_root_sharded_dev(field: c_int, inst: c_int, arity: c_int, d_local_leaves: *const u64, n_local: usize, nccl_comm: *mut c_void, d_scratch: *mut u64, d_root: *mut u64, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_merkle_sharded_scratch_felts(arity: c_int, n_local: usize, nranks: c_int) -> usize;
    pub fn anemoi_b200_nccl_version() -> c_int;
    pub fn anemoi_b200_comm_unique_id(id128: *mut u8) -> c_int;
    pub fn anemoi_b200_comm_init_rank(id128: *const u8, nranks: c_int, rank: c_int, comm: *mut *mut c_void) -> c_int;
    pub fn anemoi_b200_comm_info(comm: *mut c_void, nranks: *mut c_int, rank: *mut c_int) -> c_int;
    pub fn anemoi_b200_comm_destroy(comm: *mut c_void) -> c_int;
    pub fn anemoi_b200_merkle_tree_felts(arity: c_int, n_leaves: usize) -> usize;
    pub fn anemoi_b200_merkle_tree_dev(field: c_int, inst: c_int, arity: c_int, d_leaves: *const u64, n_leaves: usize, d_tree: *mut u64, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_merkle_open_dev(field: c_int, inst: c_int, arity: c_int, d_leaves: *const u64, d_tree: *const u64, n_leaves: usize, d_indices: *const u64, n_idx: usize, d_paths: *mut u64, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_merkle_verify_dev(field: c_int, inst: c_int, arity: c_int, d_leaf_values: *const u64, d_indices: *const u64, d_paths: *const u64, height: c_int, n_idx: usize, d_scratch: *mut u64, d_roots: *mut u64, stream: *mut c_void) -> c_int;
    pub fn anemoi_b200_merkle_open(field: c_int, inst: c_int, arity: c_int, leaves: *const u64, n_leaves: usize, indices: *const u64, n_idx: usize, root: *mut u64, paths: *mut u64, device: c_int) -> c_int;
    pub fn anemoi_b200_merkle_verify(field: c_int, inst: c_int, arity: c_int, leaf_values: *const u64, indices: *const u64, paths: *const u64, height: c_int, n_idx: usize, roots: *mut u64, device: c_int) -> c_int;
    pub fn anemoi_b200_imad_peak(variant: c_int, ops_per_s: *mut f64, sm_mhz: *mut f64) -> c_int;
    // END GENERATED FFI
}

/// Field ids: module order of src/lib.rs:27-64.
pub const FIELD_BLS12_377: c_int = 0;
pub const FIELD_BLS12_381: c_int = 1;
pub const FIELD_BN_254: c_int = 2;
pub const FIELD_ED_ON_BLS12_377: c_int = 3;
pub const FIELD_JUBJUB: c_int = 4;
pub const FIELD_PALLAS: c_int = 5;
pub const FIELD_VESTA: c_int = 6;
pub const INST_2_1: c_int = 0;
pub const INST_4_3: c_int = 1;

/// The reference panics on misuse (`assert!` in hasher.rs:97,107; 4-3 :149,163-165); the batched API keeps that contract.
fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { core::ffi::CStr::from_ptr(anemoi_b200_strerror(rc)) }.to_string_lossy().into_owned();
        let detail = unsafe { core::ffi::CStr::from_ptr(anemoi_b200_last_cuda_error()) }.to_string_lossy().into_owned();
        panic!("anemoi_b200 ({}): {} {}", rc, msg, detail);
    }
}

/// Batched forms of `Jive` / `Sponge` / `Anemoi::permutation` for one (field, instantiation).
/// `F` is the module's `Felt`, `Digest` its `AnemoiDigest` (src/<field>/anemoi_*/digest.rs:13).
pub trait B200Batch<F: PrimeField> {
    type Digest: Copy;
    const FIELD: c_int;
    const INST: c_int;
    const STATE_WIDTH: usize;
    /// CUDA device the host-pointer calls run on.
    const DEVICE: c_int = 0;

    fn digest_from_felt(e: F) -> Self::Digest;
    fn digest_to_felt(d: &Self::Digest) -> F;
    /// `AnemoiDigest::digests_to_elements` (digest.rs:32-39).
    fn digests_to_felts(digests: &[Self::Digest]) -> Vec<F> { digests.iter().map(Self::digest_to_felt).collect() }
    fn felts_to_digests(elems: Vec<F>) -> Vec<Self::Digest> { elems.into_iter().map(Self::digest_from_felt).collect() }

    /// `Jive::compress_k` on `elems.len() / STATE_WIDTH` states (src/traits.rs:29-33).
    fn compress_k_batch(elems: &[F], k: usize) -> Vec<F> {
        assert!(elems.len() % Self::STATE_WIDTH == 0);
        assert!(k != 0 && Self::STATE_WIDTH % k == 0 && k % 2 == 0);
        let n = elems.len() / Self::STATE_WIDTH;
        let mut out = vec![F::zero(); n * Self::STATE_WIDTH / k];
        check(unsafe {
            anemoi_b200_compress(Self::FIELD, Self::INST, k as c_int, elems.as_ptr() as *const u64,
                                 out.as_mut_ptr() as *mut u64, n, Self::DEVICE)
        });
        out
    }
    /// `Jive::compress` (src/traits.rs:23-27).
    fn compress_batch(elems: &[F]) -> Vec<F> { Self::compress_k_batch(elems, 2) }
    /// Same, the batch split over `n_gpus` devices of this process (independent states: no collective).
    fn compress_k_batch_multi(elems: &[F], k: usize, n_gpus: usize) -> Vec<F> {
        assert!(elems.len() % Self::STATE_WIDTH == 0);
        assert!(k != 0 && Self::STATE_WIDTH % k == 0 && k % 2 == 0);
        let n = elems.len() / Self::STATE_WIDTH;
        let mut out = vec![F::zero(); n * Self::STATE_WIDTH / k];
        check(unsafe {
            anemoi_b200_compress_multi(Self::FIELD, Self::INST, k as c_int, elems.as_ptr() as *const u64,
                                       out.as_mut_ptr() as *mut u64, n, n_gpus as c_int)
        });
        out
    }

    /// `Anemoi::permutation` on a slice of states, in place (src/traits.rs:370-378).
    fn permutation_batch(states: &mut [F]) {
        assert!(states.len() % Self::STATE_WIDTH == 0);
        check(unsafe {
            anemoi_b200_permute(Self::FIELD, Self::INST, states.as_mut_ptr() as *mut u64,
                                states.len() / Self::STATE_WIDTH, Self::DEVICE)
        });
    }

    /// `Sponge::hash_field` on many messages of different lengths: one launch (src/traits.rs:16).
    fn hash_field_batch(msgs: &[&[F]]) -> Vec<Self::Digest> {
        let mut offsets: Vec<u64> = Vec::with_capacity(msgs.len() + 1);
        let mut flat: Vec<F> = Vec::with_capacity(msgs.iter().map(|m| m.len()).sum());
        offsets.push(0);
        for m in msgs {
            flat.extend_from_slice(m);
            offsets.push(flat.len() as u64);
        }
        let mut out = vec![F::zero(); msgs.len()];
        check(unsafe {
            anemoi_b200_hash_field_ragged(Self::FIELD, Self::INST, flat.as_ptr() as *const u64, offsets.as_ptr(),
                                          msgs.len(), out.as_mut_ptr() as *mut u64, Self::DEVICE)
        });
        Self::felts_to_digests(out)
    }
    /// `Sponge::hash_field` on `elems.len() / felts_per_msg` messages stored back to back.
    fn hash_field_batch_fixed(elems: &[F], felts_per_msg: usize) -> Vec<Self::Digest> {
        assert!(felts_per_msg != 0 && elems.len() % felts_per_msg == 0);
        let n = elems.len() / felts_per_msg;
        let mut out = vec![F::zero(); n];
        check(unsafe {
            anemoi_b200_hash_field(Self::FIELD, Self::INST, elems.as_ptr() as *const u64, n, felts_per_msg,
                                   out.as_mut_ptr() as *mut u64, Self::DEVICE)
        });
        Self::felts_to_digests(out)
    }

    /// `Sponge::hash` on many byte strings (src/traits.rs:13): chunking, padding and the Montgomery conversion run on the device.
    fn hash_batch(msgs: &[&[u8]]) -> Vec<Self::Digest> {
        let mut offsets: Vec<u64> = Vec::with_capacity(msgs.len() + 1);
        let mut flat: Vec<u8> = Vec::with_capacity(msgs.iter().map(|m| m.len()).sum());
        offsets.push(0);
        for m in msgs {
            flat.extend_from_slice(m);
            offsets.push(flat.len() as u64);
        }
        let mut out = vec![F::zero(); msgs.len()];
        check(unsafe {
            anemoi_b200_hash_bytes_ragged(Self::FIELD, Self::INST, flat.as_ptr(), offsets.as_ptr(), msgs.len(),
                                          out.as_mut_ptr() as *mut u64, Self::DEVICE)
        });
        Self::felts_to_digests(out)
    }
    /// `Sponge::hash` on equal-length byte strings stored back to back.
    fn hash_batch_fixed(bytes: &[u8], bytes_per_msg: usize) -> Vec<Self::Digest> {
        assert!(bytes_per_msg != 0 && bytes.len() % bytes_per_msg == 0);
        let n = bytes.len() / bytes_per_msg;
        let mut out = vec![F::zero(); n];
        check(unsafe {
            anemoi_b200_hash_bytes(Self::FIELD, Self::INST, bytes.as_ptr(), n, bytes_per_msg,
                                   out.as_mut_ptr() as *mut u64, Self::DEVICE)
        });
        Self::felts_to_digests(out)
    }

    /// `Sponge::merge` on many digest pairs (src/traits.rs:19). 2-1: Jive; 4-3: the reference's sponge merge, which
    /// reads digests[0] only (anemoi_4_3/hasher.rs:131-144) -- reproduced as written.
    fn merge_batch(pairs: &[[Self::Digest; 2]]) -> Vec<Self::Digest> {
        let mut flat: Vec<F> = Vec::with_capacity(2 * pairs.len());
        for p in pairs {
            flat.push(Self::digest_to_felt(&p[0]));
            flat.push(Self::digest_to_felt(&p[1]));
        }
        let mut out = vec![F::zero(); pairs.len()];
        check(unsafe {
            anemoi_b200_merge(Self::FIELD, Self::INST, flat.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64,
                              pairs.len(), Self::DEVICE)
        });
        Self::felts_to_digests(out)
    }

    /// `AnemoiDigest::to_bytes` on many digests (digest.rs:42-46): canonical little-endian, size_of::<F>() bytes each.
    fn digests_to_bytes(digests: &[Self::Digest]) -> Vec<u8> {
        let flat = Self::digests_to_felts(digests);
        let mut out = vec![0u8; flat.len() * core::mem::size_of::<F>()];
        check(unsafe {
            anemoi_b200_digest_to_bytes(Self::FIELD, flat.as_ptr() as *const u64, out.as_mut_ptr(), flat.len(), Self::DEVICE)
        });
        out
    }

    /// Jive Merkle root of arity STATE_WIDTH (2-1: `compress`; 4-3: `compress_k(.,4)`) over `leaves.len()` = arity^h
    /// leaf digests; `n_gpus` > 1 shards the leaves over that many devices (one NCCL all-gather of the partial roots).
    fn merkle_root(leaves: &[Self::Digest], n_gpus: usize) -> Self::Digest {
        let flat = Self::digests_to_felts(leaves);
        let mut root = F::zero();
        check(unsafe {
            anemoi_b200_merkle_root(Self::FIELD, Self::INST, Self::STATE_WIDTH as c_int, flat.as_ptr() as *const u64,
                                    flat.len(), &mut root as *mut F as *mut u64, n_gpus as c_int)
        });
        Self::digest_from_felt(root)
    }

    /// Build the tree and open `indices`: (root, paths) with `paths[q]` = height * (arity - 1) siblings, leaf level first.
    fn merkle_open(leaves: &[Self::Digest], indices: &[u64]) -> (Self::Digest, Vec<Vec<Self::Digest>>) {
        let flat = Self::digests_to_felts(leaves);
        let arity = Self::STATE_WIDTH;
        let mut height = 0usize;
        let mut m = flat.len();
        while m > 1 {
            assert!(m % arity == 0);
            m /= arity;
            height += 1;
        }
        let per = height * (arity - 1);
        let mut root = F::zero();
        let mut paths = vec![F::zero(); indices.len() * per];
        check(unsafe {
            anemoi_b200_merkle_open(Self::FIELD, Self::INST, arity as c_int, flat.as_ptr() as *const u64, flat.len(),
                                    indices.as_ptr(), indices.len(), &mut root as *mut F as *mut u64,
                                    paths.as_mut_ptr() as *mut u64, Self::DEVICE)
        });
        let out = if per == 0 {
            vec![Vec::new(); indices.len()]
        } else {
            paths.chunks(per).map(|c| Self::felts_to_digests(c.to_vec())).collect()
        };
        (Self::digest_from_felt(root), out)
    }

    /// Recompute the root implied by each (leaf, index, path); the caller compares with the committed root.
    fn merkle_verify(leaf_values: &[Self::Digest], indices: &[u64], paths: &[Vec<Self::Digest>]) -> Vec<Self::Digest> {
        assert!(leaf_values.len() == indices.len() && paths.len() == indices.len());
        let arity = Self::STATE_WIDTH;
        let per = paths.first().map(|p| p.len()).unwrap_or(0);
        assert!(per % (arity - 1) == 0 && paths.iter().all(|p| p.len() == per));
        let vals = Self::digests_to_felts(leaf_values);
        let flat: Vec<F> = paths.iter().flat_map(|p| p.iter().map(Self::digest_to_felt)).collect();
        let mut roots = vec![F::zero(); indices.len()];
        check(unsafe {
            anemoi_b200_merkle_verify(Self::FIELD, Self::INST, arity as c_int, vals.as_ptr() as *const u64, indices.as_ptr(),
                                      flat.as_ptr() as *const u64, (per / (arity - 1)) as c_int, indices.len(),
                                      roots.as_mut_ptr() as *mut u64, Self::DEVICE)
        });
        Self::felts_to_digests(roots)
    }
}

/// Grows the library's device-memory pool (buffers of the host-pointer calls) ahead of the first big call.
pub fn pool_reserve(device: usize, bytes: usize) {
    check(unsafe { anemoi_b200_pool_reserve(device as c_int, bytes) });
}

/// Hands the pool's cached device memory back to the driver, keeping at most `keep_bytes`.
pub fn pool_trim(device: usize, keep_bytes: usize) {
    check(unsafe { anemoi_b200_pool_trim(device as c_int, keep_bytes) });
}

/// Device-resident, stream-ordered forms: raw device pointers to `[F]` / `[u64; N]` limb arrays on the CURRENT CUDA
/// device and a `cudaStream_t` (as `*mut c_void`; null = legacy default stream). They only enqueue work.
pub mod dev {
    use super::*;

    /// An NCCL communicator owned by libanemoi_b200.so (`anemoi_b200_comm_*`): rank 0 draws the id, the host ships the 128
    /// bytes to the other ranks by its own transport, every rank joins with its device current.
    pub struct Comm(pub *mut c_void);
    impl Comm {
        pub fn unique_id() -> [u8; 128] {
            let mut id = [0u8; 128];
            check(unsafe { anemoi_b200_comm_unique_id(id.as_mut_ptr()) });
            id
        }
        pub fn init_rank(id: &[u8; 128], nranks: usize, rank: usize) -> Self {
            let mut c: *mut c_void = core::ptr::null_mut();
            check(unsafe { anemoi_b200_comm_init_rank(id.as_ptr(), nranks as c_int, rank as c_int, &mut c) });
            Comm(c)
        }
        pub fn info(&self) -> (usize, usize) {
            let (mut n, mut r) = (0 as c_int, 0 as c_int);
            check(unsafe { anemoi_b200_comm_info(self.0, &mut n, &mut r) });
            (n as usize, r as usize)
        }
    }
    impl Drop for Comm {
        fn drop(&mut self) { unsafe { anemoi_b200_comm_destroy(self.0); } }
    }

    pub unsafe fn permutation_batch<F: PrimeField, H: B200Batch<F>>(d_states: *mut F, n: usize, stream: *mut c_void) {
        check(anemoi_b200_permute_dev(H::FIELD, H::INST, d_states as *mut u64, n, stream));
    }
    pub unsafe fn compress_k_batch<F: PrimeField, H: B200Batch<F>>(d_in: *const F, d_out: *mut F, n: usize, k: usize, stream: *mut c_void) {
        check(anemoi_b200_compress_dev(H::FIELD, H::INST, k as c_int, d_in as *const u64, d_out as *mut u64, n, stream));
    }
    pub unsafe fn hash_field_batch_fixed<F: PrimeField, H: B200Batch<F>>(d_elems: *const F, n_msgs: usize, felts_per_msg: usize,
                                                                         d_digests: *mut F, stream: *mut c_void) {
        check(anemoi_b200_hash_field_dev(H::FIELD, H::INST, d_elems as *const u64, n_msgs, felts_per_msg, d_digests as *mut u64, stream));
    }
    pub unsafe fn hash_field_batch<F: PrimeField, H: B200Batch<F>>(d_elems: *const F, d_offsets: *const u64, n_msgs: usize,
                                                                   d_digests: *mut F, stream: *mut c_void) {
        check(anemoi_b200_hash_field_ragged_dev(H::FIELD, H::INST, d_elems as *const u64, d_offsets, n_msgs, d_digests as *mut u64, stream));
    }
    pub unsafe fn hash_batch_fixed<F: PrimeField, H: B200Batch<F>>(d_bytes: *const u8, n_msgs: usize, bytes_per_msg: usize,
                                                                   d_digests: *mut F, stream: *mut c_void) {
        check(anemoi_b200_hash_bytes_dev(H::FIELD, H::INST, d_bytes, n_msgs, bytes_per_msg, d_digests as *mut u64, stream));
    }
    pub unsafe fn hash_batch<F: PrimeField, H: B200Batch<F>>(d_bytes: *const u8, d_offsets: *const u64, n_msgs: usize,
                                                             d_digests: *mut F, stream: *mut c_void) {
        check(anemoi_b200_hash_bytes_ragged_dev(H::FIELD, H::INST, d_bytes, d_offsets, n_msgs, d_digests as *mut u64, stream));
    }
    pub unsafe fn merge_batch<F: PrimeField, H: B200Batch<F>>(d_pairs: *const F, d_out: *mut F, n: usize, stream: *mut c_void) {
        check(anemoi_b200_merge_dev(H::FIELD, H::INST, d_pairs as *const u64, d_out as *mut u64, n, stream));
    }
    pub unsafe fn digests_to_bytes<F: PrimeField, H: B200Batch<F>>(d_digests: *const F, d_bytes: *mut u8, n: usize, stream: *mut c_void) {
        check(anemoi_b200_digest_to_bytes_dev(H::FIELD, d_digests as *const u64, d_bytes, n, stream));
    }
    /// `levels` levels of a tree (or of many equal trees stored back to back): n_leaves -> n_leaves / arity^levels.
    pub unsafe fn merkle_reduce<F: PrimeField, H: B200Batch<F>>(d_leaves: *const F, n_leaves: usize, levels: usize, d_scratch: *mut F,
                                                                d_out: *mut F, stream: *mut c_void) {
        check(anemoi_b200_merkle_reduce_dev(H::FIELD, H::INST, H::STATE_WIDTH as c_int, d_leaves as *const u64, n_leaves,
                                            levels as c_int, d_scratch as *mut u64, d_out as *mut u64, stream));
    }
    /// The sharded root (one rank per GPU): sub-tree, ONE ncclAllGather of the partial roots, top levels; every rank gets the root.
    pub unsafe fn merkle_root_sharded<F: PrimeField, H: B200Batch<F>>(d_local_leaves: *const F, n_local: usize, comm: Option<&Comm>,
                                                                      d_root: *mut F, stream: *mut c_void) {
        check(anemoi_b200_merkle_root_sharded_dev(H::FIELD, H::INST, H::STATE_WIDTH as c_int, d_local_leaves as *const u64, n_local,
                                                  comm.map(|c| c.0).unwrap_or(core::ptr::null_mut()), core::ptr::null_mut(),
                                                  d_root as *mut u64, stream));
    }
    pub unsafe fn merkle_tree<F: PrimeField, H: B200Batch<F>>(d_leaves: *const F, n_leaves: usize, d_tree: *mut F, stream: *mut c_void) {
        check(anemoi_b200_merkle_tree_dev(H::FIELD, H::INST, H::STATE_WIDTH as c_int, d_leaves as *const u64, n_leaves, d_tree as *mut u64, stream));
    }
    pub unsafe fn merkle_open<F: PrimeField, H: B200Batch<F>>(d_leaves: *const F, d_tree: *const F, n_leaves: usize, d_indices: *const u64,
                                                              n_idx: usize, d_paths: *mut F, stream: *mut c_void) {
        check(anemoi_b200_merkle_open_dev(H::FIELD, H::INST, H::STATE_WIDTH as c_int, d_leaves as *const u64, d_tree as *const u64,
                                          n_leaves, d_indices, n_idx, d_paths as *mut u64, stream));
    }
    pub unsafe fn merkle_verify<F: PrimeField, H: B200Batch<F>>(d_leaf_values: *const F, d_indices: *const u64, d_paths: *const F,
                                                                height: usize, n_idx: usize, d_scratch: *mut F, d_roots: *mut F,
                                                                stream: *mut c_void) {
        check(anemoi_b200_merkle_verify_dev(H::FIELD, H::INST, H::STATE_WIDTH as c_int, d_leaf_values as *const u64, d_indices,
                                            d_paths as *const u64, height as c_int, n_idx, d_scratch as *mut u64, d_roots as *mut u64, stream));
    }
}

macro_rules! impl_b200 {
    ($module:path, $ty:ident, $felt:path, $field:expr, $inst:expr, $w:expr, $n64:expr) => {
        // an arkworks Fp must be exactly its N Montgomery limbs for the pointer casts above to be sound
        const _: () = assert!(core::mem::size_of::<$felt>() == 8 * $n64 && core::mem::align_of::<$felt>() == 8);
        impl B200Batch<$felt> for $module::$ty {
            type Digest = $module::AnemoiDigest;
            const FIELD: c_int = $field;
            const INST: c_int = $inst;
            const STATE_WIDTH: usize = $w;
            fn digest_from_felt(e: $felt) -> Self::Digest { <$module::AnemoiDigest>::new([e]) }
            fn digest_to_felt(d: &Self::Digest) -> $felt { d.as_elements()[0] }
        }
    };
}

// One line per reference marker type (src/<field>/anemoi_{2_1,4_3}/mod.rs:38).
impl_b200!(crate::bls12_377::anemoi_2_1, AnemoiBls12_377_2_1, crate::bls12_377::Felt, FIELD_BLS12_377, INST_2_1, 2, 6);
impl_b200!(crate::bls12_377::anemoi_4_3, AnemoiBls12_377_4_3, crate::bls12_377::Felt, FIELD_BLS12_377, INST_4_3, 4, 6);
impl_b200!(crate::bls12_381::anemoi_2_1, AnemoiBls12_381_2_1, crate::bls12_381::Felt, FIELD_BLS12_381, INST_2_1, 2, 6);
impl_b200!(crate::bls12_381::anemoi_4_3, AnemoiBls12_381_4_3, crate::bls12_381::Felt, FIELD_BLS12_381, INST_4_3, 4, 6);
impl_b200!(crate::bn_254::anemoi_2_1, AnemoiBn254_2_1, crate::bn_254::Felt, FIELD_BN_254, INST_2_1, 2, 4);
impl_b200!(crate::bn_254::anemoi_4_3, AnemoiBn254_4_3, crate::bn_254::Felt, FIELD_BN_254, INST_4_3, 4, 4);
impl_b200!(crate::ed_on_bls12_377::anemoi_2_1, AnemoiEdOnBls12_377_2_1, crate::ed_on_bls12_377::Felt, FIELD_ED_ON_BLS12_377, INST_2_1, 2, 4);
impl_b200!(crate::ed_on_bls12_377::anemoi_4_3, AnemoiEdOnBls12_377_4_3, crate::ed_on_bls12_377::Felt, FIELD_ED_ON_BLS12_377, INST_4_3, 4, 4);
impl_b200!(crate::jubjub::anemoi_2_1, AnemoiJubjub_2_1, crate::jubjub::Felt, FIELD_JUBJUB, INST_2_1, 2, 4);
impl_b200!(crate::jubjub::anemoi_4_3, AnemoiJubjub_4_3, crate::jubjub::Felt, FIELD_JUBJUB, INST_4_3, 4, 4);
impl_b200!(crate::pallas::anemoi_2_1, AnemoiPallas_2_1, crate::pallas::Felt, FIELD_PALLAS, INST_2_1, 2, 4);
impl_b200!(crate::pallas::anemoi_4_3, AnemoiPallas_4_3, crate::pallas::Felt, FIELD_PALLAS, INST_4_3, 4, 4);
impl_b200!(crate::vesta::anemoi_2_1, AnemoiVesta_2_1, crate::vesta::Felt, FIELD_VESTA, INST_2_1, 2, 4);
impl_b200!(crate::vesta::anemoi_4_3, AnemoiVesta_4_3, crate::vesta::Felt, FIELD_VESTA, INST_4_3, 4, 4);

#[cfg(test)]
mod tests {
    //! What `cargo test --features b200` adds on a machine with cargo + a B200: the layout pin, and the reference's own
    //! per-item functions against the batched ones on the same inputs (bit-exact by `==` on `Felt`).
    use super::*;
    use crate::{Jive, Sponge};
    use ark_ff::{One, UniformRand};

    #[test]
    fn layout_is_montgomery_limbs() {
        // R mod p of BLS12-381 Fq: the in-memory form of Felt::one()
        let one = crate::bls12_381::Felt::one();
        let limbs: [u64; 6] = unsafe { core::mem::transmute(one) };
        assert_eq!(limbs, [0x760900000002fffd, 0xebf4000bc40c0002, 0x5f48985753c758ba, 0x77ce585370525745, 0x5c071a97a256ec6d, 0x15f65ec3fa80e493]);
        let one = crate::pallas::Felt::one();
        let limbs: [u64; 4] = unsafe { core::mem::transmute(one) };
        assert_eq!(limbs, [0x34786d38fffffffd, 0x992c350be41914ad, 0xffffffffffffffff, 0x3fffffffffffffff]);
    }

    #[test]
    fn batched_equals_per_item() {
        use crate::bls12_381::anemoi_2_1::AnemoiBls12_381_2_1 as H;
        use crate::bls12_381::Felt;
        let mut rng = ark_std::test_rng();
        let elems: Vec<Felt> = (0..2 * 257).map(|_| Felt::rand(&mut rng)).collect();
        let batched = H::compress_batch(&elems);
        for (i, pair) in elems.chunks(2).enumerate() {
            assert_eq!(batched[i], <H as Jive<Felt>>::compress(pair)[0]);
        }
        let msgs: Vec<&[Felt]> = vec![&elems[..0], &elems[..1], &elems[..7], &elems[..331]];
        let digests = H::hash_field_batch(&msgs);
        for (m, d) in msgs.iter().zip(digests.iter()) {
            assert_eq!(*d, <H as Sponge<Felt>>::hash_field(m));
        }
        let bytes: Vec<&[u8]> = vec![b"", b"a", &[7u8; 47], &[9u8; 10240]];
        for (m, d) in bytes.iter().zip(H::hash_batch(&bytes).iter()) {
            assert_eq!(*d, <H as Sponge<Felt>>::hash(m));
        }
        let d0 = digests[1];
        let d1 = digests[2];
        assert_eq!(H::merge_batch(&[[d0, d1]])[0], <H as Sponge<Felt>>::merge(&[d0, d1]));
        assert_eq!(&H::digests_to_bytes(&[d0])[..], &d0.to_bytes()[..]);
    }
}
