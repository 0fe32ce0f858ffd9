//! Rust side of the drop-in boundary: `extern "C"` bindings of libanemoi_b200.so (include/anemoi_b200.h)
//! and batched extension traits implemented for the reference crate's own marker types.
//!
//! SOURCE ONLY: this image has no cargo/rustc, so this file has not been compiled here. It is what a
//! maintainer adds to anemoi-rust as `src/b200.rs` (+ `pub mod b200;` in src/lib.rs behind a `b200`
//! feature, and `#![allow(unsafe_code)]` on the module because src/lib.rs:13 denies unsafe crate-wide).
//! INTEGRATION.md walks through it. No conversion happens at the boundary: an arkworks
//! `Fp<MontBackend<_, N>, N>` is `#[repr(transparent)]`-equivalent to `[u64; N]` Montgomery limbs
//! (`Fp.0 .0`), which is exactly the layout the C ABI reads and writes.
#![allow(unsafe_code)]

use ark_ff::PrimeField;
use core::ffi::c_int;

#[link(name = "anemoi_b200")]
extern "C" {
    fn anemoi_b200_permute(field: c_int, inst: c_int, states: *mut u64, n: usize, device: c_int) -> c_int;
    fn anemoi_b200_compress(field: c_int, inst: c_int, k: c_int, input: *const u64, out: *mut u64, n: usize, device: c_int) -> c_int;
    fn anemoi_b200_hash_field(field: c_int, inst: c_int, elems: *const u64, n_msgs: usize, felts_per_msg: usize, digests: *mut u64, device: c_int) -> c_int;
    fn anemoi_b200_hash_field_ragged(field: c_int, inst: c_int, elems: *const u64, offsets: *const u64, n_msgs: usize, digests: *mut u64, device: c_int) -> c_int;
    fn anemoi_b200_hash_bytes(field: c_int, inst: c_int, bytes: *const u8, n_msgs: usize, bytes_per_msg: usize, digests: *mut u64, device: c_int) -> c_int;
    fn anemoi_b200_merge(field: c_int, inst: c_int, pairs: *const u64, out: *mut u64, n: usize, device: c_int) -> c_int;
    fn anemoi_b200_merkle_root(field: c_int, inst: c_int, arity: c_int, leaves: *const u64, n_leaves: usize, root: *mut u64, n_gpus: c_int) -> c_int;
    fn anemoi_b200_strerror(code: c_int) -> *const core::ffi::c_char;
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

fn check(rc: c_int) {
    // The reference panics on misuse (assert!); the batched API keeps that contract.
    if rc != 0 {
        let msg = unsafe { core::ffi::CStr::from_ptr(anemoi_b200_strerror(rc)) };
        panic!("anemoi_b200: {}", msg.to_string_lossy());
    }
}

/// Batched forms of `Jive` / `Sponge` / `Anemoi::permutation` for one (field, instantiation).
pub trait B200Batch<F: PrimeField> {
    const FIELD: c_int;
    const INST: c_int;
    const STATE_WIDTH: usize;

    /// `Jive::compress_k` on `elems.len() / STATE_WIDTH` states (src/traits.rs:23-33).
    fn compress_k_batch(elems: &[F], k: usize) -> Vec<F> {
        assert!(elems.len() % Self::STATE_WIDTH == 0);
        assert!(k != 0 && Self::STATE_WIDTH % k == 0 && k % 2 == 0);
        let n = elems.len() / Self::STATE_WIDTH;
        let mut out = vec![F::zero(); n * Self::STATE_WIDTH / k];
        check(unsafe {
            anemoi_b200_compress(Self::FIELD, Self::INST, k as c_int, elems.as_ptr() as *const u64,
                                 out.as_mut_ptr() as *mut u64, n, 0)
        });
        out
    }
    fn compress_batch(elems: &[F]) -> Vec<F> { Self::compress_k_batch(elems, 2) }

    /// `Anemoi::permutation` on a slice of states, in place (src/traits.rs:370-378).
    fn permutation_batch(states: &mut [F]) {
        assert!(states.len() % Self::STATE_WIDTH == 0);
        check(unsafe {
            anemoi_b200_permute(Self::FIELD, Self::INST, states.as_mut_ptr() as *mut u64,
                                states.len() / Self::STATE_WIDTH, 0)
        });
    }

    /// `Sponge::hash_field` on `n_msgs` messages of `felts_per_msg` elements each.
    fn hash_field_batch(elems: &[F], felts_per_msg: usize) -> Vec<F> {
        let n = if felts_per_msg == 0 { 0 } else { elems.len() / felts_per_msg };
        assert!(elems.len() == n * felts_per_msg);
        let mut out = vec![F::zero(); n];
        check(unsafe {
            anemoi_b200_hash_field(Self::FIELD, Self::INST, elems.as_ptr() as *const u64, n, felts_per_msg,
                                   out.as_mut_ptr() as *mut u64, 0)
        });
        out
    }

    /// `Sponge::hash` on equal-length byte strings.
    fn hash_batch(bytes: &[u8], bytes_per_msg: usize) -> Vec<F> {
        let n = if bytes_per_msg == 0 { 0 } else { bytes.len() / bytes_per_msg };
        let mut out = vec![F::zero(); n];
        check(unsafe {
            anemoi_b200_hash_bytes(Self::FIELD, Self::INST, bytes.as_ptr(), n, bytes_per_msg,
                                   out.as_mut_ptr() as *mut u64, 0)
        });
        out
    }

    /// `Sponge::merge` on digest pairs (2-1: Jive; 4-3: the reference's sponge merge, digests[0] only).
    fn merge_batch(pairs: &[F]) -> Vec<F> {
        assert!(pairs.len() % 2 == 0);
        let mut out = vec![F::zero(); pairs.len() / 2];
        check(unsafe {
            anemoi_b200_merge(Self::FIELD, Self::INST, pairs.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64,
                              pairs.len() / 2, 0)
        });
        out
    }

    /// Jive Merkle root of arity STATE_WIDTH (2-1: `compress`; 4-3: `compress_k(.,4)`), n = arity^h leaves.
    fn merkle_root(leaves: &[F], n_gpus: usize) -> F {
        let mut root = F::zero();
        check(unsafe {
            anemoi_b200_merkle_root(Self::FIELD, Self::INST, Self::STATE_WIDTH as c_int, leaves.as_ptr() as *const u64,
                                    leaves.len(), &mut root as *mut F as *mut u64, n_gpus as c_int)
        });
        root
    }
}

macro_rules! impl_b200 {
    ($ty:path, $felt:path, $field:expr, $inst:expr, $w:expr) => {
        impl B200Batch<$felt> for $ty {
            const FIELD: c_int = $field;
            const INST: c_int = $inst;
            const STATE_WIDTH: usize = $w;
        }
    };
}

// One line per reference marker type (src/<field>/anemoi_{2_1,4_3}/mod.rs:38).
impl_b200!(crate::bls12_377::anemoi_2_1::AnemoiBls12_377_2_1, crate::bls12_377::Felt, FIELD_BLS12_377, INST_2_1, 2);
impl_b200!(crate::bls12_377::anemoi_4_3::AnemoiBls12_377_4_3, crate::bls12_377::Felt, FIELD_BLS12_377, INST_4_3, 4);
impl_b200!(crate::bls12_381::anemoi_2_1::AnemoiBls12_381_2_1, crate::bls12_381::Felt, FIELD_BLS12_381, INST_2_1, 2);
impl_b200!(crate::bls12_381::anemoi_4_3::AnemoiBls12_381_4_3, crate::bls12_381::Felt, FIELD_BLS12_381, INST_4_3, 4);
impl_b200!(crate::bn_254::anemoi_2_1::AnemoiBn254_2_1, crate::bn_254::Felt, FIELD_BN_254, INST_2_1, 2);
impl_b200!(crate::bn_254::anemoi_4_3::AnemoiBn254_4_3, crate::bn_254::Felt, FIELD_BN_254, INST_4_3, 4);
impl_b200!(crate::ed_on_bls12_377::anemoi_2_1::AnemoiEdOnBls12_377_2_1, crate::ed_on_bls12_377::Felt, FIELD_ED_ON_BLS12_377, INST_2_1, 2);
impl_b200!(crate::ed_on_bls12_377::anemoi_4_3::AnemoiEdOnBls12_377_4_3, crate::ed_on_bls12_377::Felt, FIELD_ED_ON_BLS12_377, INST_4_3, 4);
impl_b200!(crate::jubjub::anemoi_2_1::AnemoiJubjub_2_1, crate::jubjub::Felt, FIELD_JUBJUB, INST_2_1, 2);
impl_b200!(crate::jubjub::anemoi_4_3::AnemoiJubjub_4_3, crate::jubjub::Felt, FIELD_JUBJUB, INST_4_3, 4);
impl_b200!(crate::pallas::anemoi_2_1::AnemoiPallas_2_1, crate::pallas::Felt, FIELD_PALLAS, INST_2_1, 2);
impl_b200!(crate::pallas::anemoi_4_3::AnemoiPallas_4_3, crate::pallas::Felt, FIELD_PALLAS, INST_4_3, 4);
impl_b200!(crate::vesta::anemoi_2_1::AnemoiVesta_2_1, crate::vesta::Felt, FIELD_VESTA, INST_2_1, 2);
impl_b200!(crate::vesta::anemoi_4_3::AnemoiVesta_4_3, crate::vesta::Felt, FIELD_VESTA, INST_4_3, 4);
