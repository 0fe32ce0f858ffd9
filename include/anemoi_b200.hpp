// anemoi_b200.hpp -- header-only C++ host mirror of the reference's interface for the batched path.
//
// The reference is a Rust crate (compiled code) and this image has no Rust toolchain, so the host side
// above the C ABI is mirrored in C++: one marker type per reference struct
// (AnemoiBls12_381_2_1, ... -- src/<field>/anemoi_{2_1,4_3}/mod.rs:38) with the reference's associated
// functions Jive::{compress, compress_k} (src/traits.rs:23-33), Sponge::{hash_field, hash, merge}
// (src/traits.rs:8-20), Anemoi::permutation (src/traits.rs:370), plus *_batch forms and merkle_root.
// Same argument meaning, same failure conditions: where the reference panics (assert!), these throw
// std::invalid_argument. Felt<N64> is layout-identical to arkworks' Fp<MontBackend<_, N64>, N64>
// (N64 little-endian u64 limbs, Montgomery form), so `const Felt*` is the memory of a Rust `&[Felt]`.
// rust/anemoi_b200_shim.rs is the same surface as Rust extension traits over the reference's own types.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "anemoi_b200.h"

namespace anemoi_b200 {

template <int N64>
struct Felt {
    std::array<uint64_t, N64> limbs{};  // Montgomery form, canonical
    bool operator==(const Felt& o) const { return limbs == o.limbs; }
    bool operator!=(const Felt& o) const { return !(*this == o); }
};

struct Error : std::runtime_error {
    int code;
    Error(int c) : std::runtime_error(std::string("anemoi_b200: ") + anemoi_b200_strerror(c) + " [" +
                                      anemoi_b200_last_cuda_error() + "]"), code(c) {}
};

inline void check(int rc) {
    if (rc == ANEMOI_B200_OK) return;
    if (rc == ANEMOI_B200_ERR_ARITY || rc == ANEMOI_B200_ERR_LENGTH)  // the reference's assert! panics
        throw std::invalid_argument(anemoi_b200_strerror(rc));
    throw Error(rc);
}

// The library's device-memory pool (buffers of the host-pointer calls): grow it ahead of the first big call / hand it back.
inline void pool_reserve(int device, size_t bytes) { check(anemoi_b200_pool_reserve(device, bytes)); }
inline void pool_trim(int device, size_t keep_bytes = 0) { check(anemoi_b200_pool_trim(device, keep_bytes)); }

// NCCL communicator owned by the library (anemoi_b200_comm_*): rank 0 draws the 128-byte id, the host ships it to the
// other ranks by its own transport, every rank joins with its CUDA device current. Used by merkle_root_sharded_dev.
struct Comm {
    void* c = nullptr;
    static std::array<uint8_t, 128> unique_id() {
        std::array<uint8_t, 128> id{};
        check(anemoi_b200_comm_unique_id(id.data()));
        return id;
    }
    Comm(const std::array<uint8_t, 128>& id, int nranks, int rank) { check(anemoi_b200_comm_init_rank(id.data(), nranks, rank, &c)); }
    Comm(const Comm&) = delete;
    Comm& operator=(const Comm&) = delete;
    ~Comm() { anemoi_b200_comm_destroy(c); }
};

// AnemoiDigest([Felt; 1]) -- src/<field>/anemoi_x/digest.rs:13-53
template <int FIELD, int N64>
struct AnemoiDigest {
    std::array<Felt<N64>, 1> e{};
    static AnemoiDigest new_(const std::array<Felt<N64>, 1>& v) { return AnemoiDigest{v}; }
    const std::array<Felt<N64>, 1>& as_elements() const { return e; }
    std::array<Felt<N64>, 1> to_elements() const { return e; }
    static std::vector<Felt<N64>> digests_to_elements(const std::vector<AnemoiDigest>& ds) {
        std::vector<Felt<N64>> out;
        for (const auto& d : ds) out.push_back(d.e[0]);
        return out;
    }
    // digest.rs:42-46: canonical little-endian bytes
    std::array<uint8_t, 8 * N64> to_bytes(int device = 0) const {
        std::array<uint8_t, 8 * N64> b{};
        check(anemoi_b200_digest_to_bytes(FIELD, e[0].limbs.data(), b.data(), 1, device));
        return b;
    }
    bool operator==(const AnemoiDigest& o) const { return e == o.e; }
};

template <int FIELD, int INST, int N64>
struct Anemoi {
    using F = Felt<N64>;
    using Digest = AnemoiDigest<FIELD, N64>;
    static constexpr int STATE_WIDTH = INST == ANEMOI_INST_2_1 ? 2 : 4;  // mod.rs:20
    static constexpr int RATE_WIDTH = INST == ANEMOI_INST_2_1 ? 1 : 3;   // mod.rs:22
    static constexpr int NUM_COLUMNS = STATE_WIDTH / 2;                  // mod.rs:25
    static constexpr int DIGEST_SIZE = 1;                                // mod.rs:28
    static int num_hash_rounds() { return anemoi_b200_num_rounds(FIELD, INST); }

    static const uint64_t* raw(const F* p) { return reinterpret_cast<const uint64_t*>(p); }
    static uint64_t* raw(F* p) { return reinterpret_cast<uint64_t*>(p); }

    // ---- batched forms -------------------------------------------------------------------------
    static void permutation_batch(std::vector<F>& states, int device = 0) {
        if (states.size() % STATE_WIDTH) throw std::invalid_argument("not a whole number of states");
        check(anemoi_b200_permute(FIELD, INST, raw(states.data()), states.size() / STATE_WIDTH, device));
    }
    static std::vector<F> compress_k_batch(const std::vector<F>& states, int k, int device = 0) {
        if (k <= 0 || STATE_WIDTH % k || k % 2) throw std::invalid_argument("compress_k: bad k");
        if (states.size() % STATE_WIDTH) throw std::invalid_argument("not a whole number of states");
        const size_t n = states.size() / STATE_WIDTH;
        std::vector<F> out(n * (STATE_WIDTH / k));
        check(anemoi_b200_compress(FIELD, INST, k, raw(states.data()), raw(out.data()), n, device));
        return out;
    }
    static std::vector<F> hash_field_batch(const std::vector<F>& elems, size_t n_msgs, size_t felts_per_msg, int device = 0) {
        if (elems.size() != n_msgs * felts_per_msg) throw std::invalid_argument("hash_field_batch: bad length");
        std::vector<F> out(n_msgs);
        F dummy{};
        check(anemoi_b200_hash_field(FIELD, INST, elems.empty() ? raw(&dummy) : raw(elems.data()), n_msgs, felts_per_msg,
                                     raw(out.data()), device));
        return out;
    }
    static F merkle_root(const std::vector<F>& leaves, int n_gpus = 1) {
        F root{};
        check(anemoi_b200_merkle_root(FIELD, INST, STATE_WIDTH, raw(leaves.data()), leaves.size(), raw(&root), n_gpus));
        return root;
    }

    // Sharded root, one rank per GPU (device pointers on the current device, stream-ordered): per-rank sub-tree, one
    // ncclAllGather of the partial roots issued by the library, top levels on every rank. comm == nullptr: single rank.
    static void merkle_root_sharded_dev(const F* d_local_leaves, size_t n_local, const Comm* comm, F* d_root, void* stream) {
        check(anemoi_b200_merkle_root_sharded_dev(FIELD, INST, STATE_WIDTH, raw(d_local_leaves), n_local, comm ? comm->c : nullptr,
                                                  nullptr, raw(d_root), stream));
    }

    // Openings (authentication paths) of `indices` in the tree over `leaves`; returns the root, fills `paths`
    // with indices.size() * height * (STATE_WIDTH - 1) elements (siblings per level, leaf level first).
    static F merkle_open(const std::vector<F>& leaves, const std::vector<uint64_t>& indices, std::vector<F>& paths,
                         int device = 0) {
        int height = 0;
        for (size_t m = leaves.size(); m > 1; m /= STATE_WIDTH) height++;
        paths.assign(indices.size() * (size_t)height * (STATE_WIDTH - 1), F{});
        F root{};
        check(anemoi_b200_merkle_open(FIELD, INST, STATE_WIDTH, raw(leaves.data()), leaves.size(), indices.data(),
                                      indices.size(), raw(&root), paths.empty() ? nullptr : raw(paths.data()), device));
        return root;
    }
    // Roots implied by (leaf value, index, path) triples; compare with the committed root.
    static std::vector<F> merkle_verify(const std::vector<F>& leaf_values, const std::vector<uint64_t>& indices,
                                        const std::vector<F>& paths, int height, int device = 0) {
        std::vector<F> roots(indices.size());
        check(anemoi_b200_merkle_verify(FIELD, INST, STATE_WIDTH, raw(leaf_values.data()), indices.data(),
                                        paths.empty() ? nullptr : raw(paths.data()), height, indices.size(),
                                        raw(roots.data()), device));
        return roots;
    }

    // ---- the reference's per-item API ----------------------------------------------------------
    // Anemoi::permutation(&mut [F]) -- src/traits.rs:370
    static void permutation(std::vector<F>& state) {
        if ((int)state.size() != STATE_WIDTH) throw std::invalid_argument("state.len() != WIDTH");
        permutation_batch(state);
    }
    // Jive::compress -- hasher.rs:96-103 / 4-3 :148-160
    static std::vector<F> compress(const std::vector<F>& elems) {
        if ((int)elems.size() != STATE_WIDTH) throw std::invalid_argument("elems.len() != STATE_WIDTH");
        return compress_k_batch(elems, 2);
    }
    // Jive::compress_k -- hasher.rs:105-110 / 4-3 :162-179
    static std::vector<F> compress_k(const std::vector<F>& elems, int k) {
        if (INST == ANEMOI_INST_2_1 && k != 2) throw std::invalid_argument("assert!(k == 2)");
        if ((int)elems.size() != STATE_WIDTH) throw std::invalid_argument("elems.len() != STATE_WIDTH");
        return compress_k_batch(elems, k);
    }
    // Sponge::hash_field -- hasher.rs:68-85 / 4-3 :93-129
    static Digest hash_field(const std::vector<F>& elems) {
        return Digest{{hash_field_batch(elems, 1, elems.size())[0]}};
    }
    // Sponge::hash -- hasher.rs:18-66 / 4-3 :18-91
    static Digest hash(const std::vector<uint8_t>& bytes) {
        Digest d;
        uint8_t dummy = 0;
        check(anemoi_b200_hash_bytes(FIELD, INST, bytes.empty() ? &dummy : bytes.data(), 1, bytes.size(), raw(d.e.data()), 0));
        return d;
    }
    // Sponge::merge(&[Digest; 2]) -- 2-1: Jive (hasher.rs:87-92); 4-3: reads digests[0] only (:131-144, sic)
    static Digest merge(const std::array<Digest, 2>& ds) {
        F in[2] = {ds[0].e[0], ds[1].e[0]};
        Digest d;
        check(anemoi_b200_merge(FIELD, INST, raw(in), raw(d.e.data()), 1, 0));
        return d;
    }
};

// struct names exactly as in the reference (src/<field>/anemoi_{2_1,4_3}/mod.rs:38)
using AnemoiBls12_377_2_1 = Anemoi<ANEMOI_FIELD_BLS12_377, ANEMOI_INST_2_1, 6>;
using AnemoiBls12_377_4_3 = Anemoi<ANEMOI_FIELD_BLS12_377, ANEMOI_INST_4_3, 6>;
using AnemoiBls12_381_2_1 = Anemoi<ANEMOI_FIELD_BLS12_381, ANEMOI_INST_2_1, 6>;
using AnemoiBls12_381_4_3 = Anemoi<ANEMOI_FIELD_BLS12_381, ANEMOI_INST_4_3, 6>;
using AnemoiBn254_2_1 = Anemoi<ANEMOI_FIELD_BN_254, ANEMOI_INST_2_1, 4>;
using AnemoiBn254_4_3 = Anemoi<ANEMOI_FIELD_BN_254, ANEMOI_INST_4_3, 4>;
using AnemoiEdOnBls12_377_2_1 = Anemoi<ANEMOI_FIELD_ED_ON_BLS12_377, ANEMOI_INST_2_1, 4>;
using AnemoiEdOnBls12_377_4_3 = Anemoi<ANEMOI_FIELD_ED_ON_BLS12_377, ANEMOI_INST_4_3, 4>;
using AnemoiJubjub_2_1 = Anemoi<ANEMOI_FIELD_JUBJUB, ANEMOI_INST_2_1, 4>;
using AnemoiJubjub_4_3 = Anemoi<ANEMOI_FIELD_JUBJUB, ANEMOI_INST_4_3, 4>;
using AnemoiPallas_2_1 = Anemoi<ANEMOI_FIELD_PALLAS, ANEMOI_INST_2_1, 4>;
using AnemoiPallas_4_3 = Anemoi<ANEMOI_FIELD_PALLAS, ANEMOI_INST_4_3, 4>;
using AnemoiVesta_2_1 = Anemoi<ANEMOI_FIELD_VESTA, ANEMOI_INST_2_1, 4>;
using AnemoiVesta_4_3 = Anemoi<ANEMOI_FIELD_VESTA, ANEMOI_INST_4_3, 4>;

}  // namespace anemoi_b200
