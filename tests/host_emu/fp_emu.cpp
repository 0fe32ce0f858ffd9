// Host emulation harness for anemoi_rust_b200/csrc/fp.cuh (TEST INFRASTRUCTURE).
// Compiles the device templates with the carry-flag primitives emulated in C++ so the limb algorithms
// can be checked against big integers without a GPU. Built and driven by tests/test_fp_host_emu.py.
#define ANEMOI_FP_HOST_EMU 1
#include "../../anemoi_rust_b200/csrc/fp.cuh"
#include "../../anemoi_rust_b200/csrc/generated/fields.cuh"

using namespace anemoi;

template <class F>
static int run(int op, const uint32_t* a_, const uint32_t* b_, uint32_t* r_) {
    uint32_t a[F::N], b[F::N], r[F::N];
    for (int i = 0; i < F::N; i++) { a[i] = a_[i]; b[i] = b_[i]; r[i] = 0; }
    switch (op) {
        case 0: fp::mont_mul<F, true>(r, a, b); break;
        case 1: fp::mont_sqr<F, true>(r, a); break;
        case 2: fp::mont_mul<F, false>(r, a, b); break;
        case 3: fp::mont_sqr<F, false>(r, a); break;
        case 4: fp::add_mod<F>(r, a, b); break;
        case 5: fp::sub_mod<F>(r, a, b); break;
        case 6: fp::mul_by_beta<F>(r, a); break;
        default: return -1;
    }
    for (int i = 0; i < F::N; i++) r_[i] = r[i];
    return 0;
}

// The per-kernel code-generation switch F::CARRY_CHAIN (tools/gen_params.py) changes which PTX sequences run, not the
// result. fp_emu_op_flipped runs the other setting so that every field is checked on both paths.
template <class F>
struct Flipped : F {
    static constexpr bool CARRY_CHAIN = !F::CARRY_CHAIN;
};

extern "C" int fp_emu_limbs(int field) {
    switch (field) {
        case 0: return F_bls12_377::N;
        case 1: return F_bls12_381::N;
        case 2: return F_bn_254::N;
        case 3: return F_ed_on_bls12_377::N;
        case 4: return F_jubjub::N;
        case 5: return F_pallas::N;
        case 6: return F_vesta::N;
    }
    return -1;
}

extern "C" int fp_emu_op_flipped(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    switch (field) {
        case 0: return run<Flipped<F_bls12_377>>(op, a, b, r);
        case 1: return run<Flipped<F_bls12_381>>(op, a, b, r);
        case 2: return run<Flipped<F_bn_254>>(op, a, b, r);
        case 3: return run<Flipped<F_ed_on_bls12_377>>(op, a, b, r);
        case 4: return run<Flipped<F_jubjub>>(op, a, b, r);
        case 5: return run<Flipped<F_pallas>>(op, a, b, r);
        case 6: return run<Flipped<F_vesta>>(op, a, b, r);
    }
    return -1;
}

extern "C" int fp_emu_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    switch (field) {
        case 0: return run<F_bls12_377>(op, a, b, r);
        case 1: return run<F_bls12_381>(op, a, b, r);
        case 2: return run<F_bn_254>(op, a, b, r);
        case 3: return run<F_ed_on_bls12_377>(op, a, b, r);
        case 4: return run<F_jubjub>(op, a, b, r);
        case 5: return run<F_pallas>(op, a, b, r);
        case 6: return run<F_vesta>(op, a, b, r);
    }
    return -1;
}
