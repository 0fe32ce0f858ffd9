// Montgomery arithmetic on N x 32-bit limbs held in registers (N = 8: 255-bit fields, N = 12: BLS12-377/381 Fq).
//
// Replaces layer L0 of the reference (arkworks ark-ff `Fp<MontBackend<_,N>,N>`: the + - * square double
// behind every operator in src/traits.rs:78-358 and src/<field>/sbox.rs). Same representation
// (a * 2^(32N) mod p, little-endian limbs), so values cross the C ABI unchanged.
//
// Multiplication is operand-scanning Montgomery with two 64-bit-aligned accumulators ("even" and "odd"
// columns) so that every 32x32->64 product is one PTX mad.lo.cc/madc.hi.cc pair, which ptxas fuses into
// a single IMAD.WIDE.U32(.X) with the carry in a predicate. Squaring computes the upper triangle once,
// doubles it, adds the diagonal, and then runs N reduction rows: N(N+1)/2 + N^2 + N wide multiplies
// instead of 2N^2 + N.
//
// The carry-flag primitives have a host emulation (ANEMOI_FP_HOST_EMU) so the very same templates are
// unit-tested against Python big integers on a machine without a GPU (tests/test_fp_host_emu.py).
#pragma once
#include <cstdint>

#if defined(ANEMOI_FP_HOST_EMU)
#define FPQ inline
#ifndef HD
#define HD inline
#endif
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#define FP_UNROLL
#else
#define FPQ __device__ __forceinline__
#ifndef HD
#define HD __host__ __device__ __forceinline__
#endif
#define FP_UNROLL _Pragma("unroll")
#endif

namespace anemoi {
namespace fp {

// ------------------------------------------------------------------------------------------------
// carry-flag primitives
// ------------------------------------------------------------------------------------------------
#if defined(ANEMOI_FP_HOST_EMU)
static thread_local uint32_t g_cc = 0;  // emulated PTX condition-code carry/borrow flag

FPQ void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a * b;
    lo = (uint32_t)t;
    hi = (uint32_t)(t >> 32);
}
FPQ uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
// d = a*b + c  (64-bit, d and c as lo/hi pairs), carry handling as in the PTX sequences below
FPQ void emu_mad_pair(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi, bool cin, bool cout) {
    uint64_t t = (uint64_t)a * b;
    uint64_t lo = (uint64_t)(uint32_t)t + clo + (cin ? g_cc : 0);
    uint64_t hi = (t >> 32) + chi + (lo >> 32);
    dlo = (uint32_t)lo;
    dhi = (uint32_t)hi;
    if (cout) g_cc = (uint32_t)(hi >> 32);
}
FPQ void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { emu_mad_pair(lo, hi, a, b, lo, hi, false, true); }
FPQ void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { emu_mad_pair(lo, hi, a, b, lo, hi, true, true); }
FPQ void madc_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { emu_mad_pair(lo, hi, a, b, lo, hi, true, false); }
FPQ void madc_wide_cc_to(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
    emu_mad_pair(dlo, dhi, a, b, clo, chi, true, true);
}
// {lo,hi} = a*b + {lo or 0, hi or 0} [+ carry]; carry out. lo_zero / hi_zero: that accumulator word is known to be zero.
FPQ void mac_pair(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, bool cin, bool lo_zero, bool hi_zero) {
    emu_mad_pair(lo, hi, a, b, lo_zero ? 0u : lo, hi_zero ? 0u : hi, cin, true);
}
FPQ void add_cc(uint32_t& a, uint32_t b) { uint64_t t = (uint64_t)a + b; a = (uint32_t)t; g_cc = (uint32_t)(t >> 32); }
FPQ void addc_cc(uint32_t& a, uint32_t b) { uint64_t t = (uint64_t)a + b + g_cc; a = (uint32_t)t; g_cc = (uint32_t)(t >> 32); }
FPQ void addc(uint32_t& a, uint32_t b) { a = a + b + g_cc; }
FPQ void sub_cc(uint32_t& a, uint32_t b) { uint64_t t = (uint64_t)a - b; a = (uint32_t)t; g_cc = (uint32_t)((t >> 32) & 1); }
FPQ void subc_cc(uint32_t& a, uint32_t b) { uint64_t t = (uint64_t)a - b - g_cc; a = (uint32_t)t; g_cc = (uint32_t)((t >> 32) & 1); }
FPQ void subc(uint32_t& a, uint32_t b) { a = a - b - g_cc; }
FPQ uint32_t shf_l(uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (32 - s)); }
#else
FPQ void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    // one IMAD.WIDE.U32 Rd, Ra, Rb, RZ (separate mul.lo / mul.hi would become IMAD + IMAD.HI: 2 + 6 pipe cycles)
    asm volatile("{ .reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
FPQ uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
// {lo,hi} += a*b ; carry out
FPQ void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// {lo,hi} += a*b + carry ; carry out
FPQ void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// {lo,hi} += a*b + carry ; no carry out
FPQ void madc_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// {dlo,dhi} = a*b + {clo,chi} + carry ; carry out
FPQ void madc_wide_cc_to(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                 : "=r"(dlo), "=r"(dhi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
// {lo,hi} = a*b + {lo or 0, hi or 0} [+ carry]; carry out. lo_zero / hi_zero (compile-time after unrolling): that
// accumulator word is known to be zero -- the addend is then the immediate 0 and the word is a pure output, so no register
// has to be zeroed first (ptxas does that with IMAD.MOV Rd, RZ, RZ, RZ on the FMA-heavy pipe).
FPQ void mac_pair(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, bool cin, bool lo_zero, bool hi_zero) {
    if (cin) {
        if (lo_zero && hi_zero)
            asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.cc.u32 %1, %2, %3, 0;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
        else if (lo_zero)
            asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.cc.u32 %1, %2, %3, %1;" : "=r"(lo), "+r"(hi) : "r"(a), "r"(b));
        else if (hi_zero)
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, 0;" : "+r"(lo), "=r"(hi) : "r"(a), "r"(b));
        else
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
    } else {
        if (lo_zero && hi_zero)
            asm volatile("mad.lo.cc.u32 %0, %2, %3, 0; madc.hi.cc.u32 %1, %2, %3, 0;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
        else if (lo_zero)
            asm volatile("mad.lo.cc.u32 %0, %2, %3, 0; madc.hi.cc.u32 %1, %2, %3, %1;" : "=r"(lo), "+r"(hi) : "r"(a), "r"(b));
        else if (hi_zero)
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, 0;" : "+r"(lo), "=r"(hi) : "r"(a), "r"(b));
        else
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
    }
}
FPQ void add_cc(uint32_t& a, uint32_t b) { asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(a) : "r"(b)); }
FPQ void addc_cc(uint32_t& a, uint32_t b) { asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(a) : "r"(b)); }
FPQ void addc(uint32_t& a, uint32_t b) { asm volatile("addc.u32 %0, %0, %1;" : "+r"(a) : "r"(b)); }
FPQ void sub_cc(uint32_t& a, uint32_t b) { asm volatile("sub.cc.u32 %0, %0, %1;" : "+r"(a) : "r"(b)); }
FPQ void subc_cc(uint32_t& a, uint32_t b) { asm volatile("subc.cc.u32 %0, %0, %1;" : "+r"(a) : "r"(b)); }
FPQ void subc(uint32_t& a, uint32_t b) { asm volatile("subc.u32 %0, %0, %1;" : "+r"(a) : "r"(b)); }
// funnel shift left: high 32 bits of ({hi,lo} << s), 0 < s < 32
FPQ uint32_t shf_l(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_l(lo, hi, s); }
#endif

// Carry fix-ups and the pipe they run on.
// A row of multiply-accumulates ends with "top word += carry". Its own carry-out is zero (the accumulators never
// overflow their top word), so the natural PTX is addc.u32 -- which ptxas lowers to IMAD.X Rd, RZ, RZ, Rd, P on the
// FMA-heavy pipe, the one pipe this code saturates (it does the same to plain adds: IMAD.IADD, and to moves: IMAD.MOV;
// an opaque zero or an explicit SEL + add do not help, measured in SASS). An add WITH a live carry-out, however, exists
// only as IADD3.X on the ALU pipe. So with F::CARRY_CHAIN the fix-up is addc.cc and the (zero) carry it produces is consumed by
// the first instruction of the NEXT carry chain (addc.cc / madc.lo.cc instead of add.cc / mad.lo.cc): no instruction is
// added, the fix-ups move to the idle ALU pipe. The *_after_fixup forms mark exactly those chain starts.
// The price is instruction-level parallelism: the next chain cannot start before the fix-up. Whether the trade pays is
// measured per field (F::CARRY_CHAIN, tools/gen_params.py): ptxas emits the IMAD.X forms in the 8-limb kernels and in
// bls12_377, not in bls12_381. (Also tried, in SASS: carry -> 0/1 register via SEL plus a three-input add -- ptxas splits
// that into IMAD.IADD pairs; an opaque zero addend -- IMAD.X Rd, Rz, 0x1, Rd; addc.cc with a dead carry-out -- demoted.)
template <class F>
FPQ void fixup_carry(uint32_t& a) {
    if (F::CARRY_CHAIN) addc_cc(a, 0u); else addc(a, 0u);
}
template <class F>
FPQ void add_cc_after_fixup(uint32_t& a, uint32_t b) {
    if (F::CARRY_CHAIN) addc_cc(a, b); else add_cc(a, b);
}
template <class F>
FPQ void mad_wide_cc_after_fixup(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    if (F::CARRY_CHAIN) madc_wide_cc(lo, hi, a, b); else mad_wide_cc(lo, hi, a, b);
}

// ------------------------------------------------------------------------------------------------
// modular add / sub / small multiples (canonical in, canonical out)
// ------------------------------------------------------------------------------------------------

// r = (r >= p) ? r - p : r          (r < 2p on entry; needs 2p < 2^(32N), true for all 7 fields)
template <class F>
FPQ void cond_sub_p(uint32_t (&r)[F::N]) {
    constexpr int N = F::N;
    uint32_t t[N];
    FP_UNROLL
    for (int i = 0; i < N; i++) t[i] = r[i];
    sub_cc(t[0], F::p(0));
    FP_UNROLL
    for (int i = 1; i < N; i++) subc_cc(t[i], F::p(i));
    uint32_t borrow = 0;
    subc(borrow, 0);  // borrow = 0 - 0 - b  -> 0xffffffff when r < p
    FP_UNROLL
    for (int i = 0; i < N; i++) r[i] = borrow ? r[i] : t[i];
}

template <class F>
FPQ void add_mod(uint32_t (&r)[F::N], const uint32_t (&a)[F::N], const uint32_t (&b)[F::N]) {
    constexpr int N = F::N;
    FP_UNROLL
    for (int i = 0; i < N; i++) r[i] = a[i];
    add_cc(r[0], b[0]);
    FP_UNROLL
    for (int i = 1; i < N - 1; i++) addc_cc(r[i], b[i]);
    addc(r[N - 1], b[N - 1]);
    cond_sub_p<F>(r);
}

// r = a - b mod p
template <class F>
FPQ void sub_mod(uint32_t (&r)[F::N], const uint32_t (&a)[F::N], const uint32_t (&b)[F::N]) {
    constexpr int N = F::N;
    FP_UNROLL
    for (int i = 0; i < N; i++) r[i] = a[i];
    sub_cc(r[0], b[0]);
    FP_UNROLL
    for (int i = 1; i < N; i++) subc_cc(r[i], b[i]);
    uint32_t borrow = 0;
    subc(borrow, 0);  // 0xffffffff when a < b
    // add back p & borrow
    add_cc(r[0], F::p(0) & borrow);
    FP_UNROLL
    for (int i = 1; i < N - 1; i++) addc_cc(r[i], F::p(i) & borrow);
    addc(r[N - 1], F::p(N - 1) & borrow);
}

template <class F>
FPQ void dbl_mod(uint32_t (&r)[F::N], const uint32_t (&a)[F::N]) {
    add_mod<F>(r, a, a);
}

// r = BETA * a   -- Anemoi::mul_by_generator (src/traits.rs:78-91). The reference's match arms for
// 2,3,5,7,15 are doubling chains; beta = 22 takes its generic arm (a full multiply by F::from(22)).
// All give the canonical residue beta*a mod p, so one doubling chain per beta is used here.
template <class F>
FPQ void mul_by_beta(uint32_t (&r)[F::N], const uint32_t (&a)[F::N]) {
    constexpr int N = F::N;
    uint32_t t[N], u[N];
    if (F::BETA == 2) {
        dbl_mod<F>(r, a);
    } else if (F::BETA == 3) {
        dbl_mod<F>(t, a);
        add_mod<F>(r, t, a);
    } else if (F::BETA == 5) {
        dbl_mod<F>(t, a);
        dbl_mod<F>(u, t);
        add_mod<F>(r, u, a);
    } else if (F::BETA == 7) {
        dbl_mod<F>(t, a);
        add_mod<F>(u, t, a);
        dbl_mod<F>(t, u);
        add_mod<F>(r, t, a);
    } else if (F::BETA == 15) {
        dbl_mod<F>(t, a);
        dbl_mod<F>(u, t);
        dbl_mod<F>(t, u);
        dbl_mod<F>(u, t);
        sub_mod<F>(r, u, a);
    } else if (F::BETA == 22) {  // 22 = 2 * (2 * (4 + 1) + 1)
        dbl_mod<F>(t, a);
        dbl_mod<F>(u, t);
        add_mod<F>(t, u, a);  // 5a
        dbl_mod<F>(u, t);     // 10a
        add_mod<F>(t, u, a);  // 11a
        dbl_mod<F>(r, t);     // 22a
    } else {
        // not instantiated by the reference
        FP_UNROLL
        for (int i = 0; i < N; i++) r[i] = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// Montgomery rows
// ------------------------------------------------------------------------------------------------

// acc[0..n) += a[0], a[2], ... * b   (a strided by 2), carry chain left open (carry out in CC)
template <int N>
FPQ void cmad_row(uint32_t* acc, const uint32_t* a, uint32_t b) {
    mad_wide_cc(acc[0], acc[1], a[0], b);
    FP_UNROLL
    for (int j = 2; j < N; j += 2) madc_wide_cc(acc[j], acc[j + 1], a[j], b);
}

// odd[j..j+1] = a[j]*b + odd[j+2..j+3] + carry (shifts the accumulator down by two limbs while adding)
template <int N>
FPQ void madc_row_rshift(uint32_t* odd, const uint32_t* a, uint32_t b) {
    FP_UNROLL
    for (int j = 0; j < N - 2; j += 2) madc_wide_cc_to(odd[j], odd[j + 1], a[j], b, odd[j + 2], odd[j + 3]);
    madc_wide_cc_to(odd[N - 2], odd[N - 1], a[N - 2], b, 0u, 0u);
}

// m * p rows with the modulus limbs as compile-time constants. Limbs that are 0, 1 or a power of two
// (p[0] = 1 for five of the seven fields; Pallas/Vesta: p = 2^254 + t with p[4..6] = 0, p[7] = 2^30) do not
// need a multiplier: they can become add-with-carry / shift instructions on the ALU pipe while the FMA-heavy
// pipe is saturated by IMAD.WIDE. F::special_limb(j) (tools/gen_params.py: SPECIAL_LIMBS) selects which of them
// are diverted; measured on B200, diverting p[0] = 1 helps every field that has it, diverting more does not.
HD constexpr bool fp_is_pow2(uint32_t v) { return v != 0 && (v & (v - 1)) == 0; }
HD constexpr int fp_log2(uint32_t v) { return v <= 1 ? 0 : 1 + fp_log2(v >> 1); }

template <class F>
struct ModRow {
    static constexpr int N = F::N;
    // {lo,hi} (+)= P*m [+ carry]; carry out. FIRST: no carry in. TO: write {dlo,dhi} = P*m + {clo,chi}.
    // AFTER_FIXUP (with FIRST): this chain start directly follows a fixup_carry and consumes its zero carry.
    template <bool FIRST, bool AFTER_FIXUP = false>
    static FPQ void mac(uint32_t& lo, uint32_t& hi, int j, uint32_t m) {
        const uint32_t P = F::p(j);
        constexpr bool NOCIN = FIRST && !(AFTER_FIXUP && F::CARRY_CHAIN);
        if (F::special_limb(j) && P == 0) {
            if (NOCIN) add_cc(lo, 0u); else addc_cc(lo, 0u);
            addc_cc(hi, 0u);
        } else if (F::special_limb(j) && P == 1) {
            if (NOCIN) add_cc(lo, m); else addc_cc(lo, m);
            addc_cc(hi, 0u);
        } else if (F::special_limb(j) && fp_is_pow2(P)) {
            const int k = fp_log2(P);
            if (NOCIN) add_cc(lo, m << k); else addc_cc(lo, m << k);
            addc_cc(hi, m >> (32 - k));
        } else {
            if (NOCIN) mad_wide_cc(lo, hi, P, m); else madc_wide_cc(lo, hi, P, m);
        }
    }
    static FPQ void mac_to(uint32_t& dlo, uint32_t& dhi, int j, uint32_t m, uint32_t clo, uint32_t chi) {
        const uint32_t P = F::p(j);
        if (F::special_limb(j) && P == 0) {
            dlo = clo; dhi = chi;
            addc_cc(dlo, 0u); addc_cc(dhi, 0u);
        } else if (F::special_limb(j) && P == 1) {
            dlo = clo; dhi = chi;
            addc_cc(dlo, m); addc_cc(dhi, 0u);
        } else if (F::special_limb(j) && fp_is_pow2(P)) {
            const int k = fp_log2(P);
            dlo = clo; dhi = chi;
            addc_cc(dlo, m << k); addc_cc(dhi, m >> (32 - k));
        } else {
            madc_wide_cc_to(dlo, dhi, P, m, clo, chi);
        }
    }
    template <bool AFTER_FIXUP = false>
    static FPQ void cmad_even(uint32_t* acc, uint32_t m) {
        mac<true, AFTER_FIXUP>(acc[0], acc[1], 0, m);
        FP_UNROLL
        for (int j = 2; j < N; j += 2) mac<false>(acc[j], acc[j + 1], j, m);
    }
    template <bool AFTER_FIXUP = false>
    static FPQ void cmad_odd(uint32_t* acc, uint32_t m) {
        mac<true, AFTER_FIXUP>(acc[0], acc[1], 1, m);
        FP_UNROLL
        for (int j = 2; j < N; j += 2) mac<false>(acc[j], acc[j + 1], j + 1, m);
    }
    static FPQ void mul_odd(uint32_t* acc, uint32_t m) {
        FP_UNROLL
        for (int j = 0; j < N; j += 2) mul_wide(acc[j], acc[j + 1], F::p(j + 1), m);
    }
    static FPQ void madc_odd_rshift(uint32_t* odd, uint32_t m) {
        FP_UNROLL
        for (int j = 0; j < N - 2; j += 2) mac_to(odd[j], odd[j + 1], j + 1, m, odd[j + 2], odd[j + 3]);
        mac_to(odd[N - 2], odd[N - 1], N - 1, m, 0u, 0u);
    }
};

// One operand-scanning row: (even, odd) <- ((even, odd) + a*bi + m*p) / 2^32, with the two accumulators
// exchanging roles (the caller alternates the argument order).
template <class F>
FPQ void mad_redc_row(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi, bool first) {
    constexpr int N = F::N;
    if (first) {
        FP_UNROLL
        for (int j = 0; j < N; j += 2) mul_wide(odd[j], odd[j + 1], a[j + 1], bi);
        FP_UNROLL
        for (int j = 0; j < N; j += 2) mul_wide(even[j], even[j + 1], a[j], bi);
        uint32_t m = F::quotient_digit(even[0]);
        ModRow<F>::cmad_odd(odd, m);
        ModRow<F>::cmad_even(even, m);
        fixup_carry<F>(odd[N - 1]);
    } else {
        add_cc_after_fixup<F>(even[0], odd[1]);  // every non-first row follows the previous row's closing fix-up
        madc_row_rshift<N>(odd, a + 1, bi);
        cmad_row<N>(even, a, bi);
        fixup_carry<F>(odd[N - 1]);
        uint32_t m = F::quotient_digit(even[0]);  // mul.lo / sub do not touch the carry flag
        ModRow<F>::template cmad_odd<true>(odd, m);
        ModRow<F>::cmad_even(even, m);
        fixup_carry<F>(odd[N - 1]);
    }
}

// After an even number of rows the running value is T = odd + (even << 32) with odd[0] == 0 (the last
// row ran with the roles exchanged), so T / 2^32 = (odd >> 32) + even.
template <class F, bool CANON>
FPQ void merge_even_odd(uint32_t (&r)[F::N], const uint32_t* even, const uint32_t* odd) {
    constexpr int N = F::N;
    FP_UNROLL
    for (int j = 0; j < N; j++) r[j] = even[j];
    add_cc_after_fixup<F>(r[0], odd[1]);  // follows the last row's closing fix-up
    FP_UNROLL
    for (int j = 1; j < N - 1; j++) addc_cc(r[j], odd[j + 1]);
    addc(r[N - 1], 0u);
    if (CANON) cond_sub_p<F>(r);
}

// r = a * b / R mod p.  CANON: result < p (inputs < p). !CANON (fields with >= 2 spare bits only):
// inputs < 2p, result < 2p, no final subtraction.
template <class F, bool CANON = true>
FPQ void mont_mul(uint32_t (&r)[F::N], const uint32_t (&a)[F::N], const uint32_t (&b)[F::N]) {
    constexpr int N = F::N;
    uint32_t even[N], odd[N];
    FP_UNROLL
    for (int i = 0; i < N; i += 2) {
        mad_redc_row<F>(even, odd, a, b[i], i == 0);
        mad_redc_row<F>(odd, even, a, b[i + 1], false);
    }
    merge_even_odd<F, CANON>(r, even, odd);
}

// One reduction-only row (the "multiply by 1" row): (even, odd) <- ((even, odd) + m*p) / 2^32
template <class F>
FPQ void redc_row(uint32_t* even, uint32_t* odd, bool first) {
    constexpr int N = F::N;
    if (first) {
        uint32_t m = F::quotient_digit(even[0]);
        ModRow<F>::mul_odd(odd, m);
        ModRow<F>::cmad_even(even, m);
        fixup_carry<F>(odd[N - 1]);
    } else {
        add_cc_after_fixup<F>(even[0], odd[1]);  // follows the previous row's closing fix-up
        uint32_t m = F::quotient_digit(even[0]);  // mul.lo does not touch the carry flag
        ModRow<F>::madc_odd_rshift(odd, m);
        ModRow<F>::cmad_even(even, m);
        fixup_carry<F>(odd[N - 1]);
    }
}

// r = a^2 / R mod p (same CANON contract as mont_mul)
template <class F, bool CANON = true>
FPQ void mont_sqr(uint32_t (&r)[F::N], const uint32_t (&a)[F::N]) {
    constexpr int N = F::N;
    // ---- upper triangle: sum_{i<j} a_i a_j 2^(32(i+j)) split by parity of i+j
    //   ev[k] : limb position k        (pairs (2t, 2t+1))
    //   od[k] : limb position k + 1    (pairs (2t+1, 2t+2))
    uint32_t ev[2 * N], od[2 * N];
    bool tev[2 * N], tod[2 * N];  // word already written? (compile-time after unrolling, like every index below)
    FP_UNROLL
    for (int k = 0; k < 2 * N; k++) { ev[k] = 0; od[k] = 0; tev[k] = false; tod[k] = false; }
    // `live` = the previous instruction of the carry-flag sequence was a fix-up whose (zero) carry the next chain start
    // must consume (F::CARRY_CHAIN, see above). A product into two untouched words cannot carry out, so a chain that ENDS on one
    // needs no fix-up, and one that STARTS on one (with nothing to consume) is a plain mul.wide.
    bool live = false;
    FP_UNROLL
    for (int i = 0; i < N - 1; i++) {
        FP_UNROLL
        for (int par = 0; par < 2; par++) {  // par 0: j = i+1, i+3, .. -> od[i+j-1];  par 1: j = i+2, i+4, .. -> ev[i+j]
            uint32_t* acc = par ? ev : od;
            bool* touched = par ? tev : tod;
            bool cin = live && F::CARRY_CHAIN;  // carry flag holds something the next instruction must take in
            bool pending = false;          // ... and it may be non-zero
            bool any = false;
            int top = 0;
            FP_UNROLL
            for (int j = i + 1 + par; j < N; j += 2) {
                const int k = par ? i + j : i + j - 1;
                const bool fresh = !touched[k] && !touched[k + 1];
                if (fresh && !cin) {
                    mul_wide(acc[k], acc[k + 1], a[i], a[j]);
                } else {
                    mac_pair(acc[k], acc[k + 1], a[i], a[j], cin, !touched[k], !touched[k + 1]);
                    cin = true;
                    pending = !fresh;
                }
                touched[k] = true; touched[k + 1] = true;
                any = true;
                top = k + 2;
            }
            if (any) {
                if (pending && top < 2 * N) { fixup_carry<F>(acc[top]); touched[top] = true; live = true; }
                else live = false;
            }
        }
    }
    // ---- w = ev + (od << 32)
    uint32_t w[2 * N];
    w[0] = ev[0];
    FP_UNROLL
    for (int k = 1; k < 2 * N; k++) w[k] = ev[k];
    if (live) add_cc_after_fixup<F>(w[1], od[0]);
    else add_cc(w[1], od[0]);
    FP_UNROLL
    for (int k = 2; k < 2 * N - 1; k++) addc_cc(w[k], od[k - 1]);
    // the cross-product sum is below 2^(64N-1): no carry out of the top word, which is what lets this add act as a fix-up
    if (F::CARRY_CHAIN) addc_cc(w[2 * N - 1], od[2 * N - 2]); else addc(w[2 * N - 1], od[2 * N - 2]);
    // ---- w = 2w (funnel shifts; the top bit is clear because 2*cross < a^2 < 2^(64N))
    FP_UNROLL
    for (int k = 2 * N - 1; k >= 1; k--) w[k] = shf_l(w[k - 1], w[k], 1);
    w[0] = w[0] << 1;
    // ---- w += sum_i a_i^2 2^(64 i)   (shifts do not touch the carry flag: this chain consumes the zero carry above)
    mad_wide_cc_after_fixup<F>(w[0], w[1], a[0], a[0]);
    FP_UNROLL
    for (int i = 1; i < N - 1; i++) madc_wide_cc(w[2 * i], w[2 * i + 1], a[i], a[i]);
    madc_wide(w[2 * N - 2], w[2 * N - 1], a[N - 1], a[N - 1]);
    // ---- Montgomery-reduce the low half, then add the high half
    uint32_t odd[N];
    FP_UNROLL
    for (int i = 0; i < N; i += 2) {
        redc_row<F>(&w[0], odd, i == 0);
        redc_row<F>(odd, &w[0], false);
    }
    // low half after N rows: (odd >> 32) + w[0..N)
    uint32_t u[N];
    FP_UNROLL
    for (int j = 0; j < N; j++) u[j] = w[j];
    add_cc_after_fixup<F>(u[0], odd[1]);  // follows the last row's closing fix-up
    FP_UNROLL
    for (int j = 1; j < N - 1; j++) addc_cc(u[j], odd[j + 1]);
    fixup_carry<F>(u[N - 1]);
    // + high half
    add_cc_after_fixup<F>(u[0], w[N]);
    FP_UNROLL
    for (int j = 1; j < N - 1; j++) addc_cc(u[j], w[N + j]);
    addc(u[N - 1], w[2 * N - 1]);
    FP_UNROLL
    for (int j = 0; j < N; j++) r[j] = u[j];
    if (CANON) cond_sub_p<F>(r);
}

template <class F>
FPQ void set_one(uint32_t (&r)[F::N]) {
    FP_UNROLL
    for (int i = 0; i < F::N; i++) r[i] = F::one(i);
}

template <class F>
FPQ void set_zero(uint32_t (&r)[F::N]) {
    FP_UNROLL
    for (int i = 0; i < F::N; i++) r[i] = 0;
}

}  // namespace fp
}  // namespace anemoi
