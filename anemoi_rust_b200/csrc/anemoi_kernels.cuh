// Fused Anemoi kernels for sm_100a: one thread per COLUMN of one state, the whole (x, y) column pair in
// registers across every round, one launch per batch.
//
// Replaces, per (field, instantiation), layers L1-L3 of the reference for a whole batch:
//   Anemoi::{ark_layer, mds_layer, sbox_layer, round, permutation}   src/traits.rs:113-157, 328-378
//   sbox::exp_by_inv_alpha                                           src/<field>/sbox.rs
//   Jive::{compress, compress_k}, Sponge::{hash, hash_field, merge}  src/<field>/anemoi_{2_1,4_3}/hasher.rs
//
// Mapping: Anemoi-2-1 (1 column) -> 1 thread per state; Anemoi-4-3 (2 columns) -> 2 adjacent lanes per
// state, which exchange their columns with warp shuffles for the linear layer (the two S-boxes of a
// round, > 99 % of the work, are independent). x^(1/alpha) runs a per-field addition chain on a small accumulator
// machine (any chain gives the same canonical residue as the reference's; these are searched for few multiplies and
// few live values, tools/chain_opt.py) whose slots live in per-thread local memory.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fp.cuh"
#include "kernel_args.h"

namespace anemoi {


template <int N>
FPQ void load_felt(uint32_t (&r)[N], const uint32_t* p, int vec16) {
    if (vec16) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
        for (int i = 0; i < N / 4; i++) {
            uint4 v = __ldg(q + i);
            r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
    } else {
        const uint2* q = reinterpret_cast<const uint2*>(p);
#pragma unroll
        for (int i = 0; i < N / 2; i++) {
            uint2 v = __ldg(q + i);
            r[2 * i] = v.x; r[2 * i + 1] = v.y;
        }
    }
}

template <int N>
FPQ void store_felt(uint32_t* p, const uint32_t (&r)[N], int vec16) {
    if (vec16) {
        uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
        for (int i = 0; i < N / 4; i++) q[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
    } else {
        uint2* q = reinterpret_cast<uint2*>(p);
#pragma unroll
        for (int i = 0; i < N / 2; i++) q[i] = make_uint2(r[2 * i], r[2 * i + 1]);
    }
}

// x^INV_ALPHA -- replaces sbox::exp_by_inv_alpha (src/<field>/sbox.rs). Any addition chain yields the same
// canonical residue as the reference's hard-coded one, so each field runs whichever ladder measured fastest
// (tools/gen_params.py: CHAIN_SOURCE). The ladder's slots are a dynamically indexed per-thread array, i.e. LOCAL
// memory (L1/L2-backed): one 32/48-byte entry is touched per ~5 squarings (~5 k cycles per thread), so its
// latency is irrelevant, while shared memory would cap the resident warps (measured: 2-5 % slower on every field).
// What matters is how MANY slots there are: the slot file of all resident threads should stay L2- (ideally L1-)
// resident, or its write-back reaches HBM.
//
// (1) Sliding-window ladder (round 1; still selectable): T[k] = x^(2k+1); the schedule {squarings, entry} comes from
//     constant memory (warp-uniform, no divergence). x^2 and the running odd power stay in registers while T is built.
template <class F>
FPQ void pow_window(uint32_t (&acc)[F::N], const uint32_t (&x)[F::N]) {
    constexpr int N = F::N;
    constexpr bool CANON = !F::LAZY;
    uint32_t T[F::SLOTS][N];
    {
        uint32_t x2[N], t[N];
        fp::mont_sqr<F, CANON>(x2, x);
#pragma unroll
        for (int l = 0; l < N; l++) {
            t[l] = x[l];
            T[0][l] = x[l];
        }
#pragma unroll 1
        for (int k = 1; k < F::SLOTS; k++) {
            fp::mont_mul<F, CANON>(t, t, x2);
#pragma unroll
            for (int l = 0; l < N; l++) T[k][l] = t[l];
        }
    }
#pragma unroll
    for (int l = 0; l < N; l++) acc[l] = T[F::SCHED_FIRST][l];
    const uint8_t* sched = Tables<F>::prog();
#pragma unroll 1
    for (int s = 0; s < F::SCHED_LEN; s++) {
        const int nsq = sched[2 * s];
        const int idx = sched[2 * s + 1];
#pragma unroll 1
        for (int q = 0; q < nsq; q++) fp::mont_sqr<F, CANON>(acc, acc);
        if (idx != 255) {
            uint32_t b[N];
#pragma unroll
            for (int l = 0; l < N; l++) b[l] = T[idx][l];
            fp::mont_mul<F, CANON>(acc, acc, b);
        }
    }
}

// (2) Accumulator machine: SQR n, MUL slot, LD slot, ST slot; slot 0 = x; the program is warp-uniform and comes from
//     constant memory. It runs the chains of tools/chains.json (searched by tools/chain_opt.py: 6-14 slots, 1-4 % less
//     MAC32 work than the reference's chains), or the reference crate's own chain compiled onto slots by a linear-scan
//     allocator (tools/gen_params.py; round 1: 11-28 slots). tests/test_chain_programs.py executes the generated
//     programs on the CPU.
template <class F>
FPQ void pow_program(uint32_t (&acc)[F::N], const uint32_t (&x)[F::N]) {
    constexpr int N = F::N;
    constexpr bool CANON = !F::LAZY;
    uint32_t T[F::SLOTS][N];
#pragma unroll
    for (int l = 0; l < N; l++) {
        acc[l] = x[l];
        T[0][l] = x[l];
    }
    const uint8_t* prog = Tables<F>::prog();
#pragma unroll 1
    for (int pc = 0; pc < F::PROG_LEN; pc++) {
        const int op = prog[2 * pc];
        const int arg = prog[2 * pc + 1];
        if (op == 0) {
#pragma unroll 1
            for (int q = 0; q < arg; q++) fp::mont_sqr<F, CANON>(acc, acc);
        } else if (op == 1) {
            uint32_t b[N];
#pragma unroll
            for (int l = 0; l < N; l++) b[l] = T[arg][l];
            fp::mont_mul<F, CANON>(acc, acc, b);
        } else if (op == 2) {
#pragma unroll
            for (int l = 0; l < N; l++) acc[l] = T[arg][l];
        } else {
#pragma unroll
            for (int l = 0; l < N; l++) T[arg][l] = acc[l];
        }
    }
}

template <class F>
FPQ void pow_inv_alpha(uint32_t (&r)[F::N], const uint32_t (&x)[F::N]) {
    constexpr int N = F::N;
    uint32_t acc[N];
    if constexpr (F::USE_PROGRAM) pow_program<F>(acc, x);
    else pow_window<F>(acc, x);
    if (F::LAZY) {  // lazy fields stay in [0, 2p + small) between multiplies (generated/fields.cuh)
        fp::cond_sub_p<F>(acc);
        if (F::FINAL_SUBS == 2) fp::cond_sub_p<F>(acc);
    }
#pragma unroll
    for (int l = 0; l < N; l++) r[l] = acc[l];
}

// Anemoi::sbox_layer for this thread's column (src/traits.rs:328-358):
//   x -= beta*y^2;  y -= x^(1/alpha);  x += beta*y^2 + delta
template <class F>
FPQ void sbox_column(uint32_t (&x)[F::N], uint32_t (&y)[F::N]) {
    constexpr int N = F::N;
    uint32_t t[N], g[N];
    fp::mont_sqr<F, true>(t, y);
    fp::mul_by_beta<F>(g, t);
    fp::sub_mod<F>(x, x, g);
    pow_inv_alpha<F>(t, x);
    fp::sub_mod<F>(y, y, t);
    fp::mont_sqr<F, true>(t, y);
    fp::mul_by_beta<F>(g, t);
    fp::add_mod<F>(x, x, g);
#pragma unroll
    for (int l = 0; l < N; l++) t[l] = F::delta(l);
    fp::add_mod<F>(x, x, t);
}

// Anemoi::mds_layer (src/traits.rs:129-157) incl. the PHT. COLS = 1: local. COLS = 2: the two lanes of
// a state exchange columns and both evaluate the (cheap) 2-column layer, keeping their own column.
// The 2-column layer is a real function call, on purpose. ptxas balances integer adds between the ALU pipe (IADD3) and the
// FMA pipe (IMAD.X / IMAD.IADD / IMAD.MOV) per FUNCTION: with this add-heavy layer inlined into the kernel it moves
// 10-15 adds of every multiply / squaring in the hot loops onto the FMA-heavy pipe those loops saturate (SASS: +2-5 %
// FMA-heavy cycles per squaring in every Anemoi-4-3 kernel). One call per round costs nothing measurable.
#if defined(ANEMOI_INLINE_LINEAR)
#define ANEMOI_LINEAR_Q FPQ
#else
#define ANEMOI_LINEAR_Q __device__ __noinline__
#endif
template <class F, int COLS>
ANEMOI_LINEAR_Q void linear_layer(uint32_t (&x)[F::N], uint32_t (&y)[F::N], int col, unsigned pair_mask) {
    constexpr int N = F::N;
    if (COLS == 1) {
        fp::add_mod<F>(y, y, x);
        fp::add_mod<F>(x, x, y);
    } else {
        uint32_t x0[N], x1[N], y0[N], y1[N], g[N];
#pragma unroll
        for (int l = 0; l < N; l++) {
            uint32_t xp = __shfl_xor_sync(pair_mask, x[l], 1);
            uint32_t yp = __shfl_xor_sync(pair_mask, y[l], 1);
            x0[l] = col ? xp : x[l];
            x1[l] = col ? x[l] : xp;
            y0[l] = col ? yp : y[l];
            y1[l] = col ? y[l] : yp;
        }
        fp::mul_by_beta<F>(g, x1);
        fp::add_mod<F>(x0, x0, g);  // state[0] += g * state[1]
        fp::mul_by_beta<F>(g, x0);
        fp::add_mod<F>(x1, x1, g);  // state[1] += g * state[0]
        fp::mul_by_beta<F>(g, y0);
        fp::add_mod<F>(y1, y1, g);  // state[3] += g * state[2]
        fp::mul_by_beta<F>(g, y1);
        fp::add_mod<F>(y0, y0, g);  // state[2] += g * state[3]
        // swap(state[2], state[3]) then PHT: the new y0 is the old y1 and vice versa
        if (col == 0) {
            fp::add_mod<F>(y, y1, x0);  // state[2] += state[0]
            fp::add_mod<F>(x, x0, y);   // state[0] += state[2]
        } else {
            fp::add_mod<F>(y, y0, x1);  // state[3] += state[1]
            fp::add_mod<F>(x, x1, y);   // state[1] += state[3]
        }
    }
}

template <class F, int COLS>
FPQ void permutation(uint32_t (&x)[F::N], uint32_t (&y)[F::N], int col, unsigned pair_mask) {
    constexpr int N = F::N;
    constexpr int ROUNDS = (COLS == 1) ? F::ROUNDS_2_1 : F::ROUNDS_4_3;
#pragma unroll 1
    for (int r = 0; r < ROUNDS; r++) {
        // Anemoi::ark_layer (src/traits.rs:113-125)
        const uint32_t* c = Tables<F>::ark(COLS) + ((r * COLS + col) * 2) * N;
        uint32_t k[N];
#pragma unroll
        for (int l = 0; l < N; l++) k[l] = c[l];
        fp::add_mod<F>(x, x, k);
#pragma unroll
        for (int l = 0; l < N; l++) k[l] = c[N + l];
        fp::add_mod<F>(y, y, k);
        linear_layer<F, COLS>(x, y, col, pair_mask);
        sbox_column<F>(x, y);
    }
    linear_layer<F, COLS>(x, y, col, pair_mask);
}

// One 31- / 47-byte chunk of Sponge::hash -> Montgomery felt (src/<field>/anemoi_2_1/hasher.rs:36-57,
// anemoi_4_3/hasher.rs:39-65): little-endian bytes, a 0x01 byte appended to a short last chunk.
template <class F>
FPQ void chunk_to_felt(uint32_t (&r)[F::N], const uint8_t* msg, unsigned long long nbytes, unsigned long long chunk) {
    constexpr int N = F::N;
    constexpr int B = F::BYTE_CHUNK;
    const unsigned long long start = chunk * B;
    const unsigned long long rem = nbytes - start;      // > 0
    const int clen = rem < (unsigned long long)B ? (int)rem : B;
    uint32_t v[N];
#pragma unroll
    for (int l = 0; l < N; l++) v[l] = 0;
#pragma unroll 1
    for (int i = 0; i < clen; i++) {
        uint32_t byte = msg[start + i];
#pragma unroll
        for (int l = 0; l < N; l++)
            if ((i >> 2) == l) v[l] |= byte << (8 * (i & 3));
    }
    if (clen < B) {  // only the last chunk can be short
#pragma unroll
        for (int l = 0; l < N; l++)
            if ((clen >> 2) == l) v[l] |= 1u << (8 * (clen & 3));
    }
    uint32_t r2[N];
#pragma unroll
    for (int l = 0; l < N; l++) r2[l] = F::r2(l);
    fp::mont_mul<F, true>(r, v, r2);  // canonical -> Montgomery
}

template <class F, int COLS>
FPQ void anemoi_body(const KernelArgs& a) {
    constexpr int N = F::N;
    constexpr int W = 2 * COLS;
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long unit = t / COLS;
    const int col = (int)(t % COLS);
    const bool active = unit < a.n;
    if (!active) unit = a.n - 1;  // tail lanes shadow the last unit so warp shuffles stay converged
    const unsigned lane = threadIdx.x & 31;
    const unsigned pair_mask = (COLS == 2) ? (3u << (lane & ~1u)) : 0xffffffffu;

    uint32_t x[N], y[N];
    fp::set_zero<F>(x);
    fp::set_zero<F>(y);

    const int mode = a.mode;
    unsigned long long nperm = 1;
    unsigned long long msg_len = 0;          // felts (or chunks) in this unit's message
    const uint32_t* msg_felts = nullptr;     // MODE_HASH / MODE_HASH_RAGGED
    const uint8_t* msg_bytes = nullptr;      // MODE_HASH_BYTES
    unsigned long long msg_nbytes = 0;

    if (mode == MODE_PERMUTE || mode == MODE_SBOX || mode == MODE_COMPRESS || mode == MODE_COMPRESS4) {
        const uint32_t* s = a.in + unit * (W * N);
        load_felt<N>(x, s + col * N, a.vec16);
        load_felt<N>(y, s + (COLS + col) * N, a.vec16);
    } else if (mode == MODE_MERGE43) {
        load_felt<N>(x, a.in + unit * (2 * N), a.vec16);  // both columns get digests[0]; y = 0
    } else if (mode == MODE_TO_BYTES) {
        // AnemoiDigest::to_bytes (digest.rs:42-46): canonical little-endian = Montgomery-multiply by 1
        if (col == 0) {
            uint32_t one[N];
            load_felt<N>(x, a.in + unit * N, a.vec16);
#pragma unroll
            for (int l = 0; l < N; l++) one[l] = (l == 0) ? 1u : 0u;
            fp::mont_mul<F, true>(y, x, one);
            if (active) store_felt<N>(a.out + unit * N, y, a.vec16);
        }
        return;
    } else {
        if (mode == MODE_HASH) {
            msg_len = a.len;
            msg_felts = a.in + unit * a.len * N;
        } else if (mode == MODE_HASH_RAGGED) {
            const unsigned long long o0 = a.offsets[unit], o1 = a.offsets[unit + 1];
            msg_len = o1 - o0;
            msg_felts = a.in + o0 * N;
        } else if (mode == MODE_HASH_BYTES) {
            msg_nbytes = a.len;
            msg_bytes = reinterpret_cast<const uint8_t*>(a.in) + unit * a.len;
            msg_len = (msg_nbytes + F::BYTE_CHUNK - 1) / F::BYTE_CHUNK;
        } else {  // MODE_HASH_BYTES_RAGGED
            const unsigned long long o0 = a.offsets[unit], o1 = a.offsets[unit + 1];
            msg_nbytes = o1 - o0;
            msg_bytes = reinterpret_cast<const uint8_t*>(a.in) + o0;
            msg_len = (msg_nbytes + F::BYTE_CHUNK - 1) / F::BYTE_CHUNK;
        }
        // 2-1: one permutation per element, no padding (hasher.rs:68-85).
        // 4-3: rate 3; a final padded block iff len % 3 != 0 (hasher.rs:93-129).
        nperm = (COLS == 1) ? msg_len : (msg_len + 2) / 3;
    }

#pragma unroll 1
    for (unsigned long long it = 0; it < nperm; it++) {
        if ((mode >= MODE_HASH && mode <= MODE_HASH_BYTES) || mode == MODE_HASH_BYTES_RAGGED) {
            // absorb: slot s of the rate <- element RATE*it + s; the slot just past the end gets the
            // padding 1 (4-3 only; it exists only when len % 3 != 0, i.e. sigma == 0).
            // 2-1 slots: {x}. 4-3 slots: s=0 -> col0.x, s=1 -> col1.x, s=2 -> col0.y.
            constexpr int RATE = (COLS == 1) ? 1 : 3;
#pragma unroll
            for (int s = 0; s < RATE; s++) {
                const int owner = (COLS == 1) ? 0 : (s == 1 ? 1 : 0);
                if (owner != col) continue;
                const unsigned long long idx = it * RATE + s;
                uint32_t e[N];
                bool have = false;
                if (idx < msg_len) {
                    if (mode == MODE_HASH_BYTES || mode == MODE_HASH_BYTES_RAGGED) chunk_to_felt<F>(e, msg_bytes, msg_nbytes, idx);
                    else load_felt<N>(e, msg_felts + idx * N, a.vec16);
                    have = true;
                } else if (COLS == 2 && idx == msg_len) {
                    fp::set_one<F>(e);
                    have = true;
                }
                if (have) {
                    if (s == 2) fp::add_mod<F>(y, y, e);
                    else fp::add_mod<F>(x, x, e);
                }
            }
        }
        if (mode == MODE_SBOX) sbox_column<F>(x, y);
        else permutation<F, COLS>(x, y, col, pair_mask);
    }

    // ---- epilogue
    if (mode == MODE_PERMUTE || mode == MODE_SBOX) {
        if (active) {
            uint32_t* s = a.out + unit * (W * N);
            store_felt<N>(s + col * N, x, a.vec16);
            store_felt<N>(s + (COLS + col) * N, y, a.vec16);
        }
    } else if (mode == MODE_COMPRESS || mode == MODE_COMPRESS4) {
        // Jive: sum of the inputs and of the permuted state over each group of W/k (hasher.rs:96-103,
        // 4-3 :148-179). Per column: in_x + in_y + x + y.
        uint32_t e[N], s[N];
        const uint32_t* src = a.in + unit * (W * N);
        fp::add_mod<F>(s, x, y);
        load_felt<N>(e, src + col * N, a.vec16);
        fp::add_mod<F>(s, s, e);
        load_felt<N>(e, src + (COLS + col) * N, a.vec16);
        fp::add_mod<F>(s, s, e);
        if (mode == MODE_COMPRESS4 && COLS == 2) {
#pragma unroll
            for (int l = 0; l < N; l++) e[l] = __shfl_xor_sync(pair_mask, s[l], 1);
            fp::add_mod<F>(s, s, e);
            if (active && col == 0) store_felt<N>(a.out + unit * N, s, a.vec16);
        } else {
            if (active) store_felt<N>(a.out + (unit * COLS + col) * N, s, a.vec16);
        }
    } else {
        // sponge digest = state[0] (the capacity tweak `state[W-1] += sigma/1` cannot reach it)
        if (active && col == 0) store_felt<N>(a.out + unit * N, x, a.vec16);
    }
}

// F with the code-generation switches of ONE kernel instantiation (see fp.cuh "Carry fix-ups" and tools/gen_params.py:
// CARRY_CHAIN is measured per field AND per instantiation, because ptxas balances the pipes per kernel). LATENCY selects
// the form for batches below a wave: never chained, since with a single warp per SM sub-partition nothing else hides the
// serialisation the chained fix-ups introduce (measured: Pallas 4-3 1.49 vs 1.90 ms, BLS12-377 2-1 6.83 vs 8.49 ms per
// lone launch).
template <class F, int COLS, bool LATENCY>
struct Variant : F {
    static constexpr bool CARRY_CHAIN = !LATENCY && (COLS == 1 ? F::CARRY_CHAIN_2_1 : F::CARRY_CHAIN_4_3);
};
template <class F, int COLS, bool LATENCY>
struct Tables<Variant<F, COLS, LATENCY>> : Tables<F> {};

// Throughput form: F::BLOCK-thread blocks, F::MIN_BLOCKS resident per SM (register-capped so that the FMA-heavy pipe
// always has 16+ warps to draw from).
template <class F, int COLS>
__global__ void __launch_bounds__(F::BLOCK, COLS == 1 ? F::MIN_BLOCKS_2_1 : F::MIN_BLOCKS_4_3) anemoi_kernel(KernelArgs a) {
    anemoi_body<Variant<F, COLS, false>, COLS>(a);
}

// Latency form, launched when the batch gives at most two warps per SM sub-partition (upper Merkle levels, small API
// calls): one warp per block so the few warps spread over all SMs, no register cap, unchained carries. Only instantiated
// where it differs from the throughput form.
template <class F, int COLS>
__global__ void __launch_bounds__(32, 1) anemoi_kernel_lat(KernelArgs a) {
    anemoi_body<Variant<F, COLS, true>, COLS>(a);
}

// Diagnostic kernel: ONE layer of the round function on a batch of states, in place -- the reference exposes
// ark_layer / mds_layer / sbox_layer / round as trait methods (src/traits.rs:113-157, 328-367); this lets each of
// them be compared with the oracle in isolation. a.mode = MODE_LAYER_ARK / MDS / ROUND, a.len = round index.
// (Kept apart from anemoi_kernel so that the hot kernel's code and register allocation are untouched.)
template <class F, int COLS>
__global__ void __launch_bounds__(F::BLOCK, F::MIN_BLOCKS) anemoi_layer_kernel(KernelArgs a) {
    constexpr int N = F::N;
    constexpr int W = 2 * COLS;
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long unit = t / COLS;
    const int col = (int)(t % COLS);
    const bool active = unit < a.n;
    if (!active) unit = a.n - 1;
    const unsigned lane = threadIdx.x & 31;
    const unsigned pair_mask = (COLS == 2) ? (3u << (lane & ~1u)) : 0xffffffffu;
    uint32_t x[N], y[N];
    const uint32_t* s = a.in + unit * (W * N);
    load_felt<N>(x, s + col * N, a.vec16);
    load_felt<N>(y, s + (COLS + col) * N, a.vec16);
    if (a.mode == MODE_LAYER_ARK || a.mode == MODE_LAYER_ROUND) {
        // Anemoi::ark_layer (src/traits.rs:113-125)
        const uint32_t* c = Tables<F>::ark(COLS) + (((int)a.len * COLS + col) * 2) * N;
        uint32_t k[N];
#pragma unroll
        for (int l = 0; l < N; l++) k[l] = c[l];
        fp::add_mod<F>(x, x, k);
#pragma unroll
        for (int l = 0; l < N; l++) k[l] = c[N + l];
        fp::add_mod<F>(y, y, k);
    }
    if (a.mode == MODE_LAYER_MDS || a.mode == MODE_LAYER_ROUND) linear_layer<F, COLS>(x, y, col, pair_mask);
    if (a.mode == MODE_LAYER_ROUND) sbox_column<F>(x, y);
    if (active) {
        uint32_t* d = a.out + unit * (W * N);
        store_felt<N>(d + col * N, x, a.vec16);
        store_felt<N>(d + (COLS + col) * N, y, a.vec16);
    }
}

}  // namespace anemoi
