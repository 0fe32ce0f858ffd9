/*
 * CPU oracle #2 (TEST INFRASTRUCTURE ONLY): plain-C restatement of the reference's algorithm on 64-bit
 * limbs, the way the reference computes it on a CPU -- Montgomery CIOS over u64 limbs with 128-bit
 * products (what arkworks' MontBackend does), the reference's own hard-coded addition chains for
 * x^(1/alpha), its round/mode logic line by line. It is the checker for the CUDA path and the timed
 * "cpu_baseline" of bench.py; only tests/, __graft_entry__.smoke() and bench.py may load it. The product
 * (libanemoi_b200.so) never links or calls it.
 *
 * Parity status: PINNED -- tests/test_oracle_c.py checks every entry point against all 420 known-answer
 * vectors of the reference's own tests (tests/golden/kat.json) and against the big-integer oracle
 * (oracle/anemoi_ref.py) on random inputs.
 *
 * The arithmetic dependency of the reference is NOT in the reference tree: ark-ff ^0.4.0
 * (Fp<MontBackend<_, N>, N>), ark-bls12-377 / ark-bls12-381 / ark-bn254 / ark-pallas ^0.4.0
 * (Cargo.toml:15-22, no lockfile). Its published algorithm (Montgomery multiplication, coarsely
 * integrated operand scanning, R = 2^(64 N), canonical outputs) is restated in mont_mul() below.
 *
 * Reference lines followed (relative to the reference's src/):
 *   traits.rs:78-91    mul_by_generator   -> mul_by_generator()  (incl. the generic arm for beta = 22)
 *   traits.rs:113-125  ark_layer          -> ark_layer()
 *   traits.rs:129-157  mds_layer          -> mds_layer()
 *   traits.rs:328-358  sbox_layer         -> sbox_layer()
 *   traits.rs:361-378  round, permutation -> permutation()
 *   <field>/sbox.rs    exp_by_inv_alpha   -> exp_by_inv_alpha()  (chain tables in params_gen.h)
 *   <field>/anemoi_2_1/hasher.rs:18-110, <field>/anemoi_4_3/hasher.rs:18-179 -> the oracle_* entries
 *   <field>/anemoi_x/digest.rs:42-46      -> oracle_digest_to_bytes()
 *
 * Data layout at this boundary = the product's C ABI: Montgomery form, N64 little-endian u64 limbs.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "params_gen.h"

typedef unsigned __int128 u128;
#define MAXN 6
#define INL static inline __attribute__((always_inline))

/* ---- field arithmetic (n = 4 or 6, constant after inlining) -------------------------------------- */

INL int geq(const uint64_t* a, const uint64_t* p, int n) {
    for (int i = n - 1; i >= 0; i--) {
        if (a[i] > p[i]) return 1;
        if (a[i] < p[i]) return 0;
    }
    return 1;
}

INL void sub_n(uint64_t* r, const uint64_t* a, const uint64_t* b, int n) {
    uint64_t borrow = 0;
    for (int i = 0; i < n; i++) {
        u128 t = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)t;
        borrow = (uint64_t)(t >> 64) & 1;
    }
}

INL void add_mod(uint64_t* r, const uint64_t* a, const uint64_t* b, const uint64_t* p, int n) {
    uint64_t carry = 0, t[MAXN];
    for (int i = 0; i < n; i++) {
        u128 s = (u128)a[i] + b[i] + carry;
        t[i] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
    }
    if (carry || geq(t, p, n)) sub_n(t, t, p, n);
    for (int i = 0; i < n; i++) r[i] = t[i];
}

INL void sub_mod(uint64_t* r, const uint64_t* a, const uint64_t* b, const uint64_t* p, int n) {
    uint64_t borrow = 0, t[MAXN];
    for (int i = 0; i < n; i++) {
        u128 s = (u128)a[i] - b[i] - borrow;
        t[i] = (uint64_t)s;
        borrow = (uint64_t)(s >> 64) & 1;
    }
    if (borrow) {
        uint64_t carry = 0;
        for (int i = 0; i < n; i++) {
            u128 s = (u128)t[i] + p[i] + carry;
            t[i] = (uint64_t)s;
            carry = (uint64_t)(s >> 64);
        }
    }
    for (int i = 0; i < n; i++) r[i] = t[i];
}

/* Montgomery multiplication, CIOS with the "no-carry" fusion arkworks uses when the modulus leaves the
 * top bit of the top limb free (true for all 7 fields): the a*b_i row and the m*p row run as two
 * interleaved carry chains and the running value never needs an extra limb.
 * r = a * b / 2^(64 n) mod p, canonical. */
INL void mont_mul(uint64_t* r, const uint64_t* a, const uint64_t* b, const uint64_t* p, uint64_t n0inv, int n) {
    uint64_t t[MAXN];
    for (int i = 0; i < n; i++) t[i] = 0;
    for (int i = 0; i < n; i++) {
        u128 x = (u128)a[0] * b[i] + t[0];
        uint64_t c = (uint64_t)(x >> 64);
        const uint64_t lo = (uint64_t)x;
        const uint64_t m = lo * n0inv;
        u128 y = (u128)m * p[0] + lo;
        uint64_t c2 = (uint64_t)(y >> 64);
        for (int j = 1; j < n; j++) {
            x = (u128)a[j] * b[i] + t[j] + c;
            c = (uint64_t)(x >> 64);
            y = (u128)m * p[j] + (uint64_t)x + c2;
            c2 = (uint64_t)(y >> 64);
            t[j - 1] = (uint64_t)y;
        }
        t[n - 1] = c + c2;
    }
    if (geq(t, p, n)) sub_n(t, t, p, n);
    for (int i = 0; i < n; i++) r[i] = t[i];
}

/* Montgomery squaring the way arkworks' square_in_place does it: upper-triangle products once, doubled,
 * plus the diagonal, then n word-by-word Montgomery reduction rows. Same canonical result as mont_mul(a, a). */
INL void mont_sqr(uint64_t* out, const uint64_t* a, const uint64_t* p, uint64_t n0inv, int n) {
    uint64_t r[2 * MAXN];
    for (int i = 0; i < 2 * n; i++) r[i] = 0;
    for (int i = 0; i < n - 1; i++) {
        uint64_t c = 0;
        for (int j = i + 1; j < n; j++) {
            u128 x = (u128)a[i] * a[j] + r[i + j] + c;
            r[i + j] = (uint64_t)x;
            c = (uint64_t)(x >> 64);
        }
        r[i + n] = c;
    }
    r[2 * n - 1] = r[2 * n - 2] >> 63;
    for (int i = 2 * n - 2; i >= 2; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 63);
    r[1] <<= 1;
    uint64_t c = 0;
    for (int i = 0; i < n; i++) {
        u128 x = (u128)a[i] * a[i] + r[2 * i] + c;
        r[2 * i] = (uint64_t)x;
        x = (u128)r[2 * i + 1] + (uint64_t)(x >> 64);
        r[2 * i + 1] = (uint64_t)x;
        c = (uint64_t)(x >> 64);
    }
    uint64_t carry2 = 0;
    for (int i = 0; i < n; i++) {
        const uint64_t m = r[i] * n0inv;
        u128 x = (u128)m * p[0] + r[i];
        uint64_t cc = (uint64_t)(x >> 64);
        for (int j = 1; j < n; j++) {
            x = (u128)m * p[j] + r[i + j] + cc;
            r[i + j] = (uint64_t)x;
            cc = (uint64_t)(x >> 64);
        }
        x = (u128)r[i + n] + cc + carry2;
        r[i + n] = (uint64_t)x;
        carry2 = (uint64_t)(x >> 64);
    }
    uint64_t t[MAXN];
    for (int i = 0; i < n; i++) t[i] = r[n + i];
    if (carry2 || geq(t, p, n)) sub_n(t, t, p, n);
    for (int i = 0; i < n; i++) out[i] = t[i];
}

typedef struct {
    const anemoi_field_params* f;
    int n, inst, cols, width, rate, rounds;
    uint64_t beta_mont[MAXN]; /* F::from(beta): used by the generic arm of mul_by_generator */
} ctx_t;

INL void dbl(uint64_t* r, const uint64_t* a, const ctx_t* c) { add_mod(r, a, a, c->f->p, c->n); }

/* traits.rs:78-91 */
INL void mul_by_generator(uint64_t* r, const uint64_t* x, const ctx_t* c) {
    uint64_t t[MAXN], u[MAXN];
    const uint64_t* p = c->f->p;
    const int n = c->n;
    switch (c->f->beta) {
        case 2: dbl(r, x, c); break;
        case 3: dbl(t, x, c); add_mod(r, t, x, p, n); break;
        case 5: dbl(t, x, c); dbl(u, t, c); add_mod(r, u, x, p, n); break;
        case 7: dbl(t, x, c); add_mod(u, t, x, p, n); dbl(t, u, c); add_mod(r, t, x, p, n); break;
        case 9: dbl(t, x, c); dbl(u, t, c); dbl(t, u, c); add_mod(r, t, x, p, n); break;
        case 11: dbl(t, x, c); dbl(u, t, c); add_mod(t, u, x, p, n); dbl(u, t, c); add_mod(r, u, x, p, n); break;
        case 13: dbl(t, x, c); add_mod(u, t, x, p, n); dbl(t, u, c); add_mod(u, t, x, p, n); dbl(t, u, c); add_mod(r, t, x, p, n); break;
        case 15: dbl(t, x, c); dbl(u, t, c); dbl(t, u, c); dbl(u, t, c); sub_mod(r, u, x, p, n); break;
        case 17: dbl(t, x, c); dbl(u, t, c); dbl(t, u, c); dbl(u, t, c); add_mod(r, u, x, p, n); break;
        default: mont_mul(r, c->beta_mont, x, p, c->f->n0inv, n); break; /* F::from(beta) * x */
    }
}

/* <field>/sbox.rs exp_by_inv_alpha: the reference's addition chain, value 0 = x, step i -> value i+1 */
INL void exp_by_inv_alpha(uint64_t* r, const uint64_t* x, const ctx_t* c) {
    const int n = c->n, len = c->f->chain_len;
    uint64_t v[512][MAXN];
    for (int i = 0; i < n; i++) v[0][i] = x[i];
    for (int s = 0; s < len; s++) {
        const int ia = c->f->chain[s][0], ib = c->f->chain[s][1];
        if (ia == ib) mont_sqr(v[s + 1], v[ia], c->f->p, c->f->n0inv, n); /* `.square()` steps of sbox.rs */
        else mont_mul(v[s + 1], v[ia], v[ib], c->f->p, c->f->n0inv, n);
    }
    for (int i = 0; i < n; i++) r[i] = v[len][i];
}

/* traits.rs:113-125 */
INL void ark_layer(uint64_t* s, int r, const ctx_t* c) {
    const int n = c->n, cols = c->cols;
    const uint64_t* C = c->f->arkc[c->inst] + (size_t)r * cols * n;
    const uint64_t* D = c->f->arkd[c->inst] + (size_t)r * cols * n;
    for (int i = 0; i < cols; i++) {
        add_mod(s + i * n, s + i * n, C + i * n, c->f->p, n);
        add_mod(s + (cols + i) * n, s + (cols + i) * n, D + i * n, c->f->p, n);
    }
}

/* traits.rs:129-157 */
INL void mds_layer(uint64_t* s, const ctx_t* c) {
    const int n = c->n;
    const uint64_t* p = c->f->p;
    uint64_t g[MAXN];
    if (c->cols == 1) {
        add_mod(s + n, s + n, s, p, n);
        add_mod(s, s, s + n, p, n);
    } else {
        uint64_t *s0 = s, *s1 = s + n, *s2 = s + 2 * n, *s3 = s + 3 * n, tmp[MAXN];
        mul_by_generator(g, s1, c); add_mod(s0, s0, g, p, n);
        mul_by_generator(g, s0, c); add_mod(s1, s1, g, p, n);
        mul_by_generator(g, s2, c); add_mod(s3, s3, g, p, n);
        mul_by_generator(g, s3, c); add_mod(s2, s2, g, p, n);
        for (int i = 0; i < n; i++) { tmp[i] = s2[i]; s2[i] = s3[i]; s3[i] = tmp[i]; }
        add_mod(s2, s2, s0, p, n);
        add_mod(s3, s3, s1, p, n);
        add_mod(s0, s0, s2, p, n);
        add_mod(s1, s1, s3, p, n);
    }
}

/* traits.rs:328-358 */
INL void sbox_layer(uint64_t* s, const ctx_t* c) {
    const int n = c->n, cols = c->cols;
    const uint64_t* p = c->f->p;
    for (int i = 0; i < cols; i++) {
        uint64_t *x = s + i * n, *y = s + (cols + i) * n, y2[MAXN], g[MAXN], t[MAXN];
        mont_sqr(y2, y, p, c->f->n0inv, n);
        mul_by_generator(g, y2, c);
        sub_mod(x, x, g, p, n);
        exp_by_inv_alpha(t, x, c);
        sub_mod(y, y, t, p, n);
        mont_sqr(y2, y, p, c->f->n0inv, n);
        mul_by_generator(g, y2, c);
        add_mod(x, x, g, p, n);
        add_mod(x, x, c->f->delta, p, n);
    }
}

/* traits.rs:361-378 */
INL void permutation(uint64_t* s, const ctx_t* c) {
    for (int r = 0; r < c->rounds; r++) {
        ark_layer(s, r, c);
        mds_layer(s, c);
        sbox_layer(s, c);
    }
    mds_layer(s, c);
}

static int make_ctx(ctx_t* c, int field, int inst) {
    if (field < 0 || field >= 7 || inst < 0 || inst > 1) return -1;
    c->f = &ANEMOI_FIELDS[field];
    c->n = c->f->n64;
    c->inst = inst;
    c->cols = inst == 0 ? 1 : 2;
    c->width = 2 * c->cols;
    c->rate = inst == 0 ? 1 : 3;
    c->rounds = c->f->rounds[inst];
    /* F::from(beta as u64): canonical beta -> Montgomery = beta * R^2 / R */
    uint64_t b[MAXN] = {0};
    b[0] = (uint64_t)c->f->beta;
    mont_mul(c->beta_mont, b, c->f->r2, c->f->p, c->f->n0inv, c->n);
    return 0;
}

/* One kernel per limb count so n is a compile-time constant inside the hot loops. */
#define DISPATCH(c, CALL4, CALL6) do { if ((c)->n == 4) { CALL4; } else { CALL6; } } while (0)

static void permutation4(uint64_t* s, const ctx_t* c) { ctx_t k = *c; k.n = 4; permutation(s, &k); }
static void permutation6(uint64_t* s, const ctx_t* c) { ctx_t k = *c; k.n = 6; permutation(s, &k); }
static void sbox4(uint64_t* s, const ctx_t* c) { ctx_t k = *c; k.n = 4; sbox_layer(s, &k); }
static void sbox6(uint64_t* s, const ctx_t* c) { ctx_t k = *c; k.n = 6; sbox_layer(s, &k); }
static void perm(uint64_t* s, const ctx_t* c) { DISPATCH(c, permutation4(s, c), permutation6(s, c)); }

/* Jive::compress_k on one state (anemoi_2_1/hasher.rs:96-110, anemoi_4_3/hasher.rs:148-179) */
static void compress_one(const ctx_t* c, int k, const uint64_t* in, uint64_t* out) {
    const int n = c->n, w = c->width;
    uint64_t s[4 * MAXN];
    memcpy(s, in, (size_t)w * n * 8);
    perm(s, c);
    const int cc = w / k;
    for (int i = 0; i < cc; i++) {
        uint64_t acc[MAXN] = {0};
        for (int j = 0; j < k; j++) {
            add_mod(acc, acc, in + (size_t)(i + cc * j) * n, c->f->p, n);
            add_mod(acc, acc, s + (size_t)(i + cc * j) * n, c->f->p, n);
        }
        memcpy(out + (size_t)i * n, acc, (size_t)n * 8);
    }
}

/* Sponge::hash_field on one message (anemoi_2_1/hasher.rs:68-85, anemoi_4_3/hasher.rs:93-129) */
static void hash_field_one(const ctx_t* c, const uint64_t* e, size_t len, uint64_t* digest) {
    const int n = c->n;
    const uint64_t* p = c->f->p;
    uint64_t s[4 * MAXN];
    memset(s, 0, sizeof(s));
    if (c->width == 2) {
        for (size_t k = 0; k < len; k++) {
            add_mod(s, s, e + k * n, p, n);
            perm(s, c);
        }
        add_mod(s + n, s + n, c->f->one, p, n);
    } else {
        const int sigma = (len % 3 == 0);
        int i = 0;
        for (size_t k = 0; k < len; k++) {
            add_mod(s + i * n, s + i * n, e + k * n, p, n);
            i++;
            if (i % 3 == 0) {
                perm(s, c);
                i = 0;
            }
        }
        if (sigma) add_mod(s + 3 * n, s + 3 * n, c->f->one, p, n);
        if (!sigma) {
            add_mod(s + i * n, s + i * n, c->f->one, p, n);
            perm(s, c);
        }
    }
    memcpy(digest, s, (size_t)n * 8);
}

/* byte chunk -> Montgomery felt (from_le_bytes_mod_order on the 32/48-byte buffer of Sponge::hash) */
static void chunk_to_felt(const ctx_t* c, const uint8_t* buf, uint64_t* out) {
    const int n = c->n;
    uint64_t v[MAXN];
    for (int i = 0; i < n; i++) {
        uint64_t w = 0;
        for (int b = 7; b >= 0; b--) w = (w << 8) | buf[8 * i + b];
        v[i] = w;
    }
    /* the buffer's top byte is always 0 and 2^(8(8n-1)) < p for every field, so v < p already */
    if (c->n == 4) mont_mul(out, v, c->f->r2, c->f->p, c->f->n0inv, 4);
    else mont_mul(out, v, c->f->r2, c->f->p, c->f->n0inv, 6);
}

/* ---- exported entry points (ctypes) --------------------------------------------------------------- */

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_threads(int t) {
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

int oracle_permute(int field, int inst, uint64_t* states, size_t n_states) {
    ctx_t c;
    if (make_ctx(&c, field, inst)) return -1;
    const size_t stride = (size_t)c.width * c.n;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n_states; i++) perm(states + (size_t)i * stride, &c);
    return 0;
}

int oracle_sbox_layer(int field, int inst, uint64_t* states, size_t n_states) {
    ctx_t c;
    if (make_ctx(&c, field, inst)) return -1;
    const size_t stride = (size_t)c.width * c.n;
    for (size_t i = 0; i < n_states; i++) DISPATCH(&c, sbox4(states + i * stride, &c), sbox6(states + i * stride, &c));
    return 0;
}

/* One layer in isolation: 0 = ark_layer(round), 1 = mds_layer, 2 = sbox_layer, 3 = round(round) (traits.rs:113-367) */
int oracle_layer(int field, int inst, int layer, int round, uint64_t* states, size_t n_states) {
    ctx_t c;
    if (make_ctx(&c, field, inst)) return -1;
    if (layer < 0 || layer > 3) return -1;
    if ((layer == 0 || layer == 3) && (round < 0 || round >= c.rounds)) return -5;
    const size_t stride = (size_t)c.width * c.n;
    for (size_t i = 0; i < n_states; i++) {
        uint64_t* s = states + i * stride;
        if (layer == 0 || layer == 3) ark_layer(s, round, &c);
        if (layer == 1 || layer == 3) mds_layer(s, &c);
        if (layer == 2 || layer == 3) DISPATCH(&c, sbox4(s, &c), sbox6(s, &c));
    }
    return 0;
}

int oracle_compress(int field, int inst, int k, const uint64_t* in, uint64_t* out, size_t n_states) {
    ctx_t c;
    if (make_ctx(&c, field, inst)) return -1;
    if (c.width == 2 ? (k != 2) : (k != 2 && k != 4)) return -4; /* the reference's assert!s */
    const size_t is = (size_t)c.width * c.n, os = (size_t)(c.width / k) * c.n;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n_states; i++) compress_one(&c, k, in + (size_t)i * is, out + (size_t)i * os);
    return 0;
}

int oracle_hash_field(int field, int inst, const uint64_t* elems, size_t n_msgs, size_t len, uint64_t* digests) {
    ctx_t c;
    if (make_ctx(&c, field, inst)) return -1;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n_msgs; i++)
        hash_field_one(&c, elems + (size_t)i * len * c.n, len, digests + (size_t)i * c.n);
    return 0;
}

int oracle_hash_field_ragged(int field, int inst, const uint64_t* elems, const uint64_t* offsets, size_t n_msgs,
                             uint64_t* digests) {
    ctx_t c;
    if (make_ctx(&c, field, inst)) return -1;
#pragma omp parallel for schedule(dynamic)
    for (long long i = 0; i < (long long)n_msgs; i++)
        hash_field_one(&c, elems + (size_t)offsets[i] * c.n, (size_t)(offsets[i + 1] - offsets[i]), digests + (size_t)i * c.n);
    return 0;
}

/* Sponge::hash (anemoi_2_1/hasher.rs:18-66, anemoi_4_3/hasher.rs:18-91): chunk, pad, then absorb exactly
 * like hash_field with num_elements in place of elems.len(). */
int oracle_hash_bytes(int field, int inst, const uint8_t* bytes, size_t n_msgs, size_t nbytes, uint64_t* digests) {
    ctx_t c;
    if (make_ctx(&c, field, inst)) return -1;
    const size_t B = (size_t)c.n * 8 - 1;
    const size_t nel = (nbytes + B - 1) / B;
    int rc = 0;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n_msgs; i++) {
        const uint8_t* msg = bytes + (size_t)i * nbytes;
        uint64_t felts_stack[64 * MAXN];
        uint64_t* felts = felts_stack;
        uint64_t* heap = NULL;
        if (nel > 64) {
            heap = (uint64_t*)malloc(nel * c.n * 8);
            felts = heap;
        }
        if (!felts) { rc = -8; continue; }
        for (size_t j = 0; j < nel; j++) {
            uint8_t buf[48] = {0};
            size_t start = j * B, clen = nbytes - start < B ? nbytes - start : B;
            memcpy(buf, msg + start, clen);
            if (j == nel - 1 && clen < B) buf[clen] = 1;
            chunk_to_felt(&c, buf, felts + j * c.n);
        }
        hash_field_one(&c, felts, nel, digests + (size_t)i * c.n);
        if (heap) free(heap);
    }
    return rc;
}

/* Sponge::merge: 2-1 = Jive (anemoi_2_1/hasher.rs:87-92); 4-3 copies digests[0] twice (anemoi_4_3/hasher.rs:131-144) */
int oracle_merge(int field, int inst, const uint64_t* pairs, uint64_t* out, size_t n) {
    ctx_t c;
    if (make_ctx(&c, field, inst)) return -1;
    if (c.width == 2) return oracle_compress(field, inst, 2, pairs, out, n);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; i++) {
        uint64_t s[4 * MAXN];
        memset(s, 0, sizeof(s));
        memcpy(s, pairs + (size_t)i * 2 * c.n, (size_t)c.n * 8);
        memcpy(s + c.n, pairs + (size_t)i * 2 * c.n, (size_t)c.n * 8);
        perm(s, &c);
        memcpy(out + (size_t)i * c.n, s, (size_t)c.n * 8);
    }
    return 0;
}

/* Not in the reference: iterate its node function level by level (left to right).
 * `scratch` holds n_leaves/arity + n_leaves/arity^2 + 1 felts (two ping-pong levels). */
int oracle_merkle_root(int field, int inst, int arity, const uint64_t* leaves, size_t n_leaves, uint64_t* scratch,
                       uint64_t* root) {
    ctx_t c;
    if (make_ctx(&c, field, inst)) return -1;
    if (arity != c.width) return -4;
    if (n_leaves == 0) return -5;
    uint64_t* ping = scratch;
    uint64_t* pong = scratch + (n_leaves / (size_t)arity + 1) * c.n;
    const uint64_t* src = leaves;
    size_t n = n_leaves;
    int level = 0;
    while (n > 1) {
        if (n % (size_t)arity) return -5;
        const size_t nodes = n / (size_t)arity;
        uint64_t* dst = (level & 1) ? pong : ping;
        oracle_compress(field, inst, arity, src, dst, nodes);
        src = dst;
        n = nodes;
        level++;
    }
    memcpy(root, src, (size_t)c.n * 8);
    return 0;
}

/* AnemoiDigest::to_bytes (digest.rs:42-46): canonical little-endian */
int oracle_digest_to_bytes(int field, const uint64_t* digests, uint8_t* bytes, size_t n) {
    ctx_t c;
    if (make_ctx(&c, field, 0)) return -1;
    for (size_t i = 0; i < n; i++) {
        uint64_t one[MAXN] = {1}, v[MAXN];
        if (c.n == 4) mont_mul(v, digests + i * 4, one, c.f->p, c.f->n0inv, 4);
        else mont_mul(v, digests + i * 6, one, c.f->p, c.f->n0inv, 6);
        for (int j = 0; j < c.n; j++)
            for (int b = 0; b < 8; b++) bytes[i * c.n * 8 + j * 8 + b] = (uint8_t)(v[j] >> (8 * b));
    }
    return 0;
}
