"""CPU oracle #1: big-integer restatement of the reference algorithm (TEST INFRASTRUCTURE ONLY).

This module is the *checker*. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import it; the product path (anemoi_rust_b200/) never does.

Parity status: PINNED. `tests/test_oracle_kat.py` checks every function below against all 420
known-answer vectors the reference's own unit tests hold (tests/golden/kat.json, extracted by
tools/extract_fixtures.py from the reference sources).

All values are canonical integers in [0, p). Montgomery conversion helpers are at the bottom; the
reference's arithmetic (arkworks ark-ff ^0.4.0 `Fp<MontBackend<_, N>, N>`, not vendored in the
reference tree) stores a field element a as the N64 little-endian u64 limbs of a * 2^(64*N64) mod p.

Reference lines followed (all under /root/reference/src):
  traits.rs:78-91     mul_by_generator      -> Params.g
  traits.rs:113-125   ark_layer             -> ark
  traits.rs:136-157   mds_layer (1, 2 cols) -> mds
  traits.rs:328-358   sbox_layer            -> sbox
  traits.rs:361-378   round, permutation    -> permutation
  <field>/sbox.rs     exp_by_inv_alpha      -> exp_by_inv_alpha (the reference's own addition chain)
  <field>/anemoi_2_1/hasher.rs:18-110       -> hash, hash_field, merge, compress, compress_k (2-1)
  <field>/anemoi_4_3/hasher.rs:18-179       -> hash, hash_field, merge, compress, compress_k (4-3)
  <field>/anemoi_*/digest.rs:42-46          -> digest_to_bytes
"""
import json
import os

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

FIELDS = ["bls12_377", "bls12_381", "bn_254", "ed_on_bls12_377", "jubjub", "pallas", "vesta"]
INSTS = ["anemoi_2_1", "anemoi_4_3"]

_params_cache = None


def load_params():
    global _params_cache
    if _params_cache is None:
        with open(os.path.join(_GOLDEN, "params.json")) as f:
            _params_cache = json.load(f)
    return _params_cache


class Params:
    """One (field, instantiation): constants bound exactly as <field>/anemoi_*/mod.rs:40-61 does."""

    def __init__(self, field, inst):
        fp = load_params()[field]
        ip = fp["inst"][inst]
        self.field, self.inst = field, inst
        self.p = int(fp["modulus"])
        self.n64 = fp["n64"]
        self.alpha = fp["alpha"]
        self.beta = fp["beta"]
        self.delta = int(fp["delta"])
        self.inv_alpha = int(fp["inv_alpha"])
        self.chain = fp["chain"]
        self.byte_chunk = fp["byte_chunk"]
        self.width = ip["width"]
        self.rate = ip["rate"]
        self.cols = ip["cols"]
        self.rounds = ip["rounds"]
        self.C = [int(v) for v in ip["C"]]
        self.D = [int(v) for v in ip["D"]]
        self.R = 1 << (64 * self.n64)

    # traits.rs:78-91 -- beta * x (the doubling chains and the generic arm give the same residue)
    def g(self, x):
        return (self.beta * x) % self.p


_pcache = {}


def params(field, inst):
    key = (field, inst)
    if key not in _pcache:
        _pcache[key] = Params(field, inst)
    return _pcache[key]


def exp_by_inv_alpha(P, x):
    """<field>/sbox.rs exp_by_inv_alpha: the reference's fixed addition chain (SSA pairs)."""
    v = [x % P.p]
    for a, b in P.chain:
        v.append(v[a] * v[b] % P.p)
    return v[-1]


def ark(P, s, r):  # traits.rs:113-125
    c = P.cols
    for i in range(c):
        s[i] = (s[i] + P.C[r * c + i]) % P.p
        s[c + i] = (s[c + i] + P.D[r * c + i]) % P.p


def mds(P, s):  # traits.rs:129-157
    p = P.p
    if P.cols == 1:
        s[1] = (s[1] + s[0]) % p
        s[0] = (s[0] + s[1]) % p
    elif P.cols == 2:
        s[0] = (s[0] + P.g(s[1])) % p
        s[1] = (s[1] + P.g(s[0])) % p
        s[3] = (s[3] + P.g(s[2])) % p
        s[2] = (s[2] + P.g(s[3])) % p
        s[2], s[3] = s[3], s[2]
        s[2] = (s[2] + s[0]) % p
        s[3] = (s[3] + s[1]) % p
        s[0] = (s[0] + s[2]) % p
        s[1] = (s[1] + s[3]) % p
    else:
        raise NotImplementedError("only 1- and 2-column instances are instantiated by the reference")


def sbox(P, s):  # traits.rs:328-358
    c, p = P.cols, P.p
    for i in range(c):
        x, y = s[i], s[c + i]
        x = (x - P.g(y * y % p)) % p
        y = (y - exp_by_inv_alpha(P, x)) % p
        x = (x + P.g(y * y % p) + P.delta) % p
        s[i], s[c + i] = x, y


def permutation(P, s):  # traits.rs:361-378
    assert len(s) == P.width
    for r in range(P.rounds):
        ark(P, s, r)
        mds(P, s)
        sbox(P, s)
    mds(P, s)


def compress(P, e):
    """Jive::compress -- anemoi_2_1/hasher.rs:96-103, anemoi_4_3/hasher.rs:148-160."""
    assert len(e) == P.width
    s = list(e)
    permutation(P, s)
    if P.width == 2:
        return [(s[0] + s[1] + e[0] + e[1]) % P.p]
    c = P.cols
    return [(e[i] + e[i + c] + s[i] + s[i + c]) % P.p for i in range(c)]


def compress_k(P, e, k):
    """Jive::compress_k -- anemoi_2_1/hasher.rs:105-110 (k == 2 only), anemoi_4_3/hasher.rs:162-179."""
    if P.width == 2:
        assert k == 2
        return compress(P, e)
    assert len(e) == P.width
    assert P.width % k == 0
    assert k % 2 == 0
    s = list(e)
    permutation(P, s)
    c = P.width // k
    out = [0] * c
    for i in range(c):
        for j in range(k):
            out[i] = (out[i] + e[i + c * j] + s[i + c * j]) % P.p
    return out


def hash_field(P, elems):
    """Sponge::hash_field -- anemoi_2_1/hasher.rs:68-85, anemoi_4_3/hasher.rs:93-129."""
    p = P.p
    s = [0] * P.width
    if P.width == 2:
        for e in elems:
            s[0] = (s[0] + e) % p
            permutation(P, s)
        s[1] = (s[1] + 1) % p
        return s[0]
    sigma = 1 if len(elems) % P.rate == 0 else 0
    i = 0
    for e in elems:
        s[i] = (s[i] + e) % p
        i += 1
        if i % P.rate == 0:
            permutation(P, s)
            i = 0
    s[P.width - 1] = (s[P.width - 1] + sigma) % p
    if sigma == 0:
        s[i] = (s[i] + 1) % p
        permutation(P, s)
    return s[0]


def bytes_to_felts(P, data):
    """The chunking of Sponge::hash -- anemoi_2_1/hasher.rs:18-58, anemoi_4_3/hasher.rs:18-66."""
    B = P.byte_chunk
    n = (len(data) + B - 1) // B
    out = []
    for i in range(n):
        chunk = data[i * B : (i + 1) * B]
        if i < n - 1:
            buf = bytes(chunk) + b"\0"
        else:
            buf = bytearray(B + 1)
            buf[: len(chunk)] = chunk
            if len(chunk) < B:
                buf[len(chunk)] = 1
        out.append(int.from_bytes(bytes(buf), "little") % P.p)
    return out


def hash_bytes(P, data):
    """Sponge::hash. For 2-1 this is hash_field of the chunks. For 4-3 sigma is computed from the
    chunk count (hasher.rs:22-32, 36) -- the same rule as hash_field applied to the chunks."""
    return hash_field(P, bytes_to_felts(P, data))


def merge(P, d0, d1):
    """Sponge::merge -- 2-1: Jive (hasher.rs:87-92); 4-3: sponge that copies digests[0] twice (sic,
    anemoi_4_3/hasher.rs:131-144; digests[1] is never read)."""
    if P.width == 2:
        return compress(P, [d0, d1])[0]
    s = [d0, d0, 0, 0]
    permutation(P, s)
    return s[0]


def merkle_root(P, leaves, arity):
    """NOT in the reference (it has no tree code): iterate the reference's node function level by level.
    arity 2 on a 2-1 instance -> compress; arity 4 on a 4-3 instance -> compress_k(.,4);
    arity 2 on a 4-3 instance is not defined here (compress returns 2 elements)."""
    assert arity == P.width and arity in (2, 4)
    level = list(leaves)
    n = len(level)
    assert n >= 1
    while n > 1:
        assert n % arity == 0
        level = [compress_k(P, level[i : i + arity], arity)[0] for i in range(0, n, arity)]
        n = len(level)
    return level[0]


def digest_to_bytes(P, d):
    """AnemoiDigest::to_bytes -- digest.rs:42-46: canonical little-endian, 8*N64 bytes."""
    return int(d % P.p).to_bytes(8 * P.n64, "little")


# ---- Montgomery boundary helpers (layout of an arkworks `&[Fp]` slice) ------------------------

def to_mont(P, a):
    return (a * P.R) % P.p


def from_mont(P, m):
    return (m * pow(P.R, -1, P.p)) % P.p


def to_limbs(P, m):
    """canonical-or-Montgomery integer -> list of N64 u64 limbs, little-endian."""
    return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(P.n64)]


def from_limbs(limbs):
    v = 0
    for i, l in enumerate(limbs):
        v |= int(l) << (64 * i)
    return v
