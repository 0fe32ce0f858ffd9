"""Field and instantiation metadata of the 7 x 2 Anemoi instances (host side of the boundary).

Mirrors src/<field>/mod.rs (the `Felt` alias) and src/<field>/anemoi_*/mod.rs:20-38 (sizes) of the
reference. The moduli are those of the arkworks curve crates the reference depends on. `Felt` values
cross the C ABI as Montgomery limbs (a * 2^(64*N64) mod p, little-endian u64), exactly the in-memory
form of an arkworks `Fp`; the conversions here are what `MontFp!` / `into_bigint` do on the Rust side."""
import numpy as np

FIELD_NAMES = ["bls12_377", "bls12_381", "bn_254", "ed_on_bls12_377", "jubjub", "pallas", "vesta"]

_MODULI = {
    "bls12_377": 0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001,
    "bls12_381": 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    "bn_254": 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
    "ed_on_bls12_377": 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
    "jubjub": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    "pallas": 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001,
    "vesta": 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001,
}

INST_2_1, INST_4_3 = 0, 1
_MASK64 = (1 << 64) - 1


class Field:
    def __init__(self, name):
        self.name = name
        self.id = FIELD_NAMES.index(name)
        self.p = _MODULI[name]
        self.n64 = (self.p.bit_length() + 63) // 64
        self.R = 1 << (64 * self.n64)
        self.Rinv = pow(self.R, -1, self.p)
        self.felt_bytes = 8 * self.n64
        self.byte_chunk = self.felt_bytes - 1  # 31 / 47 (hasher.rs `hash`)

    # --- Felt <-> limbs -------------------------------------------------------------------------
    def to_mont(self, a):
        return (int(a) % self.p) * self.R % self.p

    def from_mont(self, m):
        return int(m) * self.Rinv % self.p

    def encode(self, values):
        """canonical ints (any nesting flattened by the caller) -> uint64 array (len, N64), Montgomery form."""
        out = np.empty((len(values), self.n64), dtype=np.uint64)
        for i, v in enumerate(values):
            m = self.to_mont(v)
            for j in range(self.n64):
                out[i, j] = (m >> (64 * j)) & _MASK64
        return out

    def decode(self, limbs):
        """uint64 array (..., N64) in Montgomery form -> list of canonical ints (flattened)."""
        a = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, self.n64)
        out = []
        for row in a:
            m = 0
            for j in range(self.n64):
                m |= int(row[j]) << (64 * j)
            out.append(self.from_mont(m))
        return out

    def random_mont(self, n, seed):
        """n uniform canonical residues, returned directly as Montgomery limbs (uniform either way):
        the synthetic-input generator of SURVEY.md 8(d). Rejection sampling on masked 64-bit draws."""
        rng = np.random.Generator(np.random.Philox(seed))
        top_bits = self.p.bit_length() - 64 * (self.n64 - 1)
        mask = np.uint64((1 << top_bits) - 1)
        p_limbs = [(self.p >> (64 * j)) & _MASK64 for j in range(self.n64)]
        out = np.empty((n, self.n64), dtype=np.uint64)
        filled = 0
        while filled < n:
            m = max(1024, int((n - filled) * 1.6))
            cand = rng.integers(0, 1 << 64, size=(m, self.n64), dtype=np.uint64)
            cand[:, -1] &= mask
            # lexicographic compare from the top limb: keep cand < p
            lt = np.zeros(m, dtype=bool)
            eq = np.ones(m, dtype=bool)
            for j in range(self.n64 - 1, -1, -1):
                pj = np.uint64(p_limbs[j])
                lt |= eq & (cand[:, j] < pj)
                eq &= cand[:, j] == pj
            good = cand[lt]
            take = min(len(good), n - filled)
            out[filled:filled + take] = good[:take]
            filled += take
        return out


FIELDS = {name: Field(name) for name in FIELD_NAMES}
