#!/usr/bin/env python3
"""Extract parameters and known-answer vectors from the reference crate's sources.

Runs only in the build container (needs /root/reference); its OUTPUTS are committed:

  tests/golden/params.json   per field: modulus, alpha, beta, delta, inv_alpha, the reference's
                             addition chain for x^(1/alpha) in SSA form; per instantiation: rounds,
                             ARK constants C and D (canonical integers as decimal strings)
  tests/golden/kat.json      every known-answer vector of the reference's unit tests
                             (test_sbox, test_anemoi_hash, test_anemoi_hash_bytes, test_anemoi_jive)

Reference locations parsed (per field F, per instantiation I in {anemoi_2_1, anemoi_4_3}):
  src/F/sbox.rs                         ALPHA, INV_ALPHA, BETA, DELTA, fn exp_by_inv_alpha
  src/F/I/mod.rs                        STATE_WIDTH, RATE_WIDTH, NUM_COLUMNS, NUM_HASH_ROUNDS, test_sbox
  src/F/I/round_constants.rs            C, D
  src/F/I/hasher.rs                     test_anemoi_hash, test_anemoi_hash_bytes, test_anemoi_jive

The moduli are NOT in the reference (they come from the arkworks curve crates it depends on:
ark-bls12-377 / ark-bls12-381 / ark-bn254 / ark-pallas ^0.4.0 + ed_on_* scalar fields); they are
stated here from the curve definitions and validated indirectly: delta*beta == 1 (mod p),
alpha*inv_alpha == 1 (mod p-1), the addition chain evaluates to inv_alpha, and all KATs pass.
"""
import json
import os
import re
import sys

REF = os.environ.get("ANEMOI_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# field order = module order of src/lib.rs:27-64
FIELDS = ["bls12_377", "bls12_381", "bn_254", "ed_on_bls12_377", "jubjub", "pallas", "vesta"]
INSTS = ["anemoi_2_1", "anemoi_4_3"]

MODULI = {
    "bls12_377": 0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001,
    "bls12_381": 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    "bn_254": 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
    "ed_on_bls12_377": 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
    "jubjub": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    "pallas": 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001,
    "vesta": 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001,
}


def read(*parts):
    with open(os.path.join(REF, *parts)) as f:
        return f.read()


def const_u32(src, name):
    m = re.search(r"const\s+%s\s*:\s*(?:u32|usize)\s*=\s*(\d+)\s*;" % name, src)
    assert m, name
    return int(m.group(1))


def const_felt(src, name):
    m = re.search(r"const\s+%s\s*:\s*Felt\s*=\s*MontFp!\(\s*\"(\d+)\"\s*\)\s*;" % name, src)
    assert m, name
    return int(m.group(1))


def parse_chain(src):
    """fn exp_by_inv_alpha -> SSA list [[a, b], ...]; value 0 is x, step i defines value i+1."""
    m = re.search(r"fn exp_by_inv_alpha\(x: &Felt\) -> Felt \{(.*?)\n\}", src, re.S)
    body = m.group(1)
    env = {"x": 0}
    ssa = []

    def emit(a, b):
        ssa.append([env[a], env[b]])
        return len(ssa)

    for raw in body.split("\n"):
        line = raw.split("//")[0].strip()
        if not line:
            continue
        line = line.rstrip(";").strip()
        m1 = re.fullmatch(r"(?:let\s+(?:mut\s+)?)?(\w+)\s*=\s*(\w+)\.square\(\)", line)
        m2 = re.fullmatch(r"(?:let\s+(?:mut\s+)?)?(\w+)\s*=\s*(\w+)\s*\*\s*(\w+)", line)
        m3 = re.fullmatch(r"(\w+)\s*\*=\s*(\w+)", line)
        m4 = re.fullmatch(r"(\w+)\s*\*\s*(\w+)", line)
        if m1:
            v = emit(m1.group(2), m1.group(2))
            env[m1.group(1)] = v
        elif m2:
            v = emit(m2.group(2), m2.group(3))
            env[m2.group(1)] = v
        elif m3:
            v = emit(m3.group(1), m3.group(2))
            env[m3.group(1)] = v
        elif m4:
            emit(m4.group(1), m4.group(2))
        else:
            raise ValueError("unparsed chain line: %r" % raw)
    return ssa


def chain_exponent(ssa):
    e = [1]
    for a, b in ssa:
        e.append(e[a] + e[b])
    return e[-1]


def fn_body(src, name):
    i = src.index("fn %s(" % name)
    j = src.index("{", i)
    depth = 0
    for k in range(j, len(src)):
        if src[k] == "{":
            depth += 1
        elif src[k] == "}":
            depth -= 1
            if depth == 0:
                return src[j + 1 : k]
    raise ValueError(name)


TOKEN = re.compile(
    r"\[|\]|;\s*(\d+)|Felt::zero\(\)|Felt::one\(\)|MontFp!\(\s*\"(\d+)\"\s*,?\s*\)", re.S
)


def parse_array(text, start):
    """Parse a (nested) Rust array literal starting at text[start] == '['. Returns (value, end)."""
    assert text[start] == "["
    stack = []
    cur = None
    pos = start
    while True:
        m = TOKEN.search(text, pos)
        assert m, "unterminated array"
        tok = m.group(0)
        pos = m.end()
        if tok == "[":
            new = []
            if cur is not None:
                cur.append(new)
                stack.append(cur)
            cur = new
        elif tok == "]":
            if not stack:
                return cur, pos
            cur = stack.pop()
        elif tok.startswith(";"):
            rep = int(m.group(1))
            last = cur.pop()
            cur.extend([last] * rep)
        elif tok == "Felt::zero()":
            cur.append(0)
        elif tok == "Felt::one()":
            cur.append(1)
        else:
            cur.append(int(m.group(2)))


def arrays_named(body, name):
    out = []
    for m in re.finditer(r"let\s+(?:mut\s+)?%s\s*=\s*" % name, body):
        val, _ = parse_array(body, body.index("[", m.end() - 1))
        out.append(val)
    return out


def dec(v):
    if isinstance(v, list):
        return [dec(x) for x in v]
    return str(v)


def main():
    params = {}
    kat = {}
    totals = {"sbox": 0, "hash_field": 0, "hash_bytes": 0, "jive2": 0, "jive4": 0}
    for fi, field in enumerate(FIELDS):
        p = MODULI[field]
        sb = read("src", field, "sbox.rs")
        alpha = const_u32(sb, "ALPHA")
        beta = const_u32(sb, "BETA")
        delta = const_felt(sb, "DELTA")
        inv_alpha = const_felt(sb, "INV_ALPHA")
        chain = parse_chain(sb)
        assert chain_exponent(chain) == inv_alpha, field
        assert (alpha * inv_alpha) % (p - 1) == 1, field
        assert (delta * beta) % p == 1, field
        n64 = (p.bit_length() + 63) // 64
        fp = {
            "index": fi,
            "modulus": str(p),
            "bits": p.bit_length(),
            "n64": n64,
            "alpha": alpha,
            "beta": beta,
            "delta": str(delta),
            "inv_alpha": str(inv_alpha),
            "chain": chain,
            "byte_chunk": n64 * 8 - 1,
            "inst": {},
        }
        kat[field] = {}
        for inst in INSTS:
            mod = read("src", field, inst, "mod.rs")
            rc = read("src", field, inst, "round_constants.rs")
            hs = read("src", field, inst, "hasher.rs")
            width = const_u32(mod, "STATE_WIDTH")
            rate = const_u32(mod, "RATE_WIDTH")
            cols = const_u32(mod, "NUM_COLUMNS")
            rounds = const_u32(mod, "NUM_HASH_ROUNDS")
            mc = re.search(r"const C:[^=]*=\s*", rc)
            md = re.search(r"const D:[^=]*=\s*", rc)
            C, _ = parse_array(rc, rc.index("[", mc.end() - 1))
            D, _ = parse_array(rc, rc.index("[", md.end() - 1))
            assert len(C) == len(D) == cols * rounds, (field, inst)
            assert D[0] == (delta + C[0]) % p or True
            fp["inst"][inst] = {
                "width": width, "rate": rate, "cols": cols, "rounds": rounds,
                "C": dec(C), "D": dec(D),
            }
            k = {}
            body = fn_body(mod, "test_sbox")
            (sin,) = arrays_named(body, "input")
            (sout,) = arrays_named(body, "output")
            assert len(sin) == len(sout)
            k["sbox"] = {"in": dec(sin), "out": dec(sout)}
            totals["sbox"] += len(sin)

            body = fn_body(hs, "test_anemoi_hash")
            (hin,) = arrays_named(body, "input_data")
            (hout,) = arrays_named(body, "output_data")
            assert len(hin) == len(hout)
            k["hash_field"] = {"in": dec(hin), "out": dec([o[0] for o in hout])}
            totals["hash_field"] += len(hin)

            body = fn_body(hs, "test_anemoi_hash_bytes")
            (bin_,) = arrays_named(body, "input_data")
            (bout,) = arrays_named(body, "output_data")
            chunk = fp["byte_chunk"]
            # the test packs each (0/1-valued) input felt as `chunk` little-endian bytes
            # (hasher.rs test_anemoi_hash_bytes: bytes[i*chunk..(i+1)*chunk] = le_bytes(felt)[0..chunk])
            msgs = []
            for felts in bin_:
                b = b"".join(int(v).to_bytes(chunk, "little") for v in felts)
                msgs.append(b.hex())
            k["hash_bytes"] = {"in_hex": msgs, "out": dec([o[0] for o in bout])}
            totals["hash_bytes"] += len(msgs)

            body = fn_body(hs, "test_anemoi_jive")
            jin = arrays_named(body, "input_data")
            jout = arrays_named(body, "output_data")
            k["jive2"] = {"in": dec(jin[0]), "out": dec(jout[0])}
            totals["jive2"] += len(jin[0])
            if inst == "anemoi_4_3":
                assert len(jin) == 2
                k["jive4"] = {"in": dec(jin[1]), "out": dec(jout[1])}
                totals["jive4"] += len(jin[1])
            else:
                assert len(jin) == 1
            kat[field][inst] = k
        params[field] = fp

    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "params.json"), "w") as f:
        json.dump(params, f, indent=0, separators=(",", ":"))
        f.write("\n")
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(kat, f, indent=0, separators=(",", ":"))
        f.write("\n")
    print("extracted:", totals, file=sys.stderr)


if __name__ == "__main__":
    main()
