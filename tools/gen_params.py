#!/usr/bin/env python3
"""Generate the constant headers of the CUDA engine and of the C oracle from tests/golden/params.json.

  anemoi_rust_b200/csrc/generated/fields.cuh   per-field structs (modulus, Montgomery constants, delta,
                                               ARK round constants in Montgomery form as u32 limbs,
                                               sliding-window schedule for x^(1/alpha))
  oracle/params_gen.h                          the same constants as u64 limbs + the reference's own
                                               addition chain (SSA byte pairs) for the C oracle

params.json itself is produced by tools/extract_fixtures.py from the reference sources
(src/<field>/sbox.rs, src/<field>/anemoi_*/round_constants.rs, src/<field>/anemoi_*/mod.rs).
Both outputs are committed; this script needs only the repo, not the reference.
"""
import json
import os

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
FIELDS = ["bls12_377", "bls12_381", "bn_254", "ed_on_bls12_377", "jubjub", "pallas", "vesta"]
INSTS = ["anemoi_2_1", "anemoi_4_3"]

# sliding-window width per field for the GPU exponentiation (odd-power table of 2^(w-1) entries, in per-thread
# local memory: entries * N * 4 bytes per thread). w = 5 minimises squarings + multiplies for every field but
# bn_254.
WINDOW = {"bls12_377": 5, "bls12_381": 5, "bn_254": 4, "ed_on_bls12_377": 5, "jubjub": 5, "pallas": 5, "vesta": 5}


# limbs of p (value 0, 1 or 2^k) diverted from IMAD.WIDE to ALU adds/shifts in the reduction rows. Measured on
# B200 (Pallas 2-1 / 4-3, ms per 2^20 states): none 94.7 / 128.8, [7] 94.7 / 125.6, [0] 90.5 / 123.3,
# [0,7] 91.8 / 122.9, [0,4,5,7] 96.7 / 130.0, [0,4,5,6,7] 99.0 / 131.9 -- diverting more than p[0] = 1 costs more
# in ALU / dispatch pressure than it saves on the FMA-heavy pipe. p[0] = 1 alone helps every field that has it.
SPECIAL_LIMBS = {
    "bls12_377": [0], "ed_on_bls12_377": [0], "jubjub": [0], "pallas": [0], "vesta": [0],
}
# Per-field code-generation switches of fp.cuh, chosen by measurement on B200 (ptxas' ALU-vs-FMA balancing differs from
# kernel to kernel, see fp.cuh): CARRY_CHAIN = the end-of-row carry fix-ups keep their (zero) carry-out alive so that they
# stay IADD3.X on the ALU pipe (helps where ptxas would otherwise emit IMAD.X on the saturated FMA-heavy pipe; costs
# instruction-level parallelism where it would not).
# (Anemoi-2-1 kernel, Anemoi-4-3 kernel); batches below a wave always run unchained (the latency form, anemoi_kernels.cuh)
# Measured (ms per 2^20 states, chained / unchained, 2-1 and 4-3): bls12_377 262.9 / 262.4, 345.9 / 353.0; bls12_381 279.9 /
# 272.3, 370.3 / 363.2; bn_254 88.5 / 87.7, 116.4 / 117.0; ed_on_bls12_377 74.1 / 76.0, 100.5 / 102.1; jubjub 85.2 / 84.8,
# 112.8 / 111.5; pallas 79.7 / 80.8, 104.8 / 106.5; vesta 79.8 / 81.0, 105.0 / 106.8.
CARRY_CHAIN = {"bls12_377": "01", "bls12_381": "00", "bn_254": "01", "ed_on_bls12_377": "11", "jubjub": "00", "pallas": "11",
               "vesta": "11"}
CARRY_CHAIN.update({k: v for k, v in (kv.split("=") for kv in os.environ.get("ANEMOI_CARRY_CHAIN", "").split(",") if kv)})
MIN_BLOCKS = {"bn_254": (8, 7)}
if os.environ.get("ANEMOI_MIN_BLOCKS_8"):
    MIN_BLOCKS = {"__all8__": int(os.environ["ANEMOI_MIN_BLOCKS_8"])}
# Montgomery quotient digit m = t0 * (-p^-1 mod 2^32) by shift-adds where the constant allows it (bls12_381: -0x30003)
QUOTIENT_SHIFT_ADD = {"bls12_381": 0}
QUOTIENT_SHIFT_ADD.update({k: int(v) for k, v in (kv.split("=") for kv in os.environ.get("ANEMOI_QUOTIENT_SHIFT_ADD", "").split(",") if kv)})
# developer override for A/B runs: ANEMOI_SPECIAL_LIMBS="pallas=0:4:5,vesta=0:4:5"
SPECIAL_LIMBS.update({k: [int(x) for x in v.split(":") if x != ""] for k, v in
                      (kv.split("=") for kv in os.environ.get("ANEMOI_SPECIAL_LIMBS", "").split(",") if kv)})


# which exponentiation program each field runs: "window" = sliding window of WINDOW[field] bits (specialised code
# path); "reference" = the reference crate's own addition chain (src/<field>/sbox.rs) on the accumulator machine.
# Measured on B200 with the slots in local memory (ms per 2^20 states, 2-1 / 4-3):
#   bls12_381  reference 277.2 / 373.3   window-5 281.9 / 372.8      bn_254  reference 92.1 / 123.2  window-4 89.4 / 122.2
#   jubjub     reference  89.0 / 118.2   window-5  88.6 / 117.8      ed_on   reference 78.9 / 107.2  window-5 79.0 / 106.5
#   pallas / vesta: reference 82.8 / 109.5, 82.4 / 109.0 (window-5 with shared-memory slots: 88.0 / 115.9)
#
# Round 2: "searched" = a chain found by tools/chain_opt.py (dictionary-based sliding window with an annealed
# dictionary, stored in tools/chains.json and re-verified here): fewer multiplies than the reference's chain AND 10-13
# live values instead of 19-28, which keeps the slot file L2-resident (no local-memory write-back to HBM).
CHAIN_SOURCE = {"bls12_377": "searched", "bls12_381": "searched", "bn_254": "searched", "ed_on_bls12_377": "searched",
                "jubjub": "searched", "pallas": "searched", "vesta": "searched"}
CHAIN_SOURCE.update({k: v for k, v in (kv.split("=") for kv in os.environ.get("ANEMOI_CHAIN_SOURCE", "").split(",") if kv)})
CHAINS_JSON = os.environ.get("ANEMOI_CHAINS_JSON", os.path.join(ROOT, "tools", "chains.json"))


def limbs(v, n, bits):
    return [(v >> (bits * i)) & ((1 << bits) - 1) for i in range(n)]


def sliding_window(e, w):
    """Left-to-right sliding window. Returns (first_idx, [(nsq, idx or -1), ...]) with table entry
    idx holding x^(2*idx+1). Evaluating: acc = T[first]; for (nsq, idx): acc = acc^(2^nsq); if idx>=0: acc *= T[idx]."""
    bits = bin(e)[2:]
    i = 0
    n = len(bits)
    ops = []  # (nsq, value)
    pending_sq = 0
    first = None
    while i < n:
        if bits[i] == "0":
            pending_sq += 1
            i += 1
            continue
        j = min(i + w, n)
        while bits[j - 1] == "0":
            j -= 1
        val = int(bits[i:j], 2)
        if first is None:
            first = (val - 1) // 2
        else:
            ops.append((pending_sq + (j - i), (val - 1) // 2))
        pending_sq = 0
        i = j
    if pending_sq:
        ops.append((pending_sq, -1))
    # split entries whose squaring count exceeds a byte (cannot happen for these exponents, but be safe)
    out = []
    for nsq, idx in ops:
        while nsq > 255:
            out.append((255, -1))
            nsq -= 255
        out.append((nsq, idx))
    return first, out


# ---- exponentiation programs for the GPU ladder interpreter (anemoi_kernels.cuh: pow_inv_alpha) -------------
# ISA: (OP_SQR, n) acc = acc^(2^n); (OP_MUL, k) acc *= T[k]; (OP_LD, k) acc = T[k]; (OP_ST, k) T[k] = acc.
# T[] are shared-memory slots; T[0] = x on entry; the result is left in acc.
OP_SQR, OP_MUL, OP_LD, OP_ST = 0, 1, 2, 3


def program_from_window(first, ops, table):
    """Sliding-window ladder: odd powers x^(2k+1) in slots 0..table-1, x^2 in slot `table` while they are built."""
    prog = [(OP_LD, 0), (OP_SQR, 1), (OP_ST, table), (OP_LD, 0)]
    for k in range(1, table):
        prog += [(OP_MUL, table), (OP_ST, k)]
    if first != table - 1:
        prog.append((OP_LD, first))
    for nsq, idx in ops:
        prog.append((OP_SQR, nsq))
        if idx >= 0:
            prog.append((OP_MUL, idx))
    return prog, table + 1


def program_from_chain(chain):
    """The reference's own addition chain (SSA pairs, value 0 = x) compiled onto an accumulator + slots by a
    linear scan: a value is stored only if it is needed later than the very next step, slots are recycled at
    the last use."""
    uses = {}
    for i, (a, b) in enumerate(chain):
        uses.setdefault(a, []).append(i)
        uses.setdefault(b, []).append(i)
    slot_of, free, nslots, acc, prog = {0: 0}, [], 1, None, []
    for i, (a, b) in enumerate(chain):
        res = i + 1
        if acc == a and acc == b:
            prog.append((OP_SQR, 1))
        elif acc == a:
            prog.append((OP_MUL, slot_of[b]))
        elif acc == b:
            prog.append((OP_MUL, slot_of[a]))
        elif a == b:
            prog += [(OP_LD, slot_of[a]), (OP_SQR, 1)]
        else:
            prog += [(OP_LD, slot_of[a]), (OP_MUL, slot_of[b])]
        acc = res
        for v in {a, b}:
            if max(uses[v]) == i and v in slot_of:
                free.append(slot_of.pop(v))
        if any(j > i + 1 for j in uses.get(res, [])):
            if free:
                sl = free.pop()
            else:
                sl = nslots
                nslots += 1
            slot_of[res] = sl
            prog.append((OP_ST, sl))
    out = []
    for op, arg in prog:
        if op == OP_SQR and out and out[-1][0] == OP_SQR and out[-1][1] + arg <= 255:
            out[-1] = (OP_SQR, out[-1][1] + arg)
        else:
            out.append((op, arg))
    return out, nslots


def check_program(prog, slots, e):
    """Run the program on exponents (T[0] = 1 = exponent of x); returns (#squarings, #multiplies)."""
    T = [None] * slots
    T[0] = 1
    acc, nsq, nmul = None, 0, 0
    for op, arg in prog:
        if op == OP_SQR:
            acc <<= arg
            nsq += arg
        elif op == OP_MUL:
            acc += T[arg]
            nmul += 1
        elif op == OP_LD:
            acc = T[arg]
        else:
            assert arg < slots
            T[arg] = acc
    assert acc == e, "program does not compute x^e"
    return nsq, nmul


def check_schedule(e, w, first, ops):
    acc = 2 * first + 1
    for nsq, idx in ops:
        acc <<= nsq
        if idx >= 0:
            assert idx < (1 << (w - 1))
            acc += 2 * idx + 1
    assert acc == e, (hex(acc), hex(e))
    nsq = sum(o[0] for o in ops)
    nmul = sum(1 for o in ops if o[1] >= 0)
    return nsq, nmul


def switch_fn(name, words):
    body = " ".join("case %d: return 0x%08xu;" % (i, w) for i, w in enumerate(words))
    return "    HD static constexpr uint32_t %s(int i) { switch (i) { %s } return 0; }\n" % (name, body)


def main():
    with open(os.path.join(ROOT, "tests", "golden", "params.json")) as f:
        params = json.load(f)

    cu = []
    cu.append("// GENERATED by tools/gen_params.py from tests/golden/params.json -- do not edit.\n")
    cu.append("// Constants of the 7 fields x 2 instantiations in Montgomery form (R = 2^(32*N)), u32 limbs, little-endian.\n")
    cu.append("#pragma once\n#include <cstdint>\n\n#ifndef HD\n#define HD __host__ __device__ __forceinline__\n#endif\n\nnamespace anemoi {\n\n// per-field constant-memory tables (specialised inside the ANEMOI_FIELD_TABLES_<field> blocks)\ntemplate <class F> struct Tables;\n\n")

    oc = []
    oc.append("/* GENERATED by tools/gen_params.py from tests/golden/params.json -- do not edit.\n")
    oc.append(" * Constants for the C oracle: u64 limbs little-endian, Montgomery form with R = 2^(64*N64). */\n")
    oc.append("#ifndef ANEMOI_ORACLE_PARAMS_GEN_H\n#define ANEMOI_ORACLE_PARAMS_GEN_H\n#include <stdint.h>\n\n")

    summary = {}
    for field in FIELDS:
        fp = params[field]
        p = int(fp["modulus"])
        n64 = fp["n64"]
        n32 = 2 * n64
        R = 1 << (64 * n64)
        n0inv32 = (-pow(p, -1, 1 << 32)) % (1 << 32)
        n0inv64 = (-pow(p, -1, 1 << 64)) % (1 << 64)
        one = R % p
        r2 = (R * R) % p
        delta_m = (int(fp["delta"]) * R) % p
        inv_alpha = int(fp["inv_alpha"])
        spare = 64 * n64 - p.bit_length()
        w = WINDOW[field]
        first, ops = sliding_window(inv_alpha, w)
        nsq, nmul = check_schedule(inv_alpha, w, first, ops)
        table = 1 << (w - 1)
        use_program = CHAIN_SOURCE.get(field, "window") in ("reference", "searched")
        if CHAIN_SOURCE.get(field, "window") == "searched":
            with open(CHAINS_JSON) as cf:
                rec = json.load(cf)[field]
            prog, slots = [tuple(x) for x in rec["program"]], rec["slots"]
            source = "searched chain (tools/chain_opt.py, dictionary %s)" % rec["dict"]
        elif use_program:
            prog, slots = program_from_chain(fp["chain"])
            source = "reference chain (src/%s/sbox.rs)" % field
        else:
            # the window ladder has its own specialised code path in the kernel (x^2 and the running odd power stay
            # in registers while the table is built); the equivalent program is only used to count operations
            prog, _ = program_from_window(first, ops, table)
            slots = table
            source = "sliding window w = %d" % w
        psq, pmul = check_program(prog, max(slots, table + 1), inv_alpha)
        summary[field] = {"program": source, "slots": slots, "pow_sqr": psq, "pow_mul": pmul, "prog_len": len(prog),
                          "ref_chain": len(fp["chain"])}

        # -p^-1 mod 2^32 is read from constant memory on the device, NOT folded as an immediate: when ptxas sees
        # the 0xffffffff of five of the fields it rewrites m = -t0 and then splits every reduction MAC into
        # IMAD.X + IMAD.HI.U32.X (2 + 6 FMA-heavy-pipe cycles) instead of one IMAD.WIDE.U32.X (4 cycles).
        cu.append("#ifdef ANEMOI_FIELD_TABLES_%s\n__constant__ uint32_t k_n0inv_%s = 0x%08xu;\n__constant__ uint32_t k_zero_%s = 0u;\n#endif\n" % (field, field, n0inv32, field))
        cu.append("struct F_%s {\n" % field)
        cu.append("    static constexpr int ID = %d;\n" % fp["index"])
        cu.append("    static constexpr int N = %d;       // 32-bit limbs\n" % n32)
        cu.append("    static constexpr int BITS = %d;\n" % p.bit_length())
        cu.append("    static constexpr int SPARE_BITS = %d;\n" % spare)
        cu.append("    static constexpr uint32_t N0INV = 0x%08xu;  // -p^-1 mod 2^32\n" % n0inv32)
        # Lazy reduction inside the exponentiation (values kept in [0, 2p + small) with no conditional subtraction per
        # multiply). Sound when 4p <= R (>= 2 spare bits: outputs stay < 2p), and also for Pallas/Vesta where
        # 4p = R + 4t with t < 2^126: a*b/R + p < 2p + (drift of < 2^127 per dependent multiply), i.e. < 2p + 2^137
        # after the <= 600 multiplies of one S-box -- far below 2^256, so nothing overflows; two conditional
        # subtractions at the end canonicalise. (jubjub: 4p = 1.81 R, not eligible.)
        lazy = spare >= 2 or (0 <= 4 * p - R < (1 << 130))
        cu.append("    static constexpr bool LAZY = %s;\n" % ("true" if lazy else "false"))
        cu.append("    static constexpr int FINAL_SUBS = %d;\n" % (1 if spare >= 2 else 2))
        cu.append("    static constexpr int ALPHA = %d;\n" % fp["alpha"])
        cu.append("    static constexpr int BETA = %d;\n" % fp["beta"])
        cu.append("    static constexpr int BYTE_CHUNK = %d;\n" % fp["byte_chunk"])
        cu.append("    static constexpr int ROUNDS_2_1 = %d;\n" % fp["inst"]["anemoi_2_1"]["rounds"])
        cu.append("    static constexpr int ROUNDS_4_3 = %d;\n" % fp["inst"]["anemoi_4_3"]["rounds"])
        # launch geometry: (threads per block, resident blocks per SM the kernel is compiled for). The ladder's
        # slots live in local memory, so only the register file bounds residency: N = 12: 128 x 5 = 20 warps/SM
        # (<= 96 registers, no spills); N = 8: 128 x 7 = 28 warps/SM (<= 72 registers, no spills). One block fewer
        # measured 0-1.5 % slower.
        blk = 128
        # resident blocks per SM the kernels are compiled for (register cap = 65536 / (128 * blocks)): 12-limb 5 (<= 96
        # registers), 8-limb 8 (64 registers, a few words of spills outside the hot loops; measured -1 % on the 4-3 kernels
        # and -0.3..-0.8 % on the 2-1 kernels against 7 blocks / 72 registers, except bn_254 4-3: +0.7 %)
        mb = MIN_BLOCKS.get(field, (8, 8) if n32 == 8 else (5, 5))
        if "__all8__" in MIN_BLOCKS and n32 == 8:
            mb = (MIN_BLOCKS["__all8__"],) * 2
        if os.environ.get("ANEMOI_MIN_BLOCKS_12") and n32 == 12:
            mb = (int(os.environ["ANEMOI_MIN_BLOCKS_12"]),) * 2
        cu.append("    static constexpr int BLOCK = %d;\n" % blk)
        cu.append("    static constexpr int MIN_BLOCKS_2_1 = %d;\n" % mb[0])
        cu.append("    static constexpr int MIN_BLOCKS_4_3 = %d;\n" % mb[1])
        cu.append("    static constexpr int MIN_BLOCKS = MIN_BLOCKS_2_1;\n")
        cc = CARRY_CHAIN.get(field, "00")
        cu.append("    // fp.cuh: carry fix-ups chained through the carry flag, per kernel instantiation (CARRY_CHAIN itself is what a\n    // bare F gets: the diagnostic layer kernel and the host emulation)\n")
        cu.append("    static constexpr bool CARRY_CHAIN_2_1 = %s;\n" % ("true" if cc[0] == "1" else "false"))
        cu.append("    static constexpr bool CARRY_CHAIN_4_3 = %s;\n" % ("true" if cc[1] == "1" else "false"))
        cu.append("    static constexpr bool CARRY_CHAIN = CARRY_CHAIN_2_1;\n")
        cu.append("    static constexpr int SLOTS = %d;     // local-memory slots of the ladder (slot 0 = x)\n" % slots)
        cu.append("    // x^(1/alpha): %s, %d squarings + %d multiplies (reference chain: %d)\n" % (source, psq, pmul, len(fp["chain"])))
        cu.append("    static constexpr bool USE_PROGRAM = %s;\n" % ("true" if use_program else "false"))
        cu.append("    static constexpr int PROG_LEN = %d;\n" % (len(prog) if use_program else 0))
        cu.append("    static constexpr int SCHED_FIRST = %d;\n" % first)
        cu.append("    static constexpr int SCHED_LEN = %d;\n" % len(ops))
        # Montgomery quotient digit m = t0 * n0inv. When n0inv = -1 it is a negation, done on the ALU pipe as
        # (opaque zero) - t0 so that ptxas neither spends an IMAD on it nor learns that m = -t0 (see above).
        if n0inv32 == 0xFFFFFFFF:
            qd = "return k_zero_%s - t0;" % field
        elif QUOTIENT_SHIFT_ADD.get(field) and (-n0inv32) % (1 << 32) == 0x30003:
            # -p^-1 = -(3 * 0x10001) mod 2^32 (bls12_381): two shift-adds and a negation on the ALU pipe instead of an IMAD
            qd = "{ uint32_t u = t0 + (t0 << 1); uint32_t v = u + (u << 16); return k_zero_%s - v; }" % field
        else:
            qd = "return t0 * k_n0inv_%s;" % field
        cu.append("    HD static uint32_t quotient_digit(uint32_t t0) {\n#if defined(__CUDA_ARCH__) && defined(ANEMOI_FIELD_TABLES_%s)\n        %s\n#else\n        return t0 * N0INV;\n#endif\n    }\n" % (field, qd))
        # a zero that ptxas cannot see through (constant-memory load): used where a literal 0 would invite ptxas to move
        # the instruction onto the saturated FMA-heavy pipe (IMAD.X Rd, RZ, RZ, Rd / IMAD.MOV Rd, RZ) -- see fp.cuh
        cu.append("    HD static uint32_t opaque_zero() {\n#if defined(__CUDA_ARCH__) && defined(ANEMOI_FIELD_TABLES_%s)\n        return k_zero_%s;\n#else\n        return 0u;\n#endif\n    }\n" % (field, field))
        cu.append(switch_fn("p", limbs(p, n32, 32)))
        sp = SPECIAL_LIMBS.get(field, [])
        cu.append("    // modulus limbs whose m*p[j] product is done with adds/shifts on the ALU pipe (see fp.cuh ModRow)\n")
        cu.append("    HD static constexpr bool special_limb(int j) { return %s; }\n" % (" || ".join("j == %d" % j for j in sp) if sp else "false"))
        cu.append(switch_fn("one", limbs(one, n32, 32)))
        cu.append(switch_fn("r2", limbs(r2, n32, 32)))
        cu.append(switch_fn("delta", limbs(delta_m, n32, 32)))
        cu.append("};\n")
        # schedule + ARK tables as plain arrays (placed in __constant__ memory by the including TU)
        cu.append("#ifdef ANEMOI_FIELD_TABLES_%s\n" % field)
        if use_program:
            cu.append("// ladder program for x^INV_ALPHA: {op, arg}; 0 = SQR n, 1 = MUL slot, 2 = LD slot, 3 = ST slot\n")
            cu.append("__constant__ uint8_t k_prog_%s[%d][2] = {%s};\n" % (
                field, len(prog), ",".join("{%d,%d}" % (a, b) for a, b in prog)))
        else:
            cu.append("// sliding-window schedule for x^INV_ALPHA: {squarings, table index or 255}\n")
            cu.append("__constant__ uint8_t k_prog_%s[%d][2] = {%s};\n" % (
                field, len(ops), ",".join("{%d,%d}" % (a, b if b >= 0 else 255) for a, b in ops)))
        for inst in INSTS:
            ip = fp["inst"][inst]
            cols, rounds = ip["cols"], ip["rounds"]
            C = [int(v) for v in ip["C"]]
            D = [int(v) for v in ip["D"]]
            words = []
            for r in range(rounds):
                for c in range(cols):
                    words += limbs((C[r * cols + c] * R) % p, n32, 32)
                    words += limbs((D[r * cols + c] * R) % p, n32, 32)
            cu.append("// ARK constants [round][col][C,D][limb], Montgomery form (src/%s/%s/round_constants.rs)\n" % (field, inst))
            cu.append("__constant__ uint32_t k_ark_%s_%s[%d] = {\n" % (field, inst[-3:], len(words)))
            for i in range(0, len(words), 8):
                cu.append("    " + ",".join("0x%08xu" % x for x in words[i:i + 8]) + ",\n")
            cu.append("};\n")
        cu.append("template <> struct Tables<F_%s> {\n" % field)
        cu.append("    static __device__ __forceinline__ const uint8_t* prog() { return &k_prog_%s[0][0]; }\n" % field)
        cu.append("    static __device__ __forceinline__ const uint32_t* ark(int cols) { return cols == 1 ? k_ark_%s_2_1 : k_ark_%s_4_3; }\n" % (field, field))
        cu.append("};\n")
        cu.append("#endif\n\n")

        # ---- oracle
        def arr64(name, v, n=n64):
            return "static const uint64_t %s[%d] = {%s};\n" % (
                name, n, ",".join("0x%016xULL" % x for x in limbs(v, n, 64)))

        oc.append("/* ---- %s ---- */\n" % field)
        oc.append(arr64("P_%s" % field, p))
        oc.append(arr64("ONE_%s" % field, one))
        oc.append(arr64("R2_%s" % field, r2))
        oc.append(arr64("DELTA_%s" % field, delta_m))
        chain = fp["chain"]
        assert len(chain) < 65536 and max(max(a, b) for a, b in chain) < 65536
        oc.append("static const uint16_t CHAIN_%s[%d][2] = {%s};\n" % (
            field, len(chain), ",".join("{%d,%d}" % (a, b) for a, b in chain)))
        for inst in INSTS:
            ip = fp["inst"][inst]
            for nm in ("C", "D"):
                vals = [(int(v) * R) % p for v in ip[nm]]
                flat = []
                for v in vals:
                    flat += limbs(v, n64, 64)
                oc.append("static const uint64_t ARK%s_%s_%s[%d] = {%s};\n" % (
                    nm, field, inst[-3:], len(flat), ",".join("0x%016xULL" % x for x in flat)))
        oc.append("\n")

    cu.append("}  // namespace anemoi\n")

    oc.append("typedef struct {\n    const char* name; int n64; int alpha; int beta; uint64_t n0inv;\n"
              "    const uint64_t *p, *one, *r2, *delta; const uint16_t (*chain)[2]; int chain_len;\n"
              "    int rounds[2]; const uint64_t* arkc[2]; const uint64_t* arkd[2];\n} anemoi_field_params;\n\n")
    oc.append("static const anemoi_field_params ANEMOI_FIELDS[7] = {\n")
    for field in FIELDS:
        fp = params[field]
        p = int(fp["modulus"])
        n0inv64 = (-pow(p, -1, 1 << 64)) % (1 << 64)
        oc.append("    {\"%s\", %d, %d, %d, 0x%016xULL, P_%s, ONE_%s, R2_%s, DELTA_%s, CHAIN_%s, %d, {%d, %d}, "
                  "{ARKC_%s_2_1, ARKC_%s_4_3}, {ARKD_%s_2_1, ARKD_%s_4_3}},\n" % (
                      field, fp["n64"], fp["alpha"], fp["beta"], n0inv64, field, field, field, field, field,
                      len(fp["chain"]), fp["inst"]["anemoi_2_1"]["rounds"], fp["inst"]["anemoi_4_3"]["rounds"],
                      field, field, field, field))
    oc.append("};\n\n#endif\n")

    os.makedirs(os.path.join(ROOT, "anemoi_rust_b200", "csrc", "generated"), exist_ok=True)
    for path, text in ((os.path.join(ROOT, "anemoi_rust_b200", "csrc", "generated", "fields.cuh"), "".join(cu)),
                       (os.path.join(ROOT, "oracle", "params_gen.h"), "".join(oc))):
        old = open(path).read() if os.path.exists(path) else None
        if old != text:  # leave the mtime alone when nothing changed (avoids needless rebuilds)
            with open(path, "w") as f:
                f.write(text)
    for k, v in summary.items():
        print(k, v)


if __name__ == "__main__":
    main()
