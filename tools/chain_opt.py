#!/usr/bin/env python3
"""Search short, FEW-SLOT addition chains for x^(1/alpha) and compile them for the GPU ladder interpreter.

The reference crate hard-codes one addition chain per field (src/<field>/sbox.rs, generated with `addchain`): a
dictionary of ~20-33 small powers followed by a long square-and-multiply run. Any chain yields the same canonical
residue, so the GPU is free to run a different one. What the GPU wants is (i) few field multiplies, weighted by their
MAC32 cost (a squaring is ~0.78 of a multiply), and (ii) FEW LIVE VALUES: every live value is a 32/48-byte slot of
per-thread local memory, and at 28 slots x 48 B x ~95 k resident threads the slot file (127 MB) no longer fits the
126 MB L2, so it is written back to HBM (19 GB per 2^20-state launch on bls12_381 in round 1).

Method: dictionary-based sliding window with a SEARCHED dictionary.
  * dictionary D = a set of <= K odd exponents (1 in D); it is built once per S-box by a short addition sequence
    (helpers such as 2, 4, 6 allowed, freed as soon as they are dead);
  * the exponent is parsed left to right into windows whose values are in D by dynamic programming (minimum number of
    multiplies; window length <= MAXLEN bits); the parse also knows PREFIX SELF-DOUBLING -- where the next s bits spell
    the whole prefix again, ST t; SQR s; MUL t covers them with one multiply (periodic exponents: Pallas, Vesta);
  * D is optimised by simulated annealing on   cost = SQR_COST * squarings + MUL_COST * multiplies  (moves draw new
    entries from the windows that actually occur in the exponent, weighted by how often they do);
  * the result is compiled to the interpreter's ISA (SQR n / MUL slot / LD slot / ST slot) with a linear-scan slot
    allocator and VERIFIED by executing it on exponents (tools/gen_params.py: check_program).

Output: tools/chains.json  {field: {"program": [[op, arg], ...], "slots": n, "sqr": s, "mul": m, "dict": [...]}}, read by
tools/gen_params.py. The search is seeded and deterministic; the JSON is committed so builds do not depend on it.
usage: python tools/chain_opt.py [--slots K] [--iters N] [field ...]"""
import argparse
import json
import math
import os
import random

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OP_SQR, OP_MUL, OP_LD, OP_ST = 0, 1, 2, 3


def mac_costs(n32):
    return n32 * (n32 + 1) // 2 + n32 * n32 + n32, 2 * n32 * n32 + n32  # squaring, multiply


_SELF_DOUBLES = {}


def self_doubles(bits):
    """{position i: [s, ...]} such that the s bits after position i spell the first i bits again (a property of the
    exponent alone, computed once)."""
    if bits not in _SELF_DOUBLES:
        n, e_val, table = len(bits), int(bits, 2), {}
        for i in range(2, n):
            A = e_val >> (n - i)
            b = A.bit_length()
            for s in range(b, min(n - i, b + 8) + 1):
                if int(bits[i:i + s], 2) == A:
                    table.setdefault(i, []).append(s)
        _SELF_DOUBLES[bits] = table
    return _SELF_DOUBLES[bits]


def parse(bits, D, maxlen):
    """Minimum-multiply left-to-right parse of the exponent into windows with values in D.
    Returns (multiplies, squarings, [(first_value), (nsq, value), ...])."""
    n = len(bits)
    INF = 1 << 30
    dp = [INF] * (n + 1)
    ch = [None] * (n + 1)
    dp[n] = 0
    for i in range(n - 1, -1, -1):
        if bits[i] == "0":
            dp[i], ch[i] = dp[i + 1], (1, 0)
            continue
        v = 0
        for l in range(1, min(maxlen, n - i) + 1):
            v = (v << 1) | (bits[i + l - 1] == "1")
            if (v & 1) and v in D and dp[i + l] + 1 < dp[i]:
                dp[i], ch[i] = dp[i + l] + 1, (l, v)
    # self-doubling: at position i the accumulator holds x^A with A = the first i bits; if the NEXT s bits spell A again
    # (leading zeros allowed), then ST t; SQR s; MUL t covers them with one multiply. Periodic exponents (Pallas / Vesta
    # start with 126 bits of 0011...) double their prefix this way: 14 -> 30 -> 62 -> 126 bits in three multiplies.
    # (A second backward pass: the transition at i needs dp[i + s], which the first pass has already settled.)
    for i in range(n - 1, 1, -1):
        for sdbl in self_doubles(bits).get(i, ()):
            if dp[i + sdbl] + 1 < dp[i]:
                dp[i], ch[i] = dp[i + sdbl] + 1, (sdbl, -1)
        # a better dp[i] can improve the zero-bit predecessors that simply step onto it
        j = i - 1
        while j >= 0 and bits[j] == "0" and dp[j + 1] < dp[j]:
            dp[j], ch[j] = dp[j + 1], (1, 0)
            j -= 1
    # windows that END where a self-double starts may now be cheaper: one more forward-independent relaxation pass
    for i in range(n - 1, -1, -1):
        if bits[i] == "0":
            if dp[i + 1] < dp[i]:
                dp[i], ch[i] = dp[i + 1], (1, 0)
            continue
        v = 0
        for l in range(1, min(maxlen, n - i) + 1):
            v = (v << 1) | (bits[i + l - 1] == "1")
            if (v & 1) and v in D and dp[i + l] + 1 < dp[i]:
                dp[i], ch[i] = dp[i + l] + 1, (l, v)
    best = None
    v = 0
    for l in range(1, min(maxlen, n) + 1):
        v = (v << 1) | (bits[l - 1] == "1")
        if (v & 1) and v in D and dp[l] < INF and (best is None or dp[l] < best[0]):
            best = (dp[l], l, v)
    if best is None:
        return None
    mults, l0, v0 = best
    ops, i, pending = [], l0, 0
    while i < n:
        l, v = ch[i]
        if v == 0:
            pending += 1
            i += 1
        elif v == -1:  # self-double: the squarings still pending belong to the value that gets stored
            if pending:
                ops.append((pending, 0))
            ops.append((l, -1))
            pending = 0
            i += l
        else:
            ops.append((pending + l, v))
            pending = 0
            i += l
    if pending:
        ops.append((pending, 0))
    return mults, n - l0, [v0] + ops


def build_sequence(D):
    """Addition sequence producing every element of D from 1: list of (result, a, b) with a, b earlier results.
    Greedy: one step when some a + b hits the target, else one helper (preferring doublings / small helpers)."""
    have = {1}
    steps = []
    for d in sorted(D):
        if d in have:
            continue
        hit = None
        for a in sorted(have, reverse=True):
            if d - a in have:
                hit = (d, a, d - a)
                break
        if hit:
            steps.append(hit)
            have.add(d)
            continue
        done = False
        for a in sorted(have, reverse=True):
            h = d - a
            if h <= 0:
                continue
            for b in sorted(have, reverse=True):
                if h - b in have and h - b > 0:
                    steps.append((h, b, h - b))
                    have.add(h)
                    steps.append((d, a, h))
                    have.add(d)
                    done = True
                    break
            if done:
                break
        if not done:  # binary method from the largest element below d (rare)
            cur = max(x for x in have if x <= d)
            rest = d - cur
            while rest:
                a = max(x for x in have if x <= rest)
                steps.append((cur + a, cur, a))
                cur += a
                have.add(cur)
                rest -= a
    return steps


def cost_of(bits, D, maxlen, cs, cm):
    p = parse(bits, D, maxlen)
    if p is None:
        return None
    mults, sq, _ = p
    steps = build_sequence(D)
    bs = sum(1 for r, a, b in steps if a == b)
    bm = len(steps) - bs
    return cs * (sq + bs) + cm * (mults + bm), sq + bs, mults + bm


def anneal(bits, K, maxlen, cs, cm, iters, seed):
    rnd = random.Random(seed)
    # candidate dictionary entries: only odd values that OCCUR as a window of the exponent can ever be used by the parse;
    # listing each as often as it occurs biases the moves towards the frequent ones (periodic exponents have few)
    cands = []
    for i in range(len(bits)):
        if bits[i] == "1":
            v = 0
            for l in range(1, min(maxlen, len(bits) - i) + 1):
                v = (v << 1) | (bits[i + l - 1] == "1")
                if (v & 1) and v > 1:
                    cands.append(v)
    D = set(range(1, 2 * min(K, 16), 2))
    cur = cost_of(bits, D, maxlen, cs, cm)
    best, bestD = cur, set(D)
    T0 = cm * 3.0
    for it in range(iters):
        T = T0 * (1.0 - it / iters) + 1e-9
        D2 = set(D)
        r = rnd.random()
        if r < 0.45 and len(D2) > 1:
            D2.discard(rnd.choice([d for d in D2 if d != 1]))
            D2.add(rnd.choice(cands))
        elif r < 0.75:
            D2.add(rnd.choice(cands))
        elif len(D2) > 2:
            D2.discard(rnd.choice([d for d in D2 if d != 1]))
        if len(D2) > K:
            continue
        c = cost_of(bits, D2, maxlen, cs, cm)
        if c is None:
            continue
        if c[0] <= cur[0] or rnd.random() < math.exp((cur[0] - c[0]) / T):
            D, cur = D2, c
            if c[0] < best[0]:
                best, bestD = c, set(D2)
    return best, bestD


def compile_program(bits, D, maxlen):
    """Dictionary build + main run -> accumulator-machine program with linear-scan slot allocation (slot 0 = x)."""
    _, _, run = parse(bits, D, maxlen)
    steps = build_sequence(D)
    # value-level SSA: list of (result, a, b) for the build, then the run uses dictionary values
    last_use = {}
    for i, (r, a, b) in enumerate(steps):
        last_use[a] = i
        last_use[b] = i
    run_uses = {}
    for j, item in enumerate(run):
        v = item if j == 0 else item[1]
        if v and v > 0:
            run_uses[v] = len(steps) + j
    for v, t in run_uses.items():
        last_use[v] = max(last_use.get(v, -1), t)
    slot_of, free, nslots, prog = {1: 0}, [], 1, []
    acc = None

    def release(t):
        for v in [v for v in slot_of if last_use.get(v, -1) <= t and v in slot_of]:
            free.append(slot_of.pop(v))

    for i, (r, a, b) in enumerate(steps):
        if acc is not None and acc == a and acc == b:
            prog.append((OP_SQR, 1))
        elif acc is not None and acc == a and b in slot_of:
            prog.append((OP_MUL, slot_of[b]))
        elif acc is not None and acc == b and a in slot_of:
            prog.append((OP_MUL, slot_of[a]))
        elif a == b:
            prog += [(OP_LD, slot_of[a]), (OP_SQR, 1)]
        else:
            prog += [(OP_LD, slot_of[a]), (OP_MUL, slot_of[b])]
        acc = r
        release(i)
        needed_later = last_use.get(r, -1) > i + 1 or (last_use.get(r, -1) == i + 1 and not (
            i + 1 < len(steps) and r in (steps[i + 1][1], steps[i + 1][2])))
        if last_use.get(r, -1) > i and (needed_later or last_use.get(r, -1) >= len(steps)):
            sl = free.pop() if free else nslots
            if sl == nslots:
                nslots += 1
            slot_of[r] = sl
            prog.append((OP_ST, sl))
    first = run[0]
    if acc != first:
        prog.append((OP_LD, slot_of[first]))
    dbl_slot = None
    for j, (nsq, v) in enumerate(run[1:], start=1):
        if v == -1:  # self-double: the accumulator times its own 2^nsq-th power
            if dbl_slot is None:  # one scratch slot serves every self-double (each value is dead after its multiply)
                release(len(steps) + j - 1)
                if free:
                    dbl_slot = free.pop()
                else:
                    dbl_slot = nslots
                    nslots += 1
            prog += [(OP_ST, dbl_slot), (OP_SQR, nsq), (OP_MUL, dbl_slot)]
            continue
        prog.append((OP_SQR, nsq))
        if v:
            prog.append((OP_MUL, slot_of[v]))
    out = []
    for op, arg in prog:
        if op == OP_SQR and out and out[-1][0] == OP_SQR and out[-1][1] + arg <= 255:
            out[-1] = (OP_SQR, out[-1][1] + arg)
        else:
            out.append((op, arg))
    return out, nslots


def run_program(prog, slots, e):
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
            T[arg] = acc
    assert acc == e, "program does not compute x^e"
    return nsq, nmul


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("fields", nargs="*")
    ap.add_argument("--slots", type=int, default=14, help="maximum dictionary size K (slots ~ K + 1)")
    ap.add_argument("--maxlen", type=int, default=9)
    ap.add_argument("--iters", type=int, default=30000)
    ap.add_argument("--seeds", type=int, default=4)
    ap.add_argument("--seed-base", type=int, default=1000)
    ap.add_argument("--out", default=os.path.join(ROOT, "tools", "chains.json"))
    args = ap.parse_args()
    params = json.load(open(os.path.join(ROOT, "tests", "golden", "params.json")))
    table = json.load(open(args.out)) if os.path.exists(args.out) else {}
    for field in args.fields or sorted(params):
        fp = params[field]
        e = int(fp["inv_alpha"])
        bits = bin(e)[2:]
        cs, cm = mac_costs(2 * fp["n64"])
        ref_s = sum(1 for a, b in fp["chain"] if a == b)
        ref_m = len(fp["chain"]) - ref_s
        ref_cost = cs * ref_s + cm * ref_m
        best = None
        for seed in range(args.seeds):
            c, D = anneal(bits, args.slots, args.maxlen, cs, cm, args.iters, args.seed_base + seed)
            prog, nslots = compile_program(bits, D, args.maxlen)
            s, m = run_program(prog, nslots, e)
            cost = cs * s + cm * m
            if best is None or (cost, nslots) < (best["cost"], best["slots"]):
                best = {"program": [list(p) for p in prog], "slots": nslots, "sqr": s, "mul": m, "cost": cost,
                        "dict": sorted(D), "reference": {"sqr": ref_s, "mul": ref_m, "cost": ref_cost},
                        "rel_cost": round(cost / ref_cost, 5), "k": args.slots, "maxlen": args.maxlen}
        print(field, "reference %dS+%dM" % (ref_s, ref_m), "-> %dS+%dM in %d slots, cost x%.4f, dict %s" % (
            best["sqr"], best["mul"], best["slots"], best["rel_cost"], best["dict"]), flush=True)
        table[field] = best
        with open(args.out, "w") as f:
            json.dump(table, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
