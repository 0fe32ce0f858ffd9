"""Measure the other BASELINE.json configs (2-5) on one GPU, device-resident, with an oracle spot check.

  cfg2  BN-254 Anemoi-4-3 sponge hash_field, 2^18 messages x 331 felts (10 240 B at 31 B/felt, 111 permutations each)
  cfg3  Pallas / Vesta Anemoi-4-3 compress_k(4) Merkle tree over 2^26 leaves
  cfg4  BLS12-377 Fq Anemoi-2-1 Merkle tree over 2^24 leaves (alpha = 5: SURVEY D1)
  cfg5  all 7 fields x {2-1, 4-3} batched Jive compress at 2^16 .. 2^28 states, sampled oracle check

Synthetic inputs are generated on the device: random 64-bit limbs with the top limb masked to (bits - 1)
bits, i.e. uniform below 2^(bits-1) < p -- canonical residues, taken as Montgomery limbs.
One JSON line per measurement (also appended to --out). Not the contract benchmark (bench.py)."""
import argparse
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import anemoi_rust_b200 as A
from anemoi_rust_b200 import ffi, merkle
from oracle import c_oracle as C

MAC32 = {  # SURVEY.md Appendix B, squaring-aware, per permutation
    ("bls12_377", "anemoi_2_1"): 2315250, ("bls12_377", "anemoi_4_3"): 3087000,
    ("bls12_381", "anemoi_2_1"): 2346120, ("bls12_381", "anemoi_4_3"): 3128160,
    ("bn_254", "anemoi_2_1"): 728028, ("bn_254", "anemoi_4_3"): 970704,
    ("ed_on_bls12_377", "anemoi_2_1"): 657172, ("ed_on_bls12_377", "anemoi_4_3"): 899288,
    ("jubjub", "anemoi_2_1"): 727440, ("jubjub", "anemoi_4_3"): 969920,
    ("pallas", "anemoi_2_1"): 698880, ("pallas", "anemoi_4_3"): 931840,
    ("vesta", "anemoi_2_1"): 695520, ("vesta", "anemoi_4_3"): 927360,
}
PEAK = 32 * 148 * 1.965e9  # IMAD.WIDE pipe rate x SMs x max SM clock

dev = torch.device("cuda:0")


def device_random(f, n, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    x = torch.randint(-(1 << 63), (1 << 63) - 1, (n, f.n64), dtype=torch.int64, device=dev, generator=g)
    top_bits = f.p.bit_length() - 1 - 64 * (f.n64 - 1)
    x[:, f.n64 - 1] &= (1 << top_bits) - 1
    return x


def timed(fn, reps=1):
    fn()
    torch.cuda.synchronize()
    best = 1e18
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def emit(rec, out):
    line = json.dumps(rec)
    print(line, flush=True)
    if out:
        with open(out, "a") as fh:
            fh.write(line + "\n")


def ids(field, inst):
    return A.FIELD_NAMES.index(field), (0 if inst == "anemoi_2_1" else 1)


def cfg2(out, log2_msgs):
    H = A.AnemoiBn254_4_3
    f = H.FIELD
    n, L = 1 << log2_msgs, 331
    x = device_random(f, n * L, 0xA7E301 + 2)
    res = {}

    def run():
        res["d"] = H.hash_field_batch(x, felts_per_msg=L)

    ms = timed(run)
    idx = np.random.default_rng(1).choice(n, size=64, replace=False)
    xs = x.reshape(n, L, f.n64)[torch.from_numpy(idx).to(dev)].cpu().numpy().view(np.uint64)
    ok = bool(np.array_equal(res["d"].cpu().numpy().view(np.uint64)[idx], C.hash_field(2, 1, xs, 64, L)))
    perms = n * 111
    emit({"config": 2, "workload": "BN-254 Anemoi-4-3 hash_field, 2^%d messages x 331 felts (10 KB each)" % log2_msgs,
          "ms": ms, "messages_per_s": n / ms * 1e3, "permutations_per_s": perms / ms * 1e3,
          "roofline_frac": perms * MAC32[("bn_254", "anemoi_4_3")] / (ms * 1e-3) / PEAK,
          "input_GiB": n * L * 32 / 2 ** 30, "oracle_sample_ok": ok}, out)


def tree(out, cfg, field, inst, log_arity_leaves):
    H = A.HASHERS[(field, inst)]
    f, ar = H.FIELD, H.STATE_WIDTH
    fi, ii = ids(field, inst)
    n = ar ** log_arity_leaves
    leaves = device_random(f, n, 0xA7E301 + cfg)
    scratch = torch.empty((ffi.lib.anemoi_b200_merkle_scratch_felts(ar, n), f.n64), dtype=torch.int64, device=dev)
    res = {}

    def run():
        res["r"] = merkle.merkle_root_device(H, leaves, scratch=scratch)

    ms = timed(run)
    # oracle check through the decomposition property: the root of the first 2^12-ish-leaf sub-tree
    sub = ar ** (6 if ar == 4 else 11)
    sub_root = merkle.merkle_root_device(H, leaves[:sub].contiguous()).cpu().numpy().view(np.uint64)
    ok = bool(np.array_equal(sub_root, C.merkle_root(fi, ii, ar, leaves[:sub].cpu().numpy().view(np.uint64))))
    nodes = (n - 1) // (ar - 1)
    emit({"config": cfg, "workload": "%s %s arity-%d Jive Merkle root, %d^%d leaves" % (field, inst, ar, ar, log_arity_leaves),
          "ms": ms, "nodes": nodes, "nodes_per_s": nodes / ms * 1e3,
          "roofline_frac": nodes * MAC32[(field, inst)] / (ms * 1e-3) / PEAK,
          "root_limb0": int(res["r"].reshape(-1)[0].item()) & ((1 << 64) - 1), "oracle_subtree_ok": ok}, out)


def cfg5(out, sizes, only):
    for (field, inst), H in sorted(A.HASHERS.items()):
        if only and only not in field + "/" + inst:
            continue
        f, W = H.FIELD, H.STATE_WIDTH
        fi, ii = ids(field, inst)
        for lg in sizes:
            n = 1 << lg
            need = n * W * f.felt_bytes + n * f.felt_bytes
            free, _ = torch.cuda.mem_get_info()
            if need > free * 0.95:
                emit({"config": 5, "field": field, "inst": inst, "log2_states": lg, "skipped": "needs %.1f GiB" % (need / 2 ** 30)}, out)
                continue
            x = device_random(f, n * W, 0xA7E301 + 5 + lg)
            o = torch.empty((n, f.n64), dtype=torch.int64, device=dev)
            ms = timed(lambda: H.compress_k_batch(x, W, out=o), reps=1 if lg >= 24 else 3)
            m = min(n, 512)
            idx = np.sort(np.random.default_rng(lg).choice(n, size=m, replace=False))
            ti = torch.from_numpy(idx).to(dev)
            xs = x.reshape(n, W, f.n64)[ti].cpu().numpy().view(np.uint64)
            ok = bool(np.array_equal(o[ti].cpu().numpy().view(np.uint64), C.compress(fi, ii, W, xs.reshape(-1, f.n64))))
            emit({"config": 5, "field": field, "inst": inst, "k": W, "log2_states": lg, "ms": ms,
                  "compress_per_s": n / ms * 1e3, "roofline_frac": n * MAC32[(field, inst)] / (ms * 1e-3) / PEAK,
                  "oracle_sample_ok": ok}, out)
            del x, o
            torch.cuda.empty_cache()


def readme_shapes(out):
    """The four shapes the reference's README publishes single-thread CPU latencies for (README.md:77-85):
    2->1 Jive compress and `hash` of a 10 KB byte string, on BLS12-377 and Vesta, Anemoi-2-1 and 4-3."""
    published_us = {("bls12_377", "anemoi_2_1"): (429.61, 85369.0), ("bls12_377", "anemoi_4_3"): (485.99, 35937.0),
                    ("vesta", "anemoi_2_1"): (129.48, 44448.0), ("vesta", "anemoi_4_3"): (176.58, 20307.0)}
    for (field, inst), (c_us, h_us) in published_us.items():
        H = A.HASHERS[(field, inst)]
        f, W = H.FIELD, H.STATE_WIDTH
        fi, ii = ids(field, inst)
        n = 1 << 20
        x = device_random(f, n * W, 0xA7E301 + 77)
        o = torch.empty((n * (W // 2), f.n64), dtype=torch.int64, device=dev)
        ms_c = timed(lambda: H.compress_batch(x, out=o), reps=2)
        # two full waves of resident threads: 148 SMs x (512 | 768 threads) x 2 / threads-per-message
        nm, nb = 2 * 148 * (512 if f.n64 == 6 else 768) // H.NUM_COLUMNS, 10240
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        data = torch.randint(0, 256, (nm, nb), dtype=torch.uint8, device=dev, generator=g)
        dig = torch.empty((nm, f.n64), dtype=torch.int64, device=dev)

        def run_hash():
            ffi.check(ffi.lib.anemoi_b200_hash_bytes_dev(f.id, H.INST, ctypes.c_void_p(data.data_ptr()), nm, nb,
                                                        ctypes.c_void_p(dig.data_ptr()), None))

        ms_h = timed(run_hash)
        sample = data[:8].cpu().numpy()
        ok = bool(np.array_equal(dig[:8].cpu().numpy().view(np.uint64), C.hash_bytes(fi, ii, sample, 8, nb)))
        emit({"config": "readme", "field": field, "inst": inst,
              "compress_2to1_per_s": n / ms_c * 1e3, "reference_published_compress_us_1thread": c_us,
              "speedup_vs_published_1thread_compress": (n / ms_c * 1e3) * c_us * 1e-6,
              "hash_10KB_messages_per_s": nm / ms_h * 1e3, "reference_published_hash10KB_us_1thread": h_us,
              "speedup_vs_published_1thread_hash": (nm / ms_h * 1e3) * h_us * 1e-6, "oracle_sample_ok": ok}, out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="2,3,4,5")
    ap.add_argument("--sizes", default="16,20,24")
    ap.add_argument("--only", default="")
    ap.add_argument("--cfg2-log2", type=int, default=18)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    cfgs = [int(c) for c in args.configs.split(",") if c != "readme"]
    if "readme" in args.configs.split(","):
        readme_shapes(args.out)
    emit({"gpu": torch.cuda.get_device_name(0), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
          "peak_TMAC32": PEAK * 1e-12}, args.out)
    if 2 in cfgs:
        cfg2(args.out, args.cfg2_log2)
    if 3 in cfgs:
        tree(args.out, 3, "pallas", "anemoi_4_3", 13)
        tree(args.out, 3, "vesta", "anemoi_4_3", 13)
    if 4 in cfgs:
        tree(args.out, 4, "bls12_377", "anemoi_2_1", 24)
    if 5 in cfgs:
        cfg5(args.out, [int(s) for s in args.sizes.split(",")], args.only)


if __name__ == "__main__":
    main()
