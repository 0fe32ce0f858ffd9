#!/usr/bin/env python3
"""Headline benchmark (contract: python bench.py --gpus N --steps K --warmup W [--impl reference]).

Metric (BASELINE.json): Jive compressions/s on BLS12-381 Anemoi-2-1 at 1/2/4/8 B200, plus the wall time
of a 2^26-leaf arity-4 Jive Merkle root (Pallas Anemoi-4-3) as a secondary object on the same line.

A step = one pass of the hot path over one batch: Jive 2->1 compression of 2^20 random digest pairs per
GPU (BASELINE configs[0]; the batch is independent work, so N GPUs shard it with no collective and the
scaling is weak: 2^20 pairs per GPU per step). `value` is timed with CUDA events on the launching stream
with inputs resident in HBM; `e2e` is the same work through the host-pointer C-ABI call
(anemoi_b200_compress) from pinned host memory, copies inside the timed region. `roofline` reports the
fused kernel against the measured IMAD.WIDE issue peak (the path is integer-multiply bound; HBM traffic is
reported beside it to show it is negligible). `cpu_baseline` / `--impl reference` time the C restatement
of the reference's CPU algorithm (oracle/anemoi_oracle.c: the reference itself is Rust + un-vendored
arkworks and cannot be built in this image) on the box's host cores.

Secondary objects on the same JSON line (each with its own roofline): `merkle` = BASELINE configs[2] (2^26-leaf
Pallas arity-4 root, sharded over the ranks through ONE C-ABI call, anemoi_b200_merkle_root_sharded_dev, whose
ncclAllGather the library issues itself), `merkle_cfg4` = configs[3] (2^24-leaf BLS12-377 arity-2 root, same entry),
`sponge_cfg2` = configs[1] (BN-254 Anemoi-4-3 hash_field over 2^18 messages of 331 elements; N = 1 by default).
At N > 1 the run also CHECKS the multi-GPU paths: a 4^9-leaf tree built sharded must equal the single-GPU root in
every limb (`merkle.matches_single_gpu`), and rank 0 calls the single-process C-ABI forms
anemoi_b200_merkle_root(..., n_gpus = N) / anemoi_b200_compress_multi(..., N) and compares (`c_abi_multi_matches`).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2_PAIRS = 20
MAC32_PER_COMPRESS = 2_346_120       # 7 980 S x 234 + 1 596 M x 300 (reference chain, sq-aware)
HBM_BYTES_PER_COMPRESS = 144         # 96 in + 48 out
SEED = 0xA7E301
WORKLOAD = "BLS12-381 Anemoi-2-1 Jive 2->1 compress, 2^20 random digest pairs per GPU per step (BASELINE configs[0])"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-merkle", action="store_true", help="skip the 2^26-leaf Merkle-root secondary metric")
    ap.add_argument("--merkle-log4", type=int, default=13, help="Merkle leaves = 4^this (13 -> 2^26)")
    ap.add_argument("--cpu-sample-log2", type=int, default=13, help="pairs per CPU-baseline sample = 2^this")
    ap.add_argument("--configs", default="auto",
                    help="secondary BASELINE configs to measure: comma list of 2,3,4 | all | none | auto "
                         "(auto = 2,3,4 at N = 1; 3,4 at N > 1)")
    ap.add_argument("--cfg2-log2", type=int, default=18, help="config 2: messages per GPU = 2^this")
    ap.add_argument("--cfg4-log2", type=int, default=24, help="config 4: total leaves = 2^this")
    return ap.parse_args()


def load_fields_module():
    """anemoi_rust_b200/fields.py (field metadata + the synthetic-input generator) loaded BY PATH, without importing
    the package: the package import dlopens libanemoi_b200.so, which the reference arm must not map."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_anemoi_fields", os.path.join(ROOT, "anemoi_rust_b200", "fields.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def kernel_source_hash():
    """Same hash as tools/ncu_summary.py: identifies the kernel sources a profiles/kernel_traffic.json entry was taken on."""
    import hashlib

    h = hashlib.sha256()
    base = os.path.join(ROOT, "anemoi_rust_b200", "csrc")
    for name in ["fp.cuh", "anemoi_kernels.cuh", "field_tu.cuh", "kernel_args.h", "generated/fields.cuh"]:
        with open(os.path.join(base, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def measured_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture recorded in
    profiles/kernel_traffic.json (tools/ncu_summary.py) -- only if it was taken on the current kernel sources."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as f:
            rec = json.load(f)[key]
    except Exception:
        return None, "no ncu capture recorded for %s in profiles/kernel_traffic.json" % key
    if rec.get("source_hash") != kernel_source_hash():
        return None, "stale: %s was captured on kernel sources %s, this build is %s" % (rec.get("report"), rec.get("source_hash"), kernel_source_hash())
    return rec["dram_bytes_per_launch"], "%s (ncu --set full, %d units per launch, kernel sources %s)" % (
        rec.get("report"), rec.get("units_per_launch", 0), rec.get("source_hash"))


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                    power.append(float(p[3]))
                except ValueError:
                    continue
                for nm, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


def cpu_sample(log2_sample, threads=None):
    """Time the C oracle (port of the reference's CPU algorithm) on a bounded sample of the workload."""
    import numpy as np
    from oracle import c_oracle as C

    f = load_fields_module().FIELDS["bls12_381"]
    n = 1 << log2_sample
    x = f.random_mont(2 * n, SEED)
    cores = threads or (os.cpu_count() or 1)
    C.set_threads(cores)
    C.compress(1, 0, 2, x[: 2 * 64])  # warm (thread pool, page-in)
    t0 = time.perf_counter()
    out = C.compress(1, 0, 2, x)
    dt = time.perf_counter() - t0
    return n / dt, cores, n, out, x


def run_reference(args, rank):
    """--impl reference: the reference's CPU path for the same metric/config, all host threads, rank 0 only."""
    if rank != 0:
        return
    n = 1 << args.cpu_sample_log2
    for _ in range(args.warmup):
        cpu_sample(min(args.cpu_sample_log2, 9))
    t = 0.0
    cores = os.cpu_count() or 1
    for _ in range(args.steps):
        rate, cores, n, _, _ = cpu_sample(args.cpu_sample_log2)
        t += n / rate
    value = args.steps * n / t
    sample = "2^%d pairs of the same workload per step (seed 0x%X), OpenMP over all host threads" % (args.cpu_sample_log2, SEED)
    line = {
        "impl": "reference", "metric": "jive_compressions_per_s", "value": value, "unit": "compressions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": n, "note": "CPU restatement of the reference algorithm (C, 64-bit CIOS Montgomery, "
                   "the reference's addition chains); the Rust reference cannot be built here (no cargo, arkworks not vendored)"},
        "cpu_baseline": {"value": value, "unit": "compressions/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "compressions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun exactly the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    # stdout carries exactly ONE JSON line: native libraries (NCCL's version banner, for one) write to file descriptor 1
    # behind Python's back, so fd 1 is pointed at stderr for the duration of the run and the line goes to the saved fd
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist

    import anemoi_rust_b200 as A
    from anemoi_rust_b200 import ffi, merkle

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL_DEBUG=VERSION (set in some images) prints a banner there
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    H = A.AnemoiBls12_381_2_1
    f = H.FIELD
    n = 1 << LOG2_PAIRS
    # synthetic inputs: uniform canonical residues taken as Montgomery limbs; two input sets rotate so that
    # consecutive steps never re-read what the previous one left in L2 (2 x 96 MiB in + 48 MiB out > 126 MB L2)
    host = [f.random_mont(2 * n, SEED + 1000 * rank + s) for s in range(2)]
    pinned_in = [torch.from_numpy(h.view(np.int64)).pin_memory() for h in host]
    pinned_out = torch.empty((n, f.n64), dtype=torch.int64).pin_memory()
    d_in = [p.to(dev, non_blocking=True) for p in pinned_in]
    d_out = torch.empty((n, f.n64), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step(i):
        H.compress_batch(d_in[i & 1], out=d_out)   # one fused kernel launch via anemoi_b200_compress_dev

    for i in range(args.warmup):
        step(i)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    barrier()
    ms = max_over_ranks(ms)
    value = world * args.steps * n / (ms * 1e-3)
    ms_per_step = ms / args.steps

    # ---- e2e: host-pointer C-ABI call from pinned host memory, H2D + kernel + D2H inside the timed region
    in_np = [p.numpy().view(np.uint64) for p in pinned_in]
    out_np = pinned_out.numpy().view(np.uint64)
    e2e_steps = max(2, min(args.steps, 5))
    for i in range(2):
        H.device = local_rank
        H.compress_batch(in_np[i & 1], out=out_np)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        H.compress_batch(in_np[i & 1], out=out_np)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * e2e_steps * n / e2e_s
    H.compress_batch(d_in[(e2e_steps - 1) & 1], out=d_out)
    torch.cuda.synchronize()
    same = bool(np.array_equal(out_np, d_out.cpu().numpy().view(np.uint64)))

    # ---- roofline of the dominant (only) kernel: integer-multiply issue rate
    micro_ops, peak_mhz = ffi.imad_peak(2)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    # IMAD.WIDE.U32 holds the FMA-heavy pipe 4 cycles per warp instruction (ncu: sm__pipe_fmaheavy_cycles_active
    # = 94 % while this kernel retires 8.78e12 IMAD.WIDE/s, profiles/r1_ncu_full_*.json) => 32 MAC32/clk/SM.
    # The in-run microbenchmark reaches ~29/clk/SM; the stricter pipe rate is used as the denominator.
    pipe_peak = 32.0 * sms * peak_mhz * 1e6
    peak_ops = max(pipe_peak, micro_ops)
    peak_source = ("IMAD.WIDE.U32 pipe rate 32 MAC32/clk/SM x %d SMs x %.0f MHz (SM clock measured in-run by "
                   "anemoi_b200_imad_peak; pipe rate from ncu fmaheavy utilisation)" % (sms, peak_mhz))
    peaks, peaks_src = measured_peaks()

    def imad_roofline(kernel, units, mac32_per_unit, bytes_per_unit, seconds, traffic_key, launches=1):
        """Roofline object of one kernel: ALGORITHMIC MAC32 (squaring-aware, reference chain) of the
        units processed / the device time they took, against the IMAD.WIDE pipe peak; HBM beside it."""
        achieved = units * mac32_per_unit / seconds
        traffic, traffic_src = measured_traffic(traffic_key) if traffic_key else (None, "not captured")
        gbs = units * bytes_per_unit / seconds * 1e-9
        return {"bound": "imad", "achieved": achieved * 1e-12, "peak": peak_ops * 1e-12, "unit": "TMAC32/s",
                "frac": achieved / peak_ops, "traffic": traffic, "traffic_source": traffic_src,
                "kernel": kernel, "kernel_ms": seconds * 1e3 / launches, "launches": launches,
                "algorithmic_mac32_per_unit": mac32_per_unit, "algorithmic_bytes": units * bytes_per_unit,
                "peak_source": peak_source,
                "hbm": {"achieved_gbs": gbs, "peak_gbs": peaks.get("hbm_gbs"), "frac": gbs / peaks.get("hbm_gbs", 6650.0),
                        "peak_source": peaks_src + " (MEASURED_PEAKS.json)"}}

    kernel_s = ms_per_step * 1e-3                      # one launch per step: CUDA-event average over the timed region
    roofline = imad_roofline("anemoi_kernel<F_bls12_381,1>", n, MAC32_PER_COMPRESS, HBM_BYTES_PER_COMPRESS, kernel_s,
                             "bls12_381_2_1_compress_2^20")
    roofline["peak_microbenchmark"] = micro_ops * 1e-12
    roofline["frac_of_microbenchmark"] = roofline["achieved"] / (micro_ops * 1e-12)

    line = {
        "metric": "jive_compressions_per_s", "value": value, "unit": "compressions/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": n, "field": "bls12_381", "instantiation": "anemoi_2_1",
                   "parallelism": "batch sharded across GPUs, no collective",
                   "l2": "two input sets rotate between steps; in+out per step 144 MiB > 126 MB L2",
                   "seed": SEED},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "compressions/s", "h2d_bytes_per_step": 2 * n * f.felt_bytes,
                "d2h_bytes_per_step": n * f.felt_bytes, "steps": e2e_steps, "api": "anemoi_b200_compress (host pointers, pinned)",
                "matches_device_path": same},
        "gpu_launches": args.steps,
        "roofline": roofline,
    }

    if args.configs == "auto":
        cfgs = {2, 3, 4} if world == 1 else {3, 4}
    elif args.configs == "all":
        cfgs = {2, 3, 4}
    elif args.configs == "none":
        cfgs = set()
    else:
        cfgs = {int(c) for c in args.configs.split(",") if c}
    if args.no_merkle:
        cfgs.discard(3)

    def hexlimbs(t):
        return "".join("%016x" % (int(v) & ((1 << 64) - 1)) for v in reversed(t.reshape(-1).tolist()))

    def device_leaves(fld, total, seed):
        """This rank's contiguous slice of a `total`-leaf array made of 8 fixed, individually seeded chunks: rank r of N
        takes chunks [8r/N, 8(r+1)/N), so the tree -- and its root -- is the same for N = 1, 2, 4, 8. Uniform below
        2^(bits-1) < p (canonical residues), taken as Montgomery limbs."""
        chunks = 8 if (world in (1, 2, 4, 8) and total % 8 == 0) else world
        per_chunk = total // chunks
        top_bits = fld.p.bit_length() - 1 - 64 * (fld.n64 - 1)
        parts = []
        for c in range(rank * chunks // world, (rank + 1) * chunks // world):
            g = torch.Generator(device=dev)
            g.manual_seed(seed + c)
            part = torch.randint(-(1 << 63), (1 << 63) - 1, (per_chunk, fld.n64), dtype=torch.int64, device=dev, generator=g)
            part[:, fld.n64 - 1] &= (1 << top_bits) - 1
            parts.append(part)
        return torch.cat(parts) if len(parts) > 1 else parts[0]

    def sharded_merkle(Hm, total, seed, mac32_per_node, key, kernel):
        """Root of a `total`-leaf tree sharded over the ranks: ONE C-ABI call per rank
        (anemoi_b200_merkle_root_sharded_dev: sub-tree, the library's own ncclAllGather of the partial roots, top levels)."""
        fm, ar = Hm.FIELD, Hm.STATE_WIDTH
        leaves = device_leaves(fm, total, seed)
        local = total // world
        scratch = torch.empty((ffi.lib.anemoi_b200_merkle_sharded_scratch_felts(ar, local, world), fm.n64),
                              dtype=torch.int64, device=dev)
        small = leaves[: ar ** 4 * (2 if (ar == 4 and world in (2, 8)) else 1)].contiguous()
        merkle.merkle_root_distributed(Hm, small)          # warm-up: communicator, module load
        times = []
        for _ in range(2):
            barrier()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record(stream)
            root = merkle.merkle_root_distributed(Hm, leaves, scratch=scratch)
            m1.record(stream)
            torch.cuda.synchronize()
            times.append(max_over_ranks(m0.elapsed_time(m1)))
        mms = min(times)
        nodes = (total - 1) // (ar - 1)
        _, local_levels, roots_per_rank, top_levels = merkle.plan(total, ar, world)
        # kernel launches per rank in one build: one per level
        obj = {"metric": "merkle_root_ms", "value": mms, "unit": "ms", "higher_is_better": False, "runs_ms": times,
               "nodes": nodes, "nodes_per_s": nodes / (mms * 1e-3), "leaves": total,
               "api": "anemoi_b200_merkle_root_sharded_dev (device-resident leaves; NCCL all-gather issued by the library)",
               "partial_roots_per_rank": roots_per_rank, "local_levels": local_levels, "top_levels": top_levels,
               "nccl_version": ffi.lib.anemoi_b200_nccl_version() if world > 1 else None,
               "root": hexlimbs(root), "root_limb0": int(root.reshape(-1)[0].item()) & ((1 << 64) - 1),
               # whole-job roofline: all nodes of the tree / wall time x N GPUs' worth of peak
               "roofline": imad_roofline(kernel, nodes, mac32_per_node, (ar + 1) * fm.felt_bytes, mms * 1e-3 * world, key,
                                         launches=local_levels + top_levels)}
        obj["roofline"]["note"] = "all %d levels of the build (whole job; peak x %d GPU(s)); sub-wave upper levels included" % (local_levels + top_levels, world)
        del leaves, scratch
        torch.cuda.empty_cache()
        return obj

    # ---- configs[2]: 2^26-leaf arity-4 Jive Merkle root on Pallas Anemoi-4-3, sharded over the ranks
    if 3 in cfgs:
        H4 = A.AnemoiPallas_4_3
        total = 4 ** args.merkle_log4
        line["merkle"] = sharded_merkle(H4, total, SEED + 3, 931_840, "pallas_4_3_compress4_2^20", "anemoi_kernel<F_pallas,2>")
        line["merkle"]["workload"] = ("Pallas Anemoi-4-3 compress_k(4) tree, 4^%d = 2^%d leaves, sharded over %d GPU(s), one NCCL "
                                      "all-gather of partial roots (BASELINE configs[2])" % (args.merkle_log4, 2 * args.merkle_log4, world))
        if world == 1:
            # e2e: the same tree through the host-pointer C-ABI call, leaves in pinned host memory (H2D inside the call)
            host_leaves = torch.empty((total, H4.FIELD.n64), dtype=torch.int64).pin_memory()
            host_leaves.copy_(device_leaves(H4.FIELD, total, SEED + 3))
            torch.cuda.synchronize()
            hl = host_leaves.numpy().view(np.uint64)
            H4.device = local_rank
            H4.merkle_root(hl[: 4 ** 6])     # warm the stream path
            e2e_runs = []
            for _ in range(2):               # the first run also grows the library's memory pool to 3 GiB; the second reuses it
                t0 = time.perf_counter()
                r_host = H4.merkle_root(hl)
                e2e_runs.append((time.perf_counter() - t0) * 1e3)
            e2e_ms = min(e2e_runs)
            line["merkle"]["e2e"] = {"value": e2e_ms, "unit": "ms", "runs_ms": e2e_runs, "h2d_bytes": total * H4.FIELD.felt_bytes,
                                     "d2h_bytes": H4.FIELD.felt_bytes, "api": "anemoi_b200_merkle_root (host pointers, pinned)",
                                     "matches_device_path": "".join("%016x" % int(v) for v in reversed(r_host.reshape(-1).tolist())) == line["merkle"]["root"]}
            del host_leaves, hl
    # ---- configs[3]: 2^24-leaf arity-2 Jive Merkle root on BLS12-377 Anemoi-2-1
    if 4 in cfgs:
        H2 = A.AnemoiBls12_377_2_1
        line["merkle_cfg4"] = sharded_merkle(H2, 1 << args.cfg4_log2, SEED + 4, 2_315_250, "bls12_377_2_1_compress_2^20",
                                             "anemoi_kernel<F_bls12_377,1>")
        line["merkle_cfg4"]["workload"] = ("BLS12-377 Fq Anemoi-2-1 compress tree, 2^%d leaves, sharded over %d GPU(s) "
                                           "(BASELINE configs[3])" % (args.cfg4_log2, world))
    # ---- configs[1]: BN-254 Anemoi-4-3 sponge hash_field over 2^18 messages of 331 elements (10 KB each) per GPU
    if 2 in cfgs:
        Hs = A.AnemoiBn254_4_3
        fs = Hs.FIELD
        nm, L = 1 << args.cfg2_log2, 331
        g = torch.Generator(device=dev)
        g.manual_seed(SEED + 2 + 1000 * rank)
        msgs = torch.randint(-(1 << 63), (1 << 63) - 1, (nm * L, fs.n64), dtype=torch.int64, device=dev, generator=g)
        msgs[:, fs.n64 - 1] &= (1 << 61) - 1      # < 2^253 < p: canonical
        Hs.hash_field_batch(msgs[: 4096 * L], felts_per_msg=L)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        dig = Hs.hash_field_batch(msgs, felts_per_msg=L)
        s1.record(stream)
        torch.cuda.synchronize()
        sms_ = max_over_ranks(s0.elapsed_time(s1))
        perms = 111 * nm                              # 331 = 3 * 110 + 1 -> 110 full blocks + 1 padded block
        line["sponge_cfg2"] = {"metric": "messages_per_s", "value": world * nm / (sms_ * 1e-3), "unit": "messages/s",
                               "higher_is_better": True, "ms": sms_, "messages_per_gpu": nm, "felts_per_message": L,
                               "permutations_per_message": 111, "input_GiB_per_gpu": nm * L * 32 / 2 ** 30,
                               "workload": "BN-254 Anemoi-4-3 hash_field, 2^%d messages x 331 elements (10 KB) per GPU "
                                           "(BASELINE configs[1])" % args.cfg2_log2,
                               "roofline": imad_roofline("anemoi_kernel<F_bn_254,2>", perms, 970_704, (L + 1) * 32 / 111.0,
                                                         sms_ * 1e-3, "bn_254_4_3_hash_field_331")}
        if rank == 0 and world == 1:
            from oracle import c_oracle as C

            idx = np.sort(np.random.default_rng(1).choice(nm, size=32, replace=False))
            xs = msgs.reshape(nm, L, fs.n64)[torch.from_numpy(idx).to(dev)].cpu().numpy().view(np.uint64)
            line["sponge_cfg2"]["oracle_sample_ok"] = bool(np.array_equal(dig.cpu().numpy().view(np.uint64)[idx], C.hash_field(2, 1, xs, 32, L)))
        del msgs, dig
        torch.cuda.empty_cache()

    # ---- N > 1: correctness of the multi-GPU paths, every limb compared
    if world > 1:
        Hc = A.AnemoiPallas_4_3
        fc = Hc.FIELD
        total_c = 4 ** 9
        leaves_c = fc.random_mont(total_c, SEED + 9)                       # same array on every rank
        sl = total_c // world
        mine = torch.from_numpy(leaves_c[rank * sl:(rank + 1) * sl].view(np.int64).copy()).to(dev)
        sharded = merkle.merkle_root_distributed(Hc, mine)
        torch.cuda.synchronize()
        single = merkle.merkle_root_device(Hc, torch.from_numpy(leaves_c.view(np.int64)).to(dev))   # whole tree on this GPU
        torch.cuda.synchronize()
        ok = torch.tensor([1.0 if torch.equal(sharded.reshape(-1), single.reshape(-1)) else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        check = {"matches_single_gpu": bool(ok.item() == 1.0), "check_leaves": total_c, "check_root": hexlimbs(single)}
        if "merkle" in line:
            line["merkle"].update(check)
        else:
            line["merkle_check"] = check
        barrier()
        # While rank 0 drives all N GPUs from its own process, the other ranks must leave their GPUs idle: they wait on a
        # key of the rendezvous store (a CPU-side wait), not in an NCCL barrier whose kernel would spin on their device.
        try:
            store = dist.distributed_c10d._get_default_store()
        except Exception:  # private API: without it the ranks simply meet at the NCCL barrier below
            store = None
        if rank == 0:
            # the single-process C-ABI forms a Rust host would call: one host thread + stream per device inside the call,
            # NCCL all-gather between them (ncclCommInitAll)
            try:
                Hc.device = 0
                r_multi = Hc.merkle_root(leaves_c, n_gpus=world)
                ok_root = bool(np.array_equal(r_multi.reshape(-1), single.cpu().numpy().view(np.uint64).reshape(-1)))
                xc = fc.random_mont(4 * 30001, SEED + 10)                   # odd count: uneven slices
                one = Hc.compress_k_batch(xc, 4)
                ok_comp = bool(np.array_equal(Hc.compress_k_batch(xc, 4, n_gpus=world), one))
                line["c_abi_multi_matches"] = ok_root and ok_comp
                line["c_abi_multi"] = {"merkle_root_n_gpus_matches": ok_root, "compress_multi_matches": ok_comp, "n_gpus": world}
            except Exception as exc:  # report, do not lose the line
                line["c_abi_multi_matches"] = False
                line["c_abi_multi"] = {"error": str(exc)[:300]}
            if store is not None:
                store.set("anemoi_c_abi_multi_done", "1")
        elif store is not None:
            store.wait(["anemoi_c_abi_multi_done"])
        barrier()

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload
    if rank == 0 and world == 1:
        rate, cores, ns, cpu_out, cpu_in = cpu_sample(args.cpu_sample_log2)
        t = torch.from_numpy(cpu_in.view(np.int64)).to(dev)
        gpu_out = H.compress_batch(t).cpu().numpy().view(np.uint64)
        line["cpu_baseline"] = {"value": rate, "unit": "compressions/s", "cores": cores, "kind": "port",
                                "sample": "2^%d pairs of the same workload (seed 0x%X), OpenMP over all host threads" % (args.cpu_sample_log2, SEED),
                                "gpu_matches_cpu_on_sample": bool(np.array_equal(gpu_out, cpu_out))}
    if rank == 0:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
