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
    return ap.parse_args()


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
    from anemoi_rust_b200.fields import FIELDS

    f = FIELDS["bls12_381"]
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
        "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference algorithm (C, 64-bit CIOS Montgomery, "
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
    kernel_s = ms_per_step * 1e-3                      # one launch per step: CUDA-event average over the timed region
    achieved = n * MAC32_PER_COMPRESS / kernel_s       # algorithmic MAC32 per launch / launch duration
    peaks, peaks_src = measured_peaks()
    hbm_gbs = n * HBM_BYTES_PER_COMPRESS / kernel_s * 1e-9
    roofline = {
        "bound": "imad", "achieved": achieved * 1e-12, "peak": peak_ops * 1e-12, "unit": "TMAC32/s",
        "frac": achieved / peak_ops,
        # dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full (profiles/r1_ncu_full_bls12_381_*):
        # 3.66 GB read + 15.62 GB written. Far above the 0.15 GB of algorithmic bytes ON PURPOSE: the ladder's slot
        # table lives in per-thread local memory whose write-back reaches HBM (~70 GB/s, 1 % of the HBM peak, no
        # time cost); the zero-traffic alternative (table in shared memory) measured 4 % slower.
        "traffic": 19283889000,
        "kernel": "anemoi_kernel<F_bls12_381,1>", "kernel_ms": kernel_s * 1e3,
        "algorithmic_mac32_per_compress": MAC32_PER_COMPRESS,
        "algorithmic_bytes_per_launch": n * HBM_BYTES_PER_COMPRESS,
        "peak_source": "IMAD.WIDE.U32 pipe rate 32 MAC32/clk/SM x %d SMs x %.0f MHz (SM clock measured in-run by "
                       "anemoi_b200_imad_peak; pipe rate from ncu fmaheavy utilisation)" % (sms, peak_mhz),
        "peak_microbenchmark": micro_ops * 1e-12,
        "frac_of_microbenchmark": achieved / micro_ops,
        "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": peaks.get("hbm_gbs"), "frac": hbm_gbs / peaks.get("hbm_gbs", 6650.0),
                "peak_source": peaks_src + " (MEASURED_PEAKS.json)"},
    }

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

    # ---- secondary metric: 2^26-leaf arity-4 Jive Merkle root on Pallas Anemoi-4-3, sharded over the ranks
    if not args.no_merkle:
        H4 = A.AnemoiPallas_4_3
        f4 = H4.FIELD
        total = 4 ** args.merkle_log4
        local = total // world
        # the leaf array is 8 fixed, individually seeded chunks; rank r of N takes chunks [8r/N, 8(r+1)/N), so the
        # tree -- and therefore root_limb0 -- is the same for N = 1, 2, 4, 8 (end-to-end check of the sharded build)
        chunks = 8 if (world in (1, 2, 4, 8) and total % 8 == 0) else world
        per_chunk = total // chunks
        parts = []
        for c in range(rank * chunks // world, (rank + 1) * chunks // world):
            g = torch.Generator(device=dev)
            g.manual_seed(SEED + 3 + c)
            # p = 2^254 + t: every value below 2^254 is canonical, so mask the top limb to 62 bits
            part = torch.randint(-(1 << 63), (1 << 63) - 1, (per_chunk, f4.n64), dtype=torch.int64, device=dev, generator=g)
            part[:, f4.n64 - 1] &= (1 << 62) - 1
            parts.append(part)
        leaves = torch.cat(parts) if len(parts) > 1 else parts[0]
        del parts
        scratch = torch.empty((ffi.lib.anemoi_b200_merkle_scratch_felts(4, local), f4.n64), dtype=torch.int64, device=dev)
        root = merkle.merkle_root_distributed(H4, leaves, scratch=scratch)  # warm-up (also NCCL)
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record(stream)
        root = merkle.merkle_root_distributed(H4, leaves, scratch=scratch)
        m1.record(stream)
        torch.cuda.synchronize()
        mms = max_over_ranks(m0.elapsed_time(m1))
        nodes = (total - 1) // 3
        line["merkle"] = {"metric": "merkle_root_ms", "value": mms, "unit": "ms", "higher_is_better": False,
                          "workload": "Pallas Anemoi-4-3 compress_k(4) tree, 4^%d = 2^%d leaves, sharded over %d GPU(s), "
                                      "one NCCL all-gather of partial roots" % (args.merkle_log4, 2 * args.merkle_log4, world),
                          "nodes": nodes, "nodes_per_s": nodes / (mms * 1e-3),
                          "root_limb0": int(root.reshape(-1)[0].item()) & ((1 << 64) - 1)}

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload
    if rank == 0 and world == 1:
        rate, cores, ns, cpu_out, cpu_in = cpu_sample(args.cpu_sample_log2)
        t = torch.from_numpy(cpu_in.view(np.int64)).to(dev)
        gpu_out = H.compress_batch(t).cpu().numpy().view(np.uint64)
        line["cpu_baseline"] = {"value": rate, "unit": "compressions/s", "cores": cores, "kind": "port",
                                "sample": "2^%d pairs of the same workload (seed 0x%X), OpenMP over all host threads" % (args.cpu_sample_log2, SEED),
                                "gpu_matches_cpu_on_sample": bool(np.array_equal(gpu_out, cpu_out))}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
