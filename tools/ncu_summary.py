#!/usr/bin/env python3
"""Summarise an `ncu --set full` report into profiles/ and record the kernel's measured DRAM traffic for bench.py.

  python tools/ncu_summary.py <report.ncu-rep> <out.json> [--units N] [--note "..."] [--traffic-key NAME]

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU), keeps the metrics the roofline argument
rests on (pipe utilisation, issue slots, DRAM bytes, registers, stalls) for the LAST launch of each distinct kernel,
and -- with --traffic-key -- writes/updates profiles/kernel_traffic.json:

  {NAME: {"dram_bytes_per_launch": read + write, "units_per_launch": N, "kernel": "...", "source_hash": "...",
          "report": "<out.json>"}}

`source_hash` is kernel_source_hash() of the CUDA sources at the time of the capture; bench.py recomputes it and
reports `roofline.traffic` only when it still matches (a kernel change makes the number stale instead of silently
carrying it forward)."""
import argparse
import csv
import hashlib
import json
import os
import subprocess

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
KERNEL_SOURCES = ["fp.cuh", "anemoi_kernels.cuh", "field_tu.cuh", "kernel_args.h", "generated/fields.cuh"]

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__stack_size", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def kernel_source_hash():
    h = hashlib.sha256()
    base = os.path.join(ROOT, "anemoi_rust_b200", "csrc")
    for name in KERNEL_SOURCES:
        with open(os.path.join(base, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return float(value.replace(",", "")) * scale


def to_ns(value, unit):
    scale = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
    return float(value.replace(",", "")) * scale


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--units", type=int, default=0, help="units (states / messages) one launch processed")
    ap.add_argument("--note", default="")
    ap.add_argument("--traffic-key", default="")
    ap.add_argument("--kernel-filter", default="anemoi_kernel")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    short = [h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[0].isupper() else h for h in hdr]
    kernels = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if args.kernel_filter not in name:
            continue
        rec = {"Kernel Name": name, "Block Size": r[hdr.index("Block Size")], "Grid Size": r[hdr.index("Grid Size")]}
        for i, (h, s) in enumerate(zip(hdr, short)):
            for k in KEEP:
                if h == k or s == k or h.endswith("." + k):
                    v, u = r[i], units[i]
                    if v == "":
                        continue
                    if k.startswith("dram__bytes"):
                        rec[k] = to_bytes(v, u)
                    elif k == "gpu__time_duration.sum":
                        rec[k + "_ns"] = to_ns(v, u)
                    else:
                        try:
                            rec[k] = float(v.replace(",", ""))
                        except ValueError:
                            rec[k] = v
        kernels[name] = rec  # the last launch of each kernel wins (earlier ones are warm-ups)
    out = {"report": os.path.basename(args.report), "source_hash": kernel_source_hash(), "note": args.note,
           "units_per_launch": args.units, "kernels": list(kernels.values())}
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1)[:3000])
    if args.traffic_key:
        path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
        table = json.load(open(path)) if os.path.exists(path) else {}
        rec = list(kernels.values())[-1]
        table[args.traffic_key] = {
            "dram_bytes_per_launch": rec.get("dram__bytes_read.sum", 0) + rec.get("dram__bytes_write.sum", 0),
            "dram_bytes_read": rec.get("dram__bytes_read.sum"), "dram_bytes_write": rec.get("dram__bytes_write.sum"),
            "units_per_launch": args.units, "kernel": rec["Kernel Name"], "source_hash": out["source_hash"],
            "report": os.path.relpath(args.out, ROOT)}
        with open(path, "w") as f:
            json.dump(table, f, indent=1)


if __name__ == "__main__":
    main()
