#!/usr/bin/env python
"""Summarise one .ncu-rep (first kernel) into the small csv kept under profiles/.

usage: python profiles/tools/ncu_summary.py REPORT.ncu-rep "command line that was profiled" > profiles/NAME.csv

Reads the raw page for the launch metrics and the source page (SASS) for the per-instruction
counters: total warp-level instructions, stall-reason mix and the executed-instruction share of
each barrier-delimited region of the kernel.
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
]


def page(rep, name, *extra):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep, cmd = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    w = csv.writer(sys.stdout)
    w.writerow(["metric", "value", "unit"])
    w.writerow(["command", cmd, ""])
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    w.writerow(["Kernel Name", vals[col["Kernel Name"]], ""])
    for m in WANT:
        if m in col:
            w.writerow([m, vals[col[m]], units[col[m]]])
    src = page(rep, "source", "--print-source", "sass")
    h = src[1]
    ix = {n: i for i, n in enumerate(h)}
    data = [r for r in src[2:] if len(r) > ix["Instructions Executed"]]
    inst = [int(r[ix["Instructions Executed"]] or 0) for r in data]
    samp = [int(r[ix["# Samples"]] or 0) for r in data]
    w.writerow(["sass_lines", len(data), ""])
    w.writerow(["warp_instructions_executed", sum(inst), "inst"])
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = max(sum(samp), 1)
    for n in sorted(stalls, key=lambda n: -sum(int(r[ix[n]] or 0) for r in data)):
        share = 100.0 * sum(int(r[ix[n]] or 0) for r in data) / tot
        if share >= 1.0:
            w.writerow(["stall_share." + n[6:], f"{share:.1f}", "% of samples"])
    start = 0
    for i, r in enumerate(data):
        if "BAR.SYNC" in r[ix["Source"]] or i == len(data) - 1:
            e = sum(inst[start:i + 1])
            if e > 0.002 * sum(inst):
                w.writerow([f"region.sass_{start}_{i}", f"{100.0 * e / sum(inst):.1f}",
                            "% of executed instructions (regions end at a BAR.SYNC)"])
            start = i + 1


if __name__ == "__main__":
    main()
