#!/usr/bin/env python
"""Per-instruction shared-memory wavefronts of one .ncu-rep (first kernel): which SASS instructions pay
excess (bank-conflict) wavefronts, with the CUDA source line ncu maps them to.

usage: python profiles/tools/ncu_smem_hotspots.py REPORT.ncu-rep [TOP] [KERNEL_REGEX] > profiles/NAME_smem_hotspots.csv

Needs a capture taken with `--set full --import-source on` of a library built with -lineinfo.
"""
import csv
import subprocess
import sys


def page(rep, *extra):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def num(s):
    try:
        return float(s.replace(",", "")) if s else 0.0
    except ValueError:
        return 0.0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    kern = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
    rows = page(rep, "--print-source", "sass", *kern)
    # the header row is the first one that names the SASS column
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    h = rows[hi]
    ix = {n: i for i, n in enumerate(h)}
    data = [r for r in rows[hi + 1:] if len(r) == len(h)]
    wf = next((n for n in h if n.startswith("L1 Wavefronts Shared") and "Ideal" not in n and "Excessive" not in n), None)
    ideal = next((n for n in h if n.startswith("L1 Wavefronts Shared Ideal")), None)
    exc = next((n for n in h if n.startswith("L1 Wavefronts Shared Excessive")), None)
    w = csv.writer(sys.stdout)
    if wf is None:
        w.writerow(["no shared-memory wavefront columns in this report; columns:"] + h)
        return
    tot_wf = sum(num(r[ix[wf]]) for r in data)
    tot_exc = sum(num(r[ix[exc]]) for r in data) if exc else 0.0
    tot_inst = sum(num(r[ix["Instructions Executed"]]) for r in data)
    w.writerow(["total", f"wavefronts={tot_wf:.0f}", f"excessive={tot_exc:.0f}", f"warp_instructions={tot_inst:.0f}"])
    w.writerow(["sass_line", "address", "sass", "executed", "wavefronts", "ideal", "excessive", "wavefronts_per_exec"])
    key = (lambda r: num(r[ix[exc]])) if exc else (lambda r: num(r[ix[wf]]))
    order = sorted(range(len(data)), key=lambda i: -key(data[i]))[:top]
    for i in order:
        r = data[i]
        ex = num(r[ix["Instructions Executed"]])
        w.writerow([i, r[ix.get("Address", 0)], r[ix["Source"]].strip(), f"{ex:.0f}", r[ix[wf]],
                    r[ix[ideal]] if ideal else "", r[ix[exc]] if exc else "",
                    f"{num(r[ix[wf]]) / ex:.2f}" if ex else ""])
    # all shared-memory instructions by wavefronts as a second table
    w.writerow([])
    w.writerow(["by wavefronts"])
    order = sorted(range(len(data)), key=lambda i: -num(data[i][ix[wf]]))[:top]
    for i in order:
        r = data[i]
        ex = num(r[ix["Instructions Executed"]])
        w.writerow([i, r[ix.get("Address", 0)], r[ix["Source"]].strip(), f"{ex:.0f}", r[ix[wf]],
                    r[ix[ideal]] if ideal else "", r[ix[exc]] if exc else "",
                    f"{num(r[ix[wf]]) / ex:.2f}" if ex else ""])


if __name__ == "__main__":
    main()
