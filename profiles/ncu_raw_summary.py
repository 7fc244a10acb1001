#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` (one row per captured launch, ~2000 metric columns) into the handful of
numbers the kernel notes in DESIGN.md cite: duration, grid, registers, tensor-pipe active %, issue-slot %, DRAM bytes and
the top warp-stall reasons (pc-sampling counts).

    ncu -i gpurun_out/prof_K.ncu-rep --page raw --csv > raw.csv ; python profiles/ncu_raw_summary.py raw.csv > profiles/r02_K_ncu.txt
"""
import csv
import sys


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    keys = [("us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("regs", "launch__registers_per_thread"),
            ("tensor_pipe_active_%", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
            ("issue_active_%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            ("warps_active_%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
            ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum")]
    print(f"# {path}: {len(data)} launches captured with ncu --set full --clock-control none (durations are cold-cache, serialised)")
    print("# launch\tkernel\t" + "\t".join(k for k, _ in keys) + "\ttop warp-stall reasons (pc samples)")
    for i, r in enumerate(data):
        name = r[col["Kernel Name"]].replace("void unnamed>::", "")[:46]
        vals = [(r[col[m]][:9] + (" " + units[col[m]] if units[col[m]] not in ("", "%") and "byte" in units[col[m]] else ""))
                if m in col else "-" for _, m in keys]
        st = [(hdr[j].replace("smsp__pcsamp_warps_issue_stalled_", ""), f(r[j])) for j in range(len(hdr))
              if "pcsamp_warps_issue_stalled" in hdr[j] and "not_issued" not in hdr[j] and f(r[j]) > 0]
        st.sort(key=lambda t: -t[1])
        print(f"{i}\t{name}\t" + "\t".join(vals) + "\t" + ", ".join(f"{k} {int(v)}" for k, v in st[:6]))


if __name__ == "__main__":
    main(sys.argv[1])
