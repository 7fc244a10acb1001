#!/usr/bin/env python
"""Pivot an `ncu --metrics ... --csv --log-file X.csv` launch log (one row per launch and metric) into a per-kernel
summary: launches, mean duration, mean DRAM bytes, achieved HBM GB/s, tensor-pipe active %.

    python profiles/summarize_ncu.py gpurun_out/step_metrics.csv profiles/r01_step_metrics_summary.tsv \
           [profiles/r01_gemm_traffic.json]

The optional third argument receives the mean DRAM traffic per launch of the tcgen05 GEMM kernels (bench.py reports it
as roofline.traffic).  Numbers measured under ncu are never bench values: durations here are cold-cache and serialised;
they are used for shares and for bytes, not for throughput claims."""
import csv
import json
import re
import sys
from collections import OrderedDict, defaultdict


def short(name):
    m = re.search(r"(\w+_kernel|\w+Kernel\w*)", name)
    base = m.group(1) if m else name[:60]
    t = re.search(r"<([^>]*)>", name)
    return base + ("<" + t.group(1) + ">" if t and len(t.group(1)) < 30 else "")


def main(src, dst, traffic_json=None):
    rows = []
    with open(src, newline="") as f:
        lines = f.readlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rd = csv.DictReader(lines[start:])
    launches = OrderedDict()
    for r in rd:
        k = r["ID"]
        d = launches.setdefault(k, {"name": r["Kernel Name"]})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3,
                 "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r["Metric Name"]] = v * scale
    agg = defaultdict(lambda: defaultdict(float))
    for d in launches.values():
        a = agg[short(d["name"])]
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["l2"] += d.get("lts__t_bytes.sum", 0.0)
        a["tp"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                         d.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 0.0)) * d.get("gpu__time_duration.sum", 0.0)
        a["tpe"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * d.get("gpu__time_duration.sum", 0.0)
    tot = sum(a["us"] for a in agg.values()) or 1.0
    with open(dst, "w") as f:
        f.write(f"# source: {src}  ({len(launches)} launches, {tot:.0f} us under ncu: cold-cache, serialised)\n")
        f.write("kernel\tlaunches\tshare_of_time\tmean_us\tmean_dram_MB\tdram_GB/s\tL2_MB\ttensor_pipe_active_%(time-weighted)\ttensor_pipe_elapsed_%\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            mb = (a["rd"] + a["wr"]) / a["n"] / 1e6
            gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
            f.write(f"{k}\t{int(a['n'])}\t{a['us'] / tot * 100:.1f}%\t{a['us'] / a['n']:.1f}\t{mb:.1f}\t{gbs:.0f}\t"
                    f"{a['l2'] / a['n'] / 1e6:.1f}\t{a['tp'] / a['us'] if a['us'] else 0:.1f}\t{a['tpe'] / a['us'] if a['us'] else 0:.1f}\n")
    if traffic_json:
        g = [d for d in launches.values() if "gemm_tcgen05" in d["name"]]
        n = len(g) or 1
        out = {"kernel": "gemm_tcgen05_kernel (+ grouped)", "launches": len(g),
               "mean_dram_bytes_per_launch": sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in g) / n,
               "mean_us_under_ncu": sum(d.get("gpu__time_duration.sum", 0) for d in g) / n,
               "tensor_pipe_active_pct_time_weighted":
                   sum(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                             d.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", 0))
                       * d.get("gpu__time_duration.sum", 0) for d in g)
                   / (sum(d.get("gpu__time_duration.sum", 0) for d in g) or 1),
               "source": src.replace("gpurun_out/step_metrics.csv", "profiles/r02_step_metrics.csv")}
        json.dump(out, open(traffic_json, "w"), indent=1)
    print(open(dst).read())


if __name__ == "__main__":
    main(*sys.argv[1:])
