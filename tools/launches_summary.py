#!/usr/bin/env python
"""Reads an ncu launch list (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--clock-control none --csv --log-file X.csv python bench.py ...`) and prints, per kernel: launches, total time,
share of the GPU time and DRAM bytes per launch.  With --traffic OUT.json writes the per-kernel DRAM bytes per
launch that bench.py reports as roofline.traffic.
  python tools/launches_summary.py profiles/r01_launches_final.csv [--traffic profiles/traffic.json] [--cmd "..."]"""
import csv
import json
import re
import sys
from collections import defaultdict


def short(name):
    m = re.search(r"(k_[a-z0-9_]+)", name)
    return m.group(1) if m else re.sub(r"<.*", "", name.replace("void ", ""))[:60]


def main():
    path = sys.argv[1]
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    per = defaultdict(lambda: defaultdict(float))
    ids = defaultdict(set)
    for r in rows:
        k = short(r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        name = r["Metric Name"]
        if name == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        elif name.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per[k][name] += v
        ids[k].add(r["ID"])
    total = sum(p["gpu__time_duration.sum"] for p in per.values())
    print("%-28s %8s %10s %7s %14s %12s %7s %8s %8s" % ("kernel", "launches", "ms", "share", "DRAM MB/launch", "Mwinst/launch", "lanes", "issue %", "L1 wf %"))
    out = {}
    for k, p in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = len(ids[k])
        dram = (p.get("dram__bytes_read.sum", 0.0) + p.get("dram__bytes_write.sum", 0.0)) / n
        winst = p.get("smsp__inst_executed.sum", 0.0) / n
        lanes = p.get("smsp__thread_inst_executed.sum", 0.0) / p["smsp__inst_executed.sum"] if p.get("smsp__inst_executed.sum") else 0.0
        # percentages are per launch: average them weighted by nothing better than the launch count
        issue = p.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0) / n
        l1wf = p.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 0.0) / n
        print("%-28s %8d %10.3f %6.1f%% %14.1f %12.1f %7.2f %8.1f %8.1f" % (k, n, p["gpu__time_duration.sum"], 100 * p["gpu__time_duration.sum"] / total,
                                                                      dram / 1e6, winst / 1e6, lanes, issue, l1wf))
        if k.startswith("k_"):
            out[k] = {"dram_bytes_per_launch": round(dram, -5), "launches": n, "ms_per_launch_under_ncu": p["gpu__time_duration.sum"] / n}
            if winst:
                out[k].update({"warp_inst_per_launch": round(winst, -3), "lanes_per_inst": round(lanes, 2),
                               "issue_active_pct": round(issue, 1), "l1_wavefronts_pct": round(l1wf, 1)})
    if "--traffic" in sys.argv:
        dst = sys.argv[sys.argv.index("--traffic") + 1]
        cmd = sys.argv[sys.argv.index("--cmd") + 1] if "--cmd" in sys.argv else ""
        workload = sys.argv[sys.argv.index("--workload") + 1] if "--workload" in sys.argv else "practice5_dragon_100k"
        json.dump({"workload": workload, "source": "%s (%s)" % (path, cmd), "kernels": out}, open(dst, "w"), indent=1)
        print("wrote", dst)


if __name__ == "__main__":
    main()
