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
    print("%-28s %8s %10s %7s %14s" % ("kernel", "launches", "ms", "share", "DRAM MB/launch"))
    out = {}
    for k, p in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = len(ids[k])
        dram = (p.get("dram__bytes_read.sum", 0.0) + p.get("dram__bytes_write.sum", 0.0)) / n
        print("%-28s %8d %10.3f %6.1f%% %14.1f" % (k, n, p["gpu__time_duration.sum"], 100 * p["gpu__time_duration.sum"] / total, dram / 1e6))
        if k.startswith("k_"):
            out[k] = {"dram_bytes_per_launch": round(dram, -5), "launches": n}
    if "--traffic" in sys.argv:
        dst = sys.argv[sys.argv.index("--traffic") + 1]
        cmd = sys.argv[sys.argv.index("--cmd") + 1] if "--cmd" in sys.argv else ""
        json.dump({"workload": "practice5_dragon_100k", "source": "%s (%s)" % (path, cmd), "kernels": out}, open(dst, "w"), indent=1)
        print("wrote", dst)


if __name__ == "__main__":
    main()
