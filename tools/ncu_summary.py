#!/usr/bin/env python
"""Summarises an ncu report (read here, no GPU needed): key raw metrics per captured launch and,
with --source, the hottest SASS/source lines by stall samples.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__thread_inst_executed.sum"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for i, h in enumerate(hdr):
        if h in KEYS or (h.startswith("smsp__average_warp_latency_issue_stalled") and h.endswith(".ratio")) \
                or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")) or h == "Kernel Name":
            vals = [r[i] for r in data]
            try:
                if all(float(v.replace(",", "")) == 0 for v in vals):
                    continue
            except ValueError:
                pass
            print("%-95s %-10s %s" % (h, units[i], " | ".join(v[:40] for v in vals)))


def source(path, top):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    for r in rows:
        if "Source" in r and any("Sampl" in c for c in r):
            hdr = r
            break
    if hdr is None:
        print(out[:2000])
        return
    si = hdr.index("Source")
    cols = [i for i, c in enumerate(hdr) if "Sampl" in c]
    body = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
    key = cols[0]
    def num(x):
        try:
            return float(x.replace(",", ""))
        except ValueError:
            return 0.0
    tot = sum(num(r[key]) for r in body) or 1.0
    print("total samples (%s): %d over %d instructions" % (hdr[key], tot, len(body)))
    for r in sorted(body, key=lambda r: -num(r[key]))[:top]:
        print("%6.2f%%  %s" % (100 * num(r[key]) / tot, r[si][:110]))


if __name__ == "__main__":
    raw(sys.argv[1])
    if "--source" in sys.argv:
        source(sys.argv[1], int(sys.argv[sys.argv.index("--source") + 1]))
