#!/usr/bin/env python
"""Per source line (CUDA-C view of an ncu report with -lineinfo): stall samples and executed warp instructions of one
kernel, the hottest lines first.  python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_traverse [N]"""
import csv
import io
import subprocess
import sys


def main():
    path, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur = None
    hdr = None
    lines = []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            try:
                si = hdr.index("Warp Stall Sampling (All Samples)")
                ii, ti_ = hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
                lines.append((int(r[si] or 0), int(r[ii] or 0), float(r[ti_] or 0), cur, int(r[0]), r[1].strip()[:110]))
            except ValueError:
                pass
    ts = sum(l[0] for l in lines) or 1
    ti = sum(l[1] for l in lines) or 1
    print("total samples %d, warp instructions %d" % (ts, ti))
    for s, i, thr, f, n, src in sorted(lines, reverse=True)[:top]:
        print("%5.1f%% smp %5.1f%% inst %4.0f thr  %s:%d  %s" % (100 * s / ts, 100 * i / ti, thr, f, n, src))


if __name__ == "__main__":
    main()
