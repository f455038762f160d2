#!/bin/bash
# tools/gpu_quick.sh -- on the GPU box: parity tests (fail fast) then a short bench; prints a one-line summary
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_quick.log 2>&1
tail -1 gpurun_out/bench_quick.log | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print("Mpaths/s %.1f  Mrays/s %.1f  ms/step %.2f  e2e %.1f  kernels %s  top %s %.1f Munits/s" % (d["value"], d["mrays_per_s"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in r["kernel_ms_per_step"].items()}, r["kernel"], r["munits_per_s_in_kernel"]))'
