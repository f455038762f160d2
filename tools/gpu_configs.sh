#!/bin/bash
# BASELINE.json configurations 1, 2, 4 on one GPU (config 3 is the default bench; 5 needs the 8-GPU box) + the reference arm of config 1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 1 2 4; do
  timeout 600 python bench.py --config $c --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_config$c.log 2> gpurun_out/bench_config$c.err || tail -5 gpurun_out/bench_config$c.err
  tail -1 gpurun_out/bench_config$c.log | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print("config %s %s: Mpaths/s %.1f  Mrays/s %.1f  ms/step %.2f  e2e %.1f  load_s %.2f cli_s %s  kernels %s  cpu %s" % (sys.argv[1], d["config"]["workload"], d["value"], d["mrays_per_s"], d["ms_per_step"], d["e2e"]["value"], d["load_s"], d["cli_s"], {k: round(v, 2) for k, v in r["kernel_ms_per_step"].items()}, d["cpu_baseline"]))' $c
done
timeout 600 python bench.py --impl reference --config 1 --steps 2 --warmup 1 | tee gpurun_out/bench_reference_config1.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_reference_config3.log | cut -c1-400
