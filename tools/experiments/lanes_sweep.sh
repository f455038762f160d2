#!/bin/bash
# tools/experiments/lanes_sweep.sh -- on the GPU box: wavefront lanes (RTC_STREAMS) x batch size, default build
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for lanes in 2 3 4; do
  for batch in 0 8388608 4194304; do
    RTC_STREAMS=$lanes timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch-paths $batch > gpurun_out/bench_l${lanes}_b${batch}.log 2>&1
    tail -1 gpurun_out/bench_l${lanes}_b${batch}.log | python -c '
import sys, json
d = json.loads(sys.stdin.read())
print("lanes %s batch %s  Mpaths/s %.1f  ms/step %.2f  e2e %.1f" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], d["e2e"]["value"]))' $lanes $batch
  done
done
