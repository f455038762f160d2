#!/bin/bash
# the per-GPU share of the headline frame at 8 GPUs (16 spp = 4 Mi paths) on ONE GPU: frames in flight x wavefront lanes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for spp in 16 32; do
for fl in 1 2; do
  for lanes in 1 2 3 4; do
    RTC_STREAMS=$lanes timeout 200 python bench.py --spp $spp --frames-in-flight $fl --steps 30 --warmup 5 --no-cpu-baseline --no-peaks --no-cli > gpurun_out/small_${spp}_${fl}_${lanes}.log 2>&1
    tail -1 gpurun_out/small_${spp}_${fl}_${lanes}.log | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print("spp %s frames %s lanes %s: Mpaths/s %.1f  ms/step %.3f  e2e %.1f  kernels %s" % (sys.argv[1], sys.argv[2], sys.argv[3], d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in r["kernel_ms_per_step"].items()}))' $spp $fl $lanes
  done
done
done
