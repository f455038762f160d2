#!/bin/bash
# the share of one GPU of eight (16 spp) and of four (32 spp) on one GPU: one batch per render on alternating lanes vs one batch per lane
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for spp in 16 32; do
for small in 0 9000000; do
    RTC_SMALL_RENDER=$small timeout 200 python bench.py --spp $spp --frames-in-flight 2 --steps 30 --warmup 5 --no-cpu-baseline --no-peaks --no-cli > gpurun_out/smallr_${spp}_${small}.log 2>&1
    tail -1 gpurun_out/smallr_${spp}_${small}.log | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print("spp %s RTC_SMALL_RENDER=%s: Mpaths/s %.1f  ms/step %.3f  e2e %.1f  kernels %s" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in r["kernel_ms_per_step"].items()}))' $spp $small
done
done
