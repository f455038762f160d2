#!/bin/bash
# tools/experiments/run_variants.sh [bench args] -- on the GPU box: for every variants/NAME/librtc_b200.so, swap it in,
# run the golden-ray parity tests and a short bench, print one summary line per variant; restores the default build.
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
cd "$ROOT"
PKG=raytracing-course_b200
cp $PKG/librtc_b200.so /tmp/librtc_default.so
mkdir -p gpurun_out
for d in variants/*/; do
  name=$(basename "$d")
  cp "$d/librtc_b200.so" $PKG/librtc_b200.so
  ok="tests skipped"
  [ "${RUN_TESTS:-1}" = "1" ] && ok=$(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ray_intersection or primary_hits or traversals_agree or sample_exact" 2>&1 | tail -1)
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.log 2>&1
  tail -1 gpurun_out/bench_$name.log | python -c '
import sys, json
name = sys.argv[1]; ok = sys.argv[2]
try:
    d = json.loads(sys.stdin.read()); r = d["roofline"]
    print("%-14s Mpaths/s %.1f  ms/step %.2f  e2e %.1f  kernels %s | tests: %s" % (name, d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in r["kernel_ms_per_step"].items()}, ok))
except Exception as e:
    print(name, "FAILED", e, ok)' "$name" "$ok"
done
cp /tmp/librtc_default.so $PKG/librtc_b200.so
