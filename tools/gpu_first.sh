#!/bin/bash
# a quick GPU-box visit: parity suite (fail fast), lanes probe, short bench of the default build and of every build in variants/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PKG=raytracing-course_b200
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 120 python tools/experiments/lanes_probe.py | tee gpurun_out/lanes.json
summ() { tail -1 "$1" | python -c '
import sys, json
name = sys.argv[1]
try:
    d = json.loads(sys.stdin.read()); r = d["roofline"]
    print("%-10s Mpaths/s %.1f  ms/step %.2f  e2e %.1f  kernels %s fallback %s" % (name, d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in r["kernel_ms_per_step"].items()}, d["fallback_rays"]))
except Exception as e:
    print(name, "FAILED", e)' "$2"; }
timeout 200 python bench.py --steps ${STEPS:-8} --warmup 3 --no-cpu-baseline > gpurun_out/bench_first.log 2>&1; summ gpurun_out/bench_first.log base
if ls variants/*/librtc_b200.so > /dev/null 2>&1; then
  cp $PKG/librtc_b200.so /tmp/librtc_default.so
  for d in variants/*/; do
    name=$(basename "$d")
    cp "$d/librtc_b200.so" $PKG/librtc_b200.so
    timeout 200 python bench.py --steps ${STEPS:-8} --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.log 2>&1; summ gpurun_out/bench_$name.log $name
  done
  cp /tmp/librtc_default.so $PKG/librtc_b200.so
fi
