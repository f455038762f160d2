#!/bin/bash
# tools/gpu_call.sh -- one GPU-box visit of round 2: parity suite, peaks, the default build's bench, every build in
# variants/ (A/B on the same box), then optional extras given as arguments ("persist", "ncu").
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PKG=raytracing-course_b200
summ() { tail -1 "$1" | python -c '
import sys, json
name = sys.argv[1]
try:
    d = json.loads(sys.stdin.read()); r = d["roofline"]
    print("%-12s Mpaths/s %.1f  ms/step %.2f  e2e %.1f  kernels %s" % (name, d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in r["kernel_ms_per_step"].items()}))
except Exception as e:
    print(name, "FAILED", e)' "$2"; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
[ -x tools/peaks/peaks ] && timeout 120 tools/peaks/peaks 36 | tee gpurun_out/peaks.json
STEPS=${STEPS:-10}
timeout 300 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline > gpurun_out/bench_base.log 2>&1; summ gpurun_out/bench_base.log base
for arg in "$@"; do
  case $arg in
    persist) RTC_L2_PERSIST=1 timeout 300 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline > gpurun_out/bench_persist.log 2>&1; summ gpurun_out/bench_persist.log persist;;
  esac
done
if ls variants/*/librtc_b200.so > /dev/null 2>&1; then
  cp $PKG/librtc_b200.so /tmp/librtc_default.so
  for d in variants/*/; do
    name=$(basename "$d")
    cp "$d/librtc_b200.so" $PKG/librtc_b200.so
    ok=""
    [ "${RUN_TESTS:-0}" = "1" ] && ok=$(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ray_intersection or primary_hits or traversals_agree or sample_exact" 2>&1 | tail -1)
    timeout 300 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.log 2>&1; summ gpurun_out/bench_$name.log $name; [ -n "$ok" ] && echo "   tests: $ok"
  done
  cp /tmp/librtc_default.so $PKG/librtc_b200.so
fi
timeout 300 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline > gpurun_out/bench_base2.log 2>&1; summ gpurun_out/bench_base2.log base-again
for arg in "$@"; do
  case $arg in
    ncu)
      timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
        --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_traverse' -s 40 -c 2 -f -o gpurun_out/prof \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
      ls -la gpurun_out/prof.ncu-rep;;
  esac
done
