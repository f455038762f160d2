#!/bin/bash
# tools/gpu_round.sh -- one GPU-box visit: full parity suite, smoke, the bench (frames pipelined and not), the
# unspecialised k_shade variant if variants/ has one, then the ncu launch list of the bench command.
# Everything is written under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/tests.log
tail -3 gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_main.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pipeline > gpurun_out/bench_pipe.log 2>&1
for f in main pipe; do
  tail -1 gpurun_out/bench_$f.log | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print("%-8s Mpaths/s %.1f  Mrays/s %.1f  ms/step %.2f  e2e %.1f  kernels %s  frac %.3f" % (sys.argv[1], d["value"], d["mrays_per_s"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in r["kernel_ms_per_step"].items()}, r["frac"]))' $f
done
if [ -d variants ] && ls variants/*/librtc_b200.so > /dev/null 2>&1; then
  PKG=raytracing-course_b200
  cp $PKG/librtc_b200.so /tmp/librtc_default.so
  for d in variants/*/; do
    name=$(basename "$d")
    cp "$d/librtc_b200.so" $PKG/librtc_b200.so
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.log 2>&1
    tail -1 gpurun_out/bench_$name.log | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print("%-8s Mpaths/s %.1f  ms/step %.2f  e2e %.1f  kernels %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in r["kernel_ms_per_step"].items()}))' $name
  done
  cp /tmp/librtc_default.so $PKG/librtc_b200.so
fi
if [ "$1" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_traverse' -s 40 -c 2 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out/prof.ncu-rep
fi
