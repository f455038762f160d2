#!/bin/bash
# tools/gpu_round.sh [ncu] -- one GPU-box visit: full parity suite, smoke, the bench of the headline configuration, then
# (with "ncu") the launch list and one --set full capture of the two hot kernels.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv,noheader > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/tests.log
tail -3 gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_main.log 2>gpurun_out/bench_main.err || tail -5 gpurun_out/bench_main.err
tail -1 gpurun_out/bench_main.log | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print("Mpaths/s %.1f  Mrays/s %.1f  ms/step %.2f  e2e %.1f (%.2f ms)  load_s %.2f cli_s %s" % (d["value"], d["mrays_per_s"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["load_s"], d["cli_s"]))
print("kernels", {k: round(v, 2) for k, v in r["kernel_ms_per_step"].items()})
for k, x in d["rooflines"].items():
    print(" ", k, x["bound"], "frac %.3f" % (x["frac"] or 0), "achieved %.0f of %.0f GB/s" % (x["achieved"], x["peak"] or 0), "issue", (x.get("issue") or {}).get("frac"))
print("cpu", d["cpu_baseline"])'
if [ "$1" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-peaks --no-cli > gpurun_out/ncu_launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_traverse' -s 40 -c 2 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-peaks --no-cli > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out/prof.ncu-rep
fi
