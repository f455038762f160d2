#!/bin/bash
# tools/gpu_scaling.sh -- on an 8-GPU box: the headline configuration at 8 / 4 ranks, BASELINE configurations 4 and 5 at 8
# ranks, and the one-process multi-device call (rtc_render_u8_multi) on 1 / 2 / 4 / 8 devices
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
summ() { tail -1 "$1" | python -c '
import sys, json
try:
    d = json.loads(sys.stdin.read()); r = d["roofline"]
    print("%-14s N=%d Mpaths/s %.1f  Mrays/s %.1f  ms/step %.2f  e2e %.1f (%.2f ms)  kernels %s" % (sys.argv[1], d["n_gpus"], d["value"], d["mrays_per_s"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], {k: round(v, 2) for k, v in r["kernel_ms_per_step"].items()}))
except Exception as e:
    print(sys.argv[1], "FAILED", e)' "$2"; }
run() {  # name N extra-args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $2 --warmup 3 ${@:3} > gpurun_out/bench_$1.log 2> gpurun_out/bench_$1.err || tail -8 gpurun_out/bench_$1.err
  summ gpurun_out/bench_$1.log $1
}
run 8gpu 8 --steps 20
run 4gpu 4 --steps 20
run 2gpu 2 --steps 20
run config4_glass_8gpu 8 --config 4 --steps 20
run config5_metal_4k_8gpu 8 --config 5 --steps 3
timeout 300 python tools/experiments/multi_probe.py 8 | tee gpurun_out/multi_probe_8.json
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_device" 2>&1 | tail -2
