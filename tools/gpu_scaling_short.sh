#!/bin/bash
# tools/gpu_scaling_short.sh N [N2 ...] -- on a multi-GPU box: the headline configuration at each of the given rank counts
# (torchrun, one rank per GPU), one summary line each; logs in gpurun_out/bench_<N>gpu.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for n in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --warmup 3 --steps 20 --no-cli > gpurun_out/bench_${n}gpu.log 2> gpurun_out/bench_${n}gpu.err || tail -8 gpurun_out/bench_${n}gpu.err
  tail -1 gpurun_out/bench_${n}gpu.log | python -c '
import sys, json
try:
    d = json.loads(sys.stdin.read()); r = d["roofline"]
    print("N=%d Mpaths/s %.1f  Mrays/s %.1f  ms/step %.2f  e2e %.1f (%.2f ms, h2d %d B)  kernels %s" % (d["n_gpus"], d["value"], d["mrays_per_s"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], {k: round(v, 2) for k, v in r["kernel_ms_per_step"].items()}))
except Exception as e:
    print("FAILED", e)'
done
