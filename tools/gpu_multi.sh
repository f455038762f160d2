#!/bin/bash
# tools/gpu_multi.sh N -- on a box with N GPUs: the multi-device tests, the torchrun bench at N ranks, and the
# single-process multi-device CLI (RTC_DEVICES) against one device
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_device or frames_in_flight" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err || tail -20 gpurun_out/bench_${N}gpu.err
tail -1 gpurun_out/bench_${N}gpu.log | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print("N=%d Mpaths/s %.1f  ms/step %.2f  e2e %.1f (%.2f ms)  kernels %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], {k: round(v, 2) for k, v in r["kernel_ms_per_step"].items()}))'
DEVS=$(python -c "print(','.join(str(i) for i in range($N)))")
for d in 0 $DEVS; do
  /usr/bin/time -f "run.sh RTC_DEVICES=$d wall %e s" env RTC_DEVICES=$d ./run.sh scenes/practice5_dragon_100k.txt /tmp/out_$N.ppm 2>&1 | tail -1
done
timeout 300 python tools/experiments/multi_probe.py $N | tee gpurun_out/multi_probe_${N}.json
