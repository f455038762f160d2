#!/bin/bash
# tools/gpu_scaling.sh N [extra] -- on an N-GPU box: the headline workload at N ranks (the driver's own launch line);
# with "extra" also BASELINE configs 3 (glass dragon) and 4 (metal dragon, 3840x2160, 1024 spp).
cd "$(dirname "$0")/.."
N=$1
mkdir -p gpurun_out
run() {  # name, bench args...
  name=$1; shift
  if [ "$N" = "1" ]; then timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/scale_$name.log 2>&1
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
         bench.py --gpus $N "$@" > gpurun_out/scale_$name.log 2>&1; fi
  tail -1 gpurun_out/scale_$name.log | python -c '
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print("%-22s n=%d  Mpaths/s %.1f  Mrays/s %.1f  ms/step %.2f  e2e %.1f" % (sys.argv[1], d["n_gpus"], d["value"], d["mrays_per_s"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)' $name
}
run ${N}gpu --steps 10 --warmup 3 --no-cpu-baseline
run ${N}gpu_pipeline --steps 10 --warmup 3 --no-cpu-baseline --pipeline
if [ "$2" = "extra" ]; then
  run config3_glass_${N}gpu --steps 10 --warmup 3 --no-cpu-baseline --scene practice5_dragon_100k_glass
  run config4_metal_4k_1024spp_${N}gpu --steps 3 --warmup 3 --no-cpu-baseline --scene practice5_dragon_100k_metal --width 3840 --height 2160 --spp 1024
fi
