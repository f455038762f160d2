#!/bin/bash
# ncu --set full on k_traverse* and k_shade (2 launches from the middle of a frame), after a plain run of the same command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_traverse' -s 20 -c 2 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/prof.ncu-rep
