cd /root/repo
RUN_TESTS=0 bash tools/experiments/run_variants.sh --steps 10
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_traverse' -s 40 -c 2 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/prof.ncu-rep
