#!/bin/bash
# Run under gpurun on one B200: launch list + full capture of the MLP kernels of one bench step.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'mlp_forward_pair_kernel|dgrad_chain_kernel|wgrad_kernel|composite_fwd_kernel|head_grads_kernel' \
    -s 24 -c 16 -o gpurun_out/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof.log 2>&1
ls -la gpurun_out
