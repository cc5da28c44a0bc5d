#!/bin/bash
# Run under gpurun on one B200: launch list + full captures of the MLP kernels (one bench step, and the render-only forward).
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'mlp_forward_pair_kernel|dgrad_pair_kernel|wgrad_kernel|composite_fwd_kernel|composite_bwd_kernel|head_grads_kernel|reduce_kernel|adam_kernel' \
    -s 30 -c 11 -o gpurun_out/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof.log 2>&1
# render-only forward (no stash), 4.2 M points; and the HBM-bound kernels at image-sized batches
ncu --set full --clock-control none -k regex:'mlp_forward_pair_kernel|composite_fwd_kernel|composite_bwd_kernel|sample_fine_kernel|sample_coarse_kernel' \
    -s 6 -c 9 -o gpurun_out/prof_render python scripts/time_kernels.py > gpurun_out/prof_render.log 2>&1
ls -la gpurun_out
# the reports are large (gpurun returns at most 64 MiB): keep the raw metric tables only
ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_render.ncu-rep --page raw --csv > gpurun_out/prof_render_raw.csv 2>/dev/null
rm -f gpurun_out/prof.ncu-rep gpurun_out/prof_render.ncu-rep
du -sh gpurun_out
