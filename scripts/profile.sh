#!/bin/bash
# Run under gpurun on one B200: launch list + full captures of the kernels of one bench step (eager launches: --no-graph),
# and of the HBM-bound stage kernels / render-only forward at image-sized batches.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
KERNELS='mlp_forward_pair_kernel|backward_fused_kernel|composite_fwd4_kernel|composite_bwd4_kernel|head_grads_kernel|reduce_kernel|adam_kernel|sample_fine64_kernel|sample_coarse_warp_kernel|rays_pack_kernel|pack_kernel'
ncu --set full --clock-control none --import-source on -k regex:"$KERNELS" \
    -s 40 -c 20 -o gpurun_out/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/prof.log 2>&1
ncu --set full --clock-control none -k regex:'mlp_forward_pair_kernel|composite_fwd4_kernel|composite_bwd4_kernel|sample_fine64_kernel|sample_coarse_warp_kernel|rays_kernel' \
    -s 9 -c 9 -o gpurun_out/prof_render python scripts/prof_hbm_stages.py > gpurun_out/prof_render.log 2>&1
ls -la gpurun_out
# the reports are large (gpurun returns at most 64 MiB): keep the raw metric tables only
ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_render.ncu-rep --page raw --csv > gpurun_out/prof_render_raw.csv 2>/dev/null
rm -f gpurun_out/prof.ncu-rep gpurun_out/prof_render.ncu-rep
du -sh gpurun_out
