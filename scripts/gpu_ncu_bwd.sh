#!/bin/bash
# DRAM bytes + duration of backward_fused_kernel alone at P = 524,288 (STAGGER from the environment)
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed -k regex:backward_fused -c 3 --csv --log-file gpurun_out/ncu_bwd.csv python scripts/prof_fused.py ${1:-524288} > /dev/null 2>&1
grep backward_fused gpurun_out/ncu_bwd.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tail -5
