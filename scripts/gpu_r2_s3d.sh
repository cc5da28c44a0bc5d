#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py -q -x -p no:cacheprovider 2>&1 | tail -5
for st in 0 350 716 1400; do
  echo "== stagger $st"; STAGGER=$st python scripts/prof_fused.py 524288 2>&1 | grep -v "^head\|^reduce"
done | tee gpurun_out/s3d_prof.txt
STAGGER=716 python scripts/prof_fused.py 262144 2>&1 | grep -v "^head\|^reduce" | tee -a gpurun_out/s3d_prof.txt
for st in 0 716; do
STAGGER=$st timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:backward_fused -c 3 --csv --log-file gpurun_out/s3d_ncu_$st.csv python scripts/prof_fused.py 524288 > /dev/null 2>&1
grep backward_fused gpurun_out/s3d_ncu_$st.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tail -3
done
