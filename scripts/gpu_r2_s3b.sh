#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_fused.py 524288 2>&1 | tee gpurun_out/s3b_prof.txt
python scripts/prof_fused.py 262144 2>&1 | tee -a gpurun_out/s3b_prof.txt
TRACE=1 MVIP_LIB=variants/trace/libmvip_nerf.so python scripts/prof_fused.py 524288 2>&1 | tee gpurun_out/s3b_trace.txt
bash scripts/profile.sh > gpurun_out/s3b_profile.log 2>&1
tail -3 gpurun_out/s3b_profile.log
