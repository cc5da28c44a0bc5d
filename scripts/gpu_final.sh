#!/bin/bash
mkdir -p gpurun_out
python scripts/hbm_stages.py > gpurun_out/r02_hbm_stages.txt 2>&1; head -8 gpurun_out/r02_hbm_stages.txt
bash scripts/gpu_validate.sh
