#!/bin/bash
mkdir -p gpurun_out
./scripts/ubench/l2_bw > gpurun_out/r2b_l2bw.txt 2>&1
python -m pytest tests/test_gpu_train_parity.py tests/test_gpu_insitu.py -q -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/r2b_tests.log
python tests/insitu_worker.py > gpurun_out/r2b_insitu.log 2>&1
tail -5 gpurun_out/r2b_tests.log; cat gpurun_out/r2b_l2bw.txt
