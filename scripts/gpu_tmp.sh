timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "sample_fine or sample_pdf" 2>&1 | tail -3
timeout 100 python scripts/hbm_stages.py 2>&1 | grep "sample_fine" | grep -v "^{"
