timeout 200 python -m pytest tests/test_gpu_mlp.py -q -x -p no:cacheprovider -k "repack or empty_batch" 2>&1 | tail -15
