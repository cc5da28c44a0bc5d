timeout 200 python -m pytest tests/test_gpu_guidance.py -q -x -p no:cacheprovider -k "accepts" 2>&1 | tail -25
