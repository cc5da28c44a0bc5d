timeout 600 python -m pytest tests/test_gpu_mlp.py -q -x -p no:cacheprovider 2>&1 | tail -5
for P in 524288 262144; do python scripts/prof_fused.py $P 2>&1 | grep "best"; done
timeout 300 python scripts/stress_bwd.py 2>&1 | tail -2
