timeout 300 python -m pytest tests/test_gpu_mlp.py -q -x -p no:cacheprovider 2>&1 | tail -2
for i in 1 2; do python scripts/prof_fused.py 524288 2>&1 | grep "backward_fused"; done
python scripts/prof_fused.py 262144 2>&1 | grep "backward_fused"
MVIP_LIB=variants/pc9/libmvip_nerf.so python scripts/prof_fused.py 524288 2>&1 | grep "backward_fused\|issuer\|producer"
python scripts/bwd_timeline.py 524288 2>&1 | grep "span\|mean"
