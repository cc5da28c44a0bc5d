timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "sample_coarse" 2>&1 | tail -2
for v in "" variants/cb3/libmvip_nerf.so variants/cb2/libmvip_nerf.so; do echo "== $v"; MVIP_LIB=$v timeout 300 python scripts/hbm_stages.py 2>&1 | grep "composite_bwd\|sample_coarse" | grep -v "^{"; done
