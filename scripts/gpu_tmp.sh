timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "composite" 2>&1 | tail -2
for v in "" variants/cf3/libmvip_nerf.so variants/cf4/libmvip_nerf.so variants/cf6/libmvip_nerf.so; do echo "== $v"; MVIP_LIB=$v timeout 300 python scripts/hbm_stages.py 2>&1 | grep "composite" | grep -v "^{"; done
