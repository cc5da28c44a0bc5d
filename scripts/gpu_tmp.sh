timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "sample" 2>&1 | tail -3
for v in "" variants/sf3/libmvip_nerf.so variants/sf2/libmvip_nerf.so; do echo "== $v"; MVIP_LIB=$v timeout 300 python scripts/hbm_stages.py 2>&1 | grep "sample_fine" | grep -v "^{"; done
