timeout 300 python -m pytest tests/test_gpu_mlp.py -q -x -p no:cacheprovider 2>&1 | tail -2
for v in "" variants/wl/libmvip_nerf.so "" variants/wl/libmvip_nerf.so; do echo "== $v"; MVIP_LIB=$v python scripts/prof_fwd.py 2>&1 | grep "stash="; MVIP_LIB=$v python scripts/prof_fused.py 524288 2>&1 | grep "backward_fused"; done
