for v in "" variants/cs/libmvip_nerf.so "" variants/cs/libmvip_nerf.so; do echo "== $v"; MVIP_LIB=$v timeout 300 python scripts/hbm_stages.py 2>&1 | grep "composite" | grep -v "^{"; done
