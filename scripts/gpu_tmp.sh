mkdir -p gpurun_out
for th in "800,0" "800,30000" "500,60000" "300,100000" "1200,60000"; do echo "=== THROTTLE $th"; THROTTLE=$th python scripts/prof_fused.py 524288 2>&1 | grep "backward_fused\|chain issuer\|wgrad producer"; THROTTLE=$th python scripts/bwd_timeline.py 524288 2>&1 | grep "span\|mean"; done
THROTTLE=500,60000 bash scripts/gpu_ncu_bwd.sh
