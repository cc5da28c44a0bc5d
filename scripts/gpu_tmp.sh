mkdir -p gpurun_out
date +%s > gpurun_out/t0
timeout 420 bash scripts/profile.sh > gpurun_out/profile.log 2>&1
echo "profile.sh done after $(( $(date +%s) - $(cat gpurun_out/t0) )) s"
python scripts/prof_fused.py 524288 > gpurun_out/r02_prof_fused.txt 2>&1
python scripts/prof_fused.py 262144 >> gpurun_out/r02_prof_fused.txt 2>&1
python scripts/bwd_timeline.py 524288 > gpurun_out/r02_bwd_timeline.txt 2>&1
python scripts/hbm_stages.py > gpurun_out/r02_hbm_stages.txt 2>&1; grep -v "^{" gpurun_out/r02_hbm_stages.txt | tail -9
echo "all done after $(( $(date +%s) - $(cat gpurun_out/t0) )) s"; ls -la gpurun_out | head -30
