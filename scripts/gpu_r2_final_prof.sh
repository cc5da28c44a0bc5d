#!/bin/bash
# round-2 evidence pass: launch list + full ncu captures (scripts/profile.sh), fused-backward counters / hand-over timeline, sanitizers
mkdir -p gpurun_out
bash scripts/profile.sh > gpurun_out/profile.log 2>&1
python scripts/prof_fused.py 524288 > gpurun_out/r02_prof_fused.txt 2>&1
python scripts/prof_fused.py 262144 >> gpurun_out/r02_prof_fused.txt 2>&1
python scripts/bwd_timeline.py 524288 > gpurun_out/r02_bwd_timeline.txt 2>&1
for tool in racecheck synccheck memcheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_mlp.py > gpurun_out/r02_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r02_$tool.log
  tail -4 gpurun_out/r02_$tool.log
done
python scripts/hbm_stages.py > gpurun_out/r02_hbm_stages.txt 2>&1; tail -15 gpurun_out/r02_hbm_stages.txt
