mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_mlp.py > gpurun_out/r02_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r02_$tool.log
  grep -c "Race reported\|Error:" gpurun_out/r02_$tool.log; grep "Race reported\|Error" gpurun_out/r02_$tool.log | sed 's/.*\(Race reported[^.]*\).*/\1/' | cut -c1-200 | sort | uniq -c | head -12; tail -3 gpurun_out/r02_$tool.log
done
