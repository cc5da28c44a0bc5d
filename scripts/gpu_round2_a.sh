#!/bin/bash
# first GPU pass of round 2: full GPU test suite, bench (both arms), sanitizer passes over the tcgen05 kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --deselect tests/test_gpu_train_parity.py::test_trained_network_render_vs_reference -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r2a_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err
for tool in racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_mlp.py > gpurun_out/r2a_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r2a_$tool.log
done
tail -5 gpurun_out/r2a_tests.log
