#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mlp.py tests/test_gpu_train_parity.py tests/test_gpu_insitu.py tests/test_gpu_render.py -q -x -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/r2c_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -5 gpurun_out/r2c_tests.log
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2c_bench.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
for k, v in d['kernels'].items(): print(k, round(v['ms_per_step'], 4), {a: round(b, 3) for a, b in v.items() if a.startswith('frac') or a in ('tflops', 'hbm_gbs')})
print(json.dumps(d['roofline'])[:900])
PY
