#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py -q -x -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2d_tests.log
tail -15 gpurun_out/r2d_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2d_bench.json') if l.startswith('{')][-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
    for k, v in d['kernels'].items(): print(k, round(v['ms_per_step'], 4), {a: round(b, 3) for a, b in v.items() if a.startswith('frac') or a in ('tflops', 'hbm_gbs')})
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2d_bench.err').read()[-1500:])
PY
