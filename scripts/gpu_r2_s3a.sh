#!/bin/bash
# session-3 first pass: full GPU suite, bench (both arms), launch list under ncu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/s3a_tests.log
tail -8 gpurun_out/s3a_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s3a_bench.json 2> gpurun_out/s3a_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s3a_ref.json 2> gpurun_out/s3a_ref.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/s3a_bench.json') if l.startswith('{')][-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
    for k, v in d['kernels'].items(): print(k, round(v['ms_per_step'], 4), {a: round(b, 3) for a, b in v.items() if a.startswith('frac') or a in ('tflops', 'hbm_gbs')})
    print(json.dumps(d['roofline'])[:1500]); print(d.get('cpu_baseline'))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/s3a_bench.err').read()[-1500:])
PY
tail -c 600 gpurun_out/s3a_ref.json; tail -c 400 gpurun_out/s3a_ref.err
