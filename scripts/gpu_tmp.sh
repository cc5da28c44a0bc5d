mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for i in 1 2 3; do python scripts/prof_fused.py 524288 2>&1 | grep "backward_fused"; done
python scripts/prof_fused.py 262144 2>&1 | grep "backward_fused"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/tmp_bench.json 2> gpurun_out/tmp_bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/tmp_bench.json') if l.startswith('{')][-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
    for k, v in d['kernels'].items(): print(k, round(v['ms_per_step'], 4), {a: round(b, 3) for a, b in v.items() if a.startswith('frac') or a in ('tflops', 'hbm_gbs')})
    print(d['clocks'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/tmp_bench.err').read()[-1500:])
PY
