timeout 600 python -m pytest tests/test_gpu_mlp.py -q -x -p no:cacheprovider 2>&1 | tail -3
for P in 524288 262144 1000; do python scripts/prof_fused.py $P 2>&1 | grep "best"; done
timeout 300 python scripts/stress_bwd.py 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/tmp_bench.json 2> gpurun_out/tmp_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/tmp_bench.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
for k, v in d['kernels'].items(): print(k, round(v['ms_per_step'], 4))
print(d['roofline']['step'], d['clocks'])
PY
