timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/tmp_bench.json 2> gpurun_out/tmp_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/tmp_bench.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
for k, v in d['hbm_stages'].items(): print(k, round(v['ms_median'], 4), round(v['frac_of_hbm_peak'], 3))
print(d['clocks'])
PY
