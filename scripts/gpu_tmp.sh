mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/final_bench.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'steps', d['steps'])
for k, v in d['kernels'].items(): print(k, round(v['ms_per_step'], 4), {a: round(b, 3) for a, b in v.items() if a.startswith('frac') or a in ('tflops', 'hbm_gbs')})
r = d['roofline']; print({k: r[k] for k in ('kernel', 'bound', 'achieved', 'peak', 'frac', 'frac_burst', 'frac_sustained', 'traffic', 'share_of_step')}); print(r['step']); print(r['secondary'])
print(d['clocks']); print(d.get('cpu_baseline')); print(d['render'], d['guidance'], d['guidance_train'])
PY
