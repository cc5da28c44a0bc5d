#!/bin/bash
# full GPU suite + both bench arms on one B200 (what the driver runs at round end)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/val_ref.json 2> gpurun_out/val_ref.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/val_bench.json') if l.startswith('{')][-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'steps', d['steps'])
    for k, v in d['kernels'].items(): print(k, round(v['ms_per_step'], 4), {a: round(b, 3) for a, b in v.items() if a.startswith('frac') or a in ('tflops', 'hbm_gbs')})
    r = d['roofline']; print({k: r[k] for k in ('kernel', 'bound', 'achieved', 'peak', 'frac', 'frac_burst', 'frac_sustained', 'traffic', 'share_of_step')}); print(r['step']); print(r['secondary'])
    print(d['clocks']); print(d.get('cpu_baseline'))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/val_bench.err').read()[-1500:])
r = json.loads([l for l in open('gpurun_out/val_ref.json') if l.startswith('{')][-1]); print('reference arm', r['value'], r['cpu_baseline'])
PY
