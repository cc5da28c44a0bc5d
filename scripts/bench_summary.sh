#!/bin/bash
# prints a compact summary of one bench.py run (used during optimisation)
python bench.py --steps ${1:-10} --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('train rays/s %.0f  ms/step %.3f  e2e %.0f  render rays/s %.0f (%.1f%% of peak)  train frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['render']['value'], 100*d['render']['frac_of_peak'], d['train_frac_of_peak']))
print('roofline', d['roofline'])
print('clocks', d['clocks'])
for k,v in d['kernels'].items(): print('  %-26s %.3f ms  %6.1f TF  %6.0f GB/s' % (k, v['ms_per_step'], v.get('tflops',0), v.get('hbm_gbs',0)))
"
