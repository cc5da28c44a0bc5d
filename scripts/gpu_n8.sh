#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo "exit $?"
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/n${N}_bench.json') if l.startswith('{')][-1])
    print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['notes']['launch'])
    print('render', d['render']['value'], 'guidance', d['guidance']['value'], 'gtrain', d['guidance_train']['value'], 'cfg4', d['train_cfg4'])
    print(d['clocks'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/n${N}_bench.err').read()[-2500:])
PY
