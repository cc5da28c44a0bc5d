#!/bin/bash
# N = 1, 2, 4, 8 back to back on one 8-GPU box (run with: gpurun --gpus 8 -- bash scripts/scale.sh)
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/scale_n$N.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale_n$N.json
  fi
  python - <<PY
import json
d=json.load(open("gpurun_out/scale_n$N.json"))
print("N=%d train %.0f rays/s (%.3f ms/step) e2e %.0f render %.0f rays/s (%.1f ms/img) clocks %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["render"]["value"], d["render"]["ms_per_image"], d["clocks"]))
PY
done
