#!/bin/bash
mkdir -p gpurun_out
timeout 300 ./scripts/ubench/l2_handoff > gpurun_out/s3c_handoff.txt 2>&1
cat gpurun_out/s3c_handoff.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/s3c_handoff_ncu.csv ./scripts/ubench/l2_handoff > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(l for l in open('gpurun_out/s3c_handoff_ncu.csv') if l.startswith('"'))]
h = rows[0]; ki, mi, vi, ui = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit')
ii = h.index('ID')
d = {}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki][:14]), {})[r[mi]] = (r[vi], r[ui])
for k in sorted(d):
    print(k, {m: v for m, v in d[k].items()})
PY
