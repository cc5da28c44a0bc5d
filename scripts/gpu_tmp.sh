# scratch: the command of the last one-off GPU experiment (kept so the run is reproducible)
for P in 524288 262144; do python scripts/prof_fused.py $P 2>&1 | grep "best"; done
