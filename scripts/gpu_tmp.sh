mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'pair_kernel<\(bool\)1' -s 1 -c 1 -o gpurun_out/fwd_src python scripts/prof_fwd.py > gpurun_out/fwd_src.log 2>&1
ls -la gpurun_out/fwd_src.ncu-rep; tail -3 gpurun_out/fwd_src.log
