#!/bin/bash
# Builds a variant of libmvip_nerf.so with extra defines into variants/<name>/libmvip_nerf.so (experiments only; git-ignored).
#   scripts/build_variant.sh trace -DMVIP_TRACE_BWD
set -e
name=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
out=$here/variants/$name
mkdir -p $out/obj
cd $here/mvip_nerf_b200/csrc
for f in api adam rays embed sampler composite normal umma_selftest mlp_pack mlp_forward mlp_backward; do
  extra=""; [ $f = sampler ] && extra="-fmad=false"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $extra "$@" -c $f.cu -o $out/obj/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libmvip_nerf.so $out/obj/*.o -lcudart
rm -rf $out/obj
echo built $out/libmvip_nerf.so
