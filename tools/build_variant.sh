#!/bin/bash
# tools/build_variant.sh <tag> [extra nvcc flags...] -> tools/scratch/libb200c_<tag>.so (A/B runs on the GPU box: B200C_LIB=<path>)
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
out=tools/scratch/libb200c_$tag.so
mkdir -p tools/scratch/obj_$tag
pids=()
for s in decode.cu merge.cu encode.cu api.cu sst_host.cc; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" \
    -x cu -c toplingdb_b200/csrc/$s -o tools/scratch/obj_$tag/${s%.*}.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o $out tools/scratch/obj_$tag/*.o -lcudart
rm -rf tools/scratch/obj_$tag
echo $out
