#!/bin/bash
# Development A/B builds: tools/build_variant.sh <name> [extra nvcc flags...]  -> build_variants/libstad_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
rm -rf build_variants/$name build_variants/libstad_$name.so
mkdir -p build_variants/$name
pids=()
for f in host rowwise gemm attention api; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c simple-tad_b200/csrc/$f.cu -o build_variants/$name/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build_variants/libstad_$name.so build_variants/$name/*.o
echo build_variants/libstad_$name.so
