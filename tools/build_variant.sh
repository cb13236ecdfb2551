#!/bin/bash
# Development A/B builds: tools/build_variant.sh <name> [extra nvcc flags...]  -> build_variants/libstad_<name>.so
# ONLY="attention gemm" recompiles just those translation units with the flags and links the rest from the shipped
# build (simple-tad_b200/build/*.o must be current: python simple-tad_b200/build.py).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
rm -rf build_variants/$name build_variants/libstad_$name.so
mkdir -p build_variants/$name
all="host rowwise gemm attention api"
only=${ONLY:-$all}
pids=()
for f in $all; do
  if [[ " $only " == *" $f "* ]]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c simple-tad_b200/csrc/$f.cu -o build_variants/$name/$f.o &
    pids+=($!)
  else
    cp simple-tad_b200/build/$f.o build_variants/$name/$f.o
  fi
done
for p in "${pids[@]}"; do wait $p; done
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build_variants/libstad_$name.so build_variants/$name/*.o
echo build_variants/libstad_$name.so
