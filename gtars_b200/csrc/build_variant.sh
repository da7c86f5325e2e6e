#!/bin/bash
# Builds an experimental variant of libgtars_gpu.so for tuning sweeps: ./build_variant.sh <name> <extra nvcc flags...>
# The result is gtars_b200/variants/libgtars_gpu_<name>.so, selected at run time with GTGPU_LIB=<path>.
set -e
cd "$(dirname "$0")"
name=$1; shift
mkdir -p ../variants /tmp/gtv_$name
for f in $(ls cuda/*.cu | xargs -n1 basename | sed "s/\.cu$//"); do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --use_fast_math "$@" \
      -Xptxas -v -c -o /tmp/gtv_$name/$f.o cuda/$f.cu 2> /tmp/gtv_$name/$f.log &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libgtars_gpu_$name.so /tmp/gtv_$name/*.o -cudart static -ldl
grep -A3 'fused_find_kernelILi[0-9]*ELb0ELb0ELb0' /tmp/gtv_$name/kernels.log | grep -E 'Used|stack' | head -2
