#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 110 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:radix_scatter_kernel<.int.8, .int.512, .int.16, .int.2, .int.2" -c 1 -f -o gpurun_out/sort_scatter_keys_final \
  python profiles/r02_sort_sweep.py 2.5e8 1 "0" > gpurun_out/sort_ncu_final.log 2>&1; echo "ncu exit=$?"
tail -2 gpurun_out/sort_ncu_final.log | cut -c1-200
