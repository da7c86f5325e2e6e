#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:igd_count_kernel -s 3 -c 1 -f -o gpurun_out/c4_igd_count_m1 \
  python bench_configs.py --configs c4 --scale 1.0 --steps 2 > gpurun_out/c4_igd_ncu.log 2>&1; echo "ncu exit=$?"
tail -2 gpurun_out/c4_igd_ncu.log | cut -c1-200
