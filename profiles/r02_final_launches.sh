#!/bin/bash
# ncu launch list of the default bench command on the final code (time only; this repo's kernels)
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gtgpu|fused_find|count_|igd_|radix_|scan_|gunzip" --csv --log-file gpurun_out/launches_bench_default_final.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --parity-files 10 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu exit=$?"
wc -l gpurun_out/launches_bench_default_final.csv
