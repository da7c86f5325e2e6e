#!/bin/bash
# shapes with larger tiles, launch list of the C5 pipeline, ncu full of the scatter kernel (shape 0)
set -u
mkdir -p gpurun_out
timeout 600 python profiles/r02_sort_sweep.py 1e9 5 "0,7,8" > gpurun_out/sort_sweep2_1e9.jsonl 2> gpurun_out/sort_sweep2_1e9.err
cat gpurun_out/sort_sweep2_1e9.jsonl
GTGPU_RS_SHAPE=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/sort_c5_launches_1e9.csv \
  python profiles/r02_sort_sweep.py 1e9 1 "0" > gpurun_out/sort_ncu_launches.log 2>&1
tail -2 gpurun_out/sort_ncu_launches.log
GTGPU_RS_SHAPE=0 timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:radix_scatter_kernel<.int.9" -c 2 -f -o gpurun_out/sort_scatter_shape0 \
  python profiles/r02_sort_sweep.py 2.5e8 1 "0" > gpurun_out/sort_ncu.log 2>&1
tail -2 gpurun_out/sort_ncu.log
ls -la gpurun_out
