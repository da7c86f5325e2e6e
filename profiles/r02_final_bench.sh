#!/bin/bash
# default bench line + the reference arm + ncu launch list of the same command
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo "bench exit=$?"
tail -c 600 gpurun_out/bench_n1_final.json; tail -3 gpurun_out/bench_n1_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref exit=$?"
cat gpurun_out/bench_ref_final.json | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(?!at::|void at::|elementwise|vectorized|index|reduce|cub|thrust|distribution).*" --csv --log-file gpurun_out/launches_bench_default_final.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --parity-files 10 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu exit=$?"
wc -l gpurun_out/launches_bench_default_final.csv
