#!/bin/bash
# table match vs ballot match, all shapes at 2.5e8, best at 1e9, sort-heavy tests under the table match
set -u
mkdir -p gpurun_out
timeout 600 python profiles/r02_sort_sweep.py 2.5e8 5 > gpurun_out/sort_sweep3_2.5e8.jsonl 2> gpurun_out/sort_sweep3_2.5e8.err
cat gpurun_out/sort_sweep3_2.5e8.jsonl; tail -3 gpurun_out/sort_sweep3_2.5e8.err
timeout 600 python profiles/r02_sort_sweep.py 1e9 5 "0,1" > gpurun_out/sort_sweep3_1e9.jsonl 2> gpurun_out/sort_sweep3_1e9.err
cat gpurun_out/sort_sweep3_1e9.jsonl; tail -3 gpurun_out/sort_sweep3_1e9.err
GTGPU_RS_SHAPE=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/sort3_pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/sort3_pytest.log
tail -3 gpurun_out/sort3_pytest.log
