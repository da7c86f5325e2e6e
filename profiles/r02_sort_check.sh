#!/bin/bash
# packed group-by vs pair sort on C5; GPU tests
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/sort4_pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/sort4_pytest.log
tail -15 gpurun_out/sort4_pytest.log
for ng in 0 1; do
GTGPU_NO_GROUP_SORT=$ng timeout 600 python profiles/r02_sort_sweep.py 1e9 5 "0" > gpurun_out/sort_sweep4_nogroup$ng.jsonl 2> gpurun_out/sort_sweep4_nogroup$ng.err
cat gpurun_out/sort_sweep4_nogroup$ng.jsonl; tail -3 gpurun_out/sort_sweep4_nogroup$ng.err
done
