#!/bin/bash
# what the driver runs at round end: GPU tests, smoke(), default bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/final_pytest.log
tail -3 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/final_smoke.log
tail -2 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1_final2.json 2> gpurun_out/bench_n1_final2.err; echo "bench exit=$?"
tail -c 300 gpurun_out/bench_n1_final2.json; tail -3 gpurun_out/bench_n1_final2.err
