#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_multidev.py tests/test_gpu_multi.py -q -x > gpurun_out/multidev_2gpu_pytest_final.log 2>&1; echo "pytest exit=$?" >> gpurun_out/multidev_2gpu_pytest_final.log
tail -4 gpurun_out/multidev_2gpu_pytest_final.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err; echo "bench exit=$?"
tail -c 400 gpurun_out/bench_n2_final.json; tail -3 gpurun_out/bench_n2_final.err
