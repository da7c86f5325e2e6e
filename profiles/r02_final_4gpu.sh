#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 > gpurun_out/bench_n4_final.json 2> gpurun_out/bench_n4_final.err; echo "bench exit=$?"
tail -c 300 gpurun_out/bench_n4_final.json; tail -2 gpurun_out/bench_n4_final.err
