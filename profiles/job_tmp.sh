source profiles/sweep.sh
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q 2>&1 | tail -2
run base
for c in 30 35 40 45 50; do run carve$c GTGPU_CARVEOUT=$c; done
run base
