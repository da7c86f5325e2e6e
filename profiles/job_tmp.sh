source profiles/sweep.sh
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q 2>&1 | tail -2
run base
GTGPU_LIB=gtars_b200/variants/libgtars_gpu_phase.so timeout 200 python profiles/phase_timing.py 2>&1 | tail -20
run base
