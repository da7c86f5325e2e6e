#!/bin/bash
# IGD tests + the C4 count kernel (bench_configs.py c4 at full scale); optional env in front, e.g. GTGPU_IGD_NO_M1=1
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "igd or lola or Igd" > gpurun_out/igd_pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/igd_pytest.log
tail -3 gpurun_out/igd_pytest.log
for v in "GTGPU_IGD_CTAS=6" "GTGPU_IGD_CTAS=5"; do
  echo "== $v"
  env $v timeout 600 python bench_configs.py --configs c4 --scale 1.0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ms_per_step','kernel_ms','hits_per_s','parity_sample_vs_oracle','region_file_hits')})"
done
