#!/bin/bash
# one gpurun call: GPU test suite on the new scatter kernel, shape sweep at 2.5e8 and 1e9 fragments, ncu of the best shape
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/sort_pytest.log 2>&1; echo "pytest exit=$?" >> gpurun_out/sort_pytest.log
tail -3 gpurun_out/sort_pytest.log
timeout 600 python profiles/r02_sort_sweep.py 2.5e8 5 > gpurun_out/sort_sweep_2.5e8.jsonl 2> gpurun_out/sort_sweep_2.5e8.err
cat gpurun_out/sort_sweep_2.5e8.jsonl
BEST=$(python - <<'P'
import json
rows=[json.loads(l) for l in open("gpurun_out/sort_sweep_2.5e8.jsonl") if l.startswith("{")]
rows=[r for r in rows if r.get("equals_first_shape", True) and r["prefetch_bucket_row"]==1]
rows.sort(key=lambda r:r["ms"])
print(",".join(str(r["shape"]) for r in rows[:3]))
P
)
echo "best shapes: $BEST"
timeout 600 python profiles/r02_sort_sweep.py 1e9 5 "0,$BEST" > gpurun_out/sort_sweep_1e9.jsonl 2> gpurun_out/sort_sweep_1e9.err
cat gpurun_out/sort_sweep_1e9.jsonl
B1=${BEST%%,*}
GTGPU_RS_SHAPE=$B1 timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:radix_scatter_kernel<9" -c 2 -f -o gpurun_out/sort_scatter_shape$B1 \
  python profiles/r02_sort_sweep.py 2.5e8 1 "$B1" > gpurun_out/sort_ncu.log 2>&1
tail -2 gpurun_out/sort_ncu.log
ls -la gpurun_out
