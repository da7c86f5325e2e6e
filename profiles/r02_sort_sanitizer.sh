#!/bin/bash
# compute-sanitizer over the radix sort / packed group-by (fragment tests at 150 k fragments, all three barcode ranges)
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 500 compute-sanitizer --tool $tool --kernel-name-exclude kns=at --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fragments_device_resident_entry_point and bits" > gpurun_out/sort_sanitizer_$tool.log 2>&1
  echo "$tool exit=$?" >> gpurun_out/sort_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|passed|failed|exit=|Race reported|hazard" gpurun_out/sort_sanitizer_$tool.log | head -8
done
