#!/bin/bash
# Tuning sweep helper (run under gpurun): prints ms_per_step for library variants / env settings.
run() { # label, env...
  label=$1; shift
  out=$(env "$@" timeout 300 python bench.py --no-e2e --no-cpu --steps 5 2>&1 | tail -1)
  echo "$label $(echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms=%.2f frac=%.3f parity=%s' % (d['ms_per_step'], d['roofline']['frac'], d['config']['parity_spot_check_first_files_vs_oracle']))" 2>&1 | tail -1)"
}
V=gtars_b200/variants
"$@"
