#!/bin/bash
# A/B the AD kernel's CTAs per SM (CSC2_AD_MINB = 2 -> 255 registers, 3 -> 168 registers)
for m in "$@"; do
  CSC2_AD_MINB=$m python bench.py --no-e2e --no-cpu --no-sweep --steps 10 --modes ad 2>&1 | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ad minb $m', round(d['modes']['ad']['ms_per_step'],4), 'ms', round(d['modes']['ad']['frac_of_hbm'],4))"
done
