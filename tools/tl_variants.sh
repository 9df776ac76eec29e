#!/bin/bash
# A/B the TL kernel ring depth (CSC2_TL_STAGES) on one B200
for st in "$@"; do
  CSC2_TL_STAGES=$st python bench.py --no-e2e --no-cpu --no-sweep --steps 10 --modes tl 2>&1 | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tl stages $st', round(d['modes']['tl']['ms_per_step'],4), 'ms', round(d['modes']['tl']['frac_of_hbm'],4))"
done
