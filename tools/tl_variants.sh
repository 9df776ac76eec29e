#!/bin/bash
# A/B the TL kernel: ring depth (CSC2_TL_STAGES) and CTAs per SM (CSC2_TL_MINB)
# usage: tools/tl_variants.sh "STAGES MINB" ...
for cfg in "$@"; do
  set -- $cfg
  CSC2_TL_STAGES=$1 CSC2_TL_MINB=${2:-2} python bench.py --no-e2e --no-cpu --no-sweep --steps 10 --modes tl 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tl stages/minb $cfg', round(d['modes']['tl']['ms_per_step'],4), 'ms', round(d['modes']['tl']['frac_of_hbm'],4))"
done
