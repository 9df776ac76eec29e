#!/bin/bash
# A/B the NL kernel variants (CSC2_NL_VARIANT) on one B200: prints ms/step and fraction of HBM peak
# usage: tools/nl_variants.sh NGPTOT v0 v1 ...
n=$1; shift
for v in "$@"; do
  CSC2_NL_VARIANT=$v python bench.py --modes nl --no-e2e --no-cpu --no-sweep --steps 20 --ngptot-per-gpu $n 2>&1 | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ngptot $n variant $v', round(d['ms_per_step'],4), 'ms', round(d['roofline']['frac'],4), d['clocks'].get('sm_mhz'), d['clocks'].get('power_w_max'))"
done
