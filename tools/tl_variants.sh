#!/bin/bash
# A/B the TL kernel: ring depth (CSC2_TL_STAGES) and CTA shape (CSC2_TL_MINB) -- experiments build only:
#   make -C tools/probes experiments && CLOUDSC2_LIB=tools/probes/libcloudsc2_b200_experiments.so tools/tl_variants.sh ...
# usage: tools/tl_variants.sh NGPTOT "STAGES MINB" ...
n=$1; shift
for cfg in "$@"; do
  set -- $cfg
  CSC2_TL_STAGES=$1 CSC2_TL_MINB=${2:-2} python bench.py --no-e2e --no-cpu --no-sweep --no-strong --steps 10 --modes tl --ngptot-per-gpu $n 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ngptot $n tl stages/minb $cfg', round(d['modes']['tl']['ms_per_step'],4), 'ms', round(d['modes']['tl']['frac_of_hbm'],4), d['clocks'].get('sm_mhz'))"
done
