#!/bin/bash
# A/B of the dynamic level-segment units (CSC2_NL_SEGMENTS: 1 = whole columns, static grid; n = n segments).
# The knob exists only with csrc/experiments/nl_dynamic_units.patch applied to csrc/cloudsc2_nl_kernel.cu (a
# rejected experiment: `git apply` it on the commit it was taken from, see profiles/r2d_nl_dynamic_units.log).
for n in "$@"; do
  for seg in 1 2 3 4 6 8; do
    CSC2_NL_SEGMENTS=$seg python bench.py --modes nl,ad --no-e2e --no-cpu --no-sweep --no-strong --steps 20 --ngptot-per-gpu $n 2>/dev/null | tail -1 | \
      python -c "
import json,sys
d=json.loads(sys.stdin.read()); m=d['modes']
print('ngptot $n segments $seg', ' '.join('%s %.4f ms %.4f' % (k, m[k]['ms_per_step'], m[k]['frac_of_hbm']) for k in ('nl','ad','ad_have_trajectory')), d['clocks'].get('sm_mhz'))"
  done
done
