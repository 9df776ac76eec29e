#!/bin/bash
# How much of the gap to the HBM roofline at 163 840 columns is wave quantisation?  Same kernels at sizes
# that are whole multiples of a full wave (592 NL CTAs of 128 columns = 75 776 columns) and at large sizes.
for n in "$@"; do
  python bench.py --modes nl,tl,ad --no-e2e --no-cpu --no-sweep --no-strong --steps 10 --ngptot-per-gpu $n 2>/dev/null | tail -1 | \
    python -c "
import json,sys
d=json.loads(sys.stdin.read()); m=d['modes']
print('ngptot $n', ' '.join('%s %.4f ms %.4f' % (k, m[k]['ms_per_step'], m[k]['frac_of_hbm']) for k in ('nl','tl','ad','ad_have_trajectory')), d['clocks'].get('sm_mhz'))"
done
