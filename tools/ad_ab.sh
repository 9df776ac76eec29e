#!/bin/bash
for v in base p6 p8 base p6 p8; do
  CLOUDSC2_LIB=tools/probes/libad_$v.so python bench.py --no-cpu --no-sweep --no-strong --no-e2e --modes ad --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['modes']
print('$v', {k:round(m[k]['ms_per_step'],4) for k in ('ad','ad_have_trajectory')})"
done
