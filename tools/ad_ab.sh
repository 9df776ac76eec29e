#!/bin/bash
# A/B of AD reverse-sweep formulations: expects tools/probes/libad_<name>.so = the product library built with a
# patched csrc/cloudsc2_ad_kernel.cu (build each variant with __graft_entry__.build_product() and copy the .so;
# results of the round-2 run: profiles/r2m_ad_ab.log).  usage: tools/ad_ab.sh name1 name2 ...
for v in "$@" "$@"; do
  CLOUDSC2_LIB=tools/probes/libad_$v.so python bench.py --no-cpu --no-sweep --no-strong --no-e2e --modes ad --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['modes']
print('$v', {k:round(m[k]['ms_per_step'],4) for k in ('ad','ad_have_trajectory')})"
done
