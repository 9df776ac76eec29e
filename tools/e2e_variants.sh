#!/bin/bash
# e2e (host-pointer cloudsc2_gpu_nl) vs staging chunk size
for mb in "$@"; do
  CSC2_E2E_CHUNK_MB=$mb python bench.py --modes nl --no-cpu --steps 5 --e2e-steps 5 2>&1 | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk_mb $mb e2e ms', round(d['e2e']['ms_per_step'],2), 'col/s', round(d['e2e']['value']))"
done
