#!/bin/bash
# r2r: ncu evidence of the FINAL library of round 2: launch list of the bench command and a full capture of the
# AD reverse sweep (its flux addressing changed after r2a; the forward sweep is now the plain NL kernel, captured
# in r2e).  One GPU, under gpurun; every ncu run is preceded by the same command exiting 0 without ncu.
set -x
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-sweep --e2e-steps 1"
$B > gpurun_out/plain_all_r2r.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_r2r.csv $B > gpurun_out/ncu_list_r2r.log 2>&1
A="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-sweep --no-strong --modes ad"
$A > gpurun_out/plain_ad_r2r.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cloudsc2_ad -s 2 -c 1 \
    -f -o gpurun_out/prof_ad_r2r $A > gpurun_out/ncu_ad_r2r.log 2>&1
T="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-sweep --no-strong --modes tl"
$T > gpurun_out/plain_tl_r2r.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cloudsc2_tl -s 2 -c 1 \
    -f -o gpurun_out/prof_tl_r2r $T > gpurun_out/ncu_tl_r2r.log 2>&1
ls -la gpurun_out/*r2r*
