#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + one --set full capture per kernel.
# Run under gpurun (1 GPU).  Follows /opt/skills/guides/B200_PROFILING.md: every ncu run is
# preceded by the same command exiting 0 without ncu.
set -x
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/plain_all.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r1b.csv $B > gpurun_out/ncu_list.log 2>&1
$B --modes nl > gpurun_out/plain_nl.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cloudsc2_nl -s 4 -c 1 \
    -f -o gpurun_out/prof_nl_r1b $B --modes nl > gpurun_out/ncu_nl.log 2>&1
$B --modes tl > gpurun_out/plain_tl.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cloudsc2_tl -s 2 -c 1 \
    -f -o gpurun_out/prof_tl_r1b $B --modes tl > gpurun_out/ncu_tl.log 2>&1
$B --modes ad > gpurun_out/plain_ad.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cloudsc2_ad -s 2 -c 1 \
    -f -o gpurun_out/prof_ad_r1b $B --modes ad > gpurun_out/ncu_ad.log 2>&1
ls -la gpurun_out/
