#!/bin/bash
# ncu evidence for profiles/ (r1c: NL at 16 warps/SM, staged Taylor kernel).  One GPU, under gpurun.
# Every ncu run is preceded by the same command exiting 0 without ncu (B200_PROFILING.md).
set -x
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/plain_all_r1c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r1c.csv $B > gpurun_out/ncu_list_r1c.log 2>&1
$B --modes nl --no-sweep > gpurun_out/plain_nl_r1c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cloudsc2_nl -s 4 -c 1 \
    -f -o gpurun_out/prof_nl_r1c $B --modes nl --no-sweep > gpurun_out/ncu_nl_r1c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_taylor_nl -c 1 \
    -f -o gpurun_out/prof_taylor_r1c dwarf-p-cloudsc2-tl-ad_b200/bin/dwarf-cloudsc2-tl 1 163840 128 > gpurun_out/ncu_taylor_r1c.log 2>&1
ls -la gpurun_out/ | tail -8
