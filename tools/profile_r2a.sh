#!/bin/bash
# r2a: fresh ncu evidence of the CURRENT default kernels (TL <1,2,0,0,2,128>, AD reverse sweep, the
# adjoint's forward sweep = NL kernel with check-points).  One GPU, under gpurun.  Every ncu run is
# preceded by the same command exiting 0 without ncu (B200_PROFILING.md).
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2a.log
python -m pytest tests/test_oracle_ref.py tests/test_f90toc.py -x -q > gpurun_out/pytest_ref_r2a.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
$B --no-sweep > gpurun_out/plain_all_r2a.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r2a.csv $B --no-sweep > gpurun_out/ncu_list_r2a.log 2>&1
$B --modes tl --no-sweep > gpurun_out/plain_tl_r2a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cloudsc2_tl -s 2 -c 1 \
    -f -o gpurun_out/prof_tl_r2a $B --modes tl --no-sweep > gpurun_out/ncu_tl_r2a.log 2>&1
$B --modes ad --no-sweep > gpurun_out/plain_ad_r2a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cloudsc2_ad -s 2 -c 1 \
    -f -o gpurun_out/prof_ad_r2a $B --modes ad --no-sweep > gpurun_out/ncu_ad_r2a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cloudsc2_nl -s 10 -c 1 \
    -f -o gpurun_out/prof_adfwd_r2a $B --modes ad --no-sweep > gpurun_out/ncu_adfwd_r2a.log 2>&1
ls -la gpurun_out/ | tail -12
