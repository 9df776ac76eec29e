#!/bin/bash
# r2e: GPU suite, the bench line, the ncu launch list of the bench command, and two full captures of the NL
# kernel: at the bench size (163 840 columns = 2.16 rounds of resident CTAs) and at 1 310 720 columns
# (17.3 rounds: the steady state, where the kernel sits on the DRAM roof).  One GPU, under gpurun.  Every ncu
# run is preceded by the same command exiting 0 without ncu (B200_PROFILING.md).
set -x
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r2e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2e.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2e.json 2>/dev/null
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-sweep --e2e-steps 1"
$B > gpurun_out/plain_all_r2e.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_r2e.csv $B > gpurun_out/ncu_list_r2e.log 2>&1
N="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-sweep --no-strong --modes nl"
$N > gpurun_out/plain_nl_r2e.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cloudsc2_nl -s 4 -c 1 \
    -f -o gpurun_out/prof_nl_r2e $N > gpurun_out/ncu_nl_r2e.log 2>&1
$N --ngptot-per-gpu 1310720 > gpurun_out/plain_nl1m_r2e.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cloudsc2_nl -s 4 -c 1 \
    -f -o gpurun_out/prof_nl1m_r2e $N --ngptot-per-gpu 1310720 > gpurun_out/ncu_nl1m_r2e.log 2>&1
ls -la gpurun_out/ | tail -12
