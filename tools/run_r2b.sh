#!/bin/bash
# r2b: device-set tests + the bench line at N = 1 and N = 2 (library communicator under torchrun).
set -x
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2b.log
tail -3 gpurun_out/pytest_gpu_r2b.log
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_r2b_n1.json 2> gpurun_out/bench_r2b_n1.err; echo "bench n1 rc=$?"
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/bench_r2b_n2.json 2> gpurun_out/bench_r2b_n2.err; echo "bench n2 rc=$?"
tail -5 gpurun_out/bench_r2b_n1.err gpurun_out/bench_r2b_n2.err
