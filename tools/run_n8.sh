#!/bin/bash
# 8-GPU evidence (BASELINE config 5).  Run as: gpurun --gpus 8 -- bash tools/run_n8.sh
#  1. the device-set tests (tests/test_gpu_multi.py, tests/test_programs.py) with 8 GPUs in ONE process
#  2. the C++ programs: ONE process drives 8 GPUs through the library (cloudsc2_gpu_init_multi: worker thread
#     and stream set per device, NCCL all-reduce of the norms / validation statistics) -- no fork, no pipes
#  3. bench.py under torchrun at N = 8 (the driver's launch): weak line + strong_1310720 / strong_5242880 blocks
set -u
out=gpurun_out; mkdir -p $out
export LD_LIBRARY_PATH=$(python -c "import nvidia.nccl, os; print(os.path.join(list(nvidia.nccl.__path__)[0], 'lib'))" 2>/dev/null):${LD_LIBRARY_PATH:-}
python -m pytest tests/test_gpu_multi.py tests/test_programs.py -q -m gpu > $out/pytest_gpu_n8.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu_n8.log
B=dwarf-p-cloudsc2-tl-ad_b200/bin
{
  for prog in nl tl ad; do
    echo "== CLOUDSC2_NGPUS=8 dwarf-cloudsc2-$prog 1 1310720 128"
    CLOUDSC2_NGPUS=8 CLOUDSC2_REPEAT=3 $B/dwarf-cloudsc2-$prog 1 1310720 128 2>&1; echo "rc=$?"
  done
  echo "== CLOUDSC2_NGPUS=8 dwarf-cloudsc2-nl 1 5242880 128"
  CLOUDSC2_NGPUS=8 CLOUDSC2_REPEAT=3 $B/dwarf-cloudsc2-nl 1 5242880 128 2>&1; echo "rc=$?"
  echo "== CLOUDSC2_NGPUS=1 dwarf-cloudsc2-nl 1 1310720 128"
  CLOUDSC2_NGPUS=1 CLOUDSC2_REPEAT=3 $B/dwarf-cloudsc2-nl 1 1310720 128 2>&1; echo "rc=$?"
  echo "== CLOUDSC2_NGPUS=8 CLOUDSC2_HOST_ARRAYS=3 dwarf-cloudsc2-nl 1 1310720 128   (host arrays, registered, sharded by the library)"
  CLOUDSC2_NGPUS=8 CLOUDSC2_HOST_ARRAYS=3 CLOUDSC2_REPEAT=2 $B/dwarf-cloudsc2-nl 1 1310720 128 2>&1; echo "rc=$?"
  echo "== NCCL_DEBUG=INFO CLOUDSC2_NGPUS=8 dwarf-cloudsc2-tl 1 800 1  (communicator evidence)"
  NCCL_DEBUG=INFO CLOUDSC2_NGPUS=8 $B/dwarf-cloudsc2-tl 1 800 1 2>&1 | grep -E "NCCL INFO (ncclCommInitAll|comm 0x|Connected|NVLS|Channel 00/|Init COMPLETE)|TEST|GPU:" | head -40
} > $out/programs_n8.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
( time $TR bench.py --gpus 8 --steps 20 --warmup 3 ) > $out/bench_n8.json 2> $out/b8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n8.json").read().strip().splitlines()[-1])
    print("N=8", d["scaling"], d["config"]["ngptot_total"], round(d["value"] / 1e6, 1), "M col/s", round(d["ms_per_step"], 4), "ms")
    print(" modes", {k: round(v["ms_per_step"], 3) for k, v in d["modes"].items()})
    print(" e2e", d["e2e"]["ms_per_step"], d["e2e"]["h2d_gbs_per_gpu"], d["e2e"]["numa"], "source", d["e2e_source"]["ms_per_step"])
    for k in d:
        if k.startswith("strong"): print(" ", k, d[k])
    print(" selftests", d["selftests"])
except Exception as e:
    print("bench_n8 failed", e)
PY
tail -4 $out/pytest_gpu_n8.log
grep -E "GPU:|TEST|rc=|^==|validation:" $out/programs_n8.log
{
  nvidia-smi topo -m 2>&1
  lscpu 2>&1 | grep -i -E "numa|socket|model name|^cpu\(s\)|thread"
  for d in /sys/bus/pci/devices/*; do
    if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist) link=$(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi
  done
  grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
  for n in /sys/devices/system/node/node*; do echo "$n cpulist=$(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
} > $out/topology_n8.log 2>&1
free -g | head -2; nproc
