#!/bin/bash
# 8-GPU evidence (BASELINE config 5): the C++ programs with 8 ranks, strong scaling of bench.py at
# NGPTOT = 1 310 720 and 5 242 880.  Run as: gpurun --gpus 8 -- bash tools/run_n8.sh
set -u
out=gpurun_out; mkdir -p $out
B=dwarf-p-cloudsc2-tl-ad_b200/bin
{
  for prog in nl tl ad; do
    echo "== CLOUDSC2_NUMPROC=8 dwarf-cloudsc2-$prog 1 1310720 128"
    CLOUDSC2_NUMPROC=8 CLOUDSC2_REPEAT=3 $B/dwarf-cloudsc2-$prog 1 1310720 128 2>&1; echo "rc=$?"
  done
  echo "== CLOUDSC2_NUMPROC=8 dwarf-cloudsc2-nl 1 5242880 128"
  CLOUDSC2_NUMPROC=8 CLOUDSC2_REPEAT=3 $B/dwarf-cloudsc2-nl 1 5242880 128 2>&1; echo "rc=$?"
} > $out/programs_n8.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
$TR bench.py --gpus 8 --steps 10 --warmup 3 --ngptot-total 5242880 --no-cpu --no-e2e 2>$out/b8s5.err | tail -1 > $out/bench_n8_strong5m.json
$TR bench.py --gpus 8 --steps 10 --warmup 3 --ngptot-total 1310720 --no-cpu --e2e-steps 2 2>$out/b8s1.err | tail -1 > $out/bench_n8_strong1m.json
python - <<'PY'
import json
for f in ("bench_n8_strong5m", "bench_n8_strong1m"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["scaling"], d["config"]["ngptot_total"], round(d["value"] / 1e6, 1), "M col/s", round(d["ms_per_step"], 3), "ms",
              {k: round(v["ms_per_step"], 3) for k, v in d["modes"].items()}, "e2e", d["e2e"] and round(d["e2e"]["value"] / 1e6, 2),
              d["selftests"].get("taylor_passed"), d["selftests"].get("adjoint_passed"), d["selftests"].get("allreduce_us_per_call"))
    except Exception as e:
        print(f, "failed", e)
PY
grep -E "GPU:|TEST|rc=|^==" $out/programs_n8.log
free -g | head -2; nproc
