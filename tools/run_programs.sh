#!/bin/bash
# Run the three dwarf programs with the reference's README command lines on the GPU box and keep
# what they print (gpurun_out/programs_r1.log); also record the host topology.
set -u
out=gpurun_out; mkdir -p $out
B=dwarf-p-cloudsc2-tl-ad_b200/bin
{
  echo "== dwarf-cloudsc2-nl 4 160000 32"; CLOUDSC2_REPEAT=5 $B/dwarf-cloudsc2-nl 4 160000 32 2>&1; echo "rc=$?"
  echo "== dwarf-cloudsc2-nl 4 160000 128"; CLOUDSC2_REPEAT=5 $B/dwarf-cloudsc2-nl 4 160000 128 2>&1; echo "rc=$?"
  echo "== CLOUDSC2_HOST_ARRAYS=1 dwarf-cloudsc2-nl 4 160000 32"; CLOUDSC2_HOST_ARRAYS=1 CLOUDSC2_REPEAT=3 $B/dwarf-cloudsc2-nl 4 160000 32 2>&1; echo "rc=$?"
  echo "== dwarf-cloudsc2-tl 1 100 1"; $B/dwarf-cloudsc2-tl 1 100 1 2>&1; echo "rc=$?"
  echo "== dwarf-cloudsc2-tl 4 160000 32"; CLOUDSC2_REPEAT=3 $B/dwarf-cloudsc2-tl 4 160000 32 2>&1; echo "rc=$?"
  echo "== dwarf-cloudsc2-ad 1 100 100"; $B/dwarf-cloudsc2-ad 1 100 100 2>&1; echo "rc=$?"
  echo "== dwarf-cloudsc2-ad 4 160000 32"; CLOUDSC2_REPEAT=3 $B/dwarf-cloudsc2-ad 4 160000 32 2>&1; echo "rc=$?"
} > $out/programs_r1.log
{
  nvidia-smi topo -m 2>&1
  lscpu 2>&1 | grep -i -E "numa|socket|model name|^cpu\(s\)|thread"
  for d in /sys/bus/pci/devices/*; do
    if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist)"; fi
  done
  grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
  which numactl && numactl --hardware
  ls /sys/devices/system/node/ | head
  for n in /sys/devices/system/node/node*; do echo "$n cpulist=$(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
} > $out/topology.log 2>&1
