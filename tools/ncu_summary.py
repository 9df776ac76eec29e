#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch, --set full) into markdown for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_nl_r1b.ncu-rep  KLEV  [title]  > profiles/....md

Reads the raw page (metrics) and the source page (per-SASS-instruction executed counts and stall
samples) with `ncu -i ... --csv`; no GPU needed.
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
klev = int(sys.argv[2]) if len(sys.argv) > 2 else 137
title = sys.argv[3] if len(sys.argv) > 3 else rep


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                          text=True).stdout


rows = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, val = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, val)}


def g(name, default="n/a"):
    return m.get(name, (default, ""))[0]


def gf(name):
    try:
        return float(g(name).replace(",", ""))
    except ValueError:
        return float("nan")


def unit(name):
    return m.get(name, ("", ""))[1]


def to_bytes(name):
    v, u = gf(name), unit(name)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


def to_ms(name):
    v, u = gf(name), unit(name)
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "usecond": 1e-3, "msecond": 1, "second": 1e3,
                "nsecond": 1e-6}.get(u, 1)


kname = g("Kernel Name")
dur = to_ms("gpu__time_duration.sum")
rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
grid, block = int(gf("launch__grid_size")), int(gf("launch__block_size"))
warps = grid * block / 32
inst = gf("smsp__inst_executed.sum")
print(f"# {title}\n")
print(f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`, one launch; times under ncu are "
      "cold-cache and serialised — use SHARES, the bench numbers come from CUDA events).\n")
print(f"* kernel: `{kname}`")
print(f"* grid {grid} x block {block}; registers/thread {g('launch__registers_per_thread')}; "
      f"dynamic smem/CTA {g('launch__shared_mem_per_block_dynamic')} {unit('launch__shared_mem_per_block_dynamic')}; "
      f"occupancy limit: registers {g('launch__occupancy_limit_registers')} / smem {g('launch__occupancy_limit_shared_mem')} CTAs per SM; "
      f"achieved warps active {gf('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} % of 64/SM")
print(f"* duration {dur:.3f} ms; SM clock {g('smsp__cycles_elapsed.avg.per_second')} {unit('smsp__cycles_elapsed.avg.per_second')}")
print(f"* DRAM read {rd / 1e9:.3f} GB + write {wr / 1e9:.3f} GB = **{(rd + wr) / 1e9:.3f} GB per launch** "
      f"({(rd + wr) / dur / 1e6:.0f} GB/s under ncu; dram throughput {gf('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} % of ncu's peak)")
print(f"* FP64 pipe: `sm__inst_executed_pipe_fp64` {gf('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.1f} % "
      f"of peak sustained (active) — the pipe issues one warp instruction per 2 cycles per sub-partition")
print(f"* issue slots busy {gf('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} %; "
      f"eligible warps/cycle {gf('smsp__warps_eligible.avg.per_cycle_active'):.2f}; "
      f"executed warp instructions {inst:.4g} = {inst / warps / klev:.0f} per warp per level")
print(f"* L2 hit rate {gf('lts__t_sector_hit_rate.pct'):.1f} %; L1 hit rate {gf('l1tex__t_sector_hit_rate.pct'):.1f} %; "
      f"local (spill) loads {g('sass__inst_executed_local_loads')} / stores {g('sass__inst_executed_local_stores')}")
print("\n## Warp stall reasons (cycles per issued instruction)\n")
st = []
for h in hdr:
    mm = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
    if mm and mm.group(1) != "selected":
        st.append((gf(h), mm.group(1)))
print("| reason | ratio |\n|---|---|")
for v, n in sorted(st, reverse=True)[:8]:
    print(f"| {n} | {v:.2f} |")

src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "sass"))))
h2 = src[1]
iS, iE, iSm = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
cnt, smp, tot = collections.Counter(), collections.Counter(), 0
for r in src[2:]:
    if len(r) <= iE:
        continue
    mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", r[iS].strip())
    if not mm:
        continue
    op = mm.group(2)
    if op in ("MUFU", "LDG", "STG", "RED", "ATOMG"):
        op += (mm.group(3) or "")
    e = int(r[iE] or 0)
    cnt[op] += e
    smp[op] += int(r[iSm] or 0)
    tot += e
print(f"\n## Dynamic instruction mix (SASS, {tot / warps / klev:.0f} warp instructions per level)\n")
print("| opcode | per warp per level | % of executed | stall samples |\n|---|---|---|---|")
for op, e in cnt.most_common(22):
    print(f"| {op} | {e / warps / klev:.1f} | {100 * e / tot:.1f} | {smp[op]} |")
fp64 = sum(e for op, e in cnt.items() if op in ("DFMA", "DMUL", "DADD", "DSETP") or op.startswith("MUFU.RCP64H") or op.startswith("MUFU.RSQ64H"))
print(f"\nFP64-pipe instructions (DFMA+DMUL+DADD+DSETP+MUFU.*64H): **{fp64 / warps / klev:.0f} per warp per level** "
      f"= {100 * fp64 / tot:.0f} % of executed instructions.")
