#!/usr/bin/env python
"""bench.py -- CLOUDSC2 NL / TL / AD throughput on B200 (columns/s, KLEV = 137) against the HBM roofline.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                      (reference CPU path on the host cores)

One "step" = one pass of the nonlinear hot path (fused SATUR + CLOUDSC2, the block loop of
CLOUDSC_DRIVER, reference src/cloudsc2_nl/cloudsc_driver_mod.F90:82-111) over this rank's columns.
Workload (BASELINE.json configs[3]/[4]): 100 synthetic source columns (seed 0) expanded ON THE
DEVICE to NGPTOT = 163 840 columns per GPU (1280 blocks of NPROMA = 128; x 8 GPUs = 1 310 720
columns, the smallest scaling size of configs[4]); blocks are sharded contiguously over the ranks,
there is no data-path collective (weak scaling).  The same timed loop is repeated for the
tangent-linear and adjoint kernels and reported under "modes".

Printed keys (one JSON line, rank 0):
  value / ms_per_step : NL columns/s over all ranks, inputs resident in HBM, CUDA events, max over ranks
  e2e                 : same metric through the host-pointer C ABI call cloudsc2_gpu_nl (pinned HOST
                        arrays in the reference layout, H2D and D2H inside the timed region)
  roofline            : NL kernel, algorithmic bytes 27 440 B/column / launch duration
  cpu_baseline        : the CPU oracle's CLOUDSC_DRIVER loop (C restatement of the reference; the
                        Fortran reference cannot be built in this image) on the host cores
  modes               : tl / ad columns/s and roofline fractions (57 072 / 84 512 B/column as written)
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

KLEV = 137
NL_BYTES_PER_COL = (14 * KLEV + (KLEV + 1) + 6 * KLEV + 4 * (KLEV + 1)) * 8           # 27 440
TL_BYTES_PER_COL = 2 * (2193 + 1374) * 8                                              # 57 072 (as written)
AD_BYTES_PER_COL = 10564 * 8                                                          # 84 512 (as written)
# what our kernels actually have to move:
#  TL: 2056 traj in (SATUR fused) + 2193 incr in + 1374 traj out + 1374 incr out
TL_BYTES_MOVED = (2056 + 2193 + 1374 + 1374) * 8
# DRAM traffic per column measured by ncu (dram__bytes_read.sum + dram__bytes_write.sum of one launch
# over 163 840 columns, profiles/r1_{nl,tl,ad}_ncu.md) and FP64-pipe utilisation of the same capture
NCU = {"nl": {"dram_bytes_per_column": 4.739e9 / 163840, "fp64_pipe_pct": 60.5, "profile": "profiles/r1c_nl_ncu.md"},
       "tl": {"dram_bytes_per_column": 9.236e9 / 163840, "fp64_pipe_pct": 56.9, "profile": "profiles/r1_tl_ncu.md"},
       # AD = forward sweep (the NL kernel, 4.74 GB; its flux outputs are the check-points) + reverse sweep
       # kernel (12.108 GB)
       "ad": {"dram_bytes_per_column": (12.108e9 + 4.741e9) / 163840, "fp64_pipe_pct": 43.3,
              "profile": "profiles/r1_ad_ncu.md"}}
METRIC = "NL columns/s (KLEV=137)"
UNIT = "columns/s"


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)),
                "power_w_max": float(max(power)), "samples": len(sm), "reasons": sorted(reasons)}


def host_threads() -> int:
    """All host cores this process may use.  Not omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1 to every rank, which would silently make the CPU arm single-threaded."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_run(pkg, ob, src, prm, nproma: int, ngptot: int, threads: int, repeats: int):
    """The reference CPU path (oracle restatement of CLOUDSC_DRIVER) on `ngptot` columns.
    Returns best columns/s over `repeats` (the reference times the block loop only)."""
    st = pkg.ArrayState(src, nproma, ngptot)
    best = float("inf")
    os.environ.setdefault("OMP_SCHEDULE", "static")
    for _ in range(repeats):
        best = min(best, ob.driver_nl(prm, src.ceta, st, numomp=threads))
    return ngptot / best, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ngptot-per-gpu", type=int, default=163840)
    ap.add_argument("--ngptot-total", type=int, default=0,
                    help="strong scaling: fixed total NGPTOT block-sharded over the ranks "
                         "(BASELINE config 5: 1310720 .. 5242880); overrides --ngptot-per-gpu")
    ap.add_argument("--nproma", type=int, default=128)
    ap.add_argument("--modes", default="nl,tl,ad")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-sweep", action="store_true", help="skip the NPROMA sweep of BASELINE config 4")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world

    pkg = importlib.import_module("dwarf-p-cloudsc2-tl-ad_b200")
    prm = pkg.default_params(lregcl=False)
    src = pkg.synth_source(seed=0, klon=100, klev=KLEV)
    nproma = args.nproma
    strong = args.ngptot_total > 0
    ngp_total = args.ngptot_total if strong else args.ngptot_per_gpu * args.gpus
    # this rank's shard: contiguous block range, same arithmetic as dwarf_cloudsc.F90:65-69
    ngp = pkg.shard_blocks(ngp_total, nproma, rank if world == args.gpus else 0, args.gpus).ngptot
    scaling = "strong" if strong else "weak"
    config = {"workload": f"CLOUDSC2 NL (SATUR+CLOUDSC2), KLEV={KLEV}, NPROMA={nproma}, "
                          f"NGPTOT={ngp} per GPU x {args.gpus} GPU(s) = {ngp_total}, "
                          "100 synthetic source columns (seed 0) expanded on device",
              "klev": KLEV, "nproma": nproma, "ngptot_per_gpu": ngp, "ngptot_total": ngp_total,
              "parallelism": f"block-sharded x{args.gpus}, no data-path collective",
              "l2": f"inputs ({ngp * 16448 / 1e9:.1f} GB per GPU) larger than L2; no flush needed"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        from tests import oracle_binding as ob
        threads = host_threads()
        sample = min(ngp, 32768)
        t_steps = []
        st = pkg.ArrayState(src, nproma, sample)
        os.environ.setdefault("OMP_SCHEDULE", "static")
        for i in range(args.warmup + args.steps):
            t = ob.driver_nl(prm, src.ceta, st, numomp=threads)
            if i >= args.warmup:
                t_steps.append(t)
        ms = 1e3 * float(np.mean(t_steps))
        val = sample / (ms * 1e-3)
        line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"{sample} of the {ngp} columns per step (NPROMA={nproma}), "
                                           "block loop only, C restatement of the reference "
                                           "(Fortran reference not buildable here)"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available() or not pkg.gpu_available():
        raise SystemExit("bench.py: no CUDA device -- the CLOUDSC2 B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    gpu = pkg.Cloudsc2(prm, KLEV, src.ceta, device=local_rank)
    sh = pkg.shard_blocks(ngp_total, nproma, rank, world)
    assert sh.ngptot == ngp
    # a non-default torch stream: the kernels are launched on it through the ABI's `stream`
    # argument, and torch.cuda.Event records on it (events only see torch's current stream)
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    # device-resident problem: upload the 100 source columns once, expand on the device
    ds = pkg.DeviceState.from_source(gpu, src, nproma, ngp, gcol0=sh.gcol0, stream=stream)
    torch.cuda.synchronize()

    peak, peak_src = load_peaks()
    sampler = ClockSampler(local_rank) if rank == 0 else None

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / steps)       # ms per step, max over ranks

    modes = [m for m in args.modes.split(",") if m]
    results = {}

    # ---- NL (the headline) ---------------------------------------------------------------
    l0 = gpu.launch_count()
    if sampler:
        sampler.start()
    ms_nl = timed(lambda: gpu.nl_dev(ds, src.ptsphy, stream=stream), args.steps, args.warmup)
    launches = gpu.launch_count() - l0 - args.warmup
    value = ngp_total / (ms_nl * 1e-3)
    nl_gbs = NL_BYTES_PER_COL * ngp / (ms_nl * 1e-3) / 1e9
    def ncu_part(mode, ms):
        n = NCU[mode]
        return {"dram_gbs_actual": n["dram_bytes_per_column"] * ngp / (ms * 1e-3) / 1e9,
                "fp64_pipe_pct_ncu": n["fp64_pipe_pct"], "ncu_profile": n["profile"]}

    results["nl"] = {"columns_per_s": value, "ms_per_step": ms_nl, "gbs_per_gpu": nl_gbs,
                     "frac_of_hbm": nl_gbs / peak, "bytes_per_column": NL_BYTES_PER_COL,
                     **ncu_part("nl", ms_nl)}

    # torch views on library memory via __cuda_array_interface__
    class _Wrap:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False),
                                             "version": 2}

    def make_increments(state):
        """The 16 + 10 increment arrays of CLOUDSC2TL / CLOUDSC2AD for a device state: dx = 0.01 x built on
        the device by scaling copies of the inputs (cloudsc_driver_tl_mod.F90:156-171), 1e-7 elsewhere."""
        m2 = state.nproma * KLEV * state.nblocks
        m2h = state.nproma * (KLEV + 1) * state.nblocks
        a = {n: gpu.malloc(8 * (m2h if n == "paph" else m2)) for n in pkg._abi.INCR_IN}
        b = {n: gpu.malloc(8 * (m2h if n.startswith("pf") else m2)) for n in pkg._abi.INCR_OUT}
        for n in ("paph", "pap", "pq", "pt", "plude", "plu", "pmfu", "pmfd", "psupsat"):
            cnt = m2h if n == "paph" else m2
            torch.as_tensor(_Wrap(a[n], cnt), device=dev).copy_(
                torch.as_tensor(_Wrap(state.ptr[n], cnt), device=dev) * 0.01)
        for n in ("pqs", "pl", "pi", "gtent", "gtenq", "gtenl", "gteni"):
            torch.as_tensor(_Wrap(a[n], m2), device=dev).fill_(1e-7)
        torch.cuda.synchronize()
        return a, b, m2, m2h

    def fill_output_adjoints(b, m2, m2h):
        for n in pkg._abi.INCR_OUT:
            torch.as_tensor(_Wrap(b[n], m2h if n.startswith("pf") else m2), device=dev).fill_(1e-6)
        torch.cuda.synchronize()

    # ---- TL / AD ------------------------------------------------------------------------------
    if "tl" in modes or "ad" in modes:
        din, dout, n2, n2h = make_increments(ds)
        if "tl" in modes:
            ms_tl = timed(lambda: gpu.tl_dev(ds, src.ptsphy, din, dout, stream=stream),
                          max(3, args.steps // 2), 3)
            gbs = TL_BYTES_PER_COL * ngp / (ms_tl * 1e-3) / 1e9
            results["tl"] = {"columns_per_s": ngp_total / (ms_tl * 1e-3), "ms_per_step": ms_tl,
                             "gbs_per_gpu": gbs, "frac_of_hbm": gbs / peak,
                             "bytes_per_column": TL_BYTES_PER_COL,
                             "note": "CLOUDSC2TL as written: 16+16 arrays in, 10+10 out",
                             **ncu_part("tl", ms_tl)}
        if "ad" in modes:
            try:
                fill_output_adjoints(dout, n2, n2h)
                ms_ad = timed(lambda: gpu.ad_dev(ds, src.ptsphy, din, dout, stream=stream),
                              max(3, args.steps // 2), 3)
                gbs = AD_BYTES_PER_COL * ngp / (ms_ad * 1e-3) / 1e9
                results["ad"] = {"columns_per_s": ngp_total / (ms_ad * 1e-3), "ms_per_step": ms_ad,
                                 "gbs_per_gpu": gbs, "frac_of_hbm": gbs / peak,
                                 "bytes_per_column": AD_BYTES_PER_COL,
                                 "note": "CLOUDSC2AD as written: traj in/out, adjoints RMW",
                                 **ncu_part("ad", ms_ad)}
                # the adjoint behind an existing trajectory (every 4D-Var inner loop; the reference's own
                # adjoint test runs CLOUDSC2TL on the same inputs first): no forward sweep, the flux
                # check-points are the PFPLSL5/PFPLSN5 already in the state.  Bytes: as written minus the
                # 1374 trajectory outputs not re-written plus the 2x137 fluxes read = 9464 values.
                gpu.nl_dev(ds, src.ptsphy, stream=stream)
                gpu.set_option("ad_have_trajectory", 1)
                try:
                    ms_ad2 = timed(lambda: gpu.ad_dev(ds, src.ptsphy, din, dout, stream=stream),
                                   max(3, args.steps // 2), 3)
                finally:
                    gpu.set_option("ad_have_trajectory", 0)
                gbs2 = 9464 * 8 * ngp / (ms_ad2 * 1e-3) / 1e9
                results["ad_have_trajectory"] = {
                    "columns_per_s": ngp_total / (ms_ad2 * 1e-3), "ms_per_step": ms_ad2, "gbs_per_gpu": gbs2,
                    "frac_of_hbm": gbs2 / peak, "bytes_per_column": 9464 * 8,
                    "note": "reverse sweep only (option ad_have_trajectory): trajectory fluxes taken from a "
                            "preceding NL/TL call on the same inputs"}
            except pkg.Cloudsc2Error as e:
                results["ad"] = {"error": str(e)}
        for p in list(din.values()) + list(dout.values()):
            gpu.free(p)

    # clocks / throttle reasons sampled over all the kernel timing loops above (NL, TL, AD)
    clocks = sampler.stop() if sampler else None

    # ---- NPROMA sweep (BASELINE config 4: NGPTOT = 160 000, NPROMA 32..256): NL, TL and AD kernels ---
    sweep = None
    if not args.no_sweep and world == 1:
        sweep = {}
        for npr in (32, 64, 128, 256):
            d2 = pkg.DeviceState.from_source(gpu, src, npr, 160000, stream=stream)
            torch.cuda.synchronize()
            ms = timed(lambda: gpu.nl_dev(d2, src.ptsphy, stream=stream), 5, 3)
            sweep[str(npr)] = {"columns_per_s": 160000 / (ms * 1e-3), "ms_per_step": ms,
                               "frac_of_hbm": NL_BYTES_PER_COL * 160000 / (ms * 1e-3) / 1e9 / peak}
            if "tl" in modes or "ad" in modes:
                a2, b2, m2, m2h = make_increments(d2)
                if "tl" in modes:
                    ms = timed(lambda: gpu.tl_dev(d2, src.ptsphy, a2, b2, stream=stream), 3, 3)
                    sweep[str(npr)]["tl"] = {"columns_per_s": 160000 / (ms * 1e-3), "ms_per_step": ms,
                                             "frac_of_hbm": TL_BYTES_PER_COL * 160000 / (ms * 1e-3) / 1e9 / peak}
                if "ad" in modes:
                    fill_output_adjoints(b2, m2, m2h)
                    ms = timed(lambda: gpu.ad_dev(d2, src.ptsphy, a2, b2, stream=stream), 3, 3)
                    sweep[str(npr)]["ad"] = {"columns_per_s": 160000 / (ms * 1e-3), "ms_per_step": ms,
                                             "frac_of_hbm": AD_BYTES_PER_COL * 160000 / (ms * 1e-3) / 1e9 / peak}
                for q in list(a2.values()) + list(b2.values()):
                    gpu.free(q)
            d2.free()

    # ---- e2e through the host-pointer C ABI call ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        st = pkg.ArrayState(src, nproma, ngp, gcol0=sh.gcol0)
        # host arrays in page-locked memory obtained from the library (cloudsc2_gpu_host_alloc);
        # BENCH_E2E_HOSTMEM=register page-locks the NumPy allocations instead (~15 % slower DMA)
        pinned, host_ptrs = [], []
        if os.environ.get("BENCH_E2E_HOSTMEM", "alloc") == "register":
            for n, a in st.a.items():
                gpu.pin(a)
                pinned.append(a)
        else:
            for n in list(st.a):
                st.a[n], p = gpu.host_alloc_like(st.a[n])
                host_ptrs.append(p)
        n2 = nproma * KLEV * st.nblocks
        n2h = nproma * (KLEV + 1) * st.nblocks
        h2d = 8 * (8 * n2 + n2h + 2 * n2 + 4 * n2)          # 8 plain + PAPH + PCLV(QL,QI) + B_CML(T,Q,QL,QI)
        d2h = 8 * (4 * n2 + n2 + 2 * n2h)                   # B_LOC(T,Q,QL,QI) + PA + PFPLSL/PFPLSN; the zero fields and
        #                                                     PFHPSL/PFHPSN = -L*flux are filled on the host (e2e_host_derive)
        gpu.nl(st)                                           # warm-up (allocates staging buffers)
        gpu.nl(st)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            gpu.nl(st)                                       # synchronous: returns with results on the host
        t_e2e = (time.perf_counter() - t0) / args.e2e_steps
        barrier()
        t_e2e = max_over_ranks(t_e2e)
        e2e = {"value": ngp_total / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e,
               "api": "cloudsc2_gpu_nl (host arrays in the reference layout, page-locked memory from "
                      "cloudsc2_gpu_host_alloc)",
               "checksum_tend_t": float(np.abs(st.a["b_loc"][:, 0]).sum())}
        for a in pinned:
            gpu.unpin(a)
        del st
        for p in host_ptrs:
            gpu.host_free(p)

    total_launches = int(sum_over_ranks(float(launches)))

    # ---- the path's only collective, on the real interconnect: block-sharded Taylor and adjoint
    # self-tests (BASELINE configs 2/3 scaled to 6400 columns per rank), norms all-reduced (MAX)
    # over the ranks with NCCL (cloudsc_driver_tl_mod.F90:125, cloudsc_driver_ad_mod.F90:107)
    selftests = None
    try:
        t0 = time.perf_counter()
        z, _ = pkg.sharded_taylor(gpu, src, nproma, 6400 * world, rank, world, device=dev)
        pen, istart = pkg.taylor_verdict(z)
        gpu_ad = pkg.Cloudsc2(pkg.default_params(lregcl=True), KLEV, src.ceta, device=local_rank)
        zn, _ = pkg.sharded_adjoint(gpu_ad, src, nproma, 6400 * world, rank, world, device=dev)
        # the reference's TL / AD PROGRAMS time their whole test loop (per block 1 NL + 1 TL + 10 perturbed
        # NL, resp. 1 TL + 1 AD; cloudsc_driver_tl_mod.F90:126-254, cloudsc_driver_ad_mod.F90:108-271):
        # the same loops as one library call each on this rank's device-resident columns
        def host_timed(fn, n=3):
            fn()
            barrier()
            t = time.perf_counter()
            for _ in range(n):
                fn()
            torch.cuda.synchronize()
            return max_over_ranks((time.perf_counter() - t) / n)
        t_ad_drv = host_timed(lambda: gpu_ad.ad_test(ds, src.ptsphy))
        gpu_ad.close()
        gpu._bind()
        t_tl_drv = host_timed(lambda: gpu.tl_taylor(ds, src.ptsphy))
        results["tl_taylor_driver"] = {"columns_per_s": ngp_total / t_tl_drv, "ms_per_step": 1e3 * t_tl_drv,
                                       "note": "dwarf-cloudsc2-tl's timed loop as one call: 1 NL + 1 TL + 10 perturbed "
                                               "NL sweeps + ERROR_NORM, max over blocks"}
        results["ad_test_driver"] = {"columns_per_s": ngp_total / t_ad_drv, "ms_per_step": 1e3 * t_ad_drv,
                                     "note": "dwarf-cloudsc2-ad's timed loop as one call: 1 TL + 1 AD + dot products"}
        # cost of the collective itself: MAX all-reduce of the ten Taylor norms, device tensor, 20 calls
        torch.cuda.synchronize()
        ta = time.perf_counter()
        for _ in range(20):
            pkg.allreduce_norms(z, "max", dev)
        torch.cuda.synchronize()
        allreduce_us = (time.perf_counter() - ta) / 20 * 1e6
        selftests = {"ngptot_total": 6400 * world, "taylor_penalty": pen, "taylor_passed": 0 <= pen <= 5,
                     "allreduce_us_per_call": allreduce_us,
                     "taylor_ratios": [float(v) for v in z], "adjoint_znormg_eps": zn,
                     "adjoint_passed": bool(pkg.adjoint_verdict(zn)),
                     "allreduce": "nccl max over %d rank(s)" % world if world > 1 else "single rank",
                     "seconds": time.perf_counter() - t0}
    except Exception as e:                          # evidence only: never lose the bench line over it
        selftests = {"error": str(e)}

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from tests import oracle_binding as ob
        threads = host_threads()
        sample = min(ngp, 65536)
        cps, _ = cpu_reference_run(pkg, ob, src, prm, nproma, sample, threads, repeats=3)
        cps4, _ = cpu_reference_run(pkg, ob, src, prm, 32, min(sample, 32768), min(4, threads), repeats=2)
        cpu = {"value": cps, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{sample} columns, NPROMA={nproma}, best of 3, block loop only "
                         "(C restatement of the reference; Fortran not buildable in this image)",
               "readme_config_4threads_nproma32": cps4}
        # TL / AD kernels of the CPU port on the same cores (CLOUDSC2TL / CLOUDSC2AD block loops,
        # increments = 1 % of the inputs), for the "modes" lines: a small sample, they are slow
        try:
            st_c = pkg.ArrayState(src, nproma, 16384)
            cpu["tl_columns_per_s"] = 16384 / min(ob.bench_tlad("tl", prm, src.ceta, st_c, numomp=threads)
                                                  for _ in range(2))
            cpu["ad_columns_per_s"] = 16384 / min(ob.bench_tlad("ad", prm, src.ceta, st_c, numomp=threads)
                                                  for _ in range(2))
            cpu["tl_ad_sample"] = "16384 columns, best of 2, block loop only"
            # the reference's TL / AD programs' own timed loops (Taylor test, adjoint test) on the CPU port
            st_d = pkg.ArrayState(src, nproma, 4096)
            cpu["tl_taylor_driver_columns_per_s"] = 4096 / ob.driver_tl(prm, src.ceta, st_d, numomp=threads)[2]
            prm_ad = pkg.default_params(lregcl=True)
            cpu["ad_test_driver_columns_per_s"] = 4096 / ob.driver_ad(prm_ad, src.ceta, st_d, numomp=threads)[2]
            cpu["driver_sample"] = "4096 columns, one pass of CLOUDSC_DRIVER_TL / CLOUDSC_DRIVER_AD"
        except Exception as e:                      # never let the baseline leg kill the bench line
            cpu["tl_ad_error"] = str(e)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_nl, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "impl": "b200",
                "roofline": {"bound": "hbm", "achieved": nl_gbs, "peak": peak, "unit": "GB/s",
                             "frac": nl_gbs / peak,
                             "traffic": NCU["nl"]["dram_bytes_per_column"] * ngp,
                             "traffic_note": "bytes per launch; ncu dram__bytes_read+write of one launch at "
                                             "163 840 columns scaled by NGPTOT (profiles/r1c_nl_ncu.md)",
                             "peak_source": peak_src, "kernel": "k_cloudsc2_nl",
                             "algorithmic_bytes_per_column": NL_BYTES_PER_COL,
                             "algorithmic_bytes_per_launch": NL_BYTES_PER_COL * ngp,
                             "fp64_pipe_pct_ncu": NCU["nl"]["fp64_pipe_pct"]},
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": total_launches, "clocks": clocks,
                "modes": results, "nproma_sweep_ngptot160000": sweep, "selftests": selftests}
        # the reference's own report lines (timer_mod.F90:114-174), on stderr
        sys.stderr.write(pkg.report.performance_table(1, ngp_total, pkg.nblocks(ngp_total, nproma), nproma,
                                                      ms_nl * 1e-3, numproc=world) + "\n")
        print(json.dumps(line))
    ds.free()
    gpu.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
