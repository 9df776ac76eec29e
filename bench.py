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
  e2e                 : same metric through the host-pointer C ABI call cloudsc2_gpu_nl on arrays the HOST
                        allocated itself (NumPy = malloc, like the Fortran ALLOCATEs of expand_mod.F90:110)
                        and page-locked once with cloudsc2_gpu_host_register (time reported); H2D and D2H
                        inside the timed region.  Beside it: e2e.pageable (no registration at all),
                        e2e.host_alloc (arrays from cloudsc2_gpu_host_alloc) and e2e_source (the dwarf's own
                        work flow in one call: 100 source columns up, device expansion, NL, device validation)
  strong_<NGPTOT>     : BASELINE config 5: a FIXED total of 1 310 720 (and, N >= 2, 5 242 880) columns
                        block-sharded over the N ranks: NL / TL / AD ms and columns/s
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
# DRAM traffic per column measured by ncu (dram__bytes_read.sum + dram__bytes_write.sum of ONE launch over
# 163 840 columns) and FP64-pipe utilisation of the same capture.  Every entry names the capture it comes
# from (kernel instantiation + profiles/ file); tools/profile_r2*.sh regenerate them.
NCU = {"nl": {"dram_bytes_per_column": 4.741e9 / 163840, "fp64_pipe_pct": 60.5, "profile": "profiles/r2e_nl_ncu.md",
              "kernel": "k_cloudsc2_nl<0,2,128,128,0,0>"},
       "tl": {"dram_bytes_per_column": 9.237e9 / 163840, "fp64_pipe_pct": 55.9, "profile": "profiles/r2r_tl_ncu.md",
              "kernel": "k_cloudsc2_tl<0,2,0,0,2,128>"},
       # AD = forward sweep (the NL kernel without the driver-level zeroing; its PFPLSL/PFPLSN outputs are the
       # flux check-points: 4.562 GB = the NL launch's 4.741 GB minus the 0.18 GB of zeros for CLD(:,:,NCLV),
       # profiles/r2e_nl_ncu.md, r2a_adfwd_ncu.md) + reverse sweep kernel (12.108 GB, profiles/r2r_ad_ncu.md)
       "ad": {"dram_bytes_per_column": (12.108e9 + 4.562e9) / 163840, "fp64_pipe_pct": 43.6,
              "profile": "profiles/r2r_ad_ncu.md + profiles/r2a_adfwd_ncu.md",
              "kernel": "k_cloudsc2_nl<0,2,128,128,0,0> (loc_last = NULL) + k_cloudsc2_ad<0,0,0,2>"}}
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


def numa_bind(gpu_index: int) -> dict:
    """Bind this process (and hence the first-touch placement of the host arrays it allocates next) to the
    CPUs of the NUMA node its GPU hangs off, as read from sysfs.  On a virtualised box the node is -1."""
    info = {"bound": False}
    try:
        bus = subprocess.run(["nvidia-smi", f"--id={gpu_index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("0000") and len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        info["pci_bus_id"] = bus
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text().strip())
        info["numa_node"] = node
        if node >= 0:
            cpus = set()
            for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["bound"] = True
                info["cpus"] = len(cpus)
    except Exception as e:
        info["error"] = repr(e)
    return info


def host_threads() -> int:
    """All host cores this process may use.  Not omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1 to every rank, which would silently make the CPU arm single-threaded."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def load_cpu_arm():
    """The CPU implementation the reference arm / cpu_baseline time: oracle/_ref -- the reference's OWN Fortran
    kernels (satur.F90, cloudsc2.F90, ...), transliterated statement by statement to C by oracle/f90toc.py from
    /root/reference/src in the build container and compiled by gcc (the .so travels to the GPU box; there is no
    Fortran compiler in this image) -- when it is there, kind "reference"; else the hand-written C restatement
    (oracle/), kind "port".  Both run the same OpenMP block loop (oracle/cloudsc2_drivers.c, the mirror of
    cloudsc_driver_mod.F90:73-119)."""
    from tests import oracle_binding as ob
    ref_lib = ROOT / "oracle" / "_ref" / "libcloudsc2_ref.so"
    if ref_lib.exists() and os.environ.get("BENCH_CPU_ARM", "ref") != "port":
        import importlib.util
        spec = importlib.util.spec_from_file_location("oracle_binding_ref", ROOT / "tests" / "oracle_binding.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.ORACLE_LIB = ref_lib
        try:
            if b"f90toc" in mod.flavour():
                return mod, "reference", ("the reference's Fortran kernels transliterated to C (oracle/f90toc.py) and "
                                          "compiled with gcc -O2 -ffp-contract=off; OpenMP block loop as in "
                                          "cloudsc_driver_mod.F90:73-119")
        except Exception:
            pass
    return ob, "port", "C restatement of the reference (oracle/); Fortran reference not buildable here"


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
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-total blocks of BASELINE config 5")
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
        # The reference's own CPU implementation of the path (load_cpu_arm: its Fortran kernels transliterated to C
        # and compiled by gcc, oracle/_ref, when that library travelled here; else the hand-written C restatement)
        # under an OpenMP block loop like cloudsc_driver_mod.F90:73-81 -- the Fortran itself cannot be built in
        # this image.  None of the repo's kernels run here: libcloudsc2_b200.so is mapped by this process
        # ONLY because the synthetic input generator and the host-side expansion (pkg.synth_source,
        # pkg.ArrayState) live in that library; no cloudsc2_gpu_* compute entry is called.
        ob, cpu_kind, cpu_desc = load_cpu_arm()
        threads = host_threads()
        sample = min(ngp, 32768)
        t_steps = []
        st = pkg.ArrayState(src, nproma, sample)
        os.environ.setdefault("OMP_SCHEDULE", "static")
        for i in range(args.warmup + args.steps):
            t = ob.driver_nl(prm, src.ceta, st, numomp=threads)
            if i >= args.warmup:
                t_steps.append(t)
        ms_sample = 1e3 * float(np.mean(t_steps))
        val = sample / (ms_sample * 1e-3)
        # one step of the stated workload = ngp columns: the sample's time scaled to it (the block loop is
        # embarrassingly parallel and the sample is >= 256 blocks per thread-team pass, so time is linear)
        ms = ms_sample * ngp / sample
        line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "ms_per_step_note": f"scaled to the stated {ngp} columns from the timed sample of {sample} "
                                    f"({ms_sample:.3f} ms per sample pass)",
                "cores": threads, "sample_columns": sample, "ms_per_sample": ms_sample,
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": cpu_kind,
                                 "sample": f"{sample} of the {ngp} columns per step (NPROMA={nproma}), "
                                           f"block loop only; {cpu_desc}"},
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

    # one process per GPU here (the driver's launch contract); the norms / validation statistics are reduced by
    # the LIBRARY's own NCCL communicator, joined after every context initialisation
    gpu = pkg.Cloudsc2(prm, KLEV, src.ceta, device=local_rank,
                       after_init=pkg.driver.torch_comm_factory(rank, world))
    comm_rank, comm_size, nccl_version = gpu.comm_info()
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
                "fp64_pipe_pct_ncu": n["fp64_pipe_pct"], "ncu_profile": n["profile"], "ncu_kernel": n["kernel"]}

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
    e2e_source = None
    if not args.no_e2e:
        numa = numa_bind(local_rank) if world > 1 or os.environ.get("BENCH_NUMA_BIND") else {"bound": False}
        st = pkg.ArrayState(src, nproma, ngp, gcol0=sh.gcol0)      # NumPy = malloc'ed, pageable: the unchanged host
        n2 = nproma * KLEV * st.nblocks
        n2h = nproma * (KLEV + 1) * st.nblocks
        h2d = 8 * (8 * n2 + n2h + 2 * n2 + 4 * n2)          # 8 plain + PAPH + PCLV(QL,QI) + B_CML(T,Q,QL,QI)
        d2h = 8 * (4 * n2 + n2 + 2 * n2h)                   # B_LOC(T,Q,QL,QI) + PA + PFPLSL/PFPLSN; the zero fields and
        #                                                     PFHPSL/PFHPSN = -L*flux are filled on the host (e2e_host_derive)

        def time_nl(steps):
            gpu.nl(st)                                       # warm-up (allocates staging buffers)
            gpu.nl(st)
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                gpu.nl(st)                                   # synchronous: returns with results on the host
            t = (time.perf_counter() - t0) / steps
            barrier()
            return max_over_ranks(t)

        # (1) no help from the host at all: pageable arrays
        t_page = time_nl(max(1, args.e2e_steps - 1))
        # (2) THE DEFAULT: the host's own arrays, page-locked once with cloudsc2_gpu_host_register
        t0 = time.perf_counter()
        pinned, reg_error = [], None
        try:
            for a in st.a.values():
                gpu.pin(a)
                pinned.append(a)
        except pkg.Cloudsc2Error as e:           # e.g. a locked-memory limit on the host: measured, not fatal
            reg_error = str(e)
        t_reg = max_over_ranks(time.perf_counter() - t0)
        registered = max_over_ranks(0.0 if reg_error is None else 1.0) == 0.0      # the same decision on every rank
        t_e2e = time_nl(args.e2e_steps) if registered else None
        checksum = float(np.abs(st.a["b_loc"][:, 0]).sum())
        t0 = time.perf_counter()
        for a in pinned:
            gpu.unpin(a)
        t_unreg = max_over_ranks(time.perf_counter() - t0)
        # (3) arrays allocated by the library (cudaHostAlloc): needs a changed host (C_F_POINTER)
        host_ptrs = []
        for n in list(st.a):
            st.a[n], p = gpu.host_alloc_like(st.a[n])
            host_ptrs.append(p)
        t_alloc = time_nl(args.e2e_steps)
        assert float(np.abs(st.a["b_loc"][:, 0]).sum()) == checksum
        api = ("cloudsc2_gpu_nl on caller-allocated (malloc) arrays in the reference layout, page-locked "
               "once with cloudsc2_gpu_host_register")
        if not registered:                       # fall back to the library-allocated arrays for the headline
            t_e2e = t_alloc
            api = ("cloudsc2_gpu_nl on arrays from cloudsc2_gpu_host_alloc -- cloudsc2_gpu_host_register failed "
                   f"on this host: {reg_error}")
        del st
        for p in host_ptrs:
            gpu.host_free(p)
        e2e = {"value": ngp_total / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e,
               "api": api, "host_registered": registered,
               "host_register_ms_once": 1e3 * t_reg, "host_unregister_ms_once": 1e3 * t_unreg,
               "h2d_gbs_per_gpu": h2d / t_e2e / 1e9, "d2h_gbs_per_gpu": d2h / t_e2e / 1e9,
               "pageable": {"value": ngp_total / t_page, "ms_per_step": 1e3 * t_page,
                            "note": "same call, arrays neither registered nor library-allocated"},
               "host_alloc": {"value": ngp_total / t_alloc, "ms_per_step": 1e3 * t_alloc,
                              "note": "arrays from cloudsc2_gpu_host_alloc (cudaHostAlloc)"},
               "numa": numa, "checksum_tend_t": checksum}

        # (4) what the dwarf itself does (dwarf_cloudsc.F90:84-122): LOAD the un-expanded columns, expand on the
        # device, CLOUDSC_DRIVER, VALIDATE on the device -- one call, nothing but the source / reference columns
        # (4 + 2.6 MB) and 50 statistics cross PCIe.  Under torchrun the ranks of the library communicator shard
        # the GLOBAL problem and the statistics are all-reduced by the library.
        try:
            st1 = pkg.ArrayState(src, src.klon, src.klon)
            gpu.nl(st1)
            ref = {"plude": st1.a["plude"][0], "pcovptot": st1.a["pcovptot"][0], "pfplsl": st1.a["pfplsl"][0],
                   "pfplsn": st1.a["pfplsn"][0], "pfhpsl": st1.a["pfhpsl"][0], "pfhpsn": st1.a["pfhpsn"][0],
                   "tend_loc": st1.a["b_loc"][0]}
            ds.free()                                            # make room: the resident state is a second copy
            ds = None
            gpu.nl_source(src, nproma, ngp_total, ref)
            barrier()
            t0 = time.perf_counter()
            tk_sum = 0.0
            for _ in range(args.e2e_steps):
                stats, tk, tt = gpu.nl_source(src, nproma, ngp_total, ref)
                tk_sum += tk
            t_src = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
            gpu.state_free()
            src_bytes = 8 * sum(int(np.asarray(v).size) for v in src.f.values())
            ref_bytes = 8 * sum(int(np.asarray(v).size) for v in ref.values())
            e2e_source = {"value": ngp_total / t_src, "unit": UNIT, "ms_per_step": 1e3 * t_src,
                          "kernel_ms_per_step": 1e3 * tk_sum / args.e2e_steps,
                          "h2d_bytes_per_step": src_bytes + ref_bytes, "d2h_bytes_per_step": 8 * int(stats.size),
                          "api": "cloudsc2_gpu_nl_source: 100 un-expanded source columns in, device-side expansion "
                                 "(expand_mod.F90:270-335), NL, device-side validation statistics out",
                          "max_abs_err_vs_unexpanded_columns": float(stats[:, 2].max()),
                          "stats_finite": bool(np.isfinite(stats).all())}
        except Exception as e:                       # evidence only: never lose the bench line over it
            e2e_source = {"error": repr(e)}
            try:
                gpu.state_free()
            except Exception:
                pass
        if ds is not None:
            ds.free()
        ds = pkg.DeviceState.from_source(gpu, src, nproma, ngp, gcol0=sh.gcol0, stream=stream)
        torch.cuda.synchronize()

    total_launches = int(sum_over_ranks(float(launches)))

    # ---- the path's only collective, on the real interconnect: block-sharded Taylor and adjoint
    # self-tests (BASELINE configs 2/3 scaled to 6400 columns per rank); the norms are all-reduced (MAX) over
    # the ranks INSIDE the library by its own NCCL communicator (cloudsc_driver_tl_mod.F90:125,
    # cloudsc_driver_ad_mod.F90:107) -- every rank gets the global ZNORMG back from the call
    selftests = None
    try:
        t0 = time.perf_counter()
        shs = pkg.shard_blocks(6400 * world, nproma, rank, world)
        st_s = pkg.ArrayState(src, nproma, shs.ngptot, gcol0=shs.gcol0)
        z, _ = gpu.tl_taylor(st_s)
        pen, istart = pkg.taylor_verdict(z)
        gpu.set_option("lregcl", 1)                  # the AD program's switch (cloudsc2_ad/dwarf_cloudsc.F90:105)
        zn, _ = gpu.ad_test(st_s)
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
        t_ad_drv = host_timed(lambda: gpu.ad_test(ds, src.ptsphy))
        gpu.set_option("lregcl", 0)
        t_tl_drv = host_timed(lambda: gpu.tl_taylor(ds, src.ptsphy))
        results["tl_taylor_driver"] = {"columns_per_s": ngp_total / t_tl_drv, "ms_per_step": 1e3 * t_tl_drv,
                                       "note": "dwarf-cloudsc2-tl's timed loop as one call: 1 NL + 1 TL + 10 perturbed "
                                               "NL sweeps + ERROR_NORM, max over blocks, all-reduced over the ranks"}
        results["ad_test_driver"] = {"columns_per_s": ngp_total / t_ad_drv, "ms_per_step": 1e3 * t_ad_drv,
                                     "note": "dwarf-cloudsc2-ad's timed loop as one call: 1 TL + 1 AD + dot products, "
                                             "ZNORMG all-reduced over the ranks"}
        # cost of the collective itself: the library's MAX all-reduce of ten device-resident norms, 20 calls
        dz = gpu.malloc(80)
        gpu.h2d(dz, np.ascontiguousarray(z))
        gpu.allreduce_dev(dz, 10, "max")
        ta = time.perf_counter()
        for _ in range(20):
            gpu.allreduce_dev(dz, 10, "max")
        allreduce_us = (time.perf_counter() - ta) / 20 * 1e6
        gpu.free(dz)
        selftests = {"ngptot_total": 6400 * world, "taylor_penalty": pen, "taylor_passed": 0 <= pen <= 5,
                     "allreduce_us_per_call": allreduce_us,
                     "taylor_ratios": [float(v) for v in z], "adjoint_znormg_eps": zn,
                     "adjoint_passed": bool(pkg.adjoint_verdict(zn)),
                     "allreduce": (f"library NCCL communicator (ncclAllReduce max, NCCL {nccl_version}), "
                                   f"rank {comm_rank} of {comm_size}") if comm_size > 1 else "single rank",
                     "comm_size": comm_size,
                     "seconds": time.perf_counter() - t0}
    except Exception as e:                          # evidence only: never lose the bench line over it
        selftests = {"error": repr(e)}

    # ---- BASELINE config 5: strong scaling, a FIXED total sharded over the ranks ---------------------
    strong_blocks = {}
    if not strong and not args.no_strong:
        ds.free()
        ds = None
        for total in (1310720, 5242880):
            if total == 5242880 and world < 2:
                continue                                        # 207 GB of state: does not fit one GPU
            key = f"strong_{total}"
            if total == ngp_total:
                strong_blocks[key] = {"note": "identical to the weak-scaling configuration of this run",
                                      "ngptot_per_gpu": ngp,
                                      **{m: {"ms_per_step": results[m]["ms_per_step"],
                                             "columns_per_s": results[m]["columns_per_s"]}
                                         for m in ("nl", "tl", "ad") if m in results and "ms_per_step" in results[m]}}
                continue
            try:
                shx = pkg.shard_blocks(total, nproma, rank, world)
                # the same decision on every rank (a rank that skipped would hang the others' barrier): sized
                # on rank 0's shard, the largest, against the device's total memory
                big = pkg.shard_blocks(total, nproma, 0, world).ngptot
                cap = 0.85 * torch.cuda.get_device_properties(dev).total_memory
                need_state = big * 36.1 * (KLEV + 1) * 8
                need_incr = big * 26.1 * (KLEV + 1) * 8
                blk = {"ngptot_per_gpu": shx.ngptot}
                if need_state > cap:
                    blk["skipped"] = f"state of {need_state / 1e9:.0f} GB does not fit {cap / 1e9:.0f} GB"
                    strong_blocks[key] = blk
                    continue
                dx = pkg.DeviceState.from_source(gpu, src, nproma, shx.ngptot, gcol0=shx.gcol0, stream=stream)
                torch.cuda.synchronize()
                ms = timed(lambda: gpu.nl_dev(dx, src.ptsphy, stream=stream), 5, 3)
                blk["nl"] = {"ms_per_step": ms, "columns_per_s": total / (ms * 1e-3),
                             "frac_of_hbm": NL_BYTES_PER_COL * shx.ngptot / (ms * 1e-3) / 1e9 / peak}
                if ("tl" in modes or "ad" in modes) and need_state + need_incr < cap:
                    ax, bx, m2, m2h = make_increments(dx)
                    if "tl" in modes:
                        ms = timed(lambda: gpu.tl_dev(dx, src.ptsphy, ax, bx, stream=stream), 3, 3)
                        blk["tl"] = {"ms_per_step": ms, "columns_per_s": total / (ms * 1e-3),
                                     "frac_of_hbm": TL_BYTES_PER_COL * shx.ngptot / (ms * 1e-3) / 1e9 / peak}
                    if "ad" in modes:
                        fill_output_adjoints(bx, m2, m2h)
                        ms = timed(lambda: gpu.ad_dev(dx, src.ptsphy, ax, bx, stream=stream), 3, 3)
                        blk["ad"] = {"ms_per_step": ms, "columns_per_s": total / (ms * 1e-3),
                                     "frac_of_hbm": AD_BYTES_PER_COL * shx.ngptot / (ms * 1e-3) / 1e9 / peak}
                    for q in list(ax.values()) + list(bx.values()):
                        gpu.free(q)
                elif "tl" in modes or "ad" in modes:
                    blk["tl_ad_skipped"] = f"increments of {need_incr / 1e9:.0f} GB do not fit beside the state"
                dx.free()
                strong_blocks[key] = blk
            except Exception as e:
                strong_blocks[key] = {"error": repr(e)}

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        ob, cpu_kind, cpu_desc = load_cpu_arm()
        threads = host_threads()
        sample = min(ngp, 65536)
        cps, _ = cpu_reference_run(pkg, ob, src, prm, nproma, sample, threads, repeats=3)
        cps4, _ = cpu_reference_run(pkg, ob, src, prm, 32, min(sample, 32768), min(4, threads), repeats=2)
        cpu = {"value": cps, "unit": UNIT, "cores": threads, "kind": cpu_kind,
               "sample": f"{sample} columns, NPROMA={nproma}, best of 3, block loop only; {cpu_desc}",
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
                                             "163 840 columns scaled by NGPTOT (profiles/r2e_nl_ncu.md)",
                             "peak_source": peak_src, "kernel": "k_cloudsc2_nl",
                             "algorithmic_bytes_per_column": NL_BYTES_PER_COL,
                             "algorithmic_bytes_per_launch": NL_BYTES_PER_COL * ngp,
                             "fp64_pipe_pct_ncu": NCU["nl"]["fp64_pipe_pct"]},
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": total_launches, "clocks": clocks,
                "e2e_source": e2e_source, **strong_blocks,
                "modes": results, "nproma_sweep_ngptot160000": sweep, "selftests": selftests}
        # the reference's own report lines (timer_mod.F90:114-174), on stderr
        sys.stderr.write(pkg.report.performance_table(1, ngp_total, pkg.nblocks(ngp_total, nproma), nproma,
                                                      ms_nl * 1e-3, numproc=world) + "\n")
        print(json.dumps(line))
    if ds is not None:
        ds.free()
    gpu.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
