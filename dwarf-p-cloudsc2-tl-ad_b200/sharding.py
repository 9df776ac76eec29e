"""Block sharding over ranks and the only collective of this path.

The reference decomposes NGPTOTG global columns into one contiguous range per MPI rank
(cloudsc2_nl/dwarf_cloudsc.F90:65-69) and every rank re-expands the same 100 source columns
(expand_mod.F90:30-46); ranks never talk during compute.  Here: one process per GPU, contiguous
ranges of whole NPROMA blocks per rank, and a single all-reduce of the scalar norms -- MAX over
the ten Taylor ratios (reduction(max:znormg), cloudsc_driver_tl_mod.F90:125) or over the adjoint
ZNORMG (cloudsc_driver_ad_mod.F90:107), SUM/MIN/MAX for validation statistics
(validate_mod.F90:197-199) -- through torch.distributed (NCCL over NVLink on GPUs, gloo in the
CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    block0: int        # first global block of this rank
    nblocks: int       # blocks owned
    gcol0: int         # first global column
    ngptot: int        # columns owned (the globally last block may be ragged)


def shard_blocks(ngptot_global: int, nproma: int, rank: int, world: int) -> Shard:
    """Contiguous block ranges: rank r owns blocks [r*ceil(NB/R), min(NB,(r+1)*ceil(NB/R)))."""
    nb = ngptot_global // nproma + min(ngptot_global % nproma, 1)
    per = -(-nb // world)
    b0 = min(nb, rank * per)
    b1 = min(nb, (rank + 1) * per)
    g0 = b0 * nproma
    g1 = min(ngptot_global, b1 * nproma)
    return Shard(rank, world, b0, b1 - b0, g0, max(0, g1 - g0))


def allreduce_norms(values, op: str = "max", device=None):
    """All-reduce a handful of FP64 scalars over the default process group (no-op without one)."""
    import torch
    import torch.distributed as dist
    v = np.atleast_1d(np.asarray(values, dtype=np.float64)).copy()
    if op == "max":
        # MAX over ranks drops NaN (std::max in gloo, no guarantee in NCCL): a rank whose kernels produced
        # non-finite norms must fail the test for everybody, so it contributes a huge sentinel instead --
        # the same mapping the library applies on the device before its own all-reduce (k_norms_prepare)
        v[~np.isfinite(v)] = 1.0e300
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return v
    t = torch.from_numpy(v.copy())
    if device is not None:
        t = t.to(device)
    rop = {"max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM}[op]
    dist.all_reduce(t, op=rop)
    return t.cpu().numpy()


def sharded_taylor(gpu, src, nproma: int, ngptot_global: int, rank: int, world: int, device=None):
    """Taylor test of a block-sharded problem: local test on this rank's blocks, then MAX
    all-reduce of ZNORMG(10).  Returns (znormg_global, shard)."""
    from .state import ArrayState
    sh = shard_blocks(ngptot_global, nproma, rank, world)
    z = np.zeros(10)
    if sh.ngptot > 0:
        st = ArrayState(src, nproma, sh.ngptot, gcol0=sh.gcol0)
        z, _ = gpu.tl_taylor(st)
    return allreduce_norms(z, "max", device), sh


def sharded_adjoint(gpu, src, nproma: int, ngptot_global: int, rank: int, world: int, device=None):
    from .state import ArrayState
    sh = shard_blocks(ngptot_global, nproma, rank, world)
    zn = 0.0
    if sh.ngptot > 0:
        st = ArrayState(src, nproma, sh.ngptot, gcol0=sh.gcol0)
        zn, _ = gpu.ad_test(st)
    return float(allreduce_norms([zn], "max", device)[0]), sh


def allreduce_validation(stats, device=None):
    """Combine per-rank validation statistics [min, max, max|err|, sum|err|, sum|ref|] the way the
    reference does (CLOUDSC_MPI_REDUCE_MIN / _MAX / _SUM, validate_mod.F90:197-199)."""
    s = np.asarray(stats, dtype=np.float64)
    out = np.empty(5)
    out[0] = allreduce_norms(s[0:1], "min", device)[0]
    out[1:3] = allreduce_norms(s[1:3], "max", device)
    out[3:5] = allreduce_norms(s[3:5], "sum", device)
    return out
