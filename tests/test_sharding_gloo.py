"""CPU suite: the N > 1 path -- contiguous block sharding (dwarf_cloudsc.F90:65-69 arithmetic), shard
expansion with a global column offset, and the only collective of this path (MAX all-reduce of
the Taylor / adjoint norms, reduction(max:znormg) in cloudsc_driver_tl_mod.F90:125 and
cloudsc_driver_ad_mod.F90:107) -- with world_size = 2 over gloo.  The compute backend of each
rank is the CPU oracle here (no GPU in this container); on the B200 box the same helpers run with
the CUDA backend (tests/test_gpu_sharding.py)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class OracleBackend:
    """Duck-types the two Cloudsc2 methods the sharding helpers call, on top of the oracle."""

    def __init__(self, ob, prm, ceta):
        self.ob, self.prm, self.ceta = ob, prm, ceta

    def tl_taylor(self, st):
        z, rb, _ = self.ob.driver_tl(self.prm, self.ceta, st, numomp=1)
        return z, rb

    def ad_test(self, st):
        zn, nc, _ = self.ob.driver_ad(self.prm, self.ceta, st, numomp=1)
        return zn, nc


def _worker(rank, world, port, nproma, ngptot, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import oracle_binding as ob
        pkg = ob.pkg
        src = pkg.synth_source(seed=0, klon=100, klev=137)
        be = OracleBackend(ob, pkg.default_params(lregcl=False), src.ceta)
        z, sh = pkg.sharded_taylor(be, src, nproma, ngptot, rank, world)
        be_ad = OracleBackend(ob, pkg.default_params(lregcl=True), src.ceta)
        zn, _ = pkg.sharded_adjoint(be_ad, src, nproma, ngptot, rank, world)
        s = pkg.allreduce_norms([float(sh.ngptot), float(sh.nblocks)], "sum")
        # validation statistics of this rank's shard of PT against the source columns, combined as
        # validate_mod.F90:197-199 does (min / max / sum over ranks)
        st = pkg.ArrayState(src, nproma, max(sh.ngptot, 1), gcol0=sh.gcol0)
        v = pkg.validate(st.a["pt"] + 1e-3 * (rank + 1), st.a["pt"], max(sh.ngptot, 1))
        loc = [v["min"], v["max"], v["max_abs_err"], v["sum_abs_err"], v["sum_abs_ref"]]
        if sh.ngptot == 0:
            loc = [np.inf, -np.inf, 0.0, 0.0, 0.0]
        vs = pkg.allreduce_validation(loc)
        q.put((rank, z, zn, s, sh, vs, loc))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nproma,ngptot", [(16, 200), (32, 100)])
def test_sharded_tests_world2_gloo(pkg, ob, src100, nproma, ngptot):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nproma, ngptot, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process truth over ALL blocks
    st = pkg.ArrayState(src100, nproma, ngptot)
    z_all, _, _ = ob.driver_tl(pkg.default_params(lregcl=False), src100.ceta, st, numomp=2)
    zn_all, _, _ = ob.driver_ad(pkg.default_params(lregcl=True), src100.ceta,
                                pkg.ArrayState(src100, nproma, ngptot), numomp=2)
    for rank, z, zn, s, sh, vs, loc in res:
        assert np.array_equal(z, z_all)              # max over blocks is order independent: exact
        assert zn == zn_all
        assert s[0] == ngptot and s[1] == st.nblocks  # shards tile the problem
    assert res[0][4].gcol0 == 0 and res[1][4].gcol0 == res[0][4].ngptot
    locs = np.array([r[6] for r in res])
    want = [locs[:, 0].min(), locs[:, 1].max(), locs[:, 2].max(), locs[:, 3].sum(), locs[:, 4].sum()]
    for r in res:
        assert np.allclose(r[5], want, rtol=1e-15, atol=0)


def test_shard_blocks_tiles_every_problem(pkg):
    for ngptot, nproma, world in [(100, 32, 2), (1310720, 128, 8), (163840, 128, 4), (5, 8, 4),
                                  (1000, 64, 3), (160000, 32, 8)]:
        nb = pkg.nblocks(ngptot, nproma)
        shards = [pkg.shard_blocks(ngptot, nproma, r, world) for r in range(world)]
        assert sum(s.nblocks for s in shards) == nb
        assert sum(s.ngptot for s in shards) == ngptot
        nxt_b, nxt_g = 0, 0
        for s in shards:
            assert s.block0 == (nxt_b if s.nblocks else s.block0)
            if s.nblocks:
                assert s.gcol0 == nxt_g == s.block0 * nproma
                nxt_b += s.nblocks
                nxt_g += s.ngptot
        # only the globally last block may be ragged
        for s in shards[:-1]:
            if s.nblocks and s.block0 + s.nblocks < nb:
                assert s.ngptot == s.nblocks * nproma


def test_shard_expansion_offset(pkg, src100):
    """A shard's local column j is global column gcol0 + j (expand_mod.F90 closed form)."""
    whole = pkg.ArrayState(src100, 16, 200)
    sh = pkg.shard_blocks(200, 16, 1, 2)
    part = pkg.ArrayState(src100, 16, sh.ngptot, gcol0=sh.gcol0)
    for n in ("pt", "paph", "pclv", "b_cml"):
        assert np.array_equal(part.a[n], whole.a[n][sh.block0:sh.block0 + sh.nblocks]), n
    assert np.allclose(pkg.allreduce_norms([1.0, 2.0]), [1.0, 2.0])   # no process group: identity


def _nan_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import oracle_binding as ob
        z = np.ones(10)
        if rank == 1:
            z[3] = np.nan          # a rank whose kernels produced NaN ratios
            z[7] = np.inf
        q.put((rank, ob.pkg.allreduce_norms(z, "max")))
    finally:
        dist.destroy_process_group()


def test_max_allreduce_does_not_drop_a_nan_rank(pkg):
    """std::max(z, NaN) returns z, so a plain MAX over ranks would print TEST PASSED from rank 0
    alone.  Non-finite norms travel as a huge sentinel and fail the verdict on every rank."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nan_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, z in res:
        assert z[3] >= 1e300 and z[7] >= 1e300 and z[0] == 1.0
        pen, _ = pkg.taylor_verdict(z)
        assert not (0 <= pen <= 5)
