"""GPU suite, rows 8f-2 / 8f-4 of the scope table: device-side validation statistics and the
drop-in Python entry points named like the reference's own (satur / cloudsc2_py)."""
from types import SimpleNamespace

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _namespaces(prm, ceta):
    g = lambda *names: SimpleNamespace(**{n: getattr(prm, n) for n in names})
    yrmcst = g("rg", "rd", "rcpd", "retv", "rlvtt", "rlstt", "rlmlt", "rtt")
    yrethf = g("r2es", "r3les", "r3ies", "r4les", "r4ies", "r5les", "r5ies", "r5alvcp", "r5alscp",
               "ralvdcp", "ralsdcp", "rtwat", "rtice", "rtwat_rtice_r", "rvtmp2")
    yrecldp = g("rclcrit", "rkconv", "rlmin", "rpecons")
    yrephli = SimpleNamespace(lphylin=True, rlptrc=prm.rlptrc)
    yrecld = SimpleNamespace(ceta=np.asarray(ceta))
    return yrmcst, yrethf, yrecldp, yrephli, yrecld


def test_python_entry_points_reproduce_reference_python_kernel(pkg, golden):
    """pkg.pyapi.satur / cloudsc2_py called exactly like the reference's cloudsc2_py.py was called to
    make the golden vectors (tests/golden/make_golden.py) must reproduce its outputs."""
    prm = pkg.default_params()
    yrmcst, yrethf, yrecldp, yrephli, yrecld = _namespaces(prm, golden["ceta"])
    x = {k[3:]: np.ascontiguousarray(golden[k]) for k in golden.files if k.startswith("in_")}
    klev, klon = x["ptm1"].shape
    pqs = np.zeros((klev, klon))
    pkg.pyapi.satur(1, klon, klon, 1, klev, True, x["papp1"], x["ptm1"], pqs, 2, yrethf, yrmcst)
    assert np.abs(pqs / golden["pqs"] - 1.0).max() < 1e-14
    y = {n: np.full((klev + (1 if n.startswith("pf") else 0), klon), 9.0)
         for n in ("ptent", "ptenq", "ptenl", "pteni", "pclc", "pfplsl", "pfplsn", "pfhpsl",
                   "pfhpsn", "pcovptot")}
    plu_pad = np.vstack([x["plu"], np.zeros((1, klon))])      # as handed to the reference kernel
    pkg.pyapi.cloudsc2_py(1, klon, klon, 1, klev, False, float(golden["ptsphy"]), x["paphp1"], x["papp1"],
                          x["pqm1"], golden["pqs"], x["ptm1"], x["pl"], x["pi"], x["plude"], plu_pad,
                          x["pmfu"], x["pmfd"], y["ptent"], x["pgtent"], y["ptenq"], x["pgtenq"],
                          y["ptenl"], x["pgtenl"], y["pteni"], x["pgteni"], x["psupsat"], y["pclc"],
                          y["pfplsl"], y["pfplsn"], y["pfhpsl"], y["pfhpsn"], y["pcovptot"],
                          yrecldp, yrecld, yrmcst, yrethf, yrephli)
    for n, v in y.items():
        r = golden["out_" + n]
        assert np.abs(v - r).max() <= 1e-11 * max(np.abs(r).max(), 1e-300), n
    with pytest.raises(NotImplementedError):
        pkg.pyapi.cloudsc2_py(1, klon, klon, 1, klev, True, 3600.0, *([None] * 31))


def test_satur_entry_matches_oracle(pkg, ob, src100, gpu_nl):
    pap, pt = src100.f["pap"], src100.f["pt"]
    got = gpu_nl.satur(pap, pt)
    want = ob.satur(gpu_nl.params, np.ascontiguousarray(pap), np.ascontiguousarray(pt))
    assert np.abs(got / want - 1.0).max() < 1e-14


@pytest.mark.parametrize("name,nproma,ngptot,gcol0", [("pt", 32, 1000, 0), ("paph", 128, 5000, 0),
                                                       ("pclv", 64, 777, 0), ("tend_cml", 16, 100, 0),
                                                       ("pq", 32, 333, 4160)])
def test_device_validation_matches_host_validate(pkg, src100, gpu_nl, name, nproma, ngptot, gcol0):
    """cloudsc2_gpu_validate_dev (reference columns read through the cyclic map on the device) ==
    VALIDATE_R2/R3 on the host against an expanded reference (validate_mod.F90:165-261)."""
    src = np.ascontiguousarray(src100.f[name])
    nlev = src.shape[-2]
    ndim = src.size // (nlev * 100)
    ref = pkg.expand(src, nproma, ngptot, gcol0=gcol0)              # (NB, [ndim,] nlev, nproma)
    rng = np.random.default_rng(5)
    fld = ref * (1.0 + 1e-6 * rng.standard_normal(ref.shape))
    fld.reshape(-1)[::97] = 0.0
    dsrc, dfld = gpu_nl.malloc(src.nbytes), gpu_nl.malloc(fld.nbytes)
    try:
        gpu_nl.h2d(dsrc, src)
        gpu_nl.h2d(dfld, fld)
        got = gpu_nl.validate_dev(dsrc, 100, dfld, nproma, nlev, ndim, ngptot, gcol0=gcol0)
    finally:
        gpu_nl.free(dsrc)
        gpu_nl.free(dfld)
    nb = ref.shape[0]
    h = pkg.validate(ref.reshape(nb, ndim * nlev, nproma), fld.reshape(nb, ndim * nlev, nproma), ngptot)
    assert got[0] == h["min"] and got[1] == h["max"] and got[2] == h["max_abs_err"]
    assert np.isclose(got[3], h["sum_abs_err"], rtol=1e-12) and np.isclose(got[4], h["sum_abs_ref"], rtol=1e-12)
    line = pkg.report.error_print(name.upper(), got, ngptot, ndim=2 if ndim == 1 else 3)
    assert line.endswith("!!!!")                                      # 1e-6 relative error is flagged


def test_nl_results_validate_clean_against_a_second_run(pkg, src100, gpu_nl):
    """End-to-end use of the device validation: NL outputs of NGPTOT = 4000 columns against the
    100-column outputs used as 'reference.h5' -- exactly zero error because expansion is cyclic."""
    small = pkg.ArrayState(src100, 100, 100)
    gpu_nl.nl(small)
    big = pkg.ArrayState(src100, 32, 4000)
    gpu_nl.nl(big)
    for n, nlev in (("pa", 137), ("pfplsl", 138), ("pfhpsn", 138)):
        ref_src = np.ascontiguousarray(small.a[n][0])                  # (nlev, 100)
        dsrc, dfld = gpu_nl.malloc(ref_src.nbytes), gpu_nl.malloc(big.a[n].nbytes)
        try:
            gpu_nl.h2d(dsrc, ref_src)
            gpu_nl.h2d(dfld, big.a[n])
            st = gpu_nl.validate_dev(dsrc, 100, dfld, 32, nlev, 1, 4000)
        finally:
            gpu_nl.free(dsrc)
            gpu_nl.free(dfld)
        assert st[2] == 0.0 and st[3] == 0.0 and st[4] > 0.0, n
        assert pkg.report.error_print(n, st, 4000).split()[1] in ("2D1",)


def _validate_outputs_against_small_run(pkg, gpu, ds, small, ngptot, nproma, names):
    for n, nlev in names:
        ref_src = np.ascontiguousarray(small.a[n][0])
        dsrc = gpu.malloc(ref_src.nbytes)
        try:
            gpu.h2d(dsrc, ref_src)
            st = gpu.validate_dev(dsrc, 100, ds.ptr[n], nproma, nlev, 1, ngptot)
        finally:
            gpu.free(dsrc)
        assert st[2] == 0.0 and st[3] == 0.0 and st[4] > 0.0, (n, st)
        assert st[0] == small.a[n][0].min() and st[1] == small.a[n][0].max(), n


def test_nl_scaling_size_on_device(pkg, src100, gpu_nl):
    """BASELINE config 5 size: NGPTOT = 1 310 720, NPROMA = 128, everything device-resident (device
    expansion, kernel, device validation).  Size-independent property: the expansion is cyclic, so
    every output must equal the 100-column run bit for bit -> zero validation error."""
    nproma, ngptot = 128, 1310720
    small = pkg.ArrayState(src100, 100, 100)
    gpu_nl.nl(small)
    ds = pkg.DeviceState.from_source(gpu_nl, src100, nproma, ngptot)
    try:
        gpu_nl.nl_dev(ds, src100.ptsphy)
        gpu_nl.sync()
        _validate_outputs_against_small_run(pkg, gpu_nl, ds, small, ngptot, nproma,
                                            (("pa", 137), ("pfplsl", 138), ("pfplsn", 138),
                                             ("pfhpsl", 138), ("pfhpsn", 138), ("pcovptot", 137)) [:5])
        # the four tendency slabs of B_LOC: slab s of block b sits at (b*8 + s) * 137 * 128
        loc = np.empty((ds.nblocks, 8, 137, nproma))
        # only the first 4 blocks are brought back (1280 x 8 slabs would be 11.5 GB)
        part = np.empty((4, 8, 137, nproma))
        gpu_nl.d2h(part, ds.ptr["b_loc"])
        del loc
        cols = np.arange(4 * nproma) % 100
        for slab in (0, 2, 3, 4):
            got = np.ascontiguousarray(np.transpose(part[:, slab], (1, 0, 2))).reshape(137, -1)
            assert np.array_equal(got, small.a["b_loc"][0, slab][:, cols]), slab
    finally:
        ds.free()


def test_taylor_and_adjoint_tests_at_throughput_size(pkg, src100):
    """The reference's two self-tests at NGPTOT = 163 840 / NPROMA = 128 (bench size), device
    resident: same verdicts, and -- the columns being cyclic copies -- the same ZNORMG as the
    100-column problems, exactly (max over identical per-block / per-column values)."""
    nproma, ngptot = 128, 163840
    with pkg.Cloudsc2(pkg.default_params(lregcl=False), 137, src100.ceta) as gpu:
        ds = pkg.DeviceState.from_source(gpu, src100, nproma, ngptot)
        try:
            z, rb = gpu.tl_taylor(ds, src100.ptsphy)
        finally:
            ds.free()
        pen, istart = pkg.taylor_verdict(z)
        assert 0 <= pen <= 5, (pen, z)
        # blocks repeat with period lcm(128, 100) / 128 = 25 blocks
        assert np.array_equal(rb[:25], rb[25:50]) and np.array_equal(rb[:25], rb[1250:1275])
    with pkg.Cloudsc2(pkg.default_params(lregcl=True), 137, src100.ceta) as gpu:
        ds = pkg.DeviceState.from_source(gpu, src100, nproma, ngptot)
        try:
            zn, nc = gpu.ad_test(ds, src100.ptsphy)
        finally:
            ds.free()
        assert pkg.adjoint_verdict(zn), zn
        small = pkg.ArrayState(src100, 100, 100)
        zn100, nc100 = gpu.ad_test(small)
        assert zn == zn100
        assert np.array_equal(nc[:100], nc100) and np.array_equal(nc[100:200], nc100)
