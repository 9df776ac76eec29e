"""GPU suite, adjoint path and the adjoint (dot-product) test.

Tolerances: AD fields |gpu - oracle| <= AD_RTOL * max|field| per field (AD_RTOL = 1e-8: the adjoint
chains many divisions by trajectory quantities; measured ~1e-12); adjoint symmetry to the
reference's own tolerance: |N1 - N2| / eps / N2 < 10000 (cloudsc_driver_ad_mod.F90:286-294).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
AD_RTOL = 1e-8

G2O_IN = {"paph": "paphp1", "pap": "papp1", "pq": "pqm1", "pqs": "pqs", "pt": "ptm1", "pl": "pl",
          "pi": "pi", "plude": "plude", "plu": "plu", "pmfu": "pmfu", "pmfd": "pmfd",
          "gtent": "pgtent", "gtenq": "pgtenq", "gtenl": "pgtenl", "gteni": "pgteni",
          "psupsat": "psupsat"}
G2O_OUT = {"tent": "ptent", "tenq": "ptenq", "tenl": "ptenl", "teni": "pteni", "pclc": "pclc",
           "pfplsl": "pfplsl", "pfplsn": "pfplsn", "pfhpsl": "pfhpsl", "pfhpsn": "pfhpsn",
           "pcovptot": "pcovptot"}


@pytest.mark.parametrize("nproma,ngptot,lregcl", [(100, 100, True), (32, 100, True), (1, 40, False),
                                                  (64, 200, False), (128, 2048, True), (256, 1000, False)])
def test_ad_fields_match_oracle(pkg, ob, src100, nproma, ngptot, lregcl):
    prm = pkg.default_params(lregcl=lregcl)
    st = pkg.ArrayState(src100, nproma, ngptot)
    nb = st.nblocks
    rng = np.random.default_rng(7)
    din, dout = pkg.driver.alloc_increments(nb, 137, nproma)
    # non-zero input adjoints on entry (they must be ACCUMULATED), random output adjoints
    for n, v in din.items():
        v[:] = 1e-3 * rng.standard_normal(v.shape)
    scale_out = {"tent": 1e3, "tenq": 1e6, "tenl": 1e7, "teni": 1e7, "pclc": 1.0, "pfplsl": 1e4,
                 "pfplsn": 1e4, "pfhpsl": 1e-2, "pfhpsn": 1e-2, "pcovptot": 1.0}
    for n, v in dout.items():
        v[:] = scale_out[n] * rng.standard_normal(v.shape)
    din0 = {k: v.copy() for k, v in din.items()}
    dout0 = {k: v.copy() for k, v in dout.items()}
    with pkg.Cloudsc2(prm, 137, src100.ceta) as gpu:
        gpu.ad(st, din, dout)
    traj = {"ptent": st.a["b_loc"][:, 0], "ptenq": st.a["b_loc"][:, 2], "pclc": st.a["pa"],
            "pfplsn": st.a["pfplsn"], "pfhpsl": st.a["pfhpsl"]}
    for b in range(nb):
        icend = min(nproma, ngptot - b * nproma)
        x5 = {k: np.ascontiguousarray(v[:, :icend]) for k, v in ob.block_inputs(st, b, prm).items()}
        xa = {o: din0[gname][b][:, :icend].copy() for gname, o in G2O_IN.items()}
        ya = {o: dout0[gname][b][:, :icend].copy() for gname, o in G2O_OUT.items()}
        y5 = ob.cloudsc2ad_block(prm, src100.ceta, st.ptsphy, x5, xa, ya)
        for gname, o in G2O_IN.items():
            ref, got = xa[o], din[gname][b][:, :icend]
            # compare the increments the adjoint added, scaled by their own magnitude
            inc_ref = ref - (0.0 if gname == "psupsat" else din0[gname][b][:, :icend])
            inc_got = got - (0.0 if gname == "psupsat" else din0[gname][b][:, :icend])
            scale = max(np.abs(inc_ref).max(), 1e-300)
            err = np.abs(inc_got - inc_ref).max() / scale
            assert err <= AD_RTOL, (gname, b, err)
        for gname, o in G2O_OUT.items():
            assert not dout[gname][b][:, :icend].any(), gname            # consumed and zeroed
            assert not ya[o].any()
            if icend < nproma:                                           # padding untouched
                assert np.array_equal(dout[gname][b][:, icend:], dout0[gname][b][:, icend:])
        for o, arr in traj.items():                                      # trajectory outputs (:842-864)
            scale = max(np.abs(y5[o]).max(), 1e-300)
            assert np.abs(arr[b][:, :icend] - y5[o]).max() <= 1e-11 * scale, o


@pytest.mark.parametrize("lregcl", [False, True])
def test_dot_product_identity_random(pkg, src100, lregcl):
    """<M' dx, y> == <dx, M'^T y> for random dx, y through the full-field TL and AD kernels."""
    prm = pkg.default_params(lregcl=lregcl)
    nproma, ngptot = 64, 256
    st = pkg.ArrayState(src100, nproma, ngptot)
    nb = st.nblocks
    rng = np.random.default_rng(11)
    a = st.a
    base = {"paph": a["paph"], "pap": a["pap"], "pq": a["pq"], "pqs": 1e-3 * np.ones_like(a["pq"]),
            "pt": a["pt"], "pl": a["pclv"][:, 0], "pi": a["pclv"][:, 1], "plude": a["plude"],
            "plu": a["plu"], "pmfu": a["pmfu"], "pmfd": a["pmfd"], "gtent": a["b_cml"][:, 0],
            "gtenq": a["b_cml"][:, 2], "gtenl": a["b_cml"][:, 3], "gteni": a["b_cml"][:, 4],
            "psupsat": a["psupsat"]}
    dx = {k: np.ascontiguousarray(0.01 * v * rng.standard_normal(v.shape)) for k, v in base.items()}
    # The reference's adjoint ASSIGNS PSUPSAT = PTSPHY*ZQP1 (cloudsc2ad.F90:1733, an extra PTSPHY
    # and no accumulation), which is not the transpose of the TL statement (:346); its own test
    # sets ZSUPSAT = 0 ("obsolete, better not use", cloudsc_driver_ad_mod.F90:139). Same here.
    dx["psupsat"][:] = 0.0
    with pkg.Cloudsc2(prm, 137, src100.ceta) as gpu:
        _, tl_out = pkg.driver.alloc_increments(nb, 137, nproma)
        gpu.tl(st, dx, tl_out)
        y = {k: np.ascontiguousarray(v * rng.uniform(0.5, 1.5, v.shape)) for k, v in tl_out.items()}
        lhs = sum(float((tl_out[k] * y[k]).sum()) for k in y)
        adj, _ = pkg.driver.alloc_increments(nb, 137, nproma)
        ycopy = {k: v.copy() for k, v in y.items()}
        gpu.ad(st, adj, ycopy)
    rhs = sum(float((dx[k] * adj[k]).sum()) for k in dx)
    assert lhs != 0.0
    assert abs(lhs - rhs) <= 1e-10 * abs(lhs), (lhs, rhs)


def test_adjoint_test_config3(pkg, ob, src100):
    """BASELINE config 3: dwarf-cloudsc2-ad 1 100 100 (LREGCL=.TRUE.)."""
    prm = pkg.default_params(lregcl=True)
    st = pkg.ArrayState(src100, nproma=100, ngptot=100)
    ref = pkg.ArrayState(src100, nproma=100, ngptot=100)
    with pkg.Cloudsc2(prm, 137, src100.ceta) as gpu:
        zn, nc = gpu.ad_test(st)
    zo, nco, _ = ob.driver_ad(prm, src100.ceta, ref)
    assert pkg.adjoint_verdict(zn) and zn < 10000.0, zn
    assert nc.shape == (100, 3)
    assert np.allclose(nc[:, 0], nco[:, 0], rtol=1e-9)      # N1 = <y, y> per column
    assert np.allclose(nc[:, 1], nco[:, 1], rtol=1e-9)      # N2 = <dx, AD y>
    assert zn == pytest.approx(nc[:, 2].max())
    assert (nc[:, 2] < 10000.0).all()
    assert ob.lib().orc_adjoint_verdict(zo) == 1


@pytest.mark.parametrize("nproma,ngptot,lregcl", [(32, 100, True), (128, 1000, False), (1, 100, True)])
def test_adjoint_test_other_blockings(pkg, src100, nproma, ngptot, lregcl):
    prm = pkg.default_params(lregcl=lregcl)
    st = pkg.ArrayState(src100, nproma, ngptot)
    with pkg.Cloudsc2(prm, 137, src100.ceta) as gpu:
        zn, nc = gpu.ad_test(st)
    assert zn < 10000.0, zn
    assert (nc[:, 0] > 0).all()
    # cyclic expansion: column g behaves exactly like column g mod 100
    if ngptot > 100:
        assert np.array_equal(nc[100:200], nc[:100])


def test_ad_with_precomputed_trajectory_equals_as_written(pkg, src100):
    """Option ad_have_trajectory: when CLOUDSC2 (or CLOUDSC2TL) has already run on the same inputs,
    cloudsc2_gpu_ad_dev restarts from the trajectory fluxes PFPLSL5 / PFPLSN5 it finds in the state
    and skips its own forward sweep.  Input adjoints must be bit-identical to the as-written path
    (trajectory recomputed inside the call, cloudsc2ad.F90:364-866)."""
    prm = pkg.default_params(lregcl=True)
    nproma, ngptot = 64, 500
    st = pkg.ArrayState(src100, nproma, ngptot)
    nb = st.nblocks
    rng = np.random.default_rng(3)
    _, seed = pkg.driver.alloc_increments(nb, 137, nproma)
    for k in seed:
        seed[k][...] = rng.standard_normal(seed[k].shape) * 1e-4

    def run(gpu, have_traj):
        ds = pkg.DeviceState(gpu, st)
        ds.zero()
        adj, _ = pkg.driver.alloc_increments(nb, 137, nproma)
        d_in = {k: gpu.malloc(v.nbytes) for k, v in adj.items()}
        d_out = {k: gpu.malloc(v.nbytes) for k, v in seed.items()}
        try:
            for k, v in adj.items():
                gpu.h2d(d_in[k], v)
            for k, v in seed.items():
                gpu.h2d(d_out[k], v)
            if have_traj:
                gpu.nl_dev(ds, st.ptsphy)                 # the trajectory run that precedes the adjoint
                gpu.set_option("ad_have_trajectory", 1)
            before = gpu.launch_count()
            gpu.ad_dev(ds, st.ptsphy, d_in, d_out)
            gpu.sync()
            launches = gpu.launch_count() - before
            for k, v in adj.items():
                gpu.d2h(v, d_in[k])
        finally:
            gpu.set_option("ad_have_trajectory", 0)
            for p in list(d_in.values()) + list(d_out.values()):
                gpu.free(p)
            ds.free()
        return adj, launches

    with pkg.Cloudsc2(prm, 137, src100.ceta) as gpu:
        ref, n_ref = run(gpu, False)
        got, n_got = run(gpu, True)
    assert (n_ref, n_got) == (2, 1)                       # forward + reverse  vs  reverse only
    for k in ref:
        assert np.array_equal(ref[k], got[k]), k
        assert np.abs(ref[k]).max() > 0 or k == "psupsat", k


def test_ad_kernel_is_the_transpose_of_the_reference_derivative(pkg, golden_fd):
    """<D, y> = <dx, M'^T y> with D = central finite differences of the REFERENCE'S Python NL kernel
    along dx = 0.01 x (tests/golden/tl_fd_pyref.npz) and M'^T y from the CUDA adjoint, called like
    CALL CLOUDSC2AD (per-block Fortran-ABI entry, PQS5 supplied): pins the GPU adjoint against
    reference code, not only against our own TL."""
    import ctypes as C
    golden, fd = golden_fd
    lib = pkg.load_library()
    x5 = {k[3:]: np.ascontiguousarray(golden[k]) for k in golden.files if k.startswith("in_")}
    x5["pqs"] = np.ascontiguousarray(golden["pqs"])
    klev, klon = x5["ptm1"].shape
    half = lambda n: klev + (1 if n == "paphp1" or n.startswith("pf") else 0)
    out10 = ("ptent", "ptenq", "ptenl", "pteni", "pclc", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn", "pcovptot")
    order26 = ("paphp1", "papp1", "pqm1", "pqs", "ptm1", "pl", "pi", "plude", "plu", "pmfu", "pmfd",
               "ptent", "pgtent", "ptenq", "pgtenq", "ptenl", "pgtenl", "pteni", "pgteni", "psupsat",
               "pclc", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn", "pcovptot")
    rng = np.random.default_rng(17)
    y = {n: np.ascontiguousarray(rng.uniform(0.5, 1.5, fd["d_" + n].shape) / max(np.abs(fd["d_" + n]).max(), 1e-30))
         for n in out10}
    lhs = sum(float((fd["d_" + n] * y[n]).sum()) for n in out10)
    traj = {**x5, **{n: np.zeros((half(n), klon)) for n in out10}}
    incr = {**{k: np.zeros_like(v) for k, v in x5.items()}, **{n: v.copy() for n, v in y.items()}}
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ref = lambda v: C.byref(C.c_int(v))
    ptsphy = float(golden["ptsphy"])
    with pkg.Cloudsc2(pkg.default_params(lregcl=False), klev, golden["ceta"]) as gpu:
        gpu._bind()
        lib.cloudsc2ad_(ref(1), ref(klon), ref(klon), ref(1), ref(klev), ref(0), C.byref(C.c_double(ptsphy)),
                        *[dp(traj[n]) for n in order26], *[dp(incr[n]) for n in order26])
    # PSUPSAT' is ASSIGNED PTSPHY*ZQP1' by the reference (cloudsc2ad.F90:1733): undo the extra PTSPHY
    rhs = sum(float((0.01 * x5[k] * incr[k]).sum()) for k in x5 if k != "psupsat")
    rhs += float((0.01 * x5["psupsat"] * incr["psupsat"]).sum()) / ptsphy
    assert abs(lhs) > 1.0 and abs(lhs - rhs) <= 1e-6 * abs(lhs), (lhs, rhs)
    for n in out10:
        assert not incr[n].any(), n                      # output adjoints consumed and zeroed


def test_host_ad_with_precomputed_trajectory_equals_as_written(pkg, src100):
    """The same through the HOST-pointer entry (chunked pipeline): with ad_have_trajectory the caller's
    PFPLSL5 / PFPLSN5 arrays (left by a cloudsc2_gpu_nl call on the same inputs) travel up with every chunk
    and the forward sweep is skipped; ragged NGPTOT, several chunks."""
    prm = pkg.default_params(lregcl=False)
    nproma, ngptot = 32, 1000                               # 32 blocks, last one ragged (8 columns)
    rng = np.random.default_rng(9)
    outs = []
    with pkg.Cloudsc2(prm, 137, src100.ceta) as gpu:
        for have in (0, 1):
            st = pkg.ArrayState(src100, nproma, ngptot)
            din, dout = pkg.driver.alloc_increments(st.nblocks, 137, nproma)
            r = np.random.default_rng(9)
            for k in din:
                din[k][...] = 1e-3 * r.standard_normal(din[k].shape)
            for k in dout:
                dout[k][...] = r.standard_normal(dout[k].shape)
            dout0 = {k: v.copy() for k, v in dout.items()}
            if have:
                gpu.nl(st)                                   # trajectory fluxes now in st.a["pfplsl"/"pfplsn"]
            gpu.set_option("ad_have_trajectory", have)
            try:
                gpu.ad(st, din, dout)
            finally:
                gpu.set_option("ad_have_trajectory", 0)
            outs.append(({k: v.copy() for k, v in din.items()}, {k: v.copy() for k, v in dout.items()}, dout0))
    for k in outs[0][0]:
        assert np.array_equal(outs[0][0][k], outs[1][0][k]), k
    for k, v in outs[1][1].items():                          # consumed and zeroed; padding columns untouched
        assert not v[:-1].any() and not v[-1][:, :8].any(), k
        assert np.array_equal(v[-1][:, 8:], outs[1][2][k][-1][:, 8:]), k
    del rng


@pytest.mark.parametrize("lregcl", [False, True])
def test_ad_fields_match_transliterated_fortran(pkg, obref, src100, lregcl):
    """The CUDA CLOUDSC2AD against the reference's OWN Fortran text (oracle/_ref: cloudsc2ad.F90,
    cuadjtqs.F90, cuadjtqsad.F90 transliterated by oracle/f90toc.py) directly: LREGCL off and on
    (cloudsc2ad.F90:1057-1059, 1308-1350, 1460, 1554-1559 -- the AD program's configuration)."""
    test_ad_fields_match_oracle(pkg, obref, src100, 100, 100, lregcl)
    test_ad_fields_match_oracle(pkg, obref, src100, 64, 200, lregcl)
