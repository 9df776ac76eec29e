"""GPU suite, tangent-linear path and Taylor test.

Tolerances: TL fields |gpu - oracle| <= TL_RTOL * max|field| (TL_RTOL = 1e-9: the TL divides by
trajectory quantities that can be tiny, amplifying the 1-2 ulp exp differences); Taylor ratios:
identical verdict, and per-column ratios equal to a lambda-dependent tolerance because
sum(F - F5) is cancellation-dominated for small lambda (SURVEY 7 "hard parts").
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TL_RTOL = 1e-9


def _tl_reference(ob, prm, src, st, scale=0.01):
    """Oracle CLOUDSC2TL block by block with dx = scale * x -> dict of stacked (NB, KLEV[+1], NPROMA)."""
    y5s, dys = [], []
    for b in range(st.nblocks):
        icend = min(st.nproma, st.ngptot - b * st.nproma)
        x5 = ob.block_inputs(st, b, prm)
        x5 = {k: np.ascontiguousarray(v[:, :icend]) for k, v in x5.items()}
        dx = {k: scale * v for k, v in x5.items()}
        y5, dy = ob.cloudsc2tl_block(prm, src.ceta, st.ptsphy, x5, dx)
        y5s.append(y5)
        dys.append(dy)
    return y5s, dys


@pytest.mark.parametrize("nproma,ngptot,lregcl", [(100, 100, False), (32, 100, False), (1, 100, False),
                                                  (64, 640, True), (128, 300, True), (128, 4096, False), (256, 3000, True)])
def test_tl_fields_match_oracle(pkg, ob, src100, nproma, ngptot, lregcl):
    prm = pkg.default_params(lregcl=lregcl)
    st = pkg.ArrayState(src100, nproma, ngptot)
    nb = st.nblocks
    din, dout = pkg.driver.alloc_increments(nb, 137, nproma)
    # increments = 1 % of the inputs (cloudsc_driver_tl_mod.F90:156-171), PQS' = 1 % of SATUR
    a = st.a
    din["paph"][:] = 0.01 * a["paph"]; din["pap"][:] = 0.01 * a["pap"]; din["pq"][:] = 0.01 * a["pq"]
    din["pt"][:] = 0.01 * a["pt"]; din["pl"][:] = 0.01 * a["pclv"][:, 0]; din["pi"][:] = 0.01 * a["pclv"][:, 1]
    din["plude"][:] = 0.01 * a["plude"]; din["plu"][:] = 0.01 * a["plu"]; din["pmfu"][:] = 0.01 * a["pmfu"]
    din["pmfd"][:] = 0.01 * a["pmfd"]; din["gtent"][:] = 0.01 * a["b_cml"][:, 0]
    din["gtenq"][:] = 0.01 * a["b_cml"][:, 2]; din["gtenl"][:] = 0.01 * a["b_cml"][:, 3]
    din["gteni"][:] = 0.01 * a["b_cml"][:, 4]; din["psupsat"][:] = 0.01 * a["psupsat"]
    for b in range(nb):
        din["pqs"][b] = 0.01 * ob.satur(prm, a["pap"][b], np.where(a["pt"][b] > 0, a["pt"][b], 250.0))
    for v in dout.values():
        v.fill(-3.5)
    with pkg.Cloudsc2(prm, 137, src100.ceta) as gpu:
        gpu.tl(st, din, dout)
    y5s, dys = _tl_reference(ob, prm, src100, st)
    names = {"tent": "ptent", "tenq": "ptenq", "tenl": "ptenl", "teni": "pteni", "pclc": "pclc",
             "pfplsl": "pfplsl", "pfplsn": "pfplsn", "pfhpsl": "pfhpsl", "pfhpsn": "pfhpsn",
             "pcovptot": "pcovptot"}
    traj = {"ptent": st.a["b_loc"][:, 0], "ptenq": st.a["b_loc"][:, 2], "ptenl": st.a["b_loc"][:, 3],
            "pteni": st.a["b_loc"][:, 4], "pclc": st.a["pa"], "pfplsl": st.a["pfplsl"],
            "pfplsn": st.a["pfplsn"], "pfhpsl": st.a["pfhpsl"], "pfhpsn": st.a["pfhpsn"]}
    for gname, oname in names.items():
        ref = np.concatenate([d[oname] for d in dys], axis=1)             # (KLEV, ngptot)
        got = np.concatenate([dout[gname][b][:, :min(nproma, ngptot - b * nproma)] for b in range(nb)], axis=1)
        scale = max(np.abs(ref).max(), 1e-300)
        # columns whose trajectory sits on a branch knife-edge may flip; count them, allow none here
        err = np.abs(got - ref).max(axis=0) / scale
        assert (err <= TL_RTOL).all(), (gname, err.max(), int((err > TL_RTOL).sum()))
    for oname, arr in traj.items():
        ref = np.concatenate([d[oname] for d in y5s], axis=1)
        got = np.concatenate([arr[b][:, :min(nproma, ngptot - b * nproma)] for b in range(nb)], axis=1)
        scale = max(np.abs(ref).max(), 1e-300)
        assert np.abs(got - ref).max() <= 1e-11 * scale, oname
    tail = ngptot - (nb - 1) * nproma
    if tail < nproma:
        assert (dout["tent"][-1, :, tail:] == -3.5).all()                  # padding untouched


def test_tl_is_linear_on_gpu(pkg, src100):
    prm = pkg.default_params()
    st = pkg.ArrayState(src100, 50, 100)
    rng = np.random.default_rng(2)
    base = {"paph": st.a["paph"], "pap": st.a["pap"], "pq": st.a["pq"], "pqs": 1e-3 * np.ones_like(st.a["pq"]),
            "pt": st.a["pt"], "pl": st.a["pclv"][:, 0], "pi": st.a["pclv"][:, 1], "plude": st.a["plude"],
            "plu": st.a["plu"], "pmfu": st.a["pmfu"], "pmfd": st.a["pmfd"], "gtent": st.a["b_cml"][:, 0],
            "gtenq": st.a["b_cml"][:, 2], "gtenl": st.a["b_cml"][:, 3], "gteni": st.a["b_cml"][:, 4],
            "psupsat": st.a["psupsat"]}
    d1 = {k: np.ascontiguousarray(0.01 * v * rng.standard_normal(v.shape)) for k, v in base.items()}
    d2 = {k: np.ascontiguousarray(0.01 * v * rng.standard_normal(v.shape)) for k, v in base.items()}
    d3 = {k: 2.0 * d1[k] - 3.0 * d2[k] for k in base}
    outs = []
    with pkg.Cloudsc2(prm, 137, src100.ceta) as gpu:
        for d in (d1, d2, d3):
            _, dout = pkg.driver.alloc_increments(st.nblocks, 137, 50)
            gpu.tl(st, d, dout)
            outs.append(dout)
    for n in outs[0]:
        want = 2.0 * outs[0][n] - 3.0 * outs[1][n]
        scale = max(np.abs(want).max(), 1e-300)
        assert np.abs(outs[2][n] - want).max() <= 1e-10 * scale, n


def test_taylor_test_config2(pkg, ob, src100, gpu_nl):
    """BASELINE config 2: dwarf-cloudsc2-tl 1 100 1 -- verdict identical to the oracle's, per-column
    ratios equal within a lambda-dependent tolerance."""
    st = pkg.ArrayState(src100, nproma=1, ngptot=100)
    ref = pkg.ArrayState(src100, nproma=1, ngptot=100)
    z, rb = gpu_nl.tl_taylor(st)
    zo, rbo, _ = ob.driver_tl(gpu_nl.params, src100.ceta, ref, numomp=2)
    pen, istart = pkg.taylor_verdict(z)
    assert (pen, istart) == ob.taylor_verdict(zo)
    assert 0 <= pen <= 5
    assert rb.shape == (100, 10) and np.isfinite(rb).all()
    lam = 10.0 ** -np.arange(1, 11)
    # |r_gpu - r_oracle| <= 1e-9 (physics) + 2e-13/lambda (cancellation in sum(F-F5), F ~ 1e-16 rel)
    tol = 1e-9 * np.maximum(1.0, np.abs(rbo)) + 2e-13 / lam * np.maximum(1.0, np.abs(rbo))
    assert (np.abs(rb - rbo) <= tol).all(), np.abs(rb - rbo).max(axis=0)
    assert np.allclose(z, rb.max(axis=0))
    # the call also leaves the trajectory outputs in the caller's arrays, like the reference
    for n, a in st.outputs().items():
        scale = max(np.abs(ref.outputs()[n]).max(), 1e-300)
        assert np.abs(a - ref.outputs()[n]).max() <= 1e-11 * scale, n


@pytest.mark.parametrize("nproma,ngptot", [(32, 100), (100, 100), (64, 1000)])
def test_taylor_test_blocked(pkg, ob, src100, gpu_nl, nproma, ngptot):
    """Per-block sums (NPROMA > 1): summation order differs (tree vs JL-fastest), verdict must agree."""
    st = pkg.ArrayState(src100, nproma, ngptot)
    ref = pkg.ArrayState(src100, nproma, ngptot)
    z, rb = gpu_nl.tl_taylor(st)
    zo, rbo, _ = ob.driver_tl(gpu_nl.params, src100.ceta, ref, numomp=2)
    assert pkg.taylor_verdict(z)[0] <= 5 and pkg.taylor_verdict(z)[0] >= 0
    assert pkg.taylor_verdict(z)[1] == ob.taylor_verdict(zo)[1]
    lam = 10.0 ** -np.arange(1, 11)
    tol = 1e-8 * np.maximum(1.0, np.abs(rbo)) + 1e-12 / lam * np.maximum(1.0, np.abs(rbo))
    assert (np.abs(rb - rbo) <= tol).all()


def test_tl_kernel_matches_finite_differences_of_reference_python_kernel(pkg, golden_fd):
    """GPU CLOUDSC2TL (with PQS5 / PQS' supplied, CLOUDSC2TL-call semantics) against the central
    finite differences of the reference's Python NL kernel (tests/golden/tl_fd_pyref.npz)."""
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
    traj = {**x5, **{n: np.zeros((half(n), klon)) for n in out10}}
    incr = {**{k: 0.01 * v for k, v in x5.items()}, **{n: np.zeros((half(n), klon)) for n in out10}}
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ref = lambda v: C.byref(C.c_int(v))
    src = pkg.synth_source(seed=0, klon=100, klev=klev)
    with pkg.Cloudsc2(pkg.default_params(lregcl=False), klev, golden["ceta"]) as gpu:
        gpu._bind()
        # the per-block Fortran-ABI entry takes PQS5 and PQS' explicitly, like CALL CLOUDSC2TL
        lib.cloudsc2tl_(ref(1), ref(klon), ref(klon), ref(1), ref(klev), ref(0),
                        C.byref(C.c_double(float(golden["ptsphy"]))),
                        *[dp(traj[n]) for n in order26], *[dp(incr[n]) for n in order26])
    del src
    for n in out10:
        d = fd["d_" + n]
        tol = (1e-6 if n == "pclc" else 1e-7) * max(np.abs(d).max(), 1e-300)   # finite-difference error: measured <= 2e-8
        assert np.abs(incr[n] - d).max() <= tol, n


ORDER26 = ("paphp1", "papp1", "pqm1", "pqs", "ptm1", "pl", "pi", "plude", "plu", "pmfu", "pmfd",
           "ptent", "pgtent", "ptenq", "pgtenq", "ptenl", "pgtenl", "pteni", "pgteni", "psupsat",
           "pclc", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn", "pcovptot")
OUT10 = ("ptent", "ptenq", "ptenl", "pteni", "pclc", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn", "pcovptot")


def _dirs_case():
    from pathlib import Path
    from tests.fd_directions import NCOL
    gdir = Path(__file__).resolve().parent / "golden"
    g, fd = np.load(gdir / "nl_pyref.npz"), np.load(gdir / "tl_fd_dirs.npz")
    x5 = {k[3:]: np.ascontiguousarray(g[k][:, :NCOL]) for k in g.files if k.startswith("in_")}
    x5["pqs"] = np.ascontiguousarray(g["pqs"][:, :NCOL])
    return g, fd, x5


def test_tl_and_ad_kernels_match_reference_derivative_input_by_input(pkg):
    """tests/golden/tl_fd_dirs.npz (finite differences of the reference's Python NL kernel along 16
    single-input and 2 random directions): every column of the CUDA TL Jacobian on its own, and every row
    of the CUDA adjoint through <D_k, y> = <dx_k, M'^T y>.  Per-block Fortran-ABI entries, PQS5 / PQS'
    supplied like CALL CLOUDSC2TL / CLOUDSC2AD."""
    import ctypes as C
    from tests.fd_directions import IN16, NAMES, OUT7, directions
    g, fd, x5 = _dirs_case()
    lib = pkg.load_library()
    klev, klon = x5["ptm1"].shape
    half = lambda n: klev + (1 if n == "paphp1" or n.startswith("pf") else 0)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ref = lambda v: C.byref(C.c_int(v))
    ptsphy = float(g["ptsphy"])
    dirs = directions(x5)
    rng = np.random.default_rng(23)
    y = {n: np.zeros((half(n), klon)) for n in OUT10}
    for n in OUT7:
        m = np.ones_like(fd[f"m00_{n}"], dtype=bool)
        for i in range(len(dirs)):
            m &= fd[f"m{i:02d}_{n}"]
        y[n] = np.ascontiguousarray(rng.uniform(0.5, 1.5, m.shape) * m)
    with pkg.Cloudsc2(pkg.default_params(lregcl=False), klev, g["ceta"]) as gpu:
        gpu._bind()
        for i, dx in enumerate(dirs):
            traj = {**x5, **{n: np.zeros((half(n), klon)) for n in OUT10}}
            incr = {**{k: v.copy() for k, v in dx.items()}, **{n: np.zeros((half(n), klon)) for n in OUT10}}
            lib.cloudsc2tl_(ref(1), ref(klon), ref(klon), ref(1), ref(klev), ref(0), C.byref(C.c_double(ptsphy)),
                            *[dp(traj[n]) for n in ORDER26], *[dp(incr[n]) for n in ORDER26])
            for n in OUT7:
                d, m = fd[f"d{i:02d}_{n}"], fd[f"m{i:02d}_{n}"]
                tol = 1e-6 * max(np.abs(d).max(), 1e-300) + 2e-10 * np.abs(traj[n]).max()
                assert (np.abs(incr[n] - d) * m).max() <= tol, (NAMES[i], n)
        traj = {**x5, **{n: np.zeros((half(n), klon)) for n in OUT10}}
        adj = {**{k: np.zeros_like(v) for k, v in x5.items()}, **{n: v.copy() for n, v in y.items()}}
        lib.cloudsc2ad_(ref(1), ref(klon), ref(klon), ref(1), ref(klev), ref(0), C.byref(C.c_double(ptsphy)),
                        *[dp(traj[n]) for n in ORDER26], *[dp(adj[n]) for n in ORDER26])
    for i, dx in enumerate(dirs):
        lhs = sum(float((fd[f"d{i:02d}_{n}"] * y[n]).sum()) for n in OUT7)
        rhs = sum(float((dx[k] * adj[k]).sum()) for k in IN16 if k != "psupsat")
        rhs += float((dx["psupsat"] * adj["psupsat"]).sum()) / ptsphy      # cloudsc2ad.F90:1733
        scale = sum(float(np.abs(fd[f"d{i:02d}_{n}"] * y[n]).sum()) for n in OUT7)
        assert scale > 0 and abs(lhs - rhs) <= 2e-6 * scale, (NAMES[i], lhs, rhs)


def test_host_tl_and_ad_pipeline_equals_device_entries_many_chunks(pkg, src100, gpu_nl):
    """The host-pointer TL / AD entries run the caller's blocks through the chunked three-stream pipeline
    (compact device layout, ramped chunk plan, increments travelling with their chunk).  On a problem large
    enough for the ramp (101 blocks of 128 columns, ragged tail) the results must equal the device-pointer
    entries on the reference layout bit for bit."""
    nproma, ngptot = 128, 100 * 128 + 57
    gpu = gpu_nl
    st = pkg.ArrayState(src100, nproma, ngptot)
    nb = st.nblocks
    r = np.random.default_rng(21)
    din, dout = pkg.driver.alloc_increments(nb, 137, nproma)
    for k in din:
        din[k][...] = 1e-3 * r.standard_normal(din[k].shape)
    for k in dout:
        dout[k][...] = 5.0                                   # must survive in the padding columns
    # device-pointer path (reference layout on the device)
    ds = pkg.DeviceState(gpu, st)
    ds.zero()
    d_in = {k: gpu.malloc(v.nbytes) for k, v in din.items()}
    d_out = {k: gpu.malloc(v.nbytes) for k, v in dout.items()}
    want_tl = {k: np.empty_like(v) for k, v in dout.items()}
    want_ad = {k: np.empty_like(v) for k, v in din.items()}
    try:
        for k, v in din.items():
            gpu.h2d(d_in[k], v)
        for k, v in dout.items():
            gpu.h2d(d_out[k], v)
        gpu.tl_dev(ds, st.ptsphy, d_in, d_out)
        gpu.sync()
        for k, v in want_tl.items():
            gpu.d2h(v, d_out[k])
        gpu.ad_dev(ds, st.ptsphy, d_in, d_out)              # adjoint of the TL outputs, accumulated into din
        gpu.sync()
        for k, v in want_ad.items():
            gpu.d2h(v, d_in[k])
    finally:
        for p in list(d_in.values()) + list(d_out.values()):
            gpu.free(p)
        ds.free()
    # host-pointer path
    gpu.tl(st, din, dout)
    for k in dout:
        assert np.array_equal(dout[k], want_tl[k]), k
        assert (dout[k][-1][:, 57:] == 5.0).all(), k         # padding of the ragged last block untouched
    gpu.ad(st, din, dout)
    for k in din:
        assert np.array_equal(din[k], want_ad[k]), k
    for k in dout:
        assert not dout[k][:-1].any() and not dout[k][-1][:, :57].any(), k


@pytest.mark.parametrize("lregcl", [False, True])
def test_tl_fields_match_transliterated_fortran(pkg, obref, src100, lregcl):
    """The CUDA CLOUDSC2TL against the reference's OWN Fortran text (oracle/_ref, cloudsc2tl.F90 +
    cuadjtqstl.F90 transliterated by oracle/f90toc.py) directly, not only through the hand oracle:
    100 columns, LREGCL off and on (the regularised statements :575-580, 657, 754-760, 794-800, 998-1000)."""
    test_tl_fields_match_oracle(pkg, obref, src100, 100, 100, lregcl)
    test_tl_fields_match_oracle(pkg, obref, src100, 64, 640, lregcl)
