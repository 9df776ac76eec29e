"""CPU suite: pins the oracle (oracle/*.c) against the golden vectors produced by the reference's
own Python kernel, and against the reference's own known-answer properties (Taylor test,
adjoint dot-product test).  No GPU, no /root/reference."""
import numpy as np
import pytest

# |oracle - reference python| <= RTOL_PY * max|field| : the two differ only by libm-vs-numpy exp and
# x**2-vs-x*x rounding (measured 2e-15).
RTOL_PY = 2e-14


def _inputs(golden):
    return {k[3:]: np.ascontiguousarray(golden[k]) for k in golden.files if k.startswith("in_")}


def test_golden_params_match_defaults(pkg, golden):
    prm = pkg.default_params()
    for n, v in zip(golden["params_names"], golden["params_values"]):
        assert float(getattr(prm, str(n))) == float(v), n


def test_satur_matches_reference_python(pkg, ob, golden):
    x = _inputs(golden)
    pqs = ob.satur(pkg.default_params(), x["papp1"], x["ptm1"])
    assert np.abs(pqs / golden["pqs"] - 1.0).max() < 1e-15 * 8


def test_cloudsc2_matches_reference_python(pkg, ob, golden):
    x = _inputs(golden)
    x["pqs"] = np.ascontiguousarray(golden["pqs"])
    y = ob.cloudsc2_block(pkg.default_params(), golden["ceta"], float(golden["ptsphy"]), x)
    for n in ob.OUT10:
        r = golden["out_" + n]
        scale = max(np.abs(r).max(), 1e-300)
        assert np.abs(y[n] - r).max() <= RTOL_PY * scale, n
    # behavioural gotchas: PCOVPTOT identically 0, -0.0 enthalpy flux at the top
    assert not y["pcovptot"].any()
    assert np.signbit(y["pfhpsl"][0]).all() and not y["pfhpsl"][0].any()


def test_golden_covers_all_cloud_cover_branches(golden):
    pclc = golden["out_pclc"]
    assert (pclc == 0).any() and (pclc == 1).any() and ((pclc > 0) & (pclc < 1)).any()
    assert golden["out_pfplsl"].max() > 0 and golden["out_pfplsn"].max() > 0


def test_every_source_column_is_active(pkg, ob, src100):
    """A clear dry column makes the reference's Taylor logic STOP (ZCOUNT==0); the synthetic
    source must not contain one."""
    st = pkg.ArrayState(src100, nproma=100, ngptot=100)
    ob.driver_nl(pkg.default_params(), src100.ceta, st)
    o = st.outputs()
    active = (np.abs(o["pa"][0]).sum(0) > 0) | (o["pfplsl"][0].sum(0) + o["pfplsn"][0].sum(0) > 0)
    assert active.all()


def test_oracle_taylor_test_passes(pkg, ob, src100):
    """dwarf-cloudsc2-tl 1 100 1 (cloudsc_driver_tl_mod.F90:273-311): TEST PASSED, penalty <= 5."""
    st = pkg.ArrayState(src100, nproma=1, ngptot=100)
    z, rb, _ = ob.driver_tl(pkg.default_params(lregcl=False), src100.ceta, st, numomp=2)
    pen, istart = ob.taylor_verdict(z)
    assert 0 <= pen <= 5 and 1 <= istart <= 4, (pen, istart, z)
    assert np.isfinite(rb).all()
    # the product's verdict function is the same logic
    assert pkg.taylor_verdict(z) == (pen, istart)
    # error is V-shaped: best lambda reaches < 1e-6
    assert np.abs(1 - z).min() < 1e-6


def test_oracle_adjoint_test_passes(pkg, ob, src100):
    """dwarf-cloudsc2-ad 1 100 100 (cloudsc_driver_ad_mod.F90:286-294): ZNORMG < 10000 eps."""
    st = pkg.ArrayState(src100, nproma=100, ngptot=100)
    zn, nc, _ = ob.driver_ad(pkg.default_params(lregcl=True), src100.ceta, st)
    assert zn < 10000.0 and pkg.adjoint_verdict(zn)
    assert (nc[:, 0] > 0).all()        # every column has a non-trivial TL response
    assert np.allclose(nc[:, 0], nc[:, 1], rtol=1e-11)


def test_oracle_adjoint_without_regularisation(pkg, ob, src100):
    st = pkg.ArrayState(src100, nproma=32, ngptot=100)      # ragged last block (4 columns)
    zn, nc, _ = ob.driver_ad(pkg.default_params(lregcl=False), src100.ceta, st, numomp=2)
    assert zn < 10000.0


def test_oracle_tl_is_linear(pkg, ob, src100):
    """CLOUDSC2TL(a dx1 + b dx2) == a CLOUDSC2TL(dx1) + b CLOUDSC2TL(dx2) (same trajectory)."""
    prm = pkg.default_params()
    st = pkg.ArrayState(src100, nproma=20, ngptot=20)
    x5 = ob.block_inputs(st, 0, prm)
    rng = np.random.default_rng(1)
    d1 = {k: v * 0.01 * rng.standard_normal(v.shape) for k, v in x5.items()}
    d2 = {k: v * 0.01 * rng.standard_normal(v.shape) for k, v in x5.items()}
    d3 = {k: 2.0 * d1[k] - 3.0 * d2[k] for k in x5}
    _, y1 = ob.cloudsc2tl_block(prm, src100.ceta, st.ptsphy, x5, d1)
    _, y2 = ob.cloudsc2tl_block(prm, src100.ceta, st.ptsphy, x5, d2)
    _, y3 = ob.cloudsc2tl_block(prm, src100.ceta, st.ptsphy, x5, d3)
    for n in ob.OUT10:
        want = 2.0 * y1[n] - 3.0 * y2[n]
        scale = max(np.abs(want).max(), 1e-300)
        assert np.abs(y3[n] - want).max() <= 1e-10 * scale, n


def test_oracle_tl_trajectory_equals_nl(pkg, ob, src100):
    prm = pkg.default_params()
    st = pkg.ArrayState(src100, nproma=50, ngptot=50)
    x5 = ob.block_inputs(st, 0, prm)
    y = ob.cloudsc2_block(prm, src100.ceta, st.ptsphy, x5)
    y5, _ = ob.cloudsc2tl_block(prm, src100.ceta, st.ptsphy, x5, {k: 0.01 * v for k, v in x5.items()})
    for n in ob.OUT10:
        scale = max(np.abs(y[n]).max(), 1e-300)
        assert np.abs(y5[n] - y[n]).max() <= 1e-13 * scale, n


def test_tl_oracle_matches_finite_differences_of_reference_python_kernel(pkg, ob, golden_fd):
    """tests/golden/tl_fd_pyref.npz holds central finite differences of the REFERENCE'S OWN Python
    nonlinear kernel along dx = 0.01 x (made by tests/golden/make_golden_tl.py in the build
    container).  The TL restatement must reproduce that derivative: this pins CLOUDSC2TL against
    reference code, not only against our own NL through the Taylor test.  Agreement is limited by
    the finite-difference error (measured 1e-10 relative; 2e-7 for PCLC next to its SQRT branch)."""
    golden, fd = golden_fd
    x5 = _inputs(golden)
    x5["pqs"] = np.ascontiguousarray(golden["pqs"])
    dx = {k: 0.01 * v for k, v in x5.items()}
    _, dy = ob.cloudsc2tl_block(pkg.default_params(lregcl=False), golden["ceta"], float(golden["ptsphy"]), x5, dx)
    for n in ob.OUT10:
        d = fd["d_" + n]
        tol = (1e-6 if n == "pclc" else 1e-7) * max(np.abs(d).max(), 1e-300)   # finite-difference error: measured <= 2e-8
        assert np.abs(dy[n] - d).max() <= tol, n
        assert fd["curv_" + n].max() <= 1e-6 * max(np.abs(d).max(), 1e-300) * float(fd["eps"]) * 1e4, n


def _dirs_case():
    from pathlib import Path
    from tests.fd_directions import NCOL
    gdir = Path(__file__).resolve().parent / "golden"
    g, fd = np.load(gdir / "nl_pyref.npz"), np.load(gdir / "tl_fd_dirs.npz")
    x5 = {k[3:]: np.ascontiguousarray(g[k][:, :NCOL]) for k in g.files if k.startswith("in_")}
    x5["pqs"] = np.ascontiguousarray(g["pqs"][:, :NCOL])
    return g, fd, x5


@pytest.mark.parametrize("obname", ["ob", "obref"])
def test_tl_matches_reference_derivative_input_by_input(pkg, request, obname):
    """tests/golden/tl_fd_dirs.npz: finite differences of the reference's Python NL kernel along 16
    single-input directions and 2 random ones.  Every column of the TL Jacobian is checked on its own
    (the one-direction golden could hide errors that compensate between inputs).  Tolerance: 1e-6 of
    the direction's largest derivative + the rounding noise of the difference quotient
    (2e-10 max|F|), on the points whose second difference is smooth.  Both the hand oracle and the
    transliterated Fortran (oracle/_ref) must pass."""
    from tests.fd_directions import NAMES, OUT7, directions
    o = request.getfixturevalue(obname)
    g, fd, x5 = _dirs_case()
    for i, dx in enumerate(directions(x5)):
        y5, dy = o.cloudsc2tl_block(pkg.default_params(lregcl=False), g["ceta"], float(g["ptsphy"]), x5, dx)
        for n in OUT7:
            d, m = fd[f"d{i:02d}_{n}"], fd[f"m{i:02d}_{n}"]
            tol = 1e-6 * max(np.abs(d).max(), 1e-300) + 2e-10 * np.abs(y5[n]).max()
            assert (np.abs(dy[n] - d) * m).max() <= tol, (NAMES[i], n)
            assert m.mean() > 0.9, (NAMES[i], n)


def test_ad_is_the_transpose_input_by_input(pkg, ob):
    """<D_k, y> = <dx_k, M'^T y> for each of the 16 single-input directions: row k of the adjoint against
    column k of the reference derivative."""
    from tests.fd_directions import IN16, NAMES, OUT7, directions
    g, fd, x5 = _dirs_case()
    klev, klon = x5["ptm1"].shape
    rng = np.random.default_rng(23)
    ptsphy = float(g["ptsphy"])
    dirs = directions(x5)
    # one adjoint run with a fixed y supported on the smooth points of every direction
    y = {n: np.zeros((klev + (1 if n.startswith("pf") else 0), klon)) for n in ob.OUT10}
    for n in OUT7:
        m = np.ones_like(fd[f"m00_{n}"], dtype=bool)
        for i in range(len(dirs)):
            m &= fd[f"m{i:02d}_{n}"]
        y[n] = np.ascontiguousarray(rng.uniform(0.5, 1.5, m.shape) * m)
    adj = ob.alloc16(klev, klon)
    ob.cloudsc2ad_block(pkg.default_params(lregcl=False), g["ceta"], ptsphy, x5, adj, {n: v.copy() for n, v in y.items()})
    for i, dx in enumerate(dirs):
        lhs = sum(float((fd[f"d{i:02d}_{n}"] * y[n]).sum()) for n in OUT7)
        rhs = sum(float((dx[k] * adj[k]).sum()) for k in IN16 if k != "psupsat")
        rhs += float((dx["psupsat"] * adj["psupsat"]).sum()) / ptsphy      # cloudsc2ad.F90:1733, see above
        scale = sum(float(np.abs(fd[f"d{i:02d}_{n}"] * y[n]).sum()) for n in OUT7)
        assert scale > 0 and abs(lhs - rhs) <= 2e-6 * scale, (NAMES[i], lhs, rhs)


def test_ad_oracle_is_the_transpose_of_the_reference_derivative(pkg, ob, golden_fd):
    """<D, y> = <dx, M'^T y> with D = finite-difference derivative of the REFERENCE'S Python kernel
    along dx (tl_fd_pyref.npz) and M'^T y from the adjoint restatement: pins CLOUDSC2AD against
    reference code through the dot-product identity the reference's own adjoint test uses
    (cloudsc_driver_ad_mod.F90:184-267).  LREGCL=.FALSE. so that the adjoint is the exact transpose."""
    golden, fd = golden_fd
    x5 = _inputs(golden)
    x5["pqs"] = np.ascontiguousarray(golden["pqs"])
    klev, klon = x5["ptm1"].shape
    dx = {k: 0.01 * v for k, v in x5.items()}
    rng = np.random.default_rng(17)
    y = {n: np.ascontiguousarray(rng.uniform(0.5, 1.5, fd["d_" + n].shape) / max(np.abs(fd["d_" + n]).max(), 1e-30))
         for n in ob.OUT10}
    lhs = sum(float((fd["d_" + n] * y[n]).sum()) for n in ob.OUT10)
    adj = ob.alloc16(klev, klon)
    ob.cloudsc2ad_block(pkg.default_params(lregcl=False), golden["ceta"], float(golden["ptsphy"]), x5, adj,
                        {n: v.copy() for n, v in y.items()})
    # The reference ASSIGNS PSUPSAT' = PTSPHY*ZQP1' (cloudsc2ad.F90:1733): an extra factor PTSPHY
    # relative to the transpose of the TL statement ZQP1 = ... + PSUPSAT (cloudsc2tl.F90:346); its own
    # test avoids the term with ZSUPSAT = 0 (cloudsc_driver_ad_mod.F90:139).  Undo the factor here.
    ptsphy = float(golden["ptsphy"])
    rhs = sum(float((dx[k] * adj[k]).sum()) for k in dx if k != "psupsat")
    rhs += float((dx["psupsat"] * adj["psupsat"]).sum()) / ptsphy
    assert abs(lhs) > 1.0
    assert abs(lhs - rhs) <= 1e-6 * abs(lhs), (lhs, rhs)


def _seed_golden(pkg, name="nl_pyref_seed5.npz"):
    from pathlib import Path
    g = np.load(Path(__file__).resolve().parent / "golden" / name)
    src = pkg.synth_source(seed=int(g["seed"]), klon=100, klev=137).subset(list(g["cols"]))
    f = src.f
    x = {"paphp1": f["paph"], "papp1": f["pap"], "pqm1": f["pq"], "ptm1": f["pt"], "pl": f["pclv"][0],
         "pi": f["pclv"][1], "plude": f["plude"], "plu": f["plu"], "pmfu": f["pmfu"], "pmfd": f["pmfd"],
         "pgtent": f["tend_cml"][0], "pgtenq": f["tend_cml"][2], "pgtenl": f["tend_cml"][3],
         "pgteni": f["tend_cml"][4], "psupsat": f["psupsat"]}
    x = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in x.items()}
    # the generator is deterministic: the regenerated inputs are the ones the golden run used
    assert np.isclose(sum(float(np.abs(v).sum()) for v in x.values()), float(g["in_checksum"]), rtol=1e-15)
    return g, src, x


def test_cloudsc2_matches_reference_python_on_a_second_atmosphere(pkg, ob):
    """Second set of golden vectors from the reference's Python kernel: 48 columns of generator
    seed 5 (tests/golden/make_golden.py 5; outputs only, inputs regenerated)."""
    g, src, x = _seed_golden(pkg)
    pqs = ob.satur(pkg.default_params(), x["papp1"], x["ptm1"])
    assert np.abs(pqs / g["pqs"] - 1.0).max() < 1e-14
    x["pqs"] = np.ascontiguousarray(g["pqs"])
    y = ob.cloudsc2_block(pkg.default_params(), g["ceta"], float(g["ptsphy"]), x)
    for n in ob.OUT10:
        r = g["out_" + n]
        assert np.abs(y[n] - r).max() <= RTOL_PY * max(np.abs(r).max(), 1e-300), n


def test_oracle_matches_reference_python_on_edge_columns(pkg, ob):
    """tests/golden/nl_pyref_edge.npz (make_golden_edge.py): KLEV = 60, columns with the first-guess
    temperature exactly ON the thresholds the scheme branches on (RTT, RTICE, RLPTRC, RTT+2), dry /
    supersaturated / no-condensate / heavy-condensate columns, PLU == ZEPS2, no mass flux, PSUPSAT > 0,
    very cold / very warm -- SATUR and CLOUDSC2 of the oracle against the reference's Python kernel."""
    from pathlib import Path
    g = np.load(Path(__file__).resolve().parent / "golden" / "nl_pyref_edge.npz")
    x = _inputs(g)
    prm = pkg.default_params()
    pqs = ob.satur(prm, x["papp1"], x["ptm1"])
    assert np.abs(pqs / g["pqs"] - 1.0).max() < 1e-15 * 8
    x["pqs"] = np.ascontiguousarray(g["pqs"])
    y = ob.cloudsc2_block(prm, g["ceta"], float(g["ptsphy"]), x)
    for n in ob.OUT10:
        r = g["out_" + n]
        scale = max(np.abs(r).max(), 1e-300)
        assert np.abs(y[n] - r).max() <= RTOL_PY * scale, n
    # the forced temperatures really sit on the thresholds
    t1 = x["ptm1"] + float(g["ptsphy"]) * x["pgtent"]
    assert (t1[:, 0] == prm.rtt).sum() >= 10 and (t1[:, 1] == prm.rtice).sum() >= 10
    assert (t1[:, 3] == prm.rtt + 2.0).sum() >= 10 and (x["plu"][:, 9] == 1.0e-10).all()
