"""GPU suite: the gfortran-ABI per-block shims satur_ / cloudsc2_ / cloudsc2tl_ / cloudsc2ad_
(include/cloudsc2_fortran.h), called exactly as the reference's Fortran drivers call SATUR /
CLOUDSC2 / CLOUDSC2TL / CLOUDSC2AD -- every argument by reference, one NPROMA block per call,
KFDIA = ICEND possibly < KLON (cloudsc_driver_mod.F90:82-111) -- against the CPU oracle's
per-block restatements of the same subroutines."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

IN16 = ("paphp1", "papp1", "pqm1", "pqs", "ptm1", "pl", "pi", "plude", "plu", "pmfu", "pmfd",
        "pgtent", "pgtenq", "pgtenl", "pgteni", "psupsat")
OUT10 = ("ptent", "ptenq", "ptenl", "pteni", "pclc", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn", "pcovptot")
# positional order of one 26-array group in cloudsc2tl.F90:10-24 / cloudsc2ad.F90:10-24
ORDER26 = ("paphp1", "papp1", "pqm1", "pqs", "ptm1", "pl", "pi", "plude", "plu", "pmfu", "pmfd",
           "ptent", "pgtent", "ptenq", "pgtenq", "ptenl", "pgtenl", "pteni", "pgteni", "psupsat",
           "pclc", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn", "pcovptot")


def _ref(v):
    return C.byref(C.c_int(v))


def _dp(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _block(pkg, ob, gpu, src100, klon, kfdia):
    """Inputs of one block of klon columns (the first kfdia valid) as (KLEV[+1], KLON) arrays."""
    st = pkg.ArrayState(src100, klon, kfdia)
    x = ob.block_inputs(st, 0, gpu.params)
    return {k: np.ascontiguousarray(v) for k, v in x.items()}


def _scalars(klon, kfdia, klev=137):
    return [_ref(1), _ref(kfdia), _ref(klon), _ref(1), _ref(klev)]


@pytest.mark.parametrize("klon,kfdia", [(32, 32), (32, 29), (100, 100)])
def test_satur_and_cloudsc2_shims(pkg, ob, src100, gpu_nl, klon, kfdia):
    lib = pkg.load_library()
    x = _block(pkg, ob, gpu_nl, src100, klon, kfdia)
    gpu_nl._bind()
    # CALL SATUR(1, ICEND, NPROMA, 1, NLEV, .TRUE., PAP, PT, ZQSAT, 2)   cloudsc_driver_mod.F90:91-92
    pqs = np.full_like(x["ptm1"], -7.0)
    lib.satur_(*_scalars(klon, kfdia), _ref(1), _dp(x["papp1"]), _dp(x["ptm1"]), _dp(pqs), _ref(2))
    want = ob.satur(gpu_nl.params, x["papp1"], np.where(x["ptm1"] > 0, x["ptm1"], 250.0))
    assert np.abs(pqs[:, :kfdia] / want[:, :kfdia] - 1.0).max() < 1e-14
    assert (pqs[:, kfdia:] == -7.0).all()
    # CALL CLOUDSC2(1, ICEND, NPROMA, 1, NLEV, LDRAIN1D, PTSPHY, ...)    cloudsc_driver_mod.F90:94-107
    x["pqs"] = np.ascontiguousarray(np.where(np.arange(klon) < kfdia, pqs, 1e-3))
    y = {n: np.full((137 + (1 if n.startswith("pf") else 0), klon), 5.5) for n in OUT10}
    lib.cloudsc2_(*_scalars(klon, kfdia), _ref(0), C.byref(C.c_double(3600.0)),
                  *[_dp(x[n]) for n in IN16[:11]],
                  _dp(y["ptent"]), _dp(x["pgtent"]), _dp(y["ptenq"]), _dp(x["pgtenq"]),
                  _dp(y["ptenl"]), _dp(x["pgtenl"]), _dp(y["pteni"]), _dp(x["pgteni"]), _dp(x["psupsat"]),
                  *[_dp(y[n]) for n in OUT10[4:]])
    xv = {k: np.ascontiguousarray(v[:, :kfdia]) for k, v in x.items()}
    yo = ob.cloudsc2_block(gpu_nl.params, src100.ceta, 3600.0, xv)
    for n in OUT10:
        scale = max(np.abs(yo[n]).max(), 1e-300)
        assert np.abs(y[n][:, :kfdia] - yo[n]).max() <= 1e-11 * scale, n
        assert (y[n][:, kfdia:] == 5.5).all(), n          # columns beyond KFDIA untouched


@pytest.mark.parametrize("klon,kfdia,lregcl", [(32, 32, False), (64, 50, True)])
def test_cloudsc2tl_and_cloudsc2ad_shims(pkg, ob, src100, klon, kfdia, lregcl):
    lib = pkg.load_library()
    prm = pkg.default_params(lregcl=lregcl)
    with pkg.Cloudsc2(prm, 137, src100.ceta) as gpu:
        x5 = _block(pkg, ob, gpu, src100, klon, kfdia)
        x5["pqs"] = np.ascontiguousarray(ob.satur(prm, x5["papp1"], np.where(x5["ptm1"] > 0, x5["ptm1"], 250.0)))
        half = lambda n: 137 + (1 if n in ("paphp1",) or n.startswith("pf") else 0)
        y5 = {n: np.zeros((half(n), klon)) for n in OUT10}
        dx = {n: 0.01 * x5[n] for n in IN16}
        dy = {n: np.full((half(n), klon), -2.0) for n in OUT10}
        traj = {**x5, **y5}
        incr = {**dx, **dy}
        gpu._bind()
        sc = _scalars(klon, kfdia) + [_ref(0), C.byref(C.c_double(3600.0))]
        lib.cloudsc2tl_(*sc, *[_dp(traj[n]) for n in ORDER26], *[_dp(incr[n]) for n in ORDER26])
        cut = lambda d: {k: np.ascontiguousarray(v[:, :kfdia]) for k, v in d.items()}
        y5o, dyo = ob.cloudsc2tl_block(prm, src100.ceta, 3600.0, cut(x5), cut(dx))
        for n in OUT10:
            for got, want in ((y5[n], y5o[n]), (dy[n], dyo[n])):
                scale = max(np.abs(want).max(), 1e-300)
                assert np.abs(got[:, :kfdia] - want).max() <= 1e-9 * scale, n
        assert (dy["ptent"][:, kfdia:] == -2.0).all()
        # adjoint: y* = the TL result, x* starts from a non-zero value to check the accumulation
        dy_ad = {n: np.ascontiguousarray(np.where(np.arange(klon) < kfdia, dy[n], 0.0)) for n in OUT10}
        dx_ad = {n: np.full((half(n), klon), 0.25) for n in IN16}
        dy_ref, dx_ref = cut(dy_ad), cut(dx_ad)
        traj = {**x5, **{n: np.zeros((half(n), klon)) for n in OUT10}}
        incr = {**dx_ad, **dy_ad}
        lib.cloudsc2ad_(*sc, *[_dp(traj[n]) for n in ORDER26], *[_dp(incr[n]) for n in ORDER26])
        ob.cloudsc2ad_block(prm, src100.ceta, 3600.0, cut(x5), dx_ref, dy_ref)
        for n in IN16:
            scale = max(np.abs(dx_ref[n] - 0.25).max(), 1e-300)
            assert np.abs(dx_ad[n][:, :kfdia] - dx_ref[n]).max() <= 1e-9 * scale, n
            assert (dx_ad[n][:, kfdia:] == 0.25).all(), n
        for n in OUT10:
            assert not dy_ad[n][:, :kfdia].any(), n       # consumed and zeroed (cloudsc2ad.F90:955-966)
