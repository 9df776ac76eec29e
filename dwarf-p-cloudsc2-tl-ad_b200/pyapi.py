"""Drop-in replacements for the reference's importable Python kernels:

    from cloudsc2_py import satur, cloudsc2_py          # reference: src/cloudsc2_nl_gt4py/cloudsc2_py.py

Same names, same positional argument lists (`satur` :12, `cloudsc2_py` :54-59), same in-place
output convention (the caller's numpy arrays `ptent, ptenq, ptenl, pteni, pclc, pfplsl, pfplsn,
pfhpsl, pfhpsn, pcovptot` / `pqsat` are overwritten), arrays shaped (klev[+1], klon) as in the
reference's driver.  The arithmetic runs on the GPU through the C ABI (one block of NPROMA = klon
columns); there is no NumPy fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .driver import Cloudsc2


def _params(yrecldp, yrmcst, yrethf, yrephli, lregcl: bool = False) -> _abi.Params:
    p = _abi.Params()
    for n in ("rg", "rd", "rcpd", "retv", "rlvtt", "rlstt", "rlmlt", "rtt"):
        setattr(p, n, float(getattr(yrmcst, n)))
    for n in ("r2es", "r3les", "r3ies", "r4les", "r4ies", "r5les", "r5ies", "r5alvcp", "r5alscp",
              "ralvdcp", "ralsdcp", "rtwat", "rtice", "rtwat_rtice_r"):
        setattr(p, n, float(getattr(yrethf, n)))
    p.rvtmp2 = float(getattr(yrethf, "rvtmp2", 0.0))
    for n in ("rclcrit", "rkconv", "rlmin", "rpecons"):
        setattr(p, n, float(getattr(yrecldp, n)))
    p.rlptrc = float(yrephli.rlptrc)
    p.lphylin = 1
    p.levapls2 = 0            # hard-coded False in the reference kernel (cloudsc2_py.py:393)
    p.lregcl = int(lregcl)
    p.ldrain1d = 0
    return p


def _dev(gpu: Cloudsc2, a: np.ndarray) -> int:
    a = np.ascontiguousarray(a, dtype=np.float64)
    p = gpu.malloc(a.nbytes)
    gpu.h2d(p, a)
    return p


def satur(kidia, kfdia, klon, ktdia, klev, ldphylin, paprsf, pt, pqsat, kflag, yrethf, yrmcst):
    """SATUR (reference cloudsc2_py.py:12-52, satur.F90:106-123): fills pqsat[ktdia-1:klev, kidia-1:kfdia]."""
    if not ldphylin:
        raise NotImplementedError("only the LDPHYLIN branch exists on the GPU path (as in the dwarf)")
    from types import SimpleNamespace
    dummy = SimpleNamespace(rclcrit=1.0, rkconv=1.0, rlmin=1.0, rpecons=1.0, rlptrc=1.0)
    prm = _params(dummy, yrmcst, yrethf, dummy)
    with Cloudsc2(prm, klev, np.linspace(0.0, 1.0, klev)) as gpu:
        qs = gpu.satur(np.asarray(paprsf, dtype=np.float64), np.asarray(pt, dtype=np.float64))
    pqsat[ktdia - 1:klev, kidia - 1:kfdia] = qs[ktdia - 1:klev, kidia - 1:kfdia]


def cloudsc2_py(kidia, kfdia, klon, ktdia, klev, ldrain1d, ptsphy, paphp1, papp1, pqm1, pqs, ptm1,
                pl, pi, plude, plu, pmfu, pmfd, ptent, pgtent, ptenq, pgtenq, ptenl, pgtenl, pteni,
                pgteni, psupsat, pclc, pfplsl, pfplsn, pfhpsl, pfhpsn, pcovptot, yrecldp, yrecld,
                yrmcst, yrethf, yrephli):
    """CLOUDSC2 (reference cloudsc2_py.py:54-612, cloudsc2.F90:10-741) for one block of klon columns,
    PQS supplied by the caller, on the GPU.  kidia..kfdia must span 1..klon."""
    if ldrain1d:
        raise NotImplementedError("LDRAIN1D=.TRUE. (evaporation branch) is not built; the dwarf forces it off")
    if kidia != 1 or kfdia != klon or ktdia != 1:
        raise ValueError("the GPU entry processes whole blocks: kidia=1, kfdia=klon, ktdia=1")
    prm = _params(yrecldp, yrmcst, yrethf, yrephli)
    ceta = np.ascontiguousarray(np.asarray(yrecld.ceta, dtype=np.float64)[:klev])
    f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    n2 = klev * klon
    with Cloudsc2(prm, klev, ceta) as gpu:
        ptrs = []
        try:
            up = lambda a: (ptrs.append(_dev(gpu, a)) or ptrs[-1])
            pclv = np.zeros((_abi.NCLV, klev, klon))
            pclv[0], pclv[1] = f8(pl), f8(pi)
            cml = np.zeros((_abi.NSTATE, klev, klon))
            cml[0], cml[2], cml[3], cml[4] = f8(pgtent), f8(pgtenq), f8(pgtenl), f8(pgteni)
            fld = _abi.Fields()
            fld.pt, fld.pq, fld.pap, fld.paph = up(f8(ptm1)), up(f8(pqm1)), up(f8(papp1)), up(f8(paphp1))
            # the reference kernel indexes plu[jk+1] and is handed klev+1 rows; only the first klev exist in Fortran
            fld.plu, fld.plude, fld.pmfu, fld.pmfd = up(f8(np.asarray(plu)[:klev])), up(f8(plude)), up(f8(pmfu)), up(f8(pmfd))
            fld.psupsat, fld.pclv, fld.b_cml = up(f8(psupsat)), up(pclv), up(cml)
            outs = {"b_loc": _abi.NSTATE * n2, "pa": n2, "pcovptot": n2, "pfplsl": n2 + klon,
                    "pfplsn": n2 + klon, "pfhpsl": n2 + klon, "pfhpsn": n2 + klon}
            optr = {}
            for n, cnt in outs.items():
                optr[n] = gpu.malloc(8 * cnt)
                ptrs.append(optr[n])
                gpu.memset(optr[n], 0, 8 * cnt)
                setattr(fld, n, optr[n])
            dqs = up(f8(pqs))
            gpu._check(gpu.lib.cloudsc2_gpu_nl_dev(klon, klev, klon, float(ptsphy), C.byref(fld), dqs, None))
            gpu.sync()
            loc = np.empty((_abi.NSTATE, klev, klon))
            gpu.d2h(loc, optr["b_loc"])
            ptent[...], ptenq[...], ptenl[...], pteni[...] = loc[0], loc[2], loc[3], loc[4]
            for arr, n in ((pclc, "pa"), (pcovptot, "pcovptot"), (pfplsl, "pfplsl"), (pfplsn, "pfplsn"),
                           (pfhpsl, "pfhpsl"), (pfhpsn, "pfhpsn")):
                tmp = np.empty(arr.shape)
                gpu.d2h(tmp, optr[n])
                arr[...] = tmp
        finally:
            for p in ptrs:
                gpu.free(p)
