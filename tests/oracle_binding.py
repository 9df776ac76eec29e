"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import importlib
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_LIB = ROOT / "oracle" / "_build" / "liboracle.so"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("dwarf-p-cloudsc2-tl-ad_b200")
_abi = pkg._abi

dp = C.POINTER(C.c_double)
IN16 = ("paphp1", "papp1", "pqm1", "pqs", "ptm1", "pl", "pi", "plude", "plu", "pmfu", "pmfd",
        "pgtent", "pgtenq", "pgtenl", "pgteni", "psupsat")
OUT10 = ("ptent", "ptenq", "ptenl", "pteni", "pclc", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn",
         "pcovptot")
HALF = ("paphp1", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn")


class In16(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in IN16]


class Out10(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in OUT10]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not ORACLE_LIB.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
    L = C.CDLL(str(ORACLE_LIB))
    P, F = C.POINTER(_abi.Params), C.POINTER(_abi.Fields)
    i, d = C.c_int, C.c_double
    L.orc_satur.restype = None
    L.orc_satur.argtypes = [P, i, i, i, i, dp, dp, dp]
    L.orc_cloudsc2.restype = i
    L.orc_cloudsc2.argtypes = [P, dp, i, i, i, i, d] + [dp] * 26
    L.orc_cloudsc2tl.restype = i
    L.orc_cloudsc2tl.argtypes = [P, dp, i, i, i, i, d, C.POINTER(In16), C.POINTER(Out10),
                                 C.POINTER(In16), C.POINTER(Out10)]
    L.orc_cloudsc2ad.restype = i
    L.orc_cloudsc2ad.argtypes = L.orc_cloudsc2tl.argtypes
    L.orc_driver_nl.restype = i
    L.orc_driver_nl.argtypes = [P, dp, i, i, i, i, d, F, dp]
    L.orc_driver_tl.restype = i
    L.orc_driver_tl.argtypes = [P, dp, i, i, i, i, d, F, dp, dp, dp]
    L.orc_driver_ad.restype = i
    L.orc_driver_ad.argtypes = [P, dp, i, i, i, i, d, F, dp, dp, dp]
    L.orc_bench_tl.restype = i
    L.orc_bench_tl.argtypes = [P, dp, i, i, i, i, d, F, dp]
    L.orc_bench_ad.restype = i
    L.orc_bench_ad.argtypes = [P, dp, i, i, i, i, d, F, dp]
    L.orc_taylor_verdict.restype = i
    L.orc_taylor_verdict.argtypes = [dp, C.POINTER(i)]
    L.orc_adjoint_verdict.restype = i
    L.orc_adjoint_verdict.argtypes = [d]
    L.orc_max_threads.restype = i
    if hasattr(L, "orc_flavour"):
        L.orc_flavour.restype = C.c_char_p
    _lib = L
    return L


def flavour() -> bytes:
    """b"" for the hand-written restatement, a description for oracle/_ref (f90toc transliteration)."""
    L = lib()
    return L.orc_flavour() if hasattr(L, "orc_flavour") else b""


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(dp)


def max_threads() -> int:
    return int(lib().orc_max_threads())


def satur(prm, pap: np.ndarray, pt: np.ndarray) -> np.ndarray:
    """pap, pt: (KLEV, KLON) -> PQSAT (KLEV, KLON)."""
    klev, klon = pt.shape
    out = np.zeros_like(pt)
    lib().orc_satur(C.byref(prm), 1, klon, klon, klev, _p(pap), _p(pt), _p(out))
    return out


def cloudsc2_block(prm, ceta, ptsphy: float, x: dict) -> dict:
    """One CLOUDSC2 call on (KLEV[+1], KLON) arrays; x has the IN16 names. Returns OUT10 dict."""
    klev, klon = x["ptm1"].shape
    y = {n: np.zeros((klev + (1 if n in HALF else 0), klon)) for n in OUT10}
    rc = lib().orc_cloudsc2(
        C.byref(prm), _p(np.ascontiguousarray(ceta)), 1, klon, klon, klev, ptsphy,
        _p(x["paphp1"]), _p(x["papp1"]), _p(x["pqm1"]), _p(x["pqs"]), _p(x["ptm1"]), _p(x["pl"]),
        _p(x["pi"]), _p(x["plude"]), _p(x["plu"]), _p(x["pmfu"]), _p(x["pmfd"]),
        _p(y["ptent"]), _p(x["pgtent"]), _p(y["ptenq"]), _p(x["pgtenq"]), _p(y["ptenl"]),
        _p(x["pgtenl"]), _p(y["pteni"]), _p(x["pgteni"]), _p(x["psupsat"]), _p(y["pclc"]),
        _p(y["pfplsl"]), _p(y["pfplsn"]), _p(y["pfhpsl"]), _p(y["pfhpsn"]), _p(y["pcovptot"]))
    if rc:
        raise RuntimeError(f"orc_cloudsc2 rc={rc}")
    return y


def _s16(x: dict) -> In16:
    s = In16()
    for n in IN16:
        setattr(s, n, x[n].ctypes.data)
    return s


def _s10(y: dict) -> Out10:
    s = Out10()
    for n in OUT10:
        setattr(s, n, y[n].ctypes.data)
    return s


def alloc16(klev, klon, fill=0.0):
    return {n: np.full((klev + (1 if n in HALF else 0), klon), fill) for n in IN16}


def alloc10(klev, klon, fill=0.0):
    return {n: np.full((klev + (1 if n in HALF else 0), klon), fill) for n in OUT10}


def cloudsc2tl_block(prm, ceta, ptsphy, x5: dict, dx: dict):
    """CLOUDSC2TL on one block -> (y5, dy)."""
    klev, klon = x5["ptm1"].shape
    y5, dy = alloc10(klev, klon), alloc10(klev, klon)
    a5, b5, a, b = _s16(x5), _s10(y5), _s16(dx), _s10(dy)
    rc = lib().orc_cloudsc2tl(C.byref(prm), _p(np.ascontiguousarray(ceta)), 1, klon, klon, klev,
                              ptsphy, C.byref(a5), C.byref(b5), C.byref(a), C.byref(b))
    if rc:
        raise RuntimeError(f"orc_cloudsc2tl rc={rc}")
    return y5, dy


def cloudsc2ad_block(prm, ceta, ptsphy, x5: dict, dx_ad: dict, dy_ad: dict):
    """CLOUDSC2AD on one block: dx_ad accumulated in place, dy_ad consumed+zeroed. Returns y5."""
    klev, klon = x5["ptm1"].shape
    y5 = alloc10(klev, klon)
    a5, b5, a, b = _s16(x5), _s10(y5), _s16(dx_ad), _s10(dy_ad)
    rc = lib().orc_cloudsc2ad(C.byref(prm), _p(np.ascontiguousarray(ceta)), 1, klon, klon, klev,
                              ptsphy, C.byref(a5), C.byref(b5), C.byref(a), C.byref(b))
    if rc:
        raise RuntimeError(f"orc_cloudsc2ad rc={rc}")
    return y5


def block_inputs(st, ibl: int, prm) -> dict:
    """IN16 dict (fresh contiguous copies) of block ibl of an ArrayState, PQS from orc_satur."""
    a = st.a
    x = {"paphp1": a["paph"][ibl], "papp1": a["pap"][ibl], "pqm1": a["pq"][ibl],
         "ptm1": a["pt"][ibl], "pl": a["pclv"][ibl, 0], "pi": a["pclv"][ibl, 1],
         "plude": a["plude"][ibl], "plu": a["plu"][ibl], "pmfu": a["pmfu"][ibl],
         "pmfd": a["pmfd"][ibl], "pgtent": a["b_cml"][ibl, 0], "pgtenq": a["b_cml"][ibl, 2],
         "pgtenl": a["b_cml"][ibl, 3], "pgteni": a["b_cml"][ibl, 4], "psupsat": a["psupsat"][ibl]}
    x = {k: np.ascontiguousarray(v).copy() for k, v in x.items()}
    x["pqs"] = satur(prm, x["papp1"], x["ptm1"])
    return x


def driver_nl(prm, ceta, st, numomp: int = 1) -> float:
    """CLOUDSC_DRIVER on an ArrayState (outputs written into st.a). Returns block-loop seconds."""
    t = C.c_double(0)
    f = st.fields()
    rc = lib().orc_driver_nl(C.byref(prm), _p(np.ascontiguousarray(ceta)), numomp, st.nproma,
                             st.klev, st.ngptot, st.ptsphy, C.byref(f), C.byref(t))
    if rc:
        raise RuntimeError(f"orc_driver_nl rc={rc}")
    return t.value


def driver_tl(prm, ceta, st, numomp: int = 1, allow_degenerate: bool = False):
    """CLOUDSC_DRIVER_TL -> (znormg[10], ratios_blk[nblocks,10], seconds)."""
    t = C.c_double(0)
    z = np.zeros(10)
    rb = np.zeros((st.nblocks, 10))
    f = st.fields()
    rc = lib().orc_driver_tl(C.byref(prm), _p(np.ascontiguousarray(ceta)), numomp, st.nproma,
                             st.klev, st.ngptot, st.ptsphy, C.byref(f), _p(z), _p(rb), C.byref(t))
    if rc and not (allow_degenerate and rc == 3):
        raise RuntimeError(f"orc_driver_tl rc={rc}")
    return z, rb, t.value


def driver_ad(prm, ceta, st, numomp: int = 1):
    """CLOUDSC_DRIVER_AD -> (znormg, norms_col[ngptot,3], seconds)."""
    t = C.c_double(0)
    zn = C.c_double(0)
    nc = np.zeros((st.ngptot, 3))
    f = st.fields()
    rc = lib().orc_driver_ad(C.byref(prm), _p(np.ascontiguousarray(ceta)), numomp, st.nproma,
                             st.klev, st.ngptot, st.ptsphy, C.byref(f), C.byref(zn), _p(nc),
                             C.byref(t))
    if rc:
        raise RuntimeError(f"orc_driver_ad rc={rc}")
    return zn.value, nc, t.value


def bench_tlad(which: str, prm, ceta, st, numomp: int = 1) -> float:
    t = C.c_double(0)
    f = st.fields()
    fn = lib().orc_bench_tl if which == "tl" else lib().orc_bench_ad
    rc = fn(C.byref(prm), _p(np.ascontiguousarray(ceta)), numomp, st.nproma, st.klev, st.ngptot,
            st.ptsphy, C.byref(f), C.byref(t))
    if rc:
        raise RuntimeError(f"orc_bench_{which} rc={rc}")
    return t.value


def taylor_verdict(z):
    z = np.ascontiguousarray(z, dtype=np.float64)
    ist = C.c_int(0)
    pen = lib().orc_taylor_verdict(_p(z), C.byref(ist))
    return int(pen), int(ist.value)
