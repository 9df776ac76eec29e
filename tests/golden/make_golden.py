"""Generate tests/golden/nl_pyref.npz: golden input/output vectors of the NL path made by the
REFERENCE'S OWN importable Python kernel (reference src/cloudsc2_nl_gt4py/cloudsc2_py.py: `satur`
:12-52 and `cloudsc2_py` :54-612), run here on the synthetic source columns.

Run in the build container only (it imports from /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py
The committed .npz is what the tests read; nothing at test time touches /root/reference.
"""
from __future__ import annotations

import importlib
import sys
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REF_PY = Path("/root/reference/src/cloudsc2_nl_gt4py")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF_PY))

pkg = importlib.import_module("dwarf-p-cloudsc2-tl-ad_b200")
import cloudsc2_py as ref   # noqa: E402  (the reference's pure-Python NL kernel)

NCOL = 32        # golden columns: every 3rd of the 100 seed-0 source columns (+ the last two)
KLEV = 137


def namespaces(prm, ceta):
    g = lambda *names: SimpleNamespace(**{n: getattr(prm, n) for n in names})
    yrmcst = g("rg", "rd", "rcpd", "retv", "rlvtt", "rlstt", "rlmlt", "rtt")
    yrethf = g("r2es", "r3les", "r3ies", "r4les", "r4ies", "r5les", "r5ies", "r5alvcp", "r5alscp",
               "ralvdcp", "ralsdcp", "rtwat", "rtice", "rtwat_rtice_r", "rvtmp2")
    yrecldp = g("rclcrit", "rkconv", "rlmin", "rpecons")
    yrephli = SimpleNamespace(lphylin=True, rlptrc=prm.rlptrc)
    yrecld = SimpleNamespace(ceta=np.asarray(ceta))
    return yrmcst, yrethf, yrecldp, yrephli, yrecld


def main():
    # usage: make_golden.py            -> nl_pyref.npz        (seed 0, inputs + outputs, 32 columns)
    #        make_golden.py SEED       -> nl_pyref_seedN.npz  (outputs only, 48 columns; the inputs
    #                                     are regenerated from the deterministic generator by the tests)
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    prm = pkg.default_params()
    src100 = pkg.synth_source(seed=seed, klon=100, klev=KLEV, params=prm)
    cols = sorted(set(list(range(0, 100, 3))[:NCOL - 2] + [98, 99]))[:NCOL]
    if seed:
        cols = list(range(1, 97, 2))
    src = src100.subset(cols)
    klon, klev = src.klon, src.klev
    yrmcst, yrethf, yrecldp, yrephli, yrecld = namespaces(prm, src.ceta)
    f = src.f
    x = {"paphp1": f["paph"], "papp1": f["pap"], "pqm1": f["pq"], "ptm1": f["pt"],
         "pl": f["pclv"][0], "pi": f["pclv"][1], "plude": f["plude"], "plu": f["plu"],
         "pmfu": f["pmfu"], "pmfd": f["pmfd"], "pgtent": f["tend_cml"][0],
         "pgtenq": f["tend_cml"][2], "pgtenl": f["tend_cml"][3], "pgteni": f["tend_cml"][4],
         "psupsat": f["psupsat"]}
    x = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in x.items()}
    pqs = np.zeros((klev, klon))
    t0 = time.time()
    # satur(kidia,kfdia,klon,ktdia,klev,ldphylin,paprsf,pt,pqsat,kflag,yrethf,yrmcst): driver call
    # cloudsc_driver_mod.F90:91 -> KFLAG=2, LDPHYLIN=.TRUE.
    ref.satur(1, klon, klon, 1, klev, True, x["papp1"], x["ptm1"], pqs, 2, yrethf, yrmcst)
    y = {n: np.zeros((klev + (1 if n.startswith("pf") else 0), klon))
         for n in ("ptent", "ptenq", "ptenl", "pteni", "pclc", "pfplsl", "pfplsn", "pfhpsl",
                   "pfhpsn", "pcovptot")}
    # cloudsc2_py.py:329 tests `jk < klev` with a 0-based jk (the Fortran's 1-based JK<KLEV,
    # cloudsc2.F90:434), so at the last level it indexes plu[klev]: hand it PLU with one extra
    # all-zero row, which makes LLO1 false there exactly as the Fortran's ELSE branch does.
    plu_pad = np.vstack([x["plu"], np.zeros((1, klon))])
    ref.cloudsc2_py(1, klon, klon, 1, klev, False, src.ptsphy, x["paphp1"], x["papp1"], x["pqm1"],
                    pqs, x["ptm1"], x["pl"], x["pi"], x["plude"], plu_pad, x["pmfu"], x["pmfd"],
                    y["ptent"], x["pgtent"], y["ptenq"], x["pgtenq"], y["ptenl"], x["pgtenl"],
                    y["pteni"], x["pgteni"], x["psupsat"], y["pclc"], y["pfplsl"], y["pfplsn"],
                    y["pfhpsl"], y["pfhpsn"], y["pcovptot"],
                    yrecldp, yrecld, yrmcst, yrethf, yrephli)
    print(f"reference python kernel: {klon} columns x {klev} levels in {time.time() - t0:.1f} s")
    for n, v in y.items():
        assert np.isfinite(v).all(), n
        print(f"  {n:9s} min {v.min(): .4e} max {v.max(): .4e} nonzero {np.count_nonzero(v)}")
    out = {"cols": np.asarray(cols), "ceta": src.ceta, "ptsphy": np.float64(src.ptsphy),
           "pqs": pqs, "seed": np.int64(seed)}
    if seed == 0:
        out.update({"in_" + k: v for k, v in x.items()})
    else:
        out["in_checksum"] = np.float64(sum(float(np.abs(v).sum()) for v in x.values()))
    out.update({"out_" + k: v for k, v in y.items()})
    out["params_names"] = np.asarray([n for n, _ in prm._fields_])
    out["params_values"] = np.asarray([float(getattr(prm, n)) for n, _ in prm._fields_])
    dst = Path(__file__).with_name("nl_pyref.npz" if seed == 0 else f"nl_pyref_seed{seed}.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
