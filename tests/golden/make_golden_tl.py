"""Generate tests/golden/tl_fd_pyref.npz: central finite differences of the REFERENCE'S OWN Python
nonlinear kernel (reference src/cloudsc2_nl_gt4py/cloudsc2_py.py) along the drivers' perturbation
direction dx = 0.01 x (cloudsc_driver_tl_mod.F90:156-171):

    D = ( F(x + eps dx) - F(x - eps dx) ) / (2 eps),     eps = 1e-4  (dx is 1 % of x -> 1e-6 relative)

for the 32 golden columns of nl_pyref.npz.  The reference ships no tangent-linear code that can be
run here (cloudsc2tl.F90 needs a Fortran compiler), so this derivative of its nonlinear Python kernel
is the independent anchor the TL restatement (oracle) and the TL CUDA kernel are tested against.
PQS is an input of CLOUDSC2 / CLOUDSC2TL with its own increment (0.01 PQS), exactly as in the
reference's Taylor test (:204), so it is perturbed, not recomputed.

Run in the build container only:   python tests/golden/make_golden_tl.py
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
sys.path.insert(0, "/root/reference/src/cloudsc2_nl_gt4py")

import cloudsc2_py as ref          # noqa: E402  (the reference's pure-Python NL kernel)
from make_golden import namespaces, pkg   # noqa: E402

EPS = 1e-4
OUT = ("ptent", "ptenq", "ptenl", "pteni", "pclc", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn", "pcovptot")
IN16 = ("paphp1", "papp1", "pqm1", "pqs", "ptm1", "pl", "pi", "plude", "plu", "pmfu", "pmfd",
        "pgtent", "pgtenq", "pgtenl", "pgteni", "psupsat")


def run(x, ptsphy, ns, klev, klon):
    yrmcst, yrethf, yrecldp, yrephli, yrecld = ns
    y = {n: np.zeros((klev + (1 if n.startswith("pf") else 0), klon)) for n in OUT}
    plu_pad = np.vstack([x["plu"], np.zeros((1, klon))])
    ref.cloudsc2_py(1, klon, klon, 1, klev, False, ptsphy, x["paphp1"], x["papp1"], x["pqm1"],
                    x["pqs"], x["ptm1"], x["pl"], x["pi"], x["plude"], plu_pad, x["pmfu"], x["pmfd"],
                    y["ptent"], x["pgtent"], y["ptenq"], x["pgtenq"], y["ptenl"], x["pgtenl"],
                    y["pteni"], x["pgteni"], x["psupsat"], y["pclc"], y["pfplsl"], y["pfplsn"],
                    y["pfhpsl"], y["pfhpsn"], y["pcovptot"], yrecldp, yrecld, yrmcst, yrethf, yrephli)
    return y


def inputs_of(g):
    """The 15 inputs of a golden set: stored (seed 0) or regenerated from the deterministic generator."""
    if "in_ptm1" in g.files:
        return {k[3:]: np.ascontiguousarray(g[k]) for k in g.files if k.startswith("in_")}
    f = pkg.synth_source(seed=int(g["seed"]), klon=100, klev=137).subset(list(g["cols"])).f
    x = {"paphp1": f["paph"], "papp1": f["pap"], "pqm1": f["pq"], "ptm1": f["pt"], "pl": f["pclv"][0],
         "pi": f["pclv"][1], "plude": f["plude"], "plu": f["plu"], "pmfu": f["pmfu"], "pmfd": f["pmfd"],
         "pgtent": f["tend_cml"][0], "pgtenq": f["tend_cml"][2], "pgtenl": f["tend_cml"][3],
         "pgteni": f["tend_cml"][4], "psupsat": f["psupsat"]}
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in x.items()}


def main():
    # usage: make_golden_tl.py          -> tl_fd_pyref.npz        (the 32 columns of nl_pyref.npz)
    #        make_golden_tl.py seed5    -> tl_fd_pyref_seed5.npz  (the 48 columns of nl_pyref_seed5.npz)
    tag = sys.argv[1] if len(sys.argv) > 1 else ""
    g = np.load(Path(__file__).with_name(f"nl_pyref_{tag}.npz" if tag else "nl_pyref.npz"))
    prm = pkg.default_params()
    x0 = inputs_of(g)
    x0["pqs"] = np.ascontiguousarray(g["pqs"])
    klev, klon = x0["ptm1"].shape
    ns = namespaces(prm, g["ceta"])
    ptsphy = float(g["ptsphy"])
    dx = {k: 0.01 * x0[k] for k in IN16}
    t0 = time.time()
    yp = run({k: x0[k] + EPS * dx[k] for k in IN16}, ptsphy, ns, klev, klon)
    ym = run({k: x0[k] - EPS * dx[k] for k in IN16}, ptsphy, ns, klev, klon)
    y0 = run(dict(x0), ptsphy, ns, klev, klon)
    print(f"3 runs of the reference python kernel: {time.time() - t0:.1f} s")
    out = {"eps": np.float64(EPS)}
    for n in OUT:
        d = (yp[n] - ym[n]) / (2 * EPS)
        # second difference: large where the path is not smooth (a branch flips inside +-eps dx)
        curv = np.abs(yp[n] - 2 * y0[n] + ym[n])
        out["d_" + n] = d
        out["curv_" + n] = curv
        print(f"  {n:9s} max|D| {np.abs(d).max():.4e}   max second difference {curv.max():.3e}")
    dst = Path(__file__).with_name(f"tl_fd_pyref_{tag}.npz" if tag else "tl_fd_pyref.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
