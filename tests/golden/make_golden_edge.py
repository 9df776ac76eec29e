"""Generate tests/golden/nl_pyref_edge.npz: golden vectors of the NL path for EDGE-CASE columns at
another vertical resolution (KLEV = 60), made by the REFERENCE'S OWN Python kernel
(reference src/cloudsc2_nl_gt4py/cloudsc2_py.py: `satur` :12-52, `cloudsc2_py` :54-612).

Columns are the seed-7 synthetic columns with one property forced each -- first-guess temperature
exactly on the thresholds the scheme branches on (RTT, RTICE, RLPTRC, RTT+2), a completely dry and a
strongly supersaturated column, no / very much condensate, no detrainment, PLU exactly ZEPS2, no mass
flux, PSUPSAT > 0, a very cold and a very warm column -- so that every data-dependent IF of
cloudsc2.F90:339-725 is decided on both sides and on its boundary.

Build container only (imports /root/reference):   python tests/golden/make_golden_edge.py
"""
from __future__ import annotations

import importlib
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/src/cloudsc2_nl_gt4py")
sys.path.insert(0, str(Path(__file__).parent))

pkg = importlib.import_module("dwarf-p-cloudsc2-tl-ad_b200")
import cloudsc2_py as ref           # noqa: E402  (the reference's pure-Python NL kernel)
from make_golden import namespaces  # noqa: E402

KLEV, NCOL, SEED = 60, 20, 7


def edge_inputs(prm):
    src = pkg.synth_source(seed=SEED, klon=NCOL, klev=KLEV, params=prm)
    f = {k: np.array(v, dtype=np.float64, copy=True) for k, v in src.f.items()}
    dt = src.ptsphy
    t1 = lambda c: f["pt"][:, c] + dt * f["tend_cml"][0][:, c]     # first-guess T of column c
    def force_t(c, value, levels):
        f["tend_cml"][0][levels, c] = 0.0
        f["pt"][levels, c] = value
    lev = np.arange(KLEV)
    force_t(0, prm.rtt, lev[20::3])                   # T == RTT      (cloudsc2.F90:350,493,528,542, cuadjtqs phase)
    force_t(1, prm.rtice, lev[10::4])                 # T == RTICE    (:404 supersaturation switch)
    force_t(2, prm.rlptrc, lev[15::5])                # T == RLPTRC   (tanh argument 0)
    force_t(3, prm.rtt + 2.0, lev[30::2])             # T == ZMELTP2  (:491 melting threshold)
    f["pq"][:, 4] = 0.0; f["tend_cml"][2][:, 4] = 0.0                      # completely dry
    f["pq"][:, 5] *= 3.0                                                  # strongly supersaturated
    f["pclv"][0][:, 6] = 0.0; f["pclv"][1][:, 6] = 0.0                    # no condensate at all
    f["tend_cml"][3][:, 6] = 0.0; f["tend_cml"][4][:, 6] = 0.0
    f["pclv"][0][:, 7] = 1.0e-3; f["pclv"][1][:, 7] = 5.0e-4              # very much condensate
    f["plude"][:, 8] = 0.0                                                # no detrainment
    f["plu"][:, 9] = 1.0e-10                                              # PLU == ZEPS2 exactly (:436)
    f["plude"][:, 9] = np.maximum(f["plude"][:, 9], 1.0e-6)
    f["plu"][:, 10] = 0.0                                                 # PLU == 0 with detrainment
    f["plude"][:, 10] = 1.0e-6
    f["pmfu"][:, 11] = 0.0; f["pmfd"][:, 11] = 0.0                        # no mass flux
    f["psupsat"][:, 12] = 1.0e-5                                          # PSUPSAT > 0
    f["pt"][:, 13] -= 40.0                                                # very cold
    f["pt"][:, 14] += 20.0                                                # very warm
    f["pq"][:, 15] = 1.0e-12                                              # q below every threshold
    f["tend_cml"][0][:, 16] *= 50.0                                       # violent temperature tendency
    return src, f


def main():
    prm = pkg.default_params()
    src, f = edge_inputs(prm)
    klon, klev = NCOL, KLEV
    yrmcst, yrethf, yrecldp, yrephli, yrecld = namespaces(prm, src.ceta)
    x = {"paphp1": f["paph"], "papp1": f["pap"], "pqm1": f["pq"], "ptm1": f["pt"],
         "pl": f["pclv"][0], "pi": f["pclv"][1], "plude": f["plude"], "plu": f["plu"],
         "pmfu": f["pmfu"], "pmfd": f["pmfd"], "pgtent": f["tend_cml"][0],
         "pgtenq": f["tend_cml"][2], "pgtenl": f["tend_cml"][3], "pgteni": f["tend_cml"][4],
         "psupsat": f["psupsat"]}
    x = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in x.items()}
    pqs = np.zeros((klev, klon))
    t0 = time.time()
    ref.satur(1, klon, klon, 1, klev, True, x["papp1"], x["ptm1"], pqs, 2, yrethf, yrmcst)
    y = {n: np.zeros((klev + (1 if n.startswith("pf") else 0), klon))
         for n in ("ptent", "ptenq", "ptenl", "pteni", "pclc", "pfplsl", "pfplsn", "pfhpsl",
                   "pfhpsn", "pcovptot")}
    plu_pad = np.vstack([x["plu"], np.zeros((1, klon))])      # see make_golden.py
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
    out = {"ceta": src.ceta, "ptsphy": np.float64(src.ptsphy), "pqs": pqs, "seed": np.int64(SEED)}
    out.update({"in_" + k: v for k, v in x.items()})
    out.update({"out_" + k: v for k, v in y.items()})
    dst = Path(__file__).with_name("nl_pyref_edge.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
