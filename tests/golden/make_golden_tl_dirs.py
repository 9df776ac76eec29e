"""Generate tests/golden/tl_fd_dirs.npz: central finite differences of the REFERENCE'S OWN Python
nonlinear kernel (reference src/cloudsc2_nl_gt4py/cloudsc2_py.py) along EIGHTEEN directions --

    k = 0..15 : only input k of the 16 inputs of CLOUDSC2TL is perturbed (dx_k = 0.01 s_k, s_k = x_k,
                or the size of a neighbouring field where x_k is identically zero: PSUPSAT)
    k = 16,17 : all 16 inputs at once with independent random signs and sizes per element
                (dx = 0.01 x r, r ~ U(-1,1), seeds 101 / 202)

for the first 8 golden columns of nl_pyref.npz.  tl_fd_pyref.npz perturbs all inputs along ONE direction
(the drivers' dx = 0.01 x), which cannot see errors that compensate between inputs; these directions
look at every column of the Jacobian separately.  The reference's TL code cannot run here (Fortran),
so this derivative of its nonlinear Python kernel is the independent anchor.

Stored per direction: the derivative of PTENT, PTENQ, PTENL, PTENI, PCLC, PFPLSL, PFPLSN (PCOVPTOT is
identically zero and PFHPSL/N = -L * PFPLSL/N exactly), and a mask of the points where the second
difference says the path is smooth (no branch flips inside +-eps dx).

Run in the build container only:   python tests/golden/make_golden_tl_dirs.py
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

from make_golden import namespaces, pkg        # noqa: E402
from make_golden_tl import IN16, run, inputs_of   # noqa: E402

from tests.fd_directions import NCOL, OUT7, directions   # noqa: E402

EPS = 1e-4


def main():
    g = np.load(Path(__file__).with_name("nl_pyref.npz"))
    prm = pkg.default_params()
    x0 = inputs_of(g)
    x0["pqs"] = np.ascontiguousarray(g["pqs"])
    x0 = {k: np.ascontiguousarray(v[:, :NCOL]) for k, v in x0.items()}
    klev, klon = x0["ptm1"].shape
    ns = namespaces(prm, g["ceta"])
    ptsphy = float(g["ptsphy"])
    t0 = time.time()
    y0 = run(dict(x0), ptsphy, ns, klev, klon)
    out = {"eps": np.float64(EPS), "ncol": np.int64(NCOL)}
    for i, dx in enumerate(directions(x0)):
        yp = run({k: x0[k] + EPS * dx[k] for k in IN16}, ptsphy, ns, klev, klon)
        ym = run({k: x0[k] - EPS * dx[k] for k in IN16}, ptsphy, ns, klev, klon)
        for n in OUT7:
            d = (yp[n] - ym[n]) / (2 * EPS)
            curv = np.abs(yp[n] - 2 * y0[n] + ym[n])
            # smooth where the second difference is at rounding level relative to the first
            smooth = curv <= 1e-5 * EPS * np.maximum(np.abs(d).max(), 1e-300) + 64 * 2.3e-16 * np.abs(y0[n]).max()
            out[f"d{i:02d}_{n}"] = d
            out[f"m{i:02d}_{n}"] = smooth
        name = IN16[i] if i < 16 else f"random{i - 16}"
        print(f"  direction {i:2d} {name:9s} max|D PTENT| {np.abs(out[f'd{i:02d}_ptent']).max():.3e} "
              f"smooth {np.mean([out[f'm{i:02d}_{n}'].mean() for n in OUT7]):.4f}")
    print(f"{2 * 18 + 1} runs of the reference python kernel: {time.time() - t0:.1f} s")
    dst = Path(__file__).with_name("tl_fd_dirs.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
