"""The eighteen perturbation directions of tests/golden/tl_fd_dirs.npz, shared by the generator
(tests/golden/make_golden_tl_dirs.py, build container only) and the CPU / GPU parity tests.
Nothing here touches /root/reference."""
import numpy as np

IN16 = ("paphp1", "papp1", "pqm1", "pqs", "ptm1", "pl", "pi", "plude", "plu", "pmfu", "pmfd",
        "pgtent", "pgtenq", "pgtenl", "pgteni", "psupsat")
OUT7 = ("ptent", "ptenq", "ptenl", "pteni", "pclc", "pfplsl", "pfplsn")
NCOL = 8
NAMES = IN16 + ("random0", "random1")


def directions(x0):
    """k = 0..15: only input k perturbed (0.01 x_k; PSUPSAT, identically zero, borrows 1e-3 PQM1 as its
    scale); k = 16, 17: all inputs with independent random factors in (-1, 1) per element."""
    zero = {k: np.zeros_like(x0[k]) for k in IN16}
    dirs = []
    for k in IN16:
        d = dict(zero)
        s = x0[k] if np.abs(x0[k]).max() > 0 else 1e-3 * x0["pqm1"]
        d[k] = 0.01 * s
        dirs.append(d)
    for seed in (101, 202):
        rng = np.random.default_rng(seed)
        d = {}
        for k in IN16:
            s = x0[k] if np.abs(x0[k]).max() > 0 else 1e-3 * x0["pqm1"]
            d[k] = 0.01 * s * rng.uniform(-1.0, 1.0, s.shape)
        # keep the half-level pressures monotone: one factor per column for PAPHP1
        d["paphp1"] = 0.01 * x0["paphp1"] * rng.uniform(-1.0, 1.0, (1, x0["paphp1"].shape[1]))
        dirs.append(d)
    return [{k: np.ascontiguousarray(v) for k, v in d.items()} for d in dirs]
