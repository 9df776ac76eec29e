"""pytest configuration: the `gpu` marker and shared fixtures.

CPU suite  : python -m pytest tests/ -x -q -m "not gpu"   (oracle vs golden vectors, host logic, ABI)
GPU suite  : python -m pytest tests/ -x -q -m gpu          (CUDA path vs oracle through the C ABI)
Nothing here reads /root/reference at run time.
"""
import importlib
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the native artefacts exist (no-op when they are up to date)."""
    import __graft_entry__ as ge
    ge.build_product()
    ge.build_programs()
    ge.build_oracle()
    return ge


@pytest.fixture(scope="session")
def pkg(built):
    return importlib.import_module("dwarf-p-cloudsc2-tl-ad_b200")


@pytest.fixture(scope="session")
def ob(built):
    from tests import oracle_binding
    return oracle_binding


@pytest.fixture(scope="session")
def obref(built):
    """The same binding on oracle/_ref/libcloudsc2_ref.so: the reference's own Fortran kernels,
    transliterated to C by oracle/f90toc.py (built in the container that has /root/reference; the
    prebuilt .so travels to the GPU box)."""
    import importlib.util
    path = built.build_oracle_ref()
    if path is None:
        pytest.skip("oracle/_ref is not built and the reference sources are not here to build it")
    spec = importlib.util.spec_from_file_location("oracle_binding_ref", ROOT / "tests" / "oracle_binding.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.ORACLE_LIB = path
    assert b"f90toc" in mod.flavour()
    return mod


@pytest.fixture(scope="session")
def src100(pkg):
    return pkg.synth_source(seed=0, klon=100, klev=137)


@pytest.fixture(scope="session")
def golden():
    return np.load(ROOT / "tests" / "golden" / "nl_pyref.npz")


@pytest.fixture(scope="session")
def gpu_nl(pkg, src100):
    """GPU context with the NL/TL program switches (LREGCL=.FALSE.)."""
    if not pkg.gpu_available():
        pytest.fail("GPU test selected but libcloudsc2_b200.so sees no CUDA device (no CPU fallback)")
    g = pkg.Cloudsc2(pkg.default_params(lregcl=False), src100.klev, src100.ceta)
    yield g
    g.close()


class _GoldenView:
    """npz-like view of a golden set whose inputs are either stored (seed 0) or regenerated from the
    deterministic generator (other seeds store outputs only)."""

    def __init__(self, g, extra):
        self._g, self._x = g, extra
        self.files = [k for k in g.files if k != "in_checksum"] + list(extra)

    def __getitem__(self, k):
        return self._x[k] if k in self._x else self._g[k]


@pytest.fixture(scope="session", params=["seed0", "seed5"])
def golden_fd(request, pkg):
    """(golden set of the reference's Python NL kernel, central finite differences of that kernel along
    dx = 0.01 x) -- the anchors of the TL / AD parity tests; two atmospheres."""
    gdir = ROOT / "tests" / "golden"
    if request.param == "seed0":
        return np.load(gdir / "nl_pyref.npz"), np.load(gdir / "tl_fd_pyref.npz")
    g = np.load(gdir / "nl_pyref_seed5.npz")
    f = pkg.synth_source(seed=int(g["seed"]), klon=100, klev=137).subset(list(g["cols"])).f
    x = {"in_paphp1": f["paph"], "in_papp1": f["pap"], "in_pqm1": f["pq"], "in_ptm1": f["pt"],
         "in_pl": f["pclv"][0], "in_pi": f["pclv"][1], "in_plude": f["plude"], "in_plu": f["plu"],
         "in_pmfu": f["pmfu"], "in_pmfd": f["pmfd"], "in_pgtent": f["tend_cml"][0],
         "in_pgtenq": f["tend_cml"][2], "in_pgtenl": f["tend_cml"][3], "in_pgteni": f["tend_cml"][4],
         "in_psupsat": f["psupsat"]}
    x = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in x.items()}
    return _GoldenView(g, x), np.load(gdir / "tl_fd_pyref_seed5.npz")
