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
