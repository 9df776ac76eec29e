"""The JSON line bench.py prints is a contract with the driver: one line, these keys, these types.
CPU: the reference arm (--impl reference, the CPU port of the reference path).  GPU: the B200 arm."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {"metric": str, "value": float, "unit": str, "n_gpus": int, "steps": int, "warmup": int,
             "ms_per_step": float, "higher_is_better": bool, "scaling": str, "dtype": str, "data": str,
             "config": dict, "e2e": dict, "gpu_launches": int}


def _line(*args, timeout=900):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                     # exactly ONE JSON line on stdout
    return json.loads(lines[0])


def _check_base(d):
    for k, t in BASE_KEYS.items():
        assert k in d, k
        assert isinstance(d[k], t), (k, type(d[k]))
    assert d["metric"] == "NL columns/s (KLEV=137)" and d["unit"] == "columns/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "vs_baseline" in d and d["vs_baseline"] is None          # BASELINE.md publishes no number
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k


def test_reference_arm_line(built):
    d = _line("--impl", "reference", "--steps", "1", "--warmup", "0", "--ngptot-per-gpu", "8192")
    _check_base(d)
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb


@pytest.mark.gpu
def test_b200_arm_line(built):
    d = _line("--steps", "3", "--warmup", "3", "--no-cpu", "--no-sweep", "--no-strong", "--e2e-steps", "1",
              "--ngptot-per-gpu", "32768")
    _check_base(d)
    assert d["impl"] == "b200" and d["n_gpus"] == 1 and d["scaling"] == "weak"
    assert d["gpu_launches"] == d["steps"] == 3                      # one NL kernel per timed step
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["peak"] > 1000
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12 and 0.2 < rf["frac"] < 1.0
    assert rf["traffic"] > 0 and rf["algorithmic_bytes_per_column"] == 27440
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert 0 < d["e2e"]["value"] < d["value"]                        # copies inside the timed region
    # round 2: the default e2e runs on caller-allocated arrays registered once; the other host situations beside it
    assert d["e2e"]["host_registered"] is True and d["e2e"]["host_register_ms_once"] > 0
    assert 0 < d["e2e"]["pageable"]["value"] <= d["e2e"]["value"] * 1.2
    assert d["e2e"]["host_alloc"]["value"] > 0
    es = d["e2e_source"]                                             # the dwarf's LOAD -> DRIVER -> VALIDATE in one call
    assert es["value"] > d["e2e"]["value"] and es["h2d_bytes_per_step"] < 1e7 and es["d2h_bytes_per_step"] == 400
    assert es["max_abs_err_vs_unexpanded_columns"] == 0.0 and es["stats_finite"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    for m in ("nl", "tl", "ad", "ad_have_trajectory", "tl_taylor_driver", "ad_test_driver"):
        assert d["modes"][m]["columns_per_s"] > 0, m
    assert d["selftests"]["taylor_passed"] and d["selftests"]["adjoint_passed"]
