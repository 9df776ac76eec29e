"""The three dwarf programs (host/dwarf_cloudsc2.cc, the C++ mirror of PROGRAM DWARF_CLOUDSC) and the
input.h5 / reference.h5 loaders behind them.

CPU part: loaders against files written by tests/h5writer.py, command-line handling, loud failure
without a GPU.  GPU part: BASELINE configs 1-3 run as the reference's own command lines
(README.md:49-61: `dwarf-cloudsc2-nl 4 160000 32`, `dwarf-cloudsc2-tl 1 100 1`,
`dwarf-cloudsc2-ad 1 100 100`) and checked through what the programs print.
"""
import ctypes as C
import os
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests.h5writer import PARAM_DATASETS, input_h5_datasets, write_h5

ROOT = Path(__file__).resolve().parent.parent


def _run(built, prog, *args, env=None, cwd=None):
    exe = ROOT / "dwarf-p-cloudsc2-tl-ad_b200" / "bin" / prog
    assert exe.exists(), exe
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([str(exe), *map(str, args)], capture_output=True, text=True, env=e, cwd=cwd,
                          timeout=600)


# ---- CPU: loaders -------------------------------------------------------------------------------

def test_input_h5_loader_roundtrip(pkg, tmp_path):
    src = pkg.synth_source(seed=3, klon=7, klev=12)
    prm = pkg.default_params()
    prm.rclcrit, prm.rtice, prm.rlptrc = 3.0e-4, 250.16, 266.0      # not the defaults
    path = tmp_path / "input.h5"
    d = input_h5_datasets(src, prm)
    assert len(d) > 32                                               # several symbol-table nodes
    write_h5(path, d)
    got, gp = pkg.load_source_h5(path)
    assert (got.klon, got.klev, got.ptsphy) == (7, 12, src.ptsphy)
    for n, a in src.f.items():
        assert np.array_equal(got.f[n], a), n
    assert np.array_equal(got.ceta, src.ceta)                        # dwarf_cloudsc.F90:100-102
    for member in PARAM_DATASETS.values():
        assert getattr(gp, member) == getattr(prm, member), member
    assert gp.rvtmp2 == 0.0 and gp.lphylin == 1 and gp.levapls2 == 0
    # a missing dataset is an error naming it (the reference aborts in LOAD_ARRAY)
    del d["PMFD"]
    write_h5(path, d)
    with pytest.raises(KeyError, match="PMFD"):
        pkg.load_source_h5(path)
    # a mis-shaped one too
    d["PMFD"] = np.zeros((3, 3))
    write_h5(path, d)
    with pytest.raises(KeyError, match="PMFD"):
        pkg.load_source_h5(path)
    with pytest.raises(KeyError):
        pkg.load_source_h5(tmp_path / "absent.h5")


def test_input_h5_writer_roundtrip(pkg, tmp_path):
    """The library's own HDF5 writer (cloudsc2_source_write_h5): everything the reference's loaders read
    is in the file, and what reaches the kernels reads back bit-identically."""
    src = pkg.synth_source(seed=1, klon=9, klev=11)
    prm = pkg.default_params()
    path = tmp_path / "input.h5"
    pkg.write_source_h5(src, prm, path)
    got, gp = pkg.load_source_h5(path)
    for n, a in src.f.items():
        assert np.array_equal(got.f[n], a), n
    for member in PARAM_DATASETS.values():
        assert getattr(gp, member) == getattr(prm, member), member
    assert got.ptsphy == src.ptsphy
    # datasets CLOUDSC2 never uses but LOAD aborts without (yoecldp.F90:244-369, yoephli.F90:81-96, ...)
    for name in ("YRECLDP_RAMID", "YRECLDP_RCL_FZRBB", "YREPHLI_RLPP00", "RKOOP2", "RV", "RALFDCP"):
        assert pkg.read_h5_f8(path, name).size == 1, name
    for name in ("YRECLDP_NBETA", "YREPHLI_LPHYLIN", "LDSLPHY", "LDMAINCALL", "YRECLDP_LAERICEAUTO"):
        assert pkg.read_h5_i4(path, name).size == 1, name
    assert pkg.read_h5_f8(path, "YRECLDP_RBETA").shape == (101,)
    assert pkg.read_h5_i4(path, "YREPHLI_LPHYLIN")[0] == 1
    assert np.isclose(pkg.read_h5_f8(path, "RV")[0], 461.5249933083879)


def test_reference_h5_loader(pkg, tmp_path):
    rng = np.random.default_rng(11)
    klon, klev = 5, 9
    d = {"KLON": np.array([klon], dtype=np.int32), "KLEV": np.array([klev], dtype=np.int32)}
    for n in ("PLUDE", "PCOVPTOT", "TENDENCY_LOC_T", "TENDENCY_LOC_A", "TENDENCY_LOC_Q"):
        d[n] = rng.standard_normal((klev, klon))
    for n in ("PFPLSL", "PFPLSN", "PFHPSL", "PFHPSN"):
        d[n] = rng.standard_normal((klev + 1, klon))
    d["TENDENCY_LOC_CLD"] = rng.standard_normal((5, klev, klon))
    path = tmp_path / "reference.h5"
    write_h5(path, d)
    lib = pkg.load_library()
    r = pkg._abi.Reference()
    assert lib.cloudsc2_reference_load_h5(C.byref(r), str(path).encode()) == 0
    try:
        assert (r.klon, r.klev) == (klon, klev)
        n = klon * klev
        assert np.array_equal(np.ctypeslib.as_array(r.pfplsn, shape=(klev + 1, klon)), d["PFPLSN"])
        tl = np.ctypeslib.as_array(r.tend_loc, shape=(8, klev, klon))
        assert np.array_equal(tl[0], d["TENDENCY_LOC_T"]) and np.array_equal(tl[1], d["TENDENCY_LOC_A"])
        assert np.array_equal(tl[2], d["TENDENCY_LOC_Q"]) and np.array_equal(tl[3:], d["TENDENCY_LOC_CLD"])
        assert n > 0
        # and the writer (WRITE_REFERENCE) reproduces the file's content
        assert lib.cloudsc2_reference_write_h5(C.byref(r), str(tmp_path / "again.h5").encode()) == 0
        for n in d:
            rd = pkg.read_h5_i4 if d[n].dtype == np.int32 else pkg.read_h5_f8
            assert np.array_equal(rd(tmp_path / "again.h5", n), d[n]), n
    finally:
        lib.cloudsc2_reference_free(C.byref(r))
    ref = Path("/root/reference/config-files/reference.h5")           # build container only
    if ref.exists():
        assert lib.cloudsc2_reference_load_h5(C.byref(r), str(ref).encode()) == 0
        assert (r.klon, r.klev) == (100, 137)
        lib.cloudsc2_reference_free(C.byref(r))


def test_shipped_reference_of_the_synthetic_columns_is_the_cpu_oracle(pkg, ob, src100):
    """config-files/reference_synth_seed0.h5 (what dwarf-cloudsc2-nl validates the default synthetic run against)
    is exactly what tests/golden/make_reference_h5.py says: the CPU restatement's CLOUDSC_DRIVER on the 100
    columns of seed 0 as one block, in the reference's reference.h5 format."""
    path = ROOT / "dwarf-p-cloudsc2-tl-ad_b200" / "config-files" / "reference_synth_seed0.h5"
    assert path.exists() and path.stat().st_size < 2_000_000
    st = pkg.ArrayState(src100, nproma=100, ngptot=100)
    ob.driver_nl(pkg.default_params(lregcl=False), src100.ceta, st, numomp=1)
    assert pkg.read_h5_i4(path, "KLON")[0] == 100 and pkg.read_h5_i4(path, "KLEV")[0] == 137
    for name, want in (("PLUDE", st.a["plude"][0]), ("PCOVPTOT", st.a["pcovptot"][0]), ("PFPLSL", st.a["pfplsl"][0]),
                       ("PFPLSN", st.a["pfplsn"][0]), ("PFHPSL", st.a["pfhpsl"][0]), ("PFHPSN", st.a["pfhpsn"][0]),
                       ("TENDENCY_LOC_T", st.a["b_loc"][0][0]), ("TENDENCY_LOC_A", st.a["b_loc"][0][1]),
                       ("TENDENCY_LOC_Q", st.a["b_loc"][0][2]), ("TENDENCY_LOC_CLD", st.a["b_loc"][0][3:])):
        assert np.array_equal(pkg.read_h5_f8(path, name), want), name
    assert np.abs(st.a["pfplsn"]).max() > 0


# ---- CPU: command line ----------------------------------------------------------------------------

def test_programs_command_line(built, pkg):
    for prog in ("dwarf-cloudsc2-nl", "dwarf-cloudsc2-tl", "dwarf-cloudsc2-ad"):
        r = _run(built, prog, "--help")
        assert r.returncode == 0 and "NUMOMP" in r.stdout
        r = _run(built, prog, "1", "abc")                             # READ(CLARG,*) would abort
        assert r.returncode == 1 and "ABOR1" in r.stderr and "NGPTOT" in r.stderr
    if not pkg.gpu_available():
        r = _run(built, "dwarf-cloudsc2-nl", "4", "1600", "32")
        assert r.returncode == 1 and "no CPU fallback" in r.stderr   # never computes on the host


# ---- GPU: the BASELINE command lines ----------------------------------------------------------------

_ERR_LINE = re.compile(r"^ (\S+)\s+(\d)D(\d)((?:\s+-?[0-9.]+E[+-]\d+){5})(\s+!!!!)?\s*$")


def _validation_lines(stdout):
    out = {}
    for line in stdout.splitlines():
        m = _ERR_LINE.match(line)
        if m:
            out[m.group(1)] = ([float(x) for x in m.group(4).split()], m.group(5) is not None)
    return out


@pytest.mark.gpu
def test_dwarf_nl_baseline_config(built, pkg):
    """dwarf-cloudsc2-nl 4 160000 32 (BASELINE config 1 / README.md:49)."""
    r = _run(built, "dwarf-cloudsc2-nl", 4, 160000, 32, env={"CLOUDSC2_REPEAT": "3"})
    assert r.returncode == 0, r.stderr
    assert "NUMPROC=1, NUMOMP=4, NGPTOTG=160000, NPROMA=32, NGPBLKS=5000" in r.stderr
    tot = [l for l in r.stderr.splitlines() if l.rstrip().endswith(": TOTAL")]
    assert len(tot) == 1
    nums = tot[0].replace("x", " ").replace(":", " ").split()
    assert nums[:6] == ["1", "4", "160000", "160000", "5000", "32"]
    # the default synthetic columns are validated against the SHIPPED reference: the CPU restatement's results
    # (config-files/reference_synth_seed0.h5, made by tests/golden/make_reference_h5.py) -- an independent check
    assert "reference: config-files/reference_synth_seed0.h5" in r.stdout and "UNVERIFIED" not in r.stdout
    v = _validation_lines(r.stdout)
    want = ["PLUDE", "PCOVPTOT", "PFPLSL", "PFPLSN", "PFHPSL", "PFHPSN", "TENDENCY_LOC%A",
            "TENDENCY_LOC%Q", "TENDENCY_LOC%T", "TENDENCY_LOC%CLD"]
    assert list(v) == want                                            # the reference's order (:239-251)
    for name, (nums5, warn) in v.items():
        assert nums5[4] < 1e-9, (name, nums5)                         # MaxRelErr-% : GPU vs CPU, last bits only
    assert v["PFPLSN"][0][1] > 0 and v["TENDENCY_LOC%T"][0][1] > 0 and v["PFPLSN"][0][2] > 0   # not trivially equal
    # CLOUDSC2_REFERENCE=self: the GPU's own one-block results as reference -- bit-identical, and labelled as what it is
    r2 = _run(built, "dwarf-cloudsc2-nl", 4, 160000, 32, env={"CLOUDSC2_REFERENCE": "self"})
    assert r2.returncode == 0 and "SELF-CONSISTENCY only" in r2.stdout and "UNVERIFIED" in r2.stdout
    for name, (nums5, warn) in _validation_lines(r2.stdout).items():
        assert not warn and nums5[2] == 0.0 and nums5[4] == 0.0, (name, nums5)   # bit-identical to 1 block
    m = re.search(r"GPU: ([0-9.]+) ms per driver call", r.stderr)
    assert m and float(m.group(1)) < 50.0


@pytest.mark.gpu
def test_dwarf_gpus_do_not_change_results(built, pkg):
    """CLOUDSC2_NGPUS=N: ONE process drives N GPUs through the library's device set (no fork, no pipes):
    a device expands from its global column offset, so the validation table, the Taylor ratios' verdict
    and the adjoint norm are those of the one-GPU run; the timer table has one line per GPU."""
    ndev = int(pkg.load_library().cloudsc2_gpu_device_count())
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    n = min(ndev, 4)
    one = _run(built, "dwarf-cloudsc2-nl", 1, 10000, 64, env={"CLOUDSC2_REFERENCE": "self"})
    many = _run(built, "dwarf-cloudsc2-nl", 1, 10000, 64, env={"CLOUDSC2_NGPUS": str(n), "CLOUDSC2_REFERENCE": "self"})
    assert one.returncode == 0 and many.returncode == 0, many.stderr
    assert f"NUMPROC=1, NUMOMP=1, NGPTOTG=10000, NPROMA=64, NGPBLKS=157   (GPUs: {n})" in many.stderr
    per_gpu = [l for l in many.stderr.splitlines() if re.search(r": GPU \d+$", l)]
    assert len(per_gpu) == n and sum(int(l.split()[2]) for l in per_gpu) == 10000
    assert f"ranks={n}" in many.stderr
    v1, vn = _validation_lines(one.stdout), _validation_lines(many.stdout)
    assert list(v1) == list(vn) and len(vn) == 10
    for name in v1:
        assert v1[name][0][:3] == vn[name][0][:3] and vn[name][0][2] == 0.0 and not vn[name][1], name
    hosted = _run(built, "dwarf-cloudsc2-nl", 1, 10000, 64, env={"CLOUDSC2_NGPUS": str(n), "CLOUDSC2_HOST_ARRAYS": "1",
                                                                   "CLOUDSC2_REFERENCE": "self"})
    assert hosted.returncode == 0 and _validation_lines(hosted.stdout) == vn
    a1 = _run(built, "dwarf-cloudsc2-ad", 1, 1000, 50)
    a2 = _run(built, "dwarf-cloudsc2-ad", 1, 1000, 50, env={"CLOUDSC2_NGPUS": "2"})
    pat = r"The maximum error is\s+([0-9.]+)"
    assert a1.returncode == 0 and a2.returncode == 0
    assert float(re.search(pat, a1.stdout).group(1)) == float(re.search(pat, a2.stdout).group(1))
    t1 = _run(built, "dwarf-cloudsc2-tl", 1, 200, 1)
    t2 = _run(built, "dwarf-cloudsc2-tl", 1, 200, 1, env={"CLOUDSC2_NGPUS": "2"})
    assert t2.returncode == 0 and "TEST PASSED" in t2.stdout
    rows = lambda r: re.findall(r"^\s+(\d+)\s+([0-9.]+)\s*$", r.stdout, flags=re.M)
    assert rows(t1) == rows(t2)


@pytest.mark.gpu
def test_dwarf_tl_baseline_config(built, pkg, ob, src100):
    """dwarf-cloudsc2-tl 1 100 1 (BASELINE config 2): same ratios as the oracle's driver, TEST PASSED."""
    r = _run(built, "dwarf-cloudsc2-tl", 1, 100, 1)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "TEST PASSED, penalty" in r.stdout
    rows = re.findall(r"^\s+(\d+)\s+([0-9.]+)\s*$", r.stdout, flags=re.M)
    assert [int(a) for a, _ in rows] == list(range(1, 11))
    z_gpu = np.array([float(b) for _, b in rows])
    prm = pkg.default_params(lregcl=False)
    st = pkg.ArrayState(src100, nproma=1, ngptot=100)
    z_cpu, _, _ = ob.driver_tl(prm, src100.ceta, st)
    # lambda = 10^-1 .. 10^-5: deterministic to many digits; beyond that cancellation-dominated
    assert np.allclose(z_gpu[:5], np.asarray(z_cpu)[:5], rtol=1e-6, atol=0)
    assert pkg.taylor_verdict(z_gpu)[0] == pkg.taylor_verdict(z_cpu)[0]


@pytest.mark.gpu
def test_dwarf_ad_baseline_config(built, pkg):
    """dwarf-cloudsc2-ad 1 100 100 (BASELINE config 3)."""
    r = _run(built, "dwarf-cloudsc2-ad", 1, 100, 100)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "=           TEST OK         =" in r.stdout
    m = re.search(r"The maximum error is\s+([0-9.]+)\s+times the zero of the machine", r.stdout)
    assert m and float(m.group(1)) < 10000.0


@pytest.mark.gpu
def test_dwarf_nl_from_input_h5_and_host_arrays(built, pkg, src100, tmp_path):
    """An input.h5 holding the synthetic columns gives the same results as the built-in generator;
    a reference.h5 made from the un-expanded GPU results validates to zero error; the host-array path
    (what the unchanged Fortran host would call) prints the same validation table."""
    prm = pkg.default_params(lregcl=False)
    write_h5(tmp_path / "input.h5", input_h5_datasets(src100, prm))
    st = pkg.ArrayState(src100, nproma=100, ngptot=100)
    with pkg.Cloudsc2(prm, src100.klev, src100.ceta) as gpu:
        gpu.nl(st)
    o = st.outputs()
    loc = st.a["b_loc"][0]                                            # (8, KLEV, 100)
    ref = {"KLON": np.array([100], dtype=np.int32), "KLEV": np.array([137], dtype=np.int32),
           "PLUDE": src100.f["plude"], "PCOVPTOT": o["pcovptot"][0], "PFPLSL": o["pfplsl"][0],
           "PFPLSN": o["pfplsn"][0], "PFHPSL": o["pfhpsl"][0], "PFHPSN": o["pfhpsn"][0],
           "TENDENCY_LOC_T": loc[0], "TENDENCY_LOC_A": loc[1], "TENDENCY_LOC_Q": loc[2],
           "TENDENCY_LOC_CLD": loc[3:]}
    write_h5(tmp_path / "reference.h5", {k: np.ascontiguousarray(v) for k, v in ref.items()})
    outs = []
    for env in ({}, {"CLOUDSC2_HOST_ARRAYS": "1"}):
        r = _run(built, "dwarf-cloudsc2-nl", 2, 4000, 64, env=env, cwd=tmp_path)   # picks up ./input.h5, ./reference.h5
        assert r.returncode == 0, r.stdout + r.stderr
        assert "input: input.h5 (KLON=100, KLEV=137)" in r.stdout and "reference: reference.h5" in r.stdout
        v = _validation_lines(r.stdout)
        assert len(v) == 10
        for name, (nums5, warn) in v.items():
            assert not warn and nums5[2] == 0.0, (name, nums5)
        outs.append(v)
    assert outs[0] == outs[1]
    # a perturbed reference is flagged with the reference's own "!!!!" marker; the exit status has its
    # own tolerance (relative L1 error 1e-11 by default): 1e-13 passes with marks, 1e-9 fails
    good = ref["PFPLSN"].copy()
    ref["PFPLSN"] = good * (1.0 + 1e-13)
    write_h5(tmp_path / "reference.h5", {k: np.ascontiguousarray(v) for k, v in ref.items()})
    r = _run(built, "dwarf-cloudsc2-nl", 1, 1000, 32, cwd=tmp_path)
    v = _validation_lines(r.stdout)
    assert r.returncode == 0 and v["PFPLSN"][1] and not v["PFPLSL"][1]
    ref["PFPLSN"] = good * (1.0 + 1e-9)
    write_h5(tmp_path / "reference.h5", {k: np.ascontiguousarray(v) for k, v in ref.items()})
    r = _run(built, "dwarf-cloudsc2-nl", 1, 1000, 32, cwd=tmp_path)
    v = _validation_lines(r.stdout)
    assert r.returncode == 3 and v["PFPLSN"][1] and not v["PFPLSL"][1]
    # a NaN in the reference (or, equally, in the results) can never pass
    ref["PFPLSN"] = good.copy()
    ref["PFPLSN"][50, 7] = np.nan
    write_h5(tmp_path / "reference.h5", {k: np.ascontiguousarray(v) for k, v in ref.items()})
    r = _run(built, "dwarf-cloudsc2-nl", 1, 1000, 32, cwd=tmp_path)
    assert r.returncode == 3, r.stdout


@pytest.mark.gpu
def test_dwarf_write_input_and_reference(built, pkg, tmp_path):
    """CLOUDSC2_WRITE_REFERENCE=1 (dwarf_cloudsc.F90:124-126: NPROMA must be KLON) and CLOUDSC2_WRITE_INPUT:
    the files the program writes are the files it reads -- a later run at another NPROMA validates
    against them with zero error."""
    r = _run(built, "dwarf-cloudsc2-nl", 1, 100, 100, cwd=tmp_path,
             env={"CLOUDSC2_WRITE_INPUT": "input.h5", "CLOUDSC2_WRITE_REFERENCE": "1"})
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "input.h5").exists() and (tmp_path / "reference.h5").exists()
    assert pkg.read_h5_f8(tmp_path / "reference.h5", "PFPLSN").shape == (138, 100)
    r = _run(built, "dwarf-cloudsc2-nl", 1, 100, 32, cwd=tmp_path, env={"CLOUDSC2_WRITE_REFERENCE": "1"})
    assert r.returncode == 1 and "NPROMA=KLON" in r.stderr            # the reference's own refusal
    r = _run(built, "dwarf-cloudsc2-nl", 2, 5000, 32, cwd=tmp_path)   # picks both files up
    assert r.returncode == 0, r.stdout + r.stderr
    assert "input: input.h5 (KLON=100, KLEV=137)" in r.stdout and "reference: reference.h5" in r.stdout
    v = _validation_lines(r.stdout)
    assert len(v) == 10 and all(nums5[2] == 0.0 and not warn for nums5, warn in v.values())
    assert v["PFPLSL"][0][1] > 0


@pytest.mark.gpu
def test_dwarf_other_vertical_resolution_and_seed(built, pkg):
    """KLEV and the synthetic atmosphere are run-time data for the programs too
    (CLOUDSC2_SYNTH_KLEV / CLOUDSC2_SYNTH_SEED): KLEV = 91, seed 3."""
    env = {"CLOUDSC2_SYNTH_KLEV": "91", "CLOUDSC2_SYNTH_SEED": "3"}
    r = _run(built, "dwarf-cloudsc2-nl", 1, 3000, 48, env=env)         # ragged: 3000 = 62*48 + 24
    assert r.returncode == 0, r.stdout + r.stderr
    assert "100 synthetic columns x 91 levels, seed 3" in r.stdout
    v = _validation_lines(r.stdout)
    assert len(v) == 10 and all(nums5[2] == 0.0 and not warn for nums5, warn in v.values())
    r = _run(built, "dwarf-cloudsc2-tl", 1, 100, 1, env=env)
    assert r.returncode == 0 and "TEST PASSED" in r.stdout, r.stdout
    r = _run(built, "dwarf-cloudsc2-ad", 1, 100, 100, env=env)
    assert r.returncode == 0 and "TEST OK" in r.stdout, r.stdout
