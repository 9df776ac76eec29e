"""CPU suite: the C-ABI library loads without a GPU, exports every symbol include/*.h declares,
refuses to compute without a device, and the host helpers (expansion, synthetic source, mini HDF5
reader, validation statistics, verdicts) behave like the reference's host code."""
import ctypes as C
import re
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_functions():
    names = []
    for h in sorted((ROOT / "include").glob("*.h")):
        txt = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        names += re.findall(r"\b(cloudsc2_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(names))


def test_header_symbols_are_exported(pkg):
    lib = pkg.load_library()
    declared = _declared_functions()
    assert len(declared) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", str(pkg.LIB_PATH)], capture_output=True,
                         text=True, check=True).stdout
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    missing = [n for n in declared if n not in exported]
    assert not missing, missing
    # Fortran-ABI per-block shims (include/cloudsc2_fortran.h)
    for n in ("satur_", "cloudsc2_", "cloudsc2tl_", "cloudsc2ad_"):
        assert n in exported, n
        assert n + "(" in (ROOT / "include" / "cloudsc2_fortran.h").read_text()
    for n in declared:
        assert getattr(lib, n) is not None
    # and the ctypes table covers exactly the header
    assert sorted(pkg._abi.EXPORTED_SYMBOLS) == declared


def test_no_cpu_fallback(pkg, src100):
    """Without a CUDA device every compute entry point must fail loudly (errors)."""
    if pkg.gpu_available():
        pytest.skip("a GPU is present; the no-device behaviour is checked on the CPU container")
    with pytest.raises(pkg.Cloudsc2Error, match="no CUDA device"):
        pkg.Cloudsc2(pkg.default_params(), src100.klev, src100.ceta)
    lib = pkg.load_library()
    f = pkg.Fields()
    rc = lib.cloudsc2_gpu_nl(32, 137, 64, 3600.0, C.byref(f), None, None)
    assert rc != 0 and b"init" in lib.cloudsc2_gpu_last_error()


def test_product_package_does_not_import_the_oracle(pkg):
    for py in (ROOT / "dwarf-p-cloudsc2-tl-ad_b200").rglob("*.py"):
        assert "oracle" not in py.read_text().replace("no CPU fallback", ""), py
    out = subprocess.run(["ldd", str(pkg.LIB_PATH)], capture_output=True, text=True).stdout
    assert "liboracle" not in out


def test_nblocks_matches_reference_formula(pkg):
    lib = pkg.load_library()
    for ngptot, nproma in [(100, 1), (100, 100), (160000, 32), (100, 32), (5, 8), (64, 64), (65, 64)]:
        want = ngptot // nproma + min(ngptot % nproma, 1)     # cloudsc_driver_mod.F90:62
        assert lib.cloudsc2_nblocks(ngptot, nproma) == want == pkg.nblocks(ngptot, nproma)


@pytest.mark.parametrize("nproma,ngptot", [(1, 100), (32, 100), (100, 100), (32, 250), (64, 1000), (7, 23)])
def test_expand_is_cyclic_with_zero_tail(pkg, src100, nproma, ngptot):
    """expand_mod.F90:270-302: column g <- source column g mod nlon, tail of last block zero."""
    for name in ("pt", "paph", "pclv"):
        src = src100.f[name]
        out = pkg.expand(src, nproma, ngptot)
        nb = pkg.nblocks(ngptot, nproma)
        assert out.shape[0] == nb and out.shape[-1] == nproma
        flat = np.moveaxis(out, 0, -2).reshape(src.shape[:-1] + (nb * nproma,))
        g = np.arange(ngptot)
        assert np.array_equal(flat[..., :ngptot], src[..., g % src.shape[-1]])
        assert not flat[..., ngptot:].any()


def test_synth_source_is_deterministic_and_plausible(pkg):
    a = pkg.synth_source(seed=3, klon=12, klev=137)
    b = pkg.synth_source(seed=3, klon=12, klev=137)
    c = pkg.synth_source(seed=4, klon=12, klev=137)
    for k in a.f:
        assert np.array_equal(a.f[k], b.f[k]), k
    assert not np.array_equal(a.f["pt"], c.f["pt"])
    f = a.f
    assert (np.diff(f["paph"], axis=0) > 0).all()                     # monotone half levels
    assert ((f["pap"] > f["paph"][:-1]) & (f["pap"] < f["paph"][1:])).all()
    assert f["pt"].min() > 150 and f["pt"].max() < 330
    assert (f["pq"] > 0).all() and (f["pmfu"] >= 0).all() and (f["pmfd"] <= 0).all()
    assert ((a.ceta > 0.1) & (a.ceta < 0.4)).sum() > 10                # tropopause window populated
    assert a.ceta[-1] < 1.0 and (np.diff(a.ceta) > 0).all()


def test_array_state_layout(pkg, src100):
    st = pkg.ArrayState(src100, nproma=32, ngptot=100)
    assert st.nblocks == 4
    assert st.a["pt"].shape == (4, 137, 32) and st.a["paph"].shape == (4, 138, 32)
    assert st.a["pclv"].shape == (4, 5, 137, 32) and st.a["b_cml"].shape == (4, 8, 137, 32)
    # Fortran (NPROMA,KLEV,NBLOCKS) element (jl,jk,ibl) at ((ibl*KLEV+jk)*NPROMA+jl)
    flat = st.a["pt"].ravel()
    assert flat[(2 * 137 + 5) * 32 + 7] == src100.f["pt"][5, (2 * 32 + 7) % 100]


def test_validate_statistics(pkg):
    rng = np.random.default_rng(0)
    ref = rng.standard_normal((3, 5, 8))
    fld = ref.copy()
    fld[1, 2, 3] += 0.5
    s = pkg.validate(ref, fld, ngptot=20)          # last block has 4 valid columns
    assert s["max_abs_err"] == pytest.approx(0.5)
    valid = np.ones_like(ref, dtype=bool)
    valid[2, :, 4:] = False
    assert s["sum_abs_ref"] == pytest.approx(np.abs(ref[valid]).sum())
    assert s["rel_err_pct"] == pytest.approx(100 * 0.5 / np.abs(ref[valid]).sum())
    assert s["flag"]
    assert not pkg.validate(ref, ref, ngptot=24)["flag"]


def test_verdicts(pkg):
    good = np.array([1.3, 1.02, 1.002, 1.0002, 1.00001, 1.0000002, 1.000003, 1.00004, 1.0005, 1.006])
    pen, istart = pkg.taylor_verdict(good)
    assert (pen, istart) == (0, 1)
    assert pkg.taylor_verdict(np.full(10, 3.0)) == (-13, 0)            # never below 0.5 -> err 13
    late = np.array([3.0, 3.0, 3.0, 3.0, 1.1, 1.01, 1.001, 1.0001, 1.00001, 1.000001])
    assert pkg.taylor_verdict(late)[0] == -13                          # ISTART > 4
    assert pkg.adjoint_verdict(9999.0) and not pkg.adjoint_verdict(10000.0)


# ---- mini HDF5 reader -------------------------------------------------------------------------

from tests.h5writer import write_h5 as _tiny_h5  # noqa: E402


def test_mini_hdf5_reader_roundtrip(pkg, tmp_path):
    rng = np.random.default_rng(5)
    d = {"PT": rng.standard_normal((7, 5)), "PAPH": rng.standard_normal((8, 5)),
         "TENDENCY_LOC_CLD": rng.standard_normal((5, 7, 5)), "KLON": np.array([5], dtype=np.int32),
         "KLEV": np.array([7], dtype=np.int32)}
    p = tmp_path / "tiny.h5"
    _tiny_h5(p, d)
    for n in ("PT", "PAPH", "TENDENCY_LOC_CLD"):
        got = pkg.read_h5_f8(p, n)
        assert got.shape == d[n].shape and np.array_equal(got, d[n]), n
    assert pkg.read_h5_i4(p, "KLON")[0] == 5 and pkg.read_h5_i4(p, "KLEV")[0] == 7
    with pytest.raises(KeyError):
        pkg.read_h5_f8(p, "NOPE")
    with pytest.raises(KeyError):
        pkg.read_h5_f8(p, "KLON")          # wrong type
    with pytest.raises(KeyError):
        pkg.read_h5_f8(tmp_path / "absent.h5", "PT")


def test_mini_hdf5_reader_on_reference_file(pkg):
    """Only where the reference checkout is mounted (build container); skipped on the GPU box."""
    ref = Path("/root/reference/config-files/reference.h5")
    if not ref.exists():
        pytest.skip("reference checkout not mounted")
    assert pkg.read_h5_i4(ref, "KLON")[0] == 100 and pkg.read_h5_i4(ref, "KLEV")[0] == 137
    fn = pkg.read_h5_f8(ref, "PFPLSN")
    assert fn.shape == (138, 100) and fn.min() >= 0 and 0 < fn.max() < 1e-4
    hn = pkg.read_h5_f8(ref, "PFHPSN")
    nz = fn > 0
    assert np.allclose(-hn[nz] / fn[nz], 2.8345e6, rtol=1e-12)      # RLSTT
    assert pkg.read_h5_f8(ref, "TENDENCY_LOC_CLD").shape == (5, 137, 100)


def test_library_block_sharding_is_the_reference_arithmetic(pkg):
    """cloudsc2_shard_blocks (what cloudsc2_gpu_init_multi shards with) == dwarf_cloudsc.F90:65-69 applied to
    NPROMA blocks == the Python helper the torchrun harness uses; shards tile the problem exactly."""
    lib = pkg.load_library()
    for ngptot, nproma, r in [(100, 1, 8), (100, 100, 8), (160000, 32, 8), (163840, 128, 3), (1000, 32, 5),
                              (7, 8, 4), (1310720, 128, 8), (4133, 128, 2), (65, 64, 2)]:
        nb = ngptot // nproma + min(ngptot % nproma, 1)
        per = (nb - 1) // r + 1                                   # NGPTOT = (NGPTOTG-1)/NUMPROC + 1 (:65)
        cols = 0
        for k in range(r):
            b0, n, ng, g0 = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
            assert lib.cloudsc2_shard_blocks(k, r, nproma, ngptot, C.byref(b0), C.byref(n), C.byref(ng), C.byref(g0)) == 0
            want_b0 = min(k * per, nb)
            want_n = min(nb, want_b0 + per) - want_b0             # the last ranks take the rest (:66-68)
            assert (b0.value, n.value, g0.value) == (want_b0, want_n, want_b0 * nproma)
            sh = pkg.shard_blocks(ngptot, nproma, k, r)
            assert (sh.block0, sh.nblocks, sh.gcol0, sh.ngptot) == (b0.value, n.value, g0.value, ng.value)
            cols += ng.value
        assert cols == ngptot
    assert lib.cloudsc2_shard_blocks(2, 2, 32, 100, None, None, None, None) == 3
