"""GPU suite, row 8e: the library's own device set.  One PROCESS drives N GPUs (one context, stream
set and host worker thread per device, blocks sharded with the arithmetic of
cloudsc2_nl/dwarf_cloudsc.F90:65-69, norms all-reduced by the library's NCCL communicator:
reduction(max:znormg) cloudsc_driver_tl_mod.F90:125, cloudsc_driver_ad_mod.F90:107; validation
statistics validate_mod.F90:197-199).  The N-device results must equal the 1-device results
bit for bit: columns are independent and every device runs the same kernels on the same values.

The resident-state entries (cloudsc2_gpu_state_*) are also covered on a single device, so the
driver's one-GPU run exercises them; the N>=2 cases skip there and are run with `gpurun --gpus 2`.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

OUT = ("b_loc", "pa", "pcovptot", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn")


def _ndev(pkg):
    return int(pkg.load_library().cloudsc2_gpu_device_count())


def _self_reference(pkg, g, src):
    """Reference columns = the NL results of the un-expanded columns (NPROMA = KLON, one block)."""
    st = pkg.ArrayState(src, src.klon, src.klon)
    g.nl(st)
    return {"plude": st.a["plude"][0], "pcovptot": st.a["pcovptot"][0], "pfplsl": st.a["pfplsl"][0],
            "pfplsn": st.a["pfplsn"][0], "pfhpsl": st.a["pfhpsl"][0], "pfhpsn": st.a["pfhpsn"][0],
            "tend_loc": st.a["b_loc"][0]}


@pytest.mark.parametrize("nproma,ngptot", [(32, 1000), (128, 4133), (100, 100)])
def test_resident_state_equals_host_pointer_entry(pkg, src100, gpu_nl, nproma, ngptot):
    """cloudsc2_gpu_state_load + _state_nl (device-side expansion) == cloudsc2_gpu_nl on the
    host-expanded arrays, bit for bit; _state_validate == the host's VALIDATE."""
    g = gpu_nl
    st = pkg.ArrayState(src100, nproma, ngptot)
    g.nl(st)
    g.state_load(src100, nproma, ngptot)
    info = g.state_info()
    assert sum(i["ngptot"] for i in info) == ngptot
    t, per = g.state_nl()
    assert t > 0 and per.max() == t
    for n in ("pt", "paph", "pclv", "b_cml"):
        assert np.array_equal(g.state_get(n), st.a[n]), n
    for n in OUT:
        assert np.array_equal(g.state_get(n), st.a[n]), n
    ref = _self_reference(pkg, g, src100)
    stats = g.state_validate(ref, src100.klon)
    # against the host implementation of VALIDATE on the expanded reference
    names = {"PLUDE": ("plude", None), "PCOVPTOT": ("pcovptot", None), "PFPLSL": ("pfplsl", None),
             "PFPLSN": ("pfplsn", None), "PFHPSL": ("pfhpsl", None), "PFHPSN": ("pfhpsn", None),
             "TENDENCY_LOC%A": ("b_loc", slice(1, 2)), "TENDENCY_LOC%Q": ("b_loc", slice(2, 3)),
             "TENDENCY_LOC%T": ("b_loc", slice(0, 1)), "TENDENCY_LOC%CLD": ("b_loc", slice(3, 8))}
    nb = st.nblocks
    for i, vn in enumerate(pkg._abi.VALIDATED_NAMES):
        arr, sl = names[vn]
        fld = st.a[arr] if sl is None else st.a[arr][:, sl]
        r = ref["tend_loc"][sl] if sl is not None else ref[arr]
        rexp = pkg.expand(np.ascontiguousarray(r), nproma, ngptot)
        h = pkg.validate(rexp.reshape(nb, -1, nproma), np.ascontiguousarray(fld).reshape(nb, -1, nproma), ngptot)
        assert stats[i, 0] == h["min"] and stats[i, 1] == h["max"], vn
        assert stats[i, 2] == h["max_abs_err"], vn
        assert np.isclose(stats[i, 3], h["sum_abs_err"], rtol=1e-12, atol=0), vn
        assert np.isclose(stats[i, 4], h["sum_abs_ref"], rtol=1e-12), vn
    # the kernel against itself on the source columns: identical up to the cyclic map
    assert stats[:, 2].max() == 0.0
    g.state_free()


def test_nl_source_one_call(pkg, src100, gpu_nl):
    ref = _self_reference(pkg, gpu_nl, src100)
    stats, tk, tt = gpu_nl.nl_source(src100, 64, 3000, ref)
    assert stats.shape == (10, 5) and tk > 0 and tt >= tk
    assert stats[:, 2].max() == 0.0 and np.isfinite(stats).all()
    gpu_nl.state_free()


def test_resident_state_tests_equal_dev_entries(pkg, src100, gpu_nl):
    g = gpu_nl
    st = pkg.ArrayState(src100, 1, 100)
    z0, _ = g.tl_taylor(st)
    g.state_load(src100, 1, 100)
    z1, t, per = g.state_tl_taylor()
    assert np.array_equal(z0, z1) and t > 0
    g.state_free()


# ---- N >= 2 devices in one process -------------------------------------------------------------------

@pytest.fixture(scope="module")
def gpu_set(pkg, src100):
    n = _ndev(pkg)
    if n < 2:
        pytest.skip("needs >= 2 GPUs in one process (gpurun --gpus 2)")
    g = pkg.Cloudsc2(pkg.default_params(lregcl=False), src100.klev, src100.ceta, ngpus=min(n, 8))
    yield g
    g.close()


def test_device_set_has_a_communicator(pkg, gpu_set):
    n = gpu_set.num_devices()
    rank, size, ver = gpu_set.comm_info()
    assert n >= 2 and size == n and rank == 0 and ver >= 20000


@pytest.mark.parametrize("nproma,ngptot", [(32, 1000), (128, 40000), (100, 100), (64, 64)])
def test_device_set_nl_equals_single_device(pkg, src100, gpu_nl, gpu_set, nproma, ngptot):
    a = pkg.ArrayState(src100, nproma, ngptot)
    b = pkg.ArrayState(src100, nproma, ngptot)
    for n in OUT:                       # padding columns / unwritten slabs must keep these values
        a.a[n][...] = 7.0
        b.a[n][...] = 7.0
    # (the two objects share the library: using one re-initialises it, so do all one-GPU work first)
    gpu_nl.nl(a)
    gpu_nl.nl(a2 := pkg.ArrayState(src100, nproma, ngptot))
    ref = _self_reference(pkg, gpu_nl, src100)
    gpu_nl.state_load(src100, nproma, ngptot)
    gpu_nl.state_nl()
    stats1 = gpu_nl.state_validate(ref, src100.klon)
    gpu_nl.state_free()
    tk, tt = gpu_set.nl(b)
    assert tk > 0 and tt >= tk
    for n in OUT:
        assert np.array_equal(a.a[n], b.a[n]), n
    # and the resident, device-expanded state
    gpu_set.state_load(src100, nproma, ngptot)
    info = gpu_set.state_info()
    assert sum(i["ngptot"] for i in info) == ngptot and len({i["device"] for i in info}) == len(info)
    gpu_set.state_nl()
    for n in OUT:
        assert np.array_equal(gpu_set.state_get(n), a2.a[n]), n
    stats = gpu_set.state_validate(ref, src100.klon)
    assert np.array_equal(stats[:, :3], stats1[:, :3])             # min, max, max|err|: exact
    assert np.allclose(stats[:, 3:], stats1[:, 3:], rtol=1e-12, atol=0)     # sums: order of addition
    gpu_set.state_free()


@pytest.mark.parametrize("nproma,ngptot", [(1, 100), (32, 1000), (16, 16)])
def test_device_set_taylor_equals_single_device(pkg, src100, gpu_nl, gpu_set, nproma, ngptot):
    a = pkg.ArrayState(src100, nproma, ngptot)
    b = pkg.ArrayState(src100, nproma, ngptot)
    z1, r1 = gpu_nl.tl_taylor(a)
    zn, rn = gpu_set.tl_taylor(b)
    assert np.array_equal(z1, zn)
    assert np.array_equal(r1, rn)
    assert pkg.taylor_verdict(zn) == pkg.taylor_verdict(z1)
    for n in OUT:
        assert np.array_equal(a.a[n], b.a[n]), n
    gpu_set.state_load(src100, nproma, ngptot)
    zs, t, per = gpu_set.state_tl_taylor()
    assert np.array_equal(zs, z1) and len(per) == gpu_set.num_devices()
    gpu_set.state_free()


def test_device_set_adjoint_equals_single_device(pkg, src100, gpu_set):
    prm = pkg.default_params(lregcl=True)
    with pkg.Cloudsc2(prm, src100.klev, src100.ceta) as g1:
        z1, c1 = g1.ad_test(pkg.ArrayState(src100, 100, 100))
        z1b, c1b = g1.ad_test(pkg.ArrayState(src100, 32, 1000))
    with pkg.Cloudsc2(prm, src100.klev, src100.ceta, ngpus=gpu_set.num_devices()) as gn:
        zn, cn = gn.ad_test(pkg.ArrayState(src100, 100, 100))       # one block: the other devices idle
        znb, cnb = gn.ad_test(pkg.ArrayState(src100, 32, 1000))
        gn.state_load(src100, 32, 1000)
        zs, t, per = gn.state_ad_test()
    assert zn == z1 and np.array_equal(cn, c1)
    assert znb == z1b and np.array_equal(cnb, c1b) and zs == z1b
    assert pkg.adjoint_verdict(znb)


def test_device_set_tl_ad_fields_equal_single_device(pkg, src100, gpu_nl, gpu_set):
    from importlib import import_module
    drv = import_module("dwarf-p-cloudsc2-tl-ad_b200").driver
    nproma, ngptot = 32, 500
    rng = np.random.default_rng(3)
    outs = []
    for g in (gpu_nl, gpu_set):
        st = pkg.ArrayState(src100, nproma, ngptot)
        din, dout = drv.alloc_increments(st.nblocks, st.klev, nproma)
        r = np.random.default_rng(3)
        for n in din:
            din[n][...] = 1e-3 * r.standard_normal(din[n].shape)
        g.tl(st, din, dout)
        tl_out = {n: v.copy() for n, v in dout.items()}
        g.ad(st, din, dout)
        outs.append((tl_out, {n: v.copy() for n, v in din.items()}))
    for n in outs[0][0]:
        assert np.array_equal(outs[0][0][n], outs[1][0][n]), n
    for n in outs[0][1]:
        assert np.array_equal(outs[0][1][n], outs[1][1][n]), n
    del rng


def test_nan_on_one_device_is_not_lost(pkg, src100, gpu_set):
    """a MAX over ranks must not drop a rank whose ratios are NaN."""
    bad = pkg.synth_source(seed=0, klon=100, klev=137)
    nproma, ngptot = 10, 100
    st = pkg.ArrayState(bad, nproma, ngptot)
    st.a["pt"][-1, 60, 3] = np.nan                 # last block -> last device
    z, _ = gpu_set.tl_taylor(st, allow_degenerate=True)
    assert not np.all(np.isfinite(z)) or z.max() >= 1e300
    pen, _ = pkg.taylor_verdict(z)
    assert not (0 <= pen <= 5)
