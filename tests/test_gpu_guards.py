"""GPU suite: out-of-bounds WRITE detection without compute-sanitizer (closed on this pool): every
device array handed to the NL / TL / AD kernels sits between two poisoned guard regions; after the
kernels ran, the guards must be untouched.  Ragged geometries (ICEND < NPROMA, NPROMA not a multiple
of the 128-column CTA) are the cases where an indexing slip would write outside."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GUARD = 4096            # doubles before and after every array
POISON = -7.0e300


class Guarded:
    def __init__(self, gpu, n_doubles, init=None):
        self.gpu, self.n = gpu, n_doubles
        self.base = gpu.malloc(8 * (n_doubles + 2 * GUARD))
        buf = np.full(n_doubles + 2 * GUARD, POISON)
        if init is not None:
            buf[GUARD:GUARD + n_doubles] = np.asarray(init, dtype=np.float64).reshape(-1)
        else:
            buf[GUARD:GUARD + n_doubles] = 0.0
        gpu.h2d(self.base, buf)
        self.ptr = self.base + 8 * GUARD

    def check(self, name):
        buf = np.empty(self.n + 2 * GUARD)
        self.gpu.d2h(buf, self.base)
        assert (buf[:GUARD] == POISON).all(), f"{name}: write below the array"
        assert (buf[GUARD + self.n:] == POISON).all(), f"{name}: write beyond the array"
        return buf[GUARD:GUARD + self.n]

    def free(self):
        self.gpu.free(self.base)


@pytest.mark.parametrize("nproma,ngptot", [(7, 23), (128, 300), (100, 250), (32, 33)])
def test_kernels_do_not_write_outside_their_arrays(pkg, src100, nproma, ngptot):
    prm = pkg.default_params(lregcl=True)
    st = pkg.ArrayState(src100, nproma, ngptot)
    nb = st.nblocks
    with pkg.Cloudsc2(prm, 137, src100.ceta) as gpu:
        g = {n: Guarded(gpu, a.size, a) for n, a in st.a.items()}
        din_h, dout_h = pkg.driver.alloc_increments(nb, 137, nproma)
        din = {n: Guarded(gpu, a.size, 0.01 * np.ones_like(a)) for n, a in din_h.items()}
        dout = {n: Guarded(gpu, a.size, 1e-6 * np.ones_like(a)) for n, a in dout_h.items()}
        try:
            f = pkg.Fields()
            for n in pkg._abi.FIELD_IN + pkg._abi.FIELD_OUT:
                setattr(f, n, g[n].ptr)
            a, b = pkg.IncrIn(), pkg.IncrOut()
            for n in pkg._abi.INCR_IN:
                setattr(a, n, din[n].ptr)
            for n in pkg._abi.INCR_OUT:
                setattr(b, n, dout[n].ptr)
            lib = gpu.lib
            gpu._check(lib.cloudsc2_gpu_nl_dev(nproma, 137, ngptot, st.ptsphy, C.byref(f), None, None))
            gpu._check(lib.cloudsc2_gpu_tl_dev(nproma, 137, ngptot, st.ptsphy, C.byref(f), C.byref(a), C.byref(b), None))
            gpu._check(lib.cloudsc2_gpu_ad_dev(nproma, 137, ngptot, st.ptsphy, C.byref(f), C.byref(a), C.byref(b), None))
            z = np.zeros(10)
            rb = np.zeros((nb, 10))
            rc = lib.cloudsc2_gpu_tl_taylor_dev(nproma, 137, ngptot, st.ptsphy, C.byref(f),
                                                z.ctypes.data_as(pkg._abi.c_double_p),
                                                rb.ctypes.data_as(pkg._abi.c_double_p))
            assert rc in (0, 3)
            zn = C.c_double(0)
            gpu._check(lib.cloudsc2_gpu_ad_test_dev(nproma, 137, ngptot, st.ptsphy, C.byref(f), C.byref(zn), None))
            gpu.sync()
            for n, x in {**g, **{"d_" + k: v for k, v in din.items()}, **{"y_" + k: v for k, v in dout.items()}}.items():
                x.check(n)
            # inputs are never written
            for n in pkg._abi.FIELD_IN:
                assert np.array_equal(g[n].check(n), st.a[n].reshape(-1)), n
        finally:
            for x in list(g.values()) + list(din.values()) + list(dout.values()):
                x.free()
