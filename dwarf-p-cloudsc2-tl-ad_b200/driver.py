"""Host-side mirror of the reference's three drivers on top of the C ABI.

    reference                                              here
    CLOUDSC_DRIVER     (cloudsc2_nl/cloudsc_driver_mod.F90:22)     Cloudsc2.nl / nl_dev
    CLOUDSC_DRIVER_TL  (cloudsc2_tl/cloudsc_driver_tl_mod.F90:33)  Cloudsc2.tl_taylor[_dev]
    CLOUDSC_DRIVER_AD  (cloudsc2_ad/cloudsc_driver_ad_mod.F90:22)  Cloudsc2.ad_test[_dev]
    CLOUDSC2TL / CLOUDSC2AD on full fields                         Cloudsc2.tl / ad [_dev]

Everything computes inside libcloudsc2_b200.so (CUDA, sm_100a).  Errors follow the reference's
convention of aborting (ABOR1): any non-zero return code raises Cloudsc2Error.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .state import ArrayState, nblocks


class Cloudsc2Error(RuntimeError):
    pass


# cloudsc2_gpu_tl_taylor*: "TL is totally wrong" (cloudsc_driver_tl_mod.F90:247-249), distinct from
# the argument errors (3)
DEGENERATE_RC = 6


def torch_comm_factory(rank: int, world: int):
    """after_init hook for process-per-GPU jobs under torch.distributed: rank 0 makes an NCCL unique id,
    torch broadcasts it (the job's existing channel, like MPI_BCAST in the Fortran host), every rank
    joins -- from then on the test norms are all-reduced INSIDE the library."""
    def hook(gpu):
        if world <= 1:
            return
        import torch.distributed as dist
        box = [gpu.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        gpu.comm_init_rank(rank, world, box[0])
    return hook


def source_struct(src) -> tuple["_abi.Source", list]:
    """struct cloudsc2_source view of a SourceColumns (the arrays are kept alive by the 2nd item)."""
    s = _abi.Source()
    s.klon, s.klev, s.ptsphy = src.klon, src.klev, src.ptsphy
    keep = []
    for n in ("pt", "pq", "pap", "paph", "plu", "plude", "pmfu", "pmfd", "pa", "psupsat", "pclv",
              "tend_cml"):
        a = np.ascontiguousarray(src.f[n], dtype=np.float64)
        keep.append(a)
        setattr(s, n, a.ctypes.data_as(_abi.c_double_p))
    c = np.ascontiguousarray(src.ceta, dtype=np.float64)
    keep.append(c)
    s.ceta = c.ctypes.data_as(_abi.c_double_p)
    return s, keep


def reference_struct(ref: dict, klon: int, klev: int) -> tuple["_abi.Reference", list]:
    """struct cloudsc2_reference from un-expanded reference columns: a dict with plude, pcovptot
    (KLEV,KLON), pfplsl, pfplsn, pfhpsl, pfhpsn (KLEV+1,KLON) and tend_loc (8,KLEV,KLON)."""
    r = _abi.Reference()
    r.klon, r.klev = klon, klev
    keep = []
    for n in ("plude", "pcovptot", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn", "tend_loc"):
        a = np.ascontiguousarray(ref[n], dtype=np.float64)
        keep.append(a)
        setattr(r, n, a.ctypes.data_as(_abi.c_double_p))
    return r, keep


def gpu_available() -> bool:
    return bool(_abi.load_library().cloudsc2_gpu_available())


def taylor_verdict(znormg) -> tuple[int, int]:
    """cloudsc_driver_tl_mod.F90:273-311 -> (penalty, istart); passed iff 0 <= penalty <= 5."""
    lib = _abi.load_library()
    z = np.ascontiguousarray(znormg, dtype=np.float64)
    istart = C.c_int(0)
    pen = lib.cloudsc2_taylor_verdict(z.ctypes.data_as(_abi.c_double_p), C.byref(istart))
    return int(pen), int(istart.value)


def adjoint_verdict(znormg: float) -> bool:
    """cloudsc_driver_ad_mod.F90:286-294."""
    return bool(_abi.load_library().cloudsc2_adjoint_verdict(float(znormg)))


INCR_IN_HALF = ("paph",)
INCR_OUT_HALF = ("pfplsl", "pfplsn", "pfhpsl", "pfhpsn")


def alloc_increments(nb: int, klev: int, nproma: int, fill: float = 0.0):
    """The 16 + 10 increment arrays of CLOUDSC2TL / CLOUDSC2AD as NumPy (NB, KLEV[+1], NPROMA)."""
    din = {n: np.full((nb, klev + (1 if n in INCR_IN_HALF else 0), nproma), fill)
           for n in _abi.INCR_IN}
    dout = {n: np.full((nb, klev + (1 if n in INCR_OUT_HALF else 0), nproma), fill)
            for n in _abi.INCR_OUT}
    return din, dout


def _incr_structs(din: dict, dout: dict):
    a, b = _abi.IncrIn(), _abi.IncrOut()
    for n in _abi.INCR_IN:
        setattr(a, n, din[n].ctypes.data if isinstance(din[n], np.ndarray) else int(din[n]))
    for n in _abi.INCR_OUT:
        setattr(b, n, dout[n].ctypes.data if isinstance(dout[n], np.ndarray) else int(dout[n]))
    return a, b


class DeviceState:
    """Device-resident copy of an ArrayState in the reference's blocked layout (library-owned
    cudaMalloc memory; no torch involved)."""

    def __init__(self, gpu: "Cloudsc2", st: ArrayState | None = None, *, nproma=None, klev=None,
                 ngptot=None):
        self.gpu = gpu
        if st is not None:
            nproma, klev, ngptot = st.nproma, st.klev, st.ngptot
        self.nproma, self.klev, self.ngptot = nproma, klev, ngptot
        self.nblocks = nblocks(ngptot, nproma)
        n2 = nproma * klev * self.nblocks
        n2h = nproma * (klev + 1) * self.nblocks
        self.sizes = {"pt": n2, "pq": n2, "pap": n2, "paph": n2h, "plu": n2, "plude": n2,
                      "pmfu": n2, "pmfd": n2, "psupsat": n2, "pclv": _abi.NCLV * n2,
                      "b_cml": _abi.NSTATE * n2, "b_loc": _abi.NSTATE * n2, "pa": n2,
                      "pcovptot": n2, "pfplsl": n2h, "pfplsn": n2h, "pfhpsl": n2h, "pfhpsn": n2h}
        self.ptr = {}
        for n, cnt in self.sizes.items():
            self.ptr[n] = gpu.malloc(cnt * 8)
        if st is not None:
            self.upload(st)

    # input.h5 dataset -> member of the blocked state (cloudsc2_array_state_mod.F90:153-203)
    SOURCE_MAP = {"pt": "pt", "pq": "pq", "pap": "pap", "paph": "paph", "plu": "plu",
                  "plude": "plude", "pmfu": "pmfu", "pmfd": "pmfd", "psupsat": "psupsat", "pa": "pa",
                  "pclv": "pclv", "b_cml": "tend_cml"}

    @classmethod
    def from_source(cls, gpu: "Cloudsc2", src, nproma: int, ngptot: int, gcol0: int = 0,
                    stream: int | None = None) -> "DeviceState":
        """CLOUDSC2_ARRAY_STATE%LOAD with the expansion done ON THE DEVICE: upload the un-expanded
        source columns (a few MB) and replicate them cyclically into NGPTOT columns
        (expand_mod.F90:270-335 -> cloudsc2_gpu_expand_shard_dev); outputs are zeroed."""
        ds = cls(gpu, nproma=nproma, klev=src.klev, ngptot=ngptot)
        for dst, name in cls.SOURCE_MAP.items():
            a = np.ascontiguousarray(src.f[name])
            nlev = a.shape[-2]
            ndim = a.size // (nlev * src.klon)
            p = gpu.malloc(a.nbytes)
            try:
                gpu.h2d(p, a)
                gpu.expand_dev(p, src.klon, nlev, ndim, ds.ptr[dst], nproma, ngptot, stream=stream,
                               gcol0=gcol0)
                gpu.sync()
            finally:
                gpu.free(p)
        ds.zero(("b_loc", "pcovptot", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn"))
        return ds

    def upload(self, st: ArrayState, names=None):
        for n in names or self.sizes:
            self.gpu.h2d(self.ptr[n], st.a[n])

    def download(self, st: ArrayState, names=_abi.FIELD_OUT):
        for n in names:
            self.gpu.d2h(st.a[n], self.ptr[n])

    def zero(self, names=_abi.FIELD_OUT):
        for n in names:
            self.gpu.memset(self.ptr[n], 0, self.sizes[n] * 8)

    def fields(self) -> _abi.Fields:
        f = _abi.Fields()
        for n in _abi.FIELD_IN + _abi.FIELD_OUT:
            setattr(f, n, self.ptr[n])
        return f

    def nbytes(self) -> int:
        return 8 * sum(self.sizes.values())

    def free(self):
        for p in self.ptr.values():
            self.gpu.free(p)
        self.ptr = {}


class Cloudsc2:
    """One set of constants/switches bound to the library's device set (one context per GPU, created by
    cloudsc2_gpu_init / _init_multi / _init_devices, released by cloudsc2_gpu_finalize).  A process has ONE
    device set at a time, so when several Cloudsc2 objects exist the one being used re-initialises
    the set with its own constants first -- like the reference, where the constants are module variables set
    before the driver call (dwarf_cloudsc.F90:105-107); only YRNCL%LREGCL can be switched without that
    (set_option("lregcl", ...))."""

    _owner = None     # the object whose constants are currently loaded in the library

    def __init__(self, params: _abi.Params, klev: int, ceta, device: int = 0, ngpus: int | None = None,
                 devices=None, after_init=None):
        """device: one GPU (cloudsc2_gpu_init, what one rank of a process-per-GPU job uses);
        ngpus / devices: a device set driven by this one process (cloudsc2_gpu_init_multi /
        _init_devices): host-pointer and resident-state entries shard the blocks over the set and
        all-reduce the norms with the library's own NCCL communicator.
        after_init(self): called after every (re-)initialisation of the library context, e.g. to join
        a job-wide communicator again (torch_comm_factory below)."""
        self.lib = _abi.load_library()
        self.params = params
        self.klev = int(klev)
        self.device = int(device)
        self.ngpus = None if ngpus is None else int(ngpus)
        self.devices = None if devices is None else [int(d) for d in devices]
        self.ceta = np.ascontiguousarray(ceta, dtype=np.float64)
        self._launches = 0
        self._open = False
        self._after_init = after_init
        self._bind()
        self._open = True

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            msg = self.lib.cloudsc2_gpu_last_error().decode(errors="replace")
            raise Cloudsc2Error(f"libcloudsc2_b200 rc={rc}: {msg}")

    def _bind(self):
        if Cloudsc2._owner is self:
            return
        prev = Cloudsc2._owner
        if prev is not None:
            prev._launches += int(self.lib.cloudsc2_gpu_launch_count())
        Cloudsc2._owner = None
        ceta = self.ceta.ctypes.data_as(_abi.c_double_p)
        if self.devices is not None:
            arr = (C.c_int * len(self.devices))(*self.devices)
            rc = self.lib.cloudsc2_gpu_init_devices(C.byref(self.params), self.klev, ceta,
                                                    len(self.devices), arr)
        elif self.ngpus is not None:
            rc = self.lib.cloudsc2_gpu_init_multi(C.byref(self.params), self.klev, ceta, self.ngpus)
        else:
            rc = self.lib.cloudsc2_gpu_init(C.byref(self.params), self.klev, ceta, self.device)
        self._check(rc)
        Cloudsc2._owner = self
        if self._after_init is not None:
            self._after_init(self)

    def close(self):
        if getattr(self, "_open", False):
            if Cloudsc2._owner is self:
                self._launches += int(self.lib.cloudsc2_gpu_launch_count())
                self.lib.cloudsc2_gpu_finalize()
                Cloudsc2._owner = None
            self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def launch_count(self) -> int:
        """Kernels launched by the library on behalf of this object."""
        live = int(self.lib.cloudsc2_gpu_launch_count()) if Cloudsc2._owner is self else 0
        return self._launches + live

    def sync(self):
        self._bind()
        self._check(self.lib.cloudsc2_gpu_sync())

    def malloc(self, nbytes: int) -> int:
        self._bind()
        p = C.c_void_p()
        self._check(self.lib.cloudsc2_gpu_malloc(C.byref(p), max(int(nbytes), 8)))
        return int(p.value)

    def free(self, ptr: int):
        self._check(self.lib.cloudsc2_gpu_free(ptr))

    def h2d(self, dptr: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        self._check(self.lib.cloudsc2_gpu_memcpy_h2d(dptr, arr.ctypes.data, arr.nbytes))

    def d2h(self, arr: np.ndarray, dptr: int):
        assert arr.flags.c_contiguous
        self._check(self.lib.cloudsc2_gpu_memcpy_d2h(arr.ctypes.data, dptr, arr.nbytes))

    def memset(self, dptr: int, value: int, nbytes: int):
        self._check(self.lib.cloudsc2_gpu_memset(dptr, value, nbytes))

    # -- nonlinear ----------------------------------------------------------------------------
    def nl(self, st: ArrayState) -> tuple[float, float]:
        """Host arrays in, host arrays out (the drop-in for CLOUDSC_DRIVER's block loop).
        Returns (kernel seconds, total seconds incl. H2D/D2H) from CUDA events."""
        self._bind()
        tk, tt = C.c_double(0), C.c_double(0)
        f = st.fields()
        self._check(self.lib.cloudsc2_gpu_nl(st.nproma, st.klev, st.ngptot, st.ptsphy, C.byref(f),
                                             C.byref(tk), C.byref(tt)))
        return tk.value, tt.value

    def nl_dev(self, ds: DeviceState, ptsphy: float, pqs: int | None = None, stream: int | None = None):
        self._bind()
        f = ds.fields()
        self._check(self.lib.cloudsc2_gpu_nl_dev(ds.nproma, ds.klev, ds.ngptot, ptsphy, C.byref(f),
                                                 pqs, stream))

    # -- TL / AD on full fields ---------------------------------------------------------------
    def tl(self, st: ArrayState, din: dict, dout: dict):
        self._bind()
        a, b = _incr_structs(din, dout)
        f = st.fields()
        self._check(self.lib.cloudsc2_gpu_tl(st.nproma, st.klev, st.ngptot, st.ptsphy, C.byref(f),
                                             C.byref(a), C.byref(b)))

    def ad(self, st: ArrayState, din: dict, dout: dict):
        self._bind()
        a, b = _incr_structs(din, dout)
        f = st.fields()
        self._check(self.lib.cloudsc2_gpu_ad(st.nproma, st.klev, st.ngptot, st.ptsphy, C.byref(f),
                                             C.byref(a), C.byref(b)))

    def tl_dev(self, ds: DeviceState, ptsphy: float, din: dict, dout: dict, stream: int | None = None):
        self._bind()
        a, b = _incr_structs(din, dout)
        f = ds.fields()
        self._check(self.lib.cloudsc2_gpu_tl_dev(ds.nproma, ds.klev, ds.ngptot, ptsphy, C.byref(f),
                                                 C.byref(a), C.byref(b), stream))

    def ad_dev(self, ds: DeviceState, ptsphy: float, din: dict, dout: dict, stream: int | None = None):
        self._bind()
        a, b = _incr_structs(din, dout)
        f = ds.fields()
        self._check(self.lib.cloudsc2_gpu_ad_dev(ds.nproma, ds.klev, ds.ngptot, ptsphy, C.byref(f),
                                                 C.byref(a), C.byref(b), stream))

    # -- self tests -----------------------------------------------------------------------------
    def tl_taylor(self, st: ArrayState | DeviceState, ptsphy: float | None = None,
                  allow_degenerate: bool = False):
        """Taylor test -> (znormg[10] raw ratios, ratios per block [nblocks, 10])."""
        self._bind()
        z = np.zeros(10)
        rb = np.zeros((st.nblocks, 10))
        f = st.fields()
        dev = isinstance(st, DeviceState)
        fn = self.lib.cloudsc2_gpu_tl_taylor_dev if dev else self.lib.cloudsc2_gpu_tl_taylor
        rc = fn(st.nproma, st.klev, st.ngptot, ptsphy if dev else st.ptsphy, C.byref(f),
                z.ctypes.data_as(_abi.c_double_p), rb.ctypes.data_as(_abi.c_double_p))
        if not (allow_degenerate and rc == DEGENERATE_RC):
            self._check(rc)
        return z, rb

    def ad_test(self, st: ArrayState | DeviceState, ptsphy: float | None = None):
        """Adjoint dot-product test -> (ZNORMG, per-column [ngptot, 3] = N1, N2, N3)."""
        self._bind()
        zn = C.c_double(0)
        nc = np.zeros((st.ngptot, 3))
        f = st.fields()
        dev = isinstance(st, DeviceState)
        fn = self.lib.cloudsc2_gpu_ad_test_dev if dev else self.lib.cloudsc2_gpu_ad_test
        self._check(fn(st.nproma, st.klev, st.ngptot, ptsphy if dev else st.ptsphy, C.byref(f),
                       C.byref(zn), nc.ctypes.data_as(_abi.c_double_p)))
        return zn.value, nc

    # -- device set / resident sharded state (cloudsc2_gpu_state_*) -------------------------------
    def num_devices(self) -> int:
        self._bind()
        return int(self.lib.cloudsc2_gpu_num_devices())

    def comm_info(self) -> tuple[int, int, int]:
        """(rank, size, NCCL version) of the library's communicator."""
        self._bind()
        r, n, v = C.c_int(0), C.c_int(0), C.c_int(0)
        self._check(self.lib.cloudsc2_gpu_comm_info(C.byref(r), C.byref(n), C.byref(v)))
        return r.value, n.value, v.value

    def select_device(self, index: int):
        self._bind()
        self._check(self.lib.cloudsc2_gpu_select_device(int(index)))

    def comm_unique_id(self) -> bytes:
        self._bind()
        buf = C.create_string_buffer(128)
        self._check(self.lib.cloudsc2_gpu_comm_unique_id(buf, 128))
        return buf.raw

    def comm_init_rank(self, rank: int, nranks: int, uid: bytes):
        """Join a job-wide NCCL communicator (one process per GPU): the test norms are then
        all-reduced inside the library."""
        self._bind()
        self._check(self.lib.cloudsc2_gpu_comm_init_rank(int(rank), int(nranks), uid, len(uid)))

    def allreduce_dev(self, dptr: int, n: int, op: str = "max"):
        """CLOUDSC_MPI_REDUCE_MAX / MIN / SUM of n device-resident doubles over the library communicator."""
        self._bind()
        self._check(self.lib.cloudsc2_gpu_allreduce_dev(dptr, int(n), {"max": 0, "min": 1, "sum": 2}[op]))

    def state_load(self, src, nproma: int, ngptot: int):
        """GLOBAL_STATE%LOAD on the devices: upload the un-expanded columns, expand every device's
        block shard there."""
        self._bind()
        s, keep = source_struct(src)
        self._check(self.lib.cloudsc2_gpu_state_load(C.byref(s), int(nproma), int(ngptot)))
        self._state_dims = (int(nproma), int(src.klev), int(ngptot))
        del keep

    def state_free(self):
        self._check(self.lib.cloudsc2_gpu_state_free())

    def state_info(self) -> list[dict]:
        out = []
        for i in range(max(1, self.num_devices())):
            dev, nb, ng, g0 = C.c_int(0), C.c_int(0), C.c_int(0), C.c_longlong(0)
            self._check(self.lib.cloudsc2_gpu_state_info(i, C.byref(dev), C.byref(nb), C.byref(ng),
                                                         C.byref(g0)))
            out.append({"device": dev.value, "nblocks": nb.value, "ngptot": ng.value, "gcol0": g0.value})
        return out

    def state_nl(self) -> tuple[float, np.ndarray]:
        """CLOUDSC_DRIVER on the resident state, all devices concurrently -> (seconds of the slowest
        device's kernel, per-device seconds)."""
        self._bind()
        t = C.c_double(0)
        per = np.zeros(max(1, self.num_devices()))
        self._check(self.lib.cloudsc2_gpu_state_nl(C.byref(t), per.ctypes.data_as(_abi.c_double_p)))
        return t.value, per

    def state_tl_taylor(self, allow_degenerate: bool = False) -> tuple[np.ndarray, float, np.ndarray]:
        self._bind()
        z = np.zeros(10)
        t = C.c_double(0)
        per = np.zeros(max(1, self.num_devices()))
        rc = self.lib.cloudsc2_gpu_state_tl_taylor(z.ctypes.data_as(_abi.c_double_p), C.byref(t),
                                                   per.ctypes.data_as(_abi.c_double_p))
        if not (allow_degenerate and rc == DEGENERATE_RC):
            self._check(rc)
        return z, t.value, per

    def state_ad_test(self) -> tuple[float, float, np.ndarray]:
        self._bind()
        zn, t = C.c_double(0), C.c_double(0)
        per = np.zeros(max(1, self.num_devices()))
        self._check(self.lib.cloudsc2_gpu_state_ad_test(C.byref(zn), C.byref(t),
                                                        per.ctypes.data_as(_abi.c_double_p)))
        return zn.value, t.value, per

    def state_validate(self, ref: dict, klon: int) -> np.ndarray:
        """GLOBAL_STATE%VALIDATE on the devices -> stats[10][5] in the order of _abi.VALIDATED_NAMES."""
        self._bind()
        r, keep = reference_struct(ref, klon, self.klev)
        out = np.zeros((_abi.NVALIDATED, 5))
        self._check(self.lib.cloudsc2_gpu_state_validate(C.byref(r), out.ctypes.data_as(_abi.c_double_p)))
        del keep
        return out

    def state_get(self, name: str) -> np.ndarray:
        """One array of the resident state, all shards in block order, as the host would own it."""
        self._bind()
        nproma, klev, ngptot = self._state_dims
        nb = nblocks(ngptot, nproma)
        shape = {"paph": (nb, klev + 1, nproma), "pfplsl": (nb, klev + 1, nproma),
                 "pfplsn": (nb, klev + 1, nproma), "pfhpsl": (nb, klev + 1, nproma),
                 "pfhpsn": (nb, klev + 1, nproma), "pclv": (nb, _abi.NCLV, klev, nproma),
                 "b_cml": (nb, _abi.NSTATE, klev, nproma), "b_loc": (nb, _abi.NSTATE, klev, nproma)
                 }.get(name, (nb, klev, nproma))
        out = np.zeros(shape)
        self._check(self.lib.cloudsc2_gpu_state_get(name.encode(), out.ctypes.data_as(_abi.c_double_p)))
        return out

    def nl_source(self, src, nproma: int, ngptot: int, ref: dict | None = None):
        """The NL program's work flow in one call (dwarf_cloudsc.F90:84-122): LOAD with device-side
        expansion, CLOUDSC_DRIVER, VALIDATE -> (stats[10][5] or None, kernel seconds, total seconds)."""
        self._bind()
        s, keep = source_struct(src)
        tk, tt = C.c_double(0), C.c_double(0)
        stats = None
        rp = None
        if ref is not None:
            r, keep2 = reference_struct(ref, src.klon, src.klev)
            rp = C.byref(r)
            stats = np.zeros((_abi.NVALIDATED, 5))
        self._check(self.lib.cloudsc2_gpu_nl_source(
            C.byref(s), rp, int(nproma), int(ngptot),
            stats.ctypes.data_as(_abi.c_double_p) if stats is not None else None, C.byref(tk), C.byref(tt)))
        self._state_dims = (int(nproma), int(src.klev), int(ngptot))
        return stats, tk.value, tt.value

    # -- expansion --------------------------------------------------------------------------------
    def expand_dev(self, src_ptr: int, nlon: int, nlev: int, ndim: int, dst_ptr: int, nproma: int,
                   ngptot: int, stream: int | None = None, gcol0: int = 0):
        """expand_mod.F90:270-302 on the device; gcol0 = first global column of this shard."""
        self._bind()
        self._check(self.lib.cloudsc2_gpu_expand_shard_dev(src_ptr, nlon, nlev, ndim, dst_ptr,
                                                           nproma, ngptot, gcol0, stream))

    def validate_dev(self, ref_src_ptr: int, nlon: int, field_ptr: int, nproma: int, nlev: int,
                     ndim: int, ngptot: int, gcol0: int = 0, blk_stride: int | None = None) -> np.ndarray:
        """validate_mod.F90:165-261 on the device -> [min, max, max|err|, sum|err|, sum|ref|].
        blk_stride (doubles between blocks) selects a slab range of an AOSOA buffer, e.g.
        TENDENCY_LOC%T = B_LOC(:,:,1,:) with blk_stride = 8*NPROMA*KLEV."""
        self._bind()
        out = np.zeros(5)
        if blk_stride is None:
            blk_stride = nproma * nlev * ndim
        self._check(self.lib.cloudsc2_gpu_validate_slabs_dev(ref_src_ptr, nlon, field_ptr, nproma, nlev,
                                                             ndim, blk_stride, ngptot, gcol0,
                                                             out.ctypes.data_as(_abi.c_double_p)))
        return out

    def satur(self, pap: np.ndarray, pt: np.ndarray) -> np.ndarray:
        """SATUR (satur.F90:106-123) elementwise on the GPU."""
        self._bind()
        pap = np.ascontiguousarray(pap, dtype=np.float64)
        pt = np.ascontiguousarray(pt, dtype=np.float64)
        assert pap.shape == pt.shape
        out = np.empty_like(pt)
        dp = _abi.c_double_p
        self._check(self.lib.cloudsc2_gpu_satur(pt.size, pap.ctypes.data_as(dp), pt.ctypes.data_as(dp),
                                                out.ctypes.data_as(dp)))
        return out

    def set_option(self, name: str, value: int):
        """cloudsc2_gpu_set_option: 'e2e_mode', 'e2e_chunk_mb', 'e2e_host_derive', 'ad_have_trajectory',
        'lregcl' (YRNCL%LREGCL without a new init)."""
        if name == "lregcl":
            self._bind()
            self.params.lregcl = int(bool(value))
        self._check(self.lib.cloudsc2_gpu_set_option(name.encode(), int(value)))

    def math_probe(self, fn: int, x: np.ndarray) -> np.ndarray:
        """Evaluate one of the kernels' elementary functions (csrc/cloudsc2_math.cuh) on the GPU."""
        self._bind()
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self._check(self.lib.cloudsc2_gpu_math_probe(fn, x.ctypes.data_as(_abi.c_double_p),
                                                     y.ctypes.data_as(_abi.c_double_p), x.size))
        return y

    def host_alloc_like(self, arr: np.ndarray) -> tuple[np.ndarray, int]:
        """A page-locked (cudaHostAlloc) copy of `arr` -> (array view, pointer to free with host_free)."""
        self._bind()
        p = C.c_void_p()
        self._check(self.lib.cloudsc2_gpu_host_alloc(C.byref(p), max(arr.nbytes, 8)))
        buf = (C.c_double * arr.size).from_address(p.value)
        out = np.frombuffer(buf, dtype=np.float64, count=arr.size).reshape(arr.shape)
        out[...] = arr
        return out, int(p.value)

    def host_free(self, ptr: int):
        self._check(self.lib.cloudsc2_gpu_host_free(ptr))

    def pin(self, arr: np.ndarray):
        self._bind()
        self._check(self.lib.cloudsc2_gpu_host_register(arr.ctypes.data, arr.nbytes))

    def unpin(self, arr: np.ndarray):
        self._check(self.lib.cloudsc2_gpu_host_unregister(arr.ctypes.data))
