"""Host-side state container: Python mirror of the reference's CLOUDSC2_ARRAY_STATE
(src/common/module/cloudsc2_array_state_mod.F90:28-203) on top of the C host helpers.

Arrays are NumPy, C-order ``(NBLOCKS, KLEV, NPROMA)`` == Fortran ``(NPROMA, KLEV, NBLOCKS)``,
so ``arr.ctypes.data`` is exactly the pointer the Fortran host would pass through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _abi

SRC_2D = ("pt", "pq", "pap", "plu", "plude", "pmfu", "pmfd", "pa", "psupsat")


def default_params(lregcl: bool = False) -> _abi.Params:
    """Constants of SURVEY Appendix E via cloudsc2_default_params (include/cloudsc2_host.h)."""
    lib = _abi.load_library()
    p = _abi.Params()
    lib.cloudsc2_default_params(C.byref(p))
    p.lregcl = int(bool(lregcl))
    return p


@dataclass
class SourceColumns:
    """Un-expanded columns in the input.h5 layout: every field C-order (NDIM, KLEV, KLON)."""
    klon: int
    klev: int
    ptsphy: float
    ceta: np.ndarray
    f: dict = field(default_factory=dict)   # name -> ndarray

    def subset(self, cols) -> "SourceColumns":
        cols = np.asarray(cols)
        g = {k: np.ascontiguousarray(v[..., cols]) for k, v in self.f.items()}
        return SourceColumns(len(cols), self.klev, self.ptsphy, self.ceta.copy(), g)


def synth_source(seed: int = 0, klon: int = 100, klev: int = 137,
                 params: _abi.Params | None = None) -> SourceColumns:
    """Seeded synthetic stand-in for config-files/input.h5 (absent from the reference)."""
    lib = _abi.load_library()
    p = params if params is not None else default_params()
    s = _abi.Source()
    rc = lib.cloudsc2_source_synth(C.byref(s), seed, klon, klev, C.byref(p))
    if rc:
        raise RuntimeError(f"cloudsc2_source_synth failed rc={rc}")
    try:
        def arr(ptr, shape):
            return np.ctypeslib.as_array(ptr, shape=shape).copy()
        f = {n: arr(getattr(s, n), (klev, klon)) for n in SRC_2D}
        f["paph"] = arr(s.paph, (klev + 1, klon))
        f["pclv"] = arr(s.pclv, (_abi.NCLV, klev, klon))
        f["tend_cml"] = arr(s.tend_cml, (_abi.NSTATE, klev, klon))
        out = SourceColumns(klon, klev, float(s.ptsphy), arr(s.ceta, (klev,)), f)
    finally:
        lib.cloudsc2_source_free(C.byref(s))
    return out


def _source_from_struct(s: "_abi.Source") -> SourceColumns:
    klon, klev = int(s.klon), int(s.klev)

    def arr(ptr, shape):
        return np.ctypeslib.as_array(ptr, shape=shape).copy()
    f = {n: arr(getattr(s, n), (klev, klon)) for n in SRC_2D}
    f["paph"] = arr(s.paph, (klev + 1, klon))
    f["pclv"] = arr(s.pclv, (_abi.NCLV, klev, klon))
    f["tend_cml"] = arr(s.tend_cml, (_abi.NSTATE, klev, klon))
    return SourceColumns(klon, klev, float(s.ptsphy), arr(s.ceta, (klev,)), f)


def load_source_h5(path) -> tuple[SourceColumns, _abi.Params]:
    """CLOUDSC2_ARRAY_STATE%LOAD's reads of input.h5 (cloudsc2_array_state_mod.F90:153-203):
    the un-expanded columns, PTSPHY and the constants that reach the kernels.  Raises on a
    missing dataset (the reference aborts)."""
    lib = _abi.load_library()
    s, p = _abi.Source(), _abi.Params()
    rc = lib.cloudsc2_source_load_h5(C.byref(s), C.byref(p), str(path).encode())
    if rc:
        raise KeyError(f"{path}: {lib.cloudsc2_input_last_error().decode()} (rc={rc})")
    try:
        return _source_from_struct(s), p
    finally:
        lib.cloudsc2_source_free(C.byref(s))


def write_source_h5(src: SourceColumns, params: _abi.Params, path) -> None:
    """Write `src` + `params` as an input.h5 with every dataset the reference's loaders read
    (cloudsc2_source_write_h5): lets the reference's own binaries run on the synthetic columns."""
    lib = _abi.load_library()
    s = _abi.Source()
    s.klon, s.klev, s.ptsphy = src.klon, src.klev, src.ptsphy
    keep = []
    for n in list(SRC_2D) + ["paph", "pclv", "tend_cml"]:
        a = np.ascontiguousarray(src.f[n], dtype=np.float64)
        keep.append(a)
        setattr(s, n, a.ctypes.data_as(_abi.c_double_p))
    ceta = np.ascontiguousarray(src.ceta, dtype=np.float64)
    s.ceta = ceta.ctypes.data_as(_abi.c_double_p)
    rc = lib.cloudsc2_source_write_h5(C.byref(s), C.byref(params), str(path).encode())
    if rc:
        raise OSError(f"cloudsc2_source_write_h5({path}) failed rc={rc}")


def nblocks(ngptot: int, nproma: int) -> int:
    return ngptot // nproma + min(ngptot % nproma, 1)


def expand(src: np.ndarray, nproma: int, ngptot: int, gcol0: int = 0) -> np.ndarray:
    """expand_mod.F90:270-335 through cloudsc2_expand_host.  src (..., NLEV, NLON).
    gcol0: first global column of a shard (the source is rotated so that local column j is
    source column (gcol0 + j) mod NLON)."""
    lib = _abi.load_library()
    src = np.asarray(src, dtype=np.float64)
    if gcol0:
        src = np.roll(src, -(gcol0 % src.shape[-1]), axis=-1)
    src = np.ascontiguousarray(src)
    nlon, nlev = src.shape[-1], src.shape[-2]
    ndim = int(np.prod(src.shape[:-2])) if src.ndim > 2 else 1
    nb = nblocks(ngptot, nproma)
    shape = (nb,) + (tuple(src.shape[:-2]) if src.ndim > 2 else ()) + (nlev, nproma)
    dst = np.empty(shape, dtype=np.float64)
    lib.cloudsc2_expand_host(src.ctypes.data_as(_abi.c_double_p), nlon, nlev, ndim,
                             dst.ctypes.data_as(_abi.c_double_p), nproma, ngptot)
    return dst


class ArrayState:
    """Blocked arrays of one problem (mirror of CLOUDSC2_ARRAY_STATE%LOAD)."""

    def __init__(self, src: SourceColumns, nproma: int, ngptot: int, gcol0: int = 0):
        self.nproma, self.klev, self.ngptot = nproma, src.klev, ngptot
        self.gcol0 = gcol0
        self.nblocks = nblocks(ngptot, nproma)
        self.ptsphy = src.ptsphy
        self.ceta = np.ascontiguousarray(src.ceta)
        a = {}
        for n in ("pt", "pq", "pap", "paph", "plu", "plude", "pmfu", "pmfd", "psupsat", "pa",
                  "pclv"):
            a[n] = expand(src.f[n], nproma, ngptot, gcol0)
        a["b_cml"] = expand(src.f["tend_cml"], nproma, ngptot, gcol0)
        nb, kl = self.nblocks, self.klev
        a["b_loc"] = np.zeros((nb, _abi.NSTATE, kl, nproma))
        a["pcovptot"] = np.zeros((nb, kl, nproma))
        for n in ("pfplsl", "pfplsn", "pfhpsl", "pfhpsn"):
            a[n] = np.zeros((nb, kl + 1, nproma))
        self.a = a

    def fields(self) -> _abi.Fields:
        f = _abi.Fields()
        for n in _abi.FIELD_IN + _abi.FIELD_OUT:
            setattr(f, n, self.a[n].ctypes.data)
        return f

    def reset_outputs(self, fill: float = 0.0):
        for n in ("b_loc", "pcovptot", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn"):
            self.a[n].fill(fill)

    # views named as the reference's outputs
    def outputs(self) -> dict:
        a = self.a
        return {"tend_loc_t": a["b_loc"][:, 0], "tend_loc_q": a["b_loc"][:, 2],
                "tend_loc_l": a["b_loc"][:, 3], "tend_loc_i": a["b_loc"][:, 4],
                "pa": a["pa"], "pfplsl": a["pfplsl"], "pfplsn": a["pfplsn"],
                "pfhpsl": a["pfhpsl"], "pfhpsn": a["pfhpsn"], "pcovptot": a["pcovptot"]}


def read_h5_f8(path: str, dataset: str) -> np.ndarray:
    """Read a contiguous little-endian f8 dataset with the library's own mini HDF5 reader
    (replaces LOAD_ARRAY, hdf5_file_mod.F90:135-164, for the files the reference ships)."""
    lib = _abi.load_library()
    dims = (C.c_int * 4)()
    nd = C.c_int(0)
    n = lib.cloudsc2_h5_read_f8(str(path).encode(), dataset.encode(), None, 0, C.byref(dims),
                                C.byref(nd))
    if n < 0:
        raise KeyError(f"{dataset!r} in {path}: mini-HDF5 reader error {n}")
    out = np.empty(int(n), dtype=np.float64)
    lib.cloudsc2_h5_read_f8(str(path).encode(), dataset.encode(),
                            out.ctypes.data_as(_abi.c_double_p), n, C.byref(dims), C.byref(nd))
    return out.reshape(tuple(dims[i] for i in range(nd.value))) if nd.value else out


def read_h5_i4(path: str, dataset: str) -> np.ndarray:
    lib = _abi.load_library()
    n = lib.cloudsc2_h5_read_i4(str(path).encode(), dataset.encode(), None, 0)
    if n < 0:
        raise KeyError(f"{dataset!r} in {path}: mini-HDF5 reader error {n}")
    out = np.empty(int(n), dtype=np.int32)
    lib.cloudsc2_h5_read_i4(str(path).encode(), dataset.encode(),
                            out.ctypes.data_as(C.POINTER(C.c_int)), n)
    return out


def validate(ref: np.ndarray, fld: np.ndarray, ngptot: int) -> dict:
    """validate_mod.F90:165-211 + ERROR_PRINT :263-296 for one blocked field (NB, NLEV, NPROMA)."""
    lib = _abi.load_library()
    ref = np.ascontiguousarray(ref, dtype=np.float64)
    fld = np.ascontiguousarray(fld, dtype=np.float64)
    nb, nlev, nproma = fld.shape
    out = (C.c_double * 5)()
    lib.cloudsc2_validate_host(ref.ctypes.data_as(_abi.c_double_p),
                               fld.ctypes.data_as(_abi.c_double_p), nproma, nlev, ngptot, out)
    flag = C.c_int(0)
    rel = lib.cloudsc2_error_rel(out, C.byref(flag))
    return {"min": out[0], "max": out[1], "max_abs_err": out[2], "sum_abs_err": out[3],
            "sum_abs_ref": out[4], "rel_err_pct": rel, "flag": bool(flag.value)}
