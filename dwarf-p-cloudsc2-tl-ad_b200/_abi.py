"""ctypes view of the C ABI declared in include/cloudsc2_b200.h and include/cloudsc2_host.h.

This is the *binding*, not an implementation: every compute call goes into
libcloudsc2_b200.so (hand-written CUDA for sm_100a).  If the shared library is missing the
import of :func:`load_library` raises -- there is no Python/NumPy/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
# CLOUDSC2_LIB: an alternative build of the same library (e.g. tools/probes/libcloudsc2_b200_experiments.so,
# the product plus the measured-slower kernel variants); never a different implementation
LIB_PATH = Path(os.environ["CLOUDSC2_LIB"]).resolve() if os.environ.get("CLOUDSC2_LIB") else \
    PKG_DIR / "csrc" / "libcloudsc2_b200.so"

NCLV = 5
NSTATE = 8
c_double_p = C.POINTER(C.c_double)


class Params(C.Structure):
    """struct cloudsc2_params (include/cloudsc2_b200.h)."""
    _fields_ = [(n, C.c_double) for n in (
        "rg", "rd", "rcpd", "retv", "rlvtt", "rlstt", "rlmlt", "rtt",
        "r2es", "r3les", "r3ies", "r4les", "r4ies", "r5les", "r5ies", "r5alvcp", "r5alscp",
        "ralvdcp", "ralsdcp", "rtwat", "rtice", "rtwat_rtice_r", "rvtmp2",
        "rclcrit", "rkconv", "rlmin", "rpecons", "rlptrc")] + [
        (n, C.c_int) for n in ("lphylin", "levapls2", "lregcl", "ldrain1d")]


FIELD_IN = ("pt", "pq", "pap", "paph", "plu", "plude", "pmfu", "pmfd", "psupsat", "pclv", "b_cml")
FIELD_OUT = ("b_loc", "pa", "pcovptot", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn")


class Fields(C.Structure):
    """struct cloudsc2_fields: raw addresses (host or device)."""
    _fields_ = [(n, C.c_void_p) for n in FIELD_IN + FIELD_OUT]


INCR_IN = ("paph", "pap", "pq", "pqs", "pt", "pl", "pi", "plude", "plu", "pmfu", "pmfd",
           "gtent", "gtenq", "gtenl", "gteni", "psupsat")
INCR_OUT = ("tent", "tenq", "tenl", "teni", "pclc", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn",
            "pcovptot")


class IncrIn(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in INCR_IN]


class IncrOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in INCR_OUT]


class Source(C.Structure):
    """struct cloudsc2_source (include/cloudsc2_host.h)."""
    _fields_ = [("klon", C.c_int), ("klev", C.c_int), ("ptsphy", C.c_double)] + [
        (n, c_double_p) for n in ("pt", "pq", "pap", "paph", "plu", "plude", "pmfu", "pmfd", "pa",
                                  "psupsat", "pclv", "tend_cml", "ceta")]


class Reference(C.Structure):
    """struct cloudsc2_reference (include/cloudsc2_host.h): un-expanded reference columns."""
    _fields_ = [("klon", C.c_int), ("klev", C.c_int)] + [
        (n, c_double_p) for n in ("plude", "pcovptot", "pfplsl", "pfplsn", "pfhpsl", "pfhpsn",
                                  "tend_loc")]


NVALIDATED = 10
VALIDATED_NAMES = ("PLUDE", "PCOVPTOT", "PFPLSL", "PFPLSN", "PFHPSL", "PFHPSN",
                   "TENDENCY_LOC%A", "TENDENCY_LOC%Q", "TENDENCY_LOC%T", "TENDENCY_LOC%CLD")


class State(C.Structure):
    """struct cloudsc2_state (include/cloudsc2_host.h)."""
    _fields_ = [("nproma", C.c_int), ("klev", C.c_int), ("ngptot", C.c_int),
                ("nblocks", C.c_int), ("f", Fields)]


_lib = None


def _declare(lib):
    i, d, vp = C.c_int, C.c_double, C.c_void_p
    P, F, II, IO = C.POINTER(Params), C.POINTER(Fields), C.POINTER(IncrIn), C.POINTER(IncrOut)
    sig = {
        # include/cloudsc2_b200.h
        "cloudsc2_gpu_init": (i, [P, i, c_double_p, i]),
        "cloudsc2_gpu_init_multi": (i, [P, i, c_double_p, i]),
        "cloudsc2_gpu_init_devices": (i, [P, i, c_double_p, i, C.POINTER(i)]),
        "cloudsc2_gpu_finalize": (i, []),
        "cloudsc2_gpu_num_devices": (i, []),
        "cloudsc2_shard_blocks": (i, [i, i, i, i, C.POINTER(i), C.POINTER(i), C.POINTER(i),
                                      C.POINTER(C.c_longlong)]),
        "cloudsc2_gpu_select_device": (i, [i]),
        "cloudsc2_gpu_comm_unique_id": (i, [vp, i]),
        "cloudsc2_gpu_comm_init_rank": (i, [i, i, vp, i]),
        "cloudsc2_gpu_comm_info": (i, [C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
        "cloudsc2_gpu_allreduce_dev": (i, [vp, i, i]),
        "cloudsc2_gpu_last_error": (C.c_char_p, []),
        "cloudsc2_gpu_available": (i, []),
        "cloudsc2_gpu_device_count": (i, []),
        "cloudsc2_gpu_launch_count": (C.c_longlong, []),
        "cloudsc2_gpu_nl": (i, [i, i, i, d, F, c_double_p, c_double_p]),
        "cloudsc2_gpu_nl_dev": (i, [i, i, i, d, F, vp, vp]),
        "cloudsc2_gpu_tl_dev": (i, [i, i, i, d, F, II, IO, vp]),
        "cloudsc2_gpu_ad_dev": (i, [i, i, i, d, F, II, IO, vp]),
        "cloudsc2_gpu_tl": (i, [i, i, i, d, F, II, IO]),
        "cloudsc2_gpu_ad": (i, [i, i, i, d, F, II, IO]),
        "cloudsc2_gpu_tl_taylor": (i, [i, i, i, d, F, c_double_p, c_double_p]),
        "cloudsc2_gpu_tl_taylor_dev": (i, [i, i, i, d, F, c_double_p, c_double_p]),
        "cloudsc2_gpu_ad_test": (i, [i, i, i, d, F, c_double_p, c_double_p]),
        "cloudsc2_gpu_ad_test_dev": (i, [i, i, i, d, F, c_double_p, c_double_p]),
        "cloudsc2_taylor_verdict": (i, [c_double_p, C.POINTER(i)]),
        "cloudsc2_adjoint_verdict": (i, [d]),
        "cloudsc2_gpu_expand_dev": (i, [vp, i, i, i, vp, i, i, vp]),
        "cloudsc2_gpu_expand_shard_dev": (i, [vp, i, i, i, vp, i, i, C.c_longlong, vp]),
        "cloudsc2_gpu_host_register": (i, [vp, C.c_ulonglong]),
        "cloudsc2_gpu_host_unregister": (i, [vp]),
        "cloudsc2_gpu_host_alloc": (i, [C.POINTER(vp), C.c_ulonglong]),
        "cloudsc2_gpu_host_free": (i, [vp]),
        "cloudsc2_gpu_malloc": (i, [C.POINTER(vp), C.c_ulonglong]),
        "cloudsc2_gpu_free": (i, [vp]),
        "cloudsc2_gpu_memcpy_h2d": (i, [vp, vp, C.c_ulonglong]),
        "cloudsc2_gpu_memcpy_d2h": (i, [vp, vp, C.c_ulonglong]),
        "cloudsc2_gpu_memset": (i, [vp, i, C.c_ulonglong]),
        "cloudsc2_gpu_sync": (i, []),
        "cloudsc2_gpu_math_probe": (i, [i, c_double_p, c_double_p, i]),
        "cloudsc2_gpu_set_option": (i, [C.c_char_p, i]),
        "cloudsc2_gpu_satur": (i, [C.c_longlong, c_double_p, c_double_p, c_double_p]),
        "cloudsc2_gpu_validate_dev": (i, [vp, i, vp, i, i, i, i, C.c_longlong, c_double_p]),
        "cloudsc2_gpu_validate_slabs_dev": (i, [vp, i, vp, i, i, i, C.c_longlong, i, C.c_longlong,
                                                c_double_p]),
        "cloudsc2_gpu_state_load": (i, [C.POINTER(Source), i, i]),
        "cloudsc2_gpu_state_free": (i, []),
        "cloudsc2_gpu_state_info": (i, [i, C.POINTER(i), C.POINTER(i), C.POINTER(i),
                                        C.POINTER(C.c_longlong)]),
        "cloudsc2_gpu_state_fields": (i, [F]),
        "cloudsc2_gpu_state_nl": (i, [c_double_p, c_double_p]),
        "cloudsc2_gpu_state_tl_taylor": (i, [c_double_p, c_double_p, c_double_p]),
        "cloudsc2_gpu_state_ad_test": (i, [c_double_p, c_double_p, c_double_p]),
        "cloudsc2_gpu_state_validate": (i, [C.POINTER(Reference), c_double_p]),
        "cloudsc2_gpu_state_get": (i, [C.c_char_p, c_double_p]),
        "cloudsc2_gpu_nl_source": (i, [C.POINTER(Source), C.POINTER(Reference), i, i, c_double_p,
                                       c_double_p, c_double_p]),
        # include/cloudsc2_host.h
        "cloudsc2_default_params": (None, [P]),
        "cloudsc2_source_synth": (i, [C.POINTER(Source), C.c_ulonglong, i, i, P]),
        "cloudsc2_source_free": (None, [C.POINTER(Source)]),
        "cloudsc2_expand_host": (None, [c_double_p, i, i, i, c_double_p, i, i]),
        "cloudsc2_state_load": (i, [C.POINTER(State), C.POINTER(Source), i, i]),
        "cloudsc2_state_free": (None, [C.POINTER(State)]),
        "cloudsc2_nblocks": (i, [i, i]),
        "cloudsc2_h5_read_f8": (C.c_longlong, [C.c_char_p, C.c_char_p, c_double_p, C.c_longlong,
                                               C.POINTER(i * 4), C.POINTER(i)]),
        "cloudsc2_h5_read_i4": (C.c_longlong, [C.c_char_p, C.c_char_p, C.POINTER(i),
                                               C.c_longlong]),
        "cloudsc2_source_load_h5": (i, [C.POINTER(Source), P, C.c_char_p]),
        "cloudsc2_reference_load_h5": (i, [C.POINTER(Reference), C.c_char_p]),
        "cloudsc2_reference_free": (None, [C.POINTER(Reference)]),
        "cloudsc2_input_last_error": (C.c_char_p, []),
        "cloudsc2_h5_write": (i, [C.c_char_p, vp, i]),
        "cloudsc2_source_write_h5": (i, [C.POINTER(Source), P, C.c_char_p]),
        "cloudsc2_reference_write_h5": (i, [C.POINTER(Reference), C.c_char_p]),
        "cloudsc2_validate_host": (None, [c_double_p, c_double_p, i, i, i, c_double_p]),
        "cloudsc2_error_rel": (d, [c_double_p, C.POINTER(i)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return sig


EXPORTED_SYMBOLS = None


def load_library(path: os.PathLike | None = None):
    """Load libcloudsc2_b200.so (built by __graft_entry__.build()).  Raises if absent."""
    global _lib, EXPORTED_SYMBOLS
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise RuntimeError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback for the CLOUDSC2 GPU path.")
    lib = C.CDLL(str(p), mode=C.RTLD_GLOBAL)
    EXPORTED_SYMBOLS = tuple(_declare(lib).keys())
    if path is None:
        _lib = lib
    return lib
