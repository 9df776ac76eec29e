"""cloudsc2-b200: B200-native (sm_100a CUDA, FP64) CLOUDSC2 NL / TL / AD column physics behind
the C ABI of include/cloudsc2_b200.h.  This package is the thin Python host side: ctypes binding
(_abi), the blocked-array state container mirroring CLOUDSC2_ARRAY_STATE (state) and the driver
objects mirroring CLOUDSC_DRIVER / CLOUDSC_DRIVER_TL / CLOUDSC_DRIVER_AD (driver).

There is no CPU fallback: every compute call goes into libcloudsc2_b200.so and fails loudly when
the library or a CUDA device is missing.
"""
from ._abi import load_library, Params, Fields, IncrIn, IncrOut, NCLV, NSTATE, LIB_PATH
from .state import (ArrayState, SourceColumns, default_params, expand, nblocks, synth_source,
                    read_h5_f8, read_h5_i4, validate, load_source_h5, write_source_h5)
from . import driver, pyapi, report, sharding
from .sharding import (Shard, allreduce_norms, allreduce_validation, shard_blocks, sharded_adjoint,
                       sharded_taylor)
from .driver import (Cloudsc2, Cloudsc2Error, DeviceState, adjoint_verdict, gpu_available,
                     taylor_verdict)

__all__ = [
    "load_library", "Params", "Fields", "IncrIn", "IncrOut", "NCLV", "NSTATE", "LIB_PATH",
    "ArrayState", "SourceColumns", "default_params", "expand", "nblocks", "synth_source",
    "read_h5_f8", "read_h5_i4", "validate", "load_source_h5", "write_source_h5",
    "Cloudsc2", "Cloudsc2Error", "DeviceState", "adjoint_verdict", "gpu_available",
    "taylor_verdict", "driver", "sharding", "Shard", "allreduce_norms", "shard_blocks",
    "sharded_adjoint", "sharded_taylor", "allreduce_validation", "pyapi", "report",
]
