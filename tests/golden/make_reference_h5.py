"""Generate dwarf-p-cloudsc2-tl-ad_b200/config-files/reference_synth_seed0.h5: the reference.h5 of the synthetic columns
(the counterpart of the reference's config-files/reference.h5).

The reference checkout ships no usable input / reference pair (config-files/input.h5 is missing and its
reference.h5 is stale for NL, README.md:22-24), so the dwarf-cloudsc2-nl program could only validate the GPU
results against themselves.  This script runs the CPU oracle -- the C restatement that is pinned
bit for bit to the reference's Fortran text (oracle/_ref) -- on the 100 synthetic columns of seed 0 as ONE
block (NPROMA = KLON = 100, KLEV = 137, exactly how the reference writes its own reference.h5:
cloudsc2_array_state_mod.F90:260-287) and stores the ten validated fields in the reference's file format
through the library's HDF5 writer.  The program picks the file up when it runs on the default synthetic
input without a reference of its own; a fixture, like the reference's reference.h5 is one.

Run in the build container:   python tests/golden/make_reference_h5.py
"""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from tests import oracle_binding as ob   # noqa: E402  (test infrastructure)

pkg = ob.pkg


def main():
    src = pkg.synth_source(seed=0, klon=100, klev=137)
    prm = pkg.default_params(lregcl=False)
    st = pkg.ArrayState(src, nproma=100, ngptot=100)
    ob.driver_nl(prm, src.ceta, st, numomp=1)
    ref = {"plude": st.a["plude"][0], "pcovptot": st.a["pcovptot"][0], "pfplsl": st.a["pfplsl"][0],
           "pfplsn": st.a["pfplsn"][0], "pfhpsl": st.a["pfhpsl"][0], "pfhpsn": st.a["pfhpsn"][0],
           "tend_loc": st.a["b_loc"][0]}
    r, keep = pkg.driver.reference_struct(ref, 100, 137)
    dst = ROOT / "dwarf-p-cloudsc2-tl-ad_b200" / "config-files" / "reference_synth_seed0.h5"
    rc = pkg.load_library().cloudsc2_reference_write_h5(C.byref(r), str(dst).encode())
    assert rc == 0, rc
    del keep
    print("wrote", dst, dst.stat().st_size, "bytes; max PFPLSN", float(np.abs(ref["pfplsn"]).max()))


if __name__ == "__main__":
    main()
