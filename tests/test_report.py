"""CPU suite: text reports compatible with the reference: ERROR_PRINT lines
(validate_mod.F90:263-296, format 1000) and the TOTAL lines of timer_mod.F90:114-174."""
import re

import numpy as np


def test_fortran_e_descriptor(pkg):
    f = pkg.report.fortran_e
    assert f(0.0) == " 0.0000000000000E+00"
    assert f(-1.5) == "-0.1500000000000E+01"
    assert f(123456.789) == " 0.1234567890000E+06"
    assert f(9.99999999999999e-5) == " 0.1000000000000E-03"       # rounding carries into the exponent
    assert f(2.5e-300) == " 0.2500000000000-299" or f(2.5e-300).endswith("E-299")
    assert all(len(f(v)) == 20 for v in (1.0, -1e-10, 3.14159e7))


def test_error_print_matches_reference_logic(pkg):
    r = pkg.report
    eps = np.finfo(float).eps
    # iopt 1: no error at all
    line = r.error_print("PCOVPTOT", [0.0, 0.0, 0.0, 0.0, 0.0], 100)
    assert line.startswith(" PCOVPTOT             2D1") and not line.endswith("!!!!")
    # iopt 3: relative error = sum|err| / sum|ref| in percent, flagged above 10 eps
    line = r.error_print("TENDENCY_LOC%T", [-1.0, 2.0, 1e-9, 4e-7, 8.0], 160000, ndim=2)
    assert " 2D3 " in line and line.endswith(" !!!!")
    nums = [float(x) for x in re.findall(r"-?0\.\d{13}E[+-]\d\d", line)]
    assert nums[:3] == [-1.0, 2.0, 1e-9] and np.isclose(nums[3], 4e-7 / 160000) and np.isclose(nums[4], 5e-6)
    assert not r.relative_error(5 * float(eps), 1.0)[2]
    # iopt 2: reference sums to (almost) nothing
    assert r.relative_error(1e-3, 0.0)[:2] == (100.0 * 1e-3, 2)
    assert len(r.error_header().split()) == 7


def test_performance_table_uses_reference_nominal_work(pkg):
    t = pkg.report.performance_table(4, 160000, 5000, 32, 0.5)
    last = t.splitlines()[-1]
    assert last.endswith(": TOTAL") and "1 x 4" in last
    # 3 996 006 flop per 100 columns (cloudsc_driver_mod.F90:58): 160000 cols in 0.5 s = 12 787 MFlops/s
    assert int(last.split(":")[1].split()[1]) == int(1e-6 * 3996006 * 1600 / 0.5)
    assert t.splitlines()[0].split()[:5] == ["NUMOMP", "NGPTOT", "#GP-cols", "#BLKS", "NPROMA"]
