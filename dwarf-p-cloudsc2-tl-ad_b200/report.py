"""Text output compatible with the reference's own reports.

* :func:`error_print` reproduces ERROR_PRINT (common/module/validate_mod.F90:263-296): one line
  per validated field, Fortran format ``(1X,A20,1X,I1,'D',I1,5(1X,E20.13),A)``.
* :func:`performance_table` reproduces the TOTAL lines of PERFORMANCE_TIMER%PRINT_PERFORMANCE
  (common/module/timer_mod.F90:114-174) with the reference's nominal work of 3 996 006 flop per
  100 columns (cloudsc2_nl/cloudsc_driver_mod.F90:58), so that GPU and CPU runs can be compared
  in the units the dwarf prints.
"""
from __future__ import annotations

import math
import sys

ZHPM_NL = 3996006.0          # cloudsc_driver_mod.F90:58
EPS = sys.float_info.epsilon


def fortran_e(x: float, width: int = 20, digits: int = 13) -> str:
    """Fortran Ew.d edit descriptor: 0.dddddddddddddE+ee, right-justified in `width`."""
    if x != x:
        return "NaN".rjust(width)
    if math.isinf(x):
        return ("Infinity" if x > 0 else "-Infinity").rjust(width)
    if x == 0.0:
        s = "0." + "0" * digits + "E+00"
        return (("-" if math.copysign(1.0, x) < 0 else "") + s).rjust(width)
    m, e = f"{abs(x):.{digits - 1}E}".split("E")      # d.ddddE+ee with `digits` significant digits
    exp = int(e) + 1
    mant = m.replace(".", "")
    s = f"0.{mant}E{'+' if exp >= 0 else '-'}{abs(exp):02d}"
    if x < 0:
        s = "-" + s
    return s.rjust(width)


def relative_error(sum_abs_err: float, sum_abs_ref: float):
    """(zrelerr in %, iopt, warn) exactly as ERROR_PRINT computes them (:273-289)."""
    if sum_abs_err < EPS:
        rel, iopt = 0.0, 1
    elif sum_abs_ref < EPS:
        rel, iopt = sum_abs_err / (1.0 + sum_abs_ref), 2
    else:
        rel, iopt = sum_abs_err / sum_abs_ref, 3
    return 100.0 * rel, iopt, rel > 10.0 * EPS


def error_print(name: str, stats, ngptot: int, ndim: int = 2) -> str:
    """stats = [min, max, max|err|, sum|err|, sum|ref|] (cloudsc2_validate_host /
    cloudsc2_gpu_validate_dev); ngptot = NGPTOTG for the average error per grid point."""
    vmin, vmax, maxerr, sumerr, sumref = (float(v) for v in stats)
    rel, iopt, warn = relative_error(sumerr, sumref)
    avg = sumerr / float(ngptot)
    nums = "".join(" " + fortran_e(v) for v in (vmin, vmax, maxerr, avg, rel))
    return f" {name:<20s} {ndim:1d}D{iopt:1d}{nums}{' !!!!' if warn else ''}"


def error_header() -> str:
    """The heading the programs print before the field lines (cloudsc2_array_state_mod.F90 VALIDATE)."""
    return (f" {'Variable':<20s} Dim {'MinValue':>20s} {'MaxValue':>20s} {'AbsMaxErr':>20s} "
            f"{'AvgAbsErr/GP':>20s} {'MaxRelErr-%':>20s}")


def performance_table(numomp: int, ngptot: int, nblocks: int, nproma: int, seconds: float,
                      numproc: int = 1, zhpm: float = ZHPM_NL) -> str:
    """Header + per-rank TOTAL + grand TOTAL lines of timer_mod.F90 formats 1000/1002/1003 for a run
    in which every rank processed ngptot/numproc columns in `seconds`."""
    mflops = int(1.0e-06 * zhpm * (ngptot / 100.0) / seconds) if seconds > 0 else 0
    msec = int(seconds * 1000.0)
    hdr = " " + "".join(f"{h:>10s}" for h in ("NUMOMP", "NGPTOT", "#GP-cols", "#BLKS", "NPROMA")) + \
          f" {'tid#':>4s} : " + "".join(f"{h:>10s}" for h in ("Time(msec)", "MFlops/s"))
    per = ngptot // numproc
    lines = [hdr]
    for r in range(numproc):
        lines.append(" " + "".join(f"{v:10d}" for v in (numomp, per, per, nblocks // numproc, nproma)) +
                     f" {-1:4d} : " + f"{msec:10d}{mflops // numproc:10d}" + f" : TOTAL @ rank#{r}")
    lines.append(f" {numproc:6d} x{numomp:2d}" + "".join(f"{v:10d}" for v in (ngptot, ngptot, nblocks, nproma)) +
                 f" {-1:4d} : " + f"{msec:10d}{mflops:10d}" + " : TOTAL")
    return "\n".join(lines)
