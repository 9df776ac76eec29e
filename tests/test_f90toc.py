"""CPU suite: the F90 -> C transliterator (oracle/f90toc.py, TEST INFRASTRUCTURE) on a Fortran
routine written here (not reference text), compiled with gcc and compared with the values Fortran
semantics give.  Independent of /root/reference: runs everywhere."""
import ctypes as C
import importlib.util
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

FUNC_H = """
REAL(KIND=JPRB) :: FSQ, FMIX
REAL(KIND=JPRB) :: PX, PY
FSQ (PX) = PX**2 + &   ! continuation inside a statement function
  & 1.0_JPRB
FMIX(PX,PY) = FSQ(PX) * PY - RCON
"""

MODULE = """
MODULE YOMTST
USE PARKIND1, ONLY : JPRB
IMPLICIT NONE
SAVE
REAL(KIND=JPRB) :: RCON
TYPE :: TTST
REAL(KIND=JPRB),ALLOCATABLE :: CVEC(:)
LOGICAL :: LSW
END TYPE TTST
TYPE(TTST), POINTER :: YRTST => NULL()
END MODULE YOMTST
"""

ROUTINE = """
SUBROUTINE TINY ( KLON, KLEV, PA, PB, &
 & POUT, KOUT )
! exercises the subset the CLOUDSC2 kernels use
USE PARKIND1 , ONLY : JPIM, JPRB
USE YOMTST   , ONLY : RCON, YRTST
IMPLICIT NONE
INTEGER(KIND=JPIM),INTENT(IN)    :: KLON
INTEGER(KIND=JPIM),INTENT(IN)    :: KLEV
REAL(KIND=JPRB)   ,INTENT(IN)    :: PA(KLON,KLEV)
REAL(KIND=JPRB)   ,INTENT(IN)    :: PB(KLON)
REAL(KIND=JPRB)   ,INTENT(OUT)   :: POUT(KLON,0:KLEV)
INTEGER(KIND=JPIM),INTENT(OUT)   :: KOUT(KLON)
INTEGER(KIND=JPIM) :: JL, JK, ISUM
REAL(KIND=JPRB) :: ZW(KLON,KLEV), ZS, ZT
REAL(KIND=JPRB) :: ZSCAL=0.5_JPRB
LOGICAL :: LLA, LLB
#include "tiny.intfb.h"
#include "tiny.func.h"
ASSOCIATE(CVEC=>YRTST%CVEC, LSW=>YRTST%LSW)
ZW(:,:) = 0.0_JPRB
POUT(:,:) = -1.0_JPRB
ISUM = 0
DO JK=KLEV,1,-1
  DO JL=1,KLON
    ZS = -PA(JL,JK)**2                      ! = -(a**2)
    ZT = 2.0_JPRB**3**2                     ! = 2**9, right-associative
    ZW(JL,JK) = ZS + ZT*1.D-3 + PA(JL,JK)**0.5_JPRB + MAX(PA(JL,JK), PB(JL), 0.25_JPRB) &
      & - MIN(1.E0_JPRB, PB(JL))/(1.0_JPRB+PB(JL))**3
    LLA = PA(JL,JK) > 0.5_JPRB .AND. .NOT. PB(JL) >= 0.9_JPRB .OR. JK == 1
    LLB = (PB(JL) /= 0.0_JPRB)
    IF (LLA) THEN
      POUT(JL,JK) = FMIX(ZW(JL,JK), CVEC(JK)) * ZSCAL
    ELSEIF (LLB .AND. LSW) THEN
      POUT(JL,JK) = ZW(JL,JK) - SIGN(2.0_JPRB, -PB(JL)) + ABS(ZS) + EXP(-PB(JL)) + TANH(ZS) &
        & + COSH(PB(JL)) + SQRT(PA(JL,JK))
    ELSE
      POUT(JL,JK) = (7/2)*ZW(JL,JK) + 7.0_JPRB/2   ! integer division, then mixed mode
    ENDIF
    IF (POUT(JL,JK) > 0.0_JPRB) ISUM=ISUM+1
  ENDDO
  POUT(1,0) = POUT(1,0) + JK                 ! loop runs KLEV..1
ENDDO
DO JL=1,KLON
  KOUT(JL) = JL*ISUM
ENDDO
IF (ISUM == 0) GO TO 100
KOUT(1) = -ISUM
100 CONTINUE
CALL TINYSUB(KLON, PA(:,2), KOUT)
END ASSOCIATE
END SUBROUTINE TINY
"""

SUB = """
SUBROUTINE TINYSUB ( KLON, PCOL, KOUT )
USE PARKIND1 , ONLY : JPIM, JPRB
IMPLICIT NONE
INTEGER(KIND=JPIM),INTENT(IN)    :: KLON
REAL(KIND=JPRB)   ,INTENT(IN)    :: PCOL(KLON)
INTEGER(KIND=JPIM),INTENT(INOUT) :: KOUT(KLON)
IF (PCOL(KLON) > 0.5_JPRB) THEN
  KOUT(KLON) = 1000
  RETURN
ENDIF
KOUT(KLON) = -1000
END SUBROUTINE TINYSUB
"""


@pytest.fixture(scope="module")
def f90toc():
    spec = importlib.util.spec_from_file_location("f90toc", ROOT / "oracle" / "f90toc.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["f90toc"] = mod
    spec.loader.exec_module(mod)
    return mod


def _expected(pa, pb, cvec, lsw, rcon):
    klev, klon = pa.shape
    pout = np.full((klev + 1, klon), -1.0)
    isum = 0
    for jk in range(klev, 0, -1):
        for jl in range(klon):
            a, b = pa[jk - 1, jl], pb[jl]
            zs = -(a * a)
            zw = zs + 512.0 * 1e-3 + a ** 0.5 + max(max(a, b), 0.25) - min(1.0, b) / ((1.0 + b) * (1.0 + b) * (1.0 + b))
            lla = (a > 0.5 and not b >= 0.9) or jk == 1
            llb = b != 0.0
            if lla:
                v = ((zw * zw + 1.0) * cvec[jk - 1] - rcon) * 0.5
            elif llb and lsw:
                v = zw - np.copysign(2.0, -b) + abs(zs) + np.exp(-b) + np.tanh(zs) + np.cosh(b) + np.sqrt(a)
            else:
                v = 3 * zw + 3.5
            pout[jk, jl] = v
            isum += v > 0.0
        pout[0, 0] += jk
    kout = np.array([(jl + 1) * isum for jl in range(klon)], dtype=np.int32)
    if isum != 0:
        kout[0] = -isum
    kout[klon - 1] = 1000 if pa[1, klon - 1] > 0.5 else -1000
    return pout, kout


@pytest.mark.parametrize("lsw", [0, 1])
def test_transliterated_routine_computes_what_fortran_semantics_say(f90toc, tmp_path, lsw):
    src = tmp_path / "src"
    (src / "common" / "include").mkdir(parents=True)
    (src / "common" / "module").mkdir(parents=True)
    (src / "common" / "include" / "tiny.func.h").write_text(FUNC_H)
    (src / "common" / "include" / "tiny.intfb.h").write_text("INTERFACE\nEND INTERFACE\n")
    (src / "common" / "module" / "yomtst.F90").write_text(MODULE)
    (src / "tiny.F90").write_text(ROUTINE)
    (src / "tinysub.F90").write_text(SUB)
    out = tmp_path / "out"
    out.mkdir()
    inc = src / "common" / "include"
    mods = f90toc.Modules()
    mods.load(src / "common" / "module" / "yomtst.F90", inc)
    h, c = mods.emit()
    (out / "ref_modules.h").write_text(h)
    (out / "ref_modules.c").write_text(c + "\nvoid set_tst(double r, double *v, int l) { RCON = r; YRTST.CVEC = v; YRTST.LSW = l; }\n")
    (out / "ref_runtime.h").write_text(f90toc.RUNTIME_H)
    protos = {}
    for rel, want in (("tinysub.F90", "TINYSUB"), ("tiny.F90", "TINY")):
        r = f90toc.Routine(mods, rel, protos)
        (out / f"ref_{want.lower()}.c").write_text(r.translate(list(f90toc.logical_lines(src / rel, inc, rel)), want))
    (out / "ref_protos.h").write_text("\n".join(p[0] + ";" for p in protos.values()) + "\n")
    so = out / "libtiny.so"
    subprocess.run(["gcc", "-O2", "-std=gnu11", "-fPIC", "-ffp-contract=off", "-w", "-shared", "-o", str(so),
                    str(out / "ref_tiny.c"), str(out / "ref_tinysub.c"), str(out / "ref_modules.c"), "-lm"], check=True)
    L = C.CDLL(str(so))
    rng = np.random.default_rng(4)
    klon, klev = 7, 5
    pa = rng.uniform(0.05, 1.0, (klev, klon))
    pb = rng.uniform(0.0, 1.0, klon)
    pb[2] = 0.0
    cvec = rng.uniform(1.0, 2.0, klev)
    pout = np.zeros((klev + 1, klon))
    kout = np.zeros(klon, dtype=np.int32)
    L.set_tst(C.c_double(0.125), cvec.ctypes.data_as(C.c_void_p), lsw)
    ci = lambda v: C.byref(C.c_int(v))
    L.ref_tiny(ci(klon), ci(klev), pa.ctypes.data_as(C.c_void_p), pb.ctypes.data_as(C.c_void_p),
               pout.ctypes.data_as(C.c_void_p), kout.ctypes.data_as(C.c_void_p))
    want_pout, want_kout = _expected(pa, pb, cvec, lsw, 0.125)
    assert np.allclose(pout, want_pout, rtol=1e-15, atol=0)
    assert np.array_equal(kout, want_kout)


def test_unsupported_fortran_is_refused_not_skipped(f90toc, tmp_path):
    inc = tmp_path
    bad = tmp_path / "bad.F90"
    for stmt in ("WHERE (PA > 0.0_JPRB) PA = 1.0_JPRB", "PA(1:2) = 0.0_JPRB", "ZX = 1.0_JPRB", "PRINT *, KLON"):
        bad.write_text("SUBROUTINE BAD(KLON, PA)\nINTEGER(KIND=JPIM),INTENT(IN) :: KLON\n"
                       "REAL(KIND=JPRB),INTENT(INOUT) :: PA(KLON)\n" + stmt + "\nEND SUBROUTINE BAD\n")
        r = f90toc.Routine(f90toc.Modules(), "bad.F90", {})
        with pytest.raises(f90toc.F90Error):
            r.translate(list(f90toc.logical_lines(bad, inc, "bad.F90")), "BAD")


def test_operator_precedence_of_the_expression_parser(f90toc):
    r = f90toc.Routine(f90toc.Modules(), "x", {})
    for n in "ABC":
        r.sym[n] = f90toc.Sym(n, "real")
    r.sym["L"] = f90toc.Sym("L", "logical")
    cx = lambda s: r.cx(f90toc.parse_expr(s))
    assert cx("-A**2") == "(-ref_pow2(A))"
    assert cx("A-B-C") == "((A - B) - C)"
    assert cx("A/B*C") == "((A / B) * C)"
    assert cx("A**B**C") == "pow(A, pow(B, C))"
    assert cx("A+B*C**2") == "(A + (B * ref_pow2(C)))"
    assert cx(".NOT.L.AND.A<B.OR.L") == "(((!L) && (A < B)) || L)"
    assert cx("1.E-12_JPRB*A") == "(1.E-12 * A)"
    assert cx("A .GE. 100._JPRB") == "(A >= 100.)"
