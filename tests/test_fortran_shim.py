"""CPU suite: mechanical check of the Fortran binding.

This image has no Fortran compiler, so fortran/cloudsc2_gpu_mod.F90 cannot be compiled here.  What CAN
be checked without one is what a compiler would NOT check either -- that the hand-written
ISO_C_BINDING view agrees with the C side:

  * every `TYPE, BIND(C)` has the fields of the C struct of the same name in include/cloudsc2_b200.h,
    same order, same kind (REAL(C_DOUBLE) <-> double, INTEGER(C_INT) <-> int, TYPE(C_PTR) <-> pointer),
    and the same as the ctypes Structure in _abi.py;
  * every `BIND(C, NAME='...')` interface names a function the header declares and the library
    exports, with the same number of arguments in the same order, each passed the same way
    (VALUE <-> by value, otherwise by reference <-> pointer) and with the same kind, and the same
    result type -- and the ctypes argtypes table says the same;
  * the three replacement driver modules keep the reference's module / subroutine names and dummy
    argument list (cloudsc_driver_mod.F90:22-30) and do not create or destroy the GPU context per call.
"""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
FDIR = ROOT / "dwarf-p-cloudsc2-tl-ad_b200" / "fortran"
HEADER = ROOT / "include" / "cloudsc2_b200.h"


# ---------------------------------------------------------------- C side
def _strip_c(txt):
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return re.sub(r"//[^\n]*", "", txt)


def _c_kind(decl):
    """'const double *x' -> ('ptr', 'double'); 'int klev' -> ('val', 'int') ..."""
    decl = decl.strip()
    if "*" in decl or "[" in decl:
        base = re.sub(r"\bconst\b|\bstruct\b", "", decl.split("*")[0].split("[")[0]).split()
        if "[" in decl and "*" not in decl:      # double znormg[10]: 'double znormg'
            base = base[:-1]
        return ("ptr", " ".join(base))
    toks = re.sub(r"\bconst\b", "", decl).split()
    return ("val", " ".join(toks[:-1]))


def c_structs():
    txt = _strip_c(HEADER.read_text())
    out = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} (\w+);", txt, flags=re.S):
        fields = []
        for stmt in m.group(2).split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            first, *rest = [s.strip() for s in stmt.split(",")]
            kind = _c_kind(first)
            name0 = re.sub(r"[\*\s]", " ", first).split()[-1]
            fields.append((name0, "ptr" if kind[0] == "ptr" else kind[1]))
            for r in rest:
                isptr = "*" in r or kind[0] == "ptr" and "*" in first and False
                fields.append((r.replace("*", "").strip(), "ptr" if "*" in r else kind[1] if not isptr else "ptr"))
        out[m.group(1)] = fields
    return out


def c_prototypes():
    txt = _strip_c(HEADER.read_text())
    protos = {}
    for m in re.finditer(r"(?:^|\n)\s*((?:const\s+)?\w[\w\s]*?[\s\*]+)(cloudsc2_\w+)\s*\(([^;{]*?)\)\s*;", txt):
        ret = m.group(1).strip()
        args = m.group(3).strip()
        alist = [] if args in ("", "void") else [_c_kind(a) for a in args.split(",")]
        protos[m.group(2)] = (("ptr", ret.replace("*", "").replace("const", "").strip()) if "*" in ret else ("val", ret), alist)
    return protos


# ---------------------------------------------------------------- Fortran side
def _join_continuations(txt):
    txt = re.sub(r"!.*", "", txt)
    txt = re.sub(r"&\s*\n\s*&?", " ", txt)
    return txt


def f_types(txt):
    out = {}
    for m in re.finditer(r"TYPE\s*,\s*BIND\(C\)\s*::\s*(\w+)(.*?)END TYPE", txt, flags=re.S | re.I):
        fields = []
        for line in m.group(2).strip().splitlines():
            line = line.strip()
            if not line:
                continue
            t, names = line.split("::")
            t = t.strip().upper().replace(" ", "")
            kind = {"REAL(C_DOUBLE)": "double", "INTEGER(C_INT)": "int", "TYPE(C_PTR)": "ptr"}[t]
            fields += [(n.strip().lower(), kind) for n in names.split(",")]
        out[m.group(1).lower()] = fields
    return out


def f_interfaces(txt):
    """name -> (result kind, [(pass, kind) per dummy argument in order])"""
    out = {}
    body = re.search(r"\n\s*INTERFACE\s*\n(.*?)\n\s*END INTERFACE", txt, flags=re.S | re.I).group(1)
    for m in re.finditer(r"((?:INTEGER\(C_INT\)\s+)?FUNCTION\s+(\w+)\s*\(([^)]*)\)\s*BIND\(C,\s*NAME='(\w+)'\)"
                         r"(?:\s*RESULT\((\w+)\))?)(.*?)END FUNCTION", body, flags=re.S | re.I):
        head, fname, dummies, cname, resvar, decls = m.groups()
        dummies = [d.strip().upper() for d in dummies.split(",") if d.strip()]
        info = {}
        for line in decls.strip().splitlines():
            line = line.strip()
            if not line or line.upper().startswith("IMPORT"):
                continue
            t, names = line.split("::")
            attrs = [a.strip().upper() for a in re.split(r",(?![^()]*\))", t)]
            base = attrs[0].replace(" ", "")
            value = "VALUE" in attrs
            for n in re.split(r",(?![^()]*\))", names):
                nm = re.sub(r"\(.*\)", "", n).strip().upper()
                info[nm] = (base, value)
        result = "int" if head.upper().startswith("INTEGER(C_INT)") else None
        if result is None:
            base, _ = info[resvar.upper()]
            result = {"TYPE(C_PTR)": "ptr"}[base]
        args = []
        for d in dummies:
            base, value = info[d]
            if value:
                kind = {"INTEGER(C_INT)": "int", "REAL(C_DOUBLE)": "double", "INTEGER(C_LONG_LONG)": "long long",
                        "TYPE(C_PTR)": "ptr"}[base]
                args.append(("val" if kind != "ptr" else "ptr", kind if kind != "ptr" else None))
            else:
                kind = {"REAL(C_DOUBLE)": "double", "INTEGER(C_INT)": "int", "CHARACTER(KIND=C_CHAR)": "char",
                        "TYPE(CLOUDSC2_PARAMS)": "cloudsc2_params", "TYPE(CLOUDSC2_FIELDS)": "cloudsc2_fields"}[base]
                args.append(("ptr", kind))
        out[cname] = (result, args)
    return out


@pytest.fixture(scope="module")
def fmod():
    return _join_continuations((FDIR / "cloudsc2_gpu_mod.F90").read_text())


def test_bind_c_types_match_the_c_structs(pkg, fmod):
    cs, fs = c_structs(), f_types(fmod)
    assert set(fs) == {"cloudsc2_params", "cloudsc2_fields"}
    for name, ffields in fs.items():
        cfields = cs[name]
        assert [n for n, _ in ffields] == [n for n, _ in cfields], name            # same order
        assert [k for _, k in ffields] == [k for _, k in cfields], name            # same kinds
    # ... and the ctypes Structures
    ct = {"cloudsc2_params": pkg.Params, "cloudsc2_fields": pkg.Fields}
    kind = {C.c_double: "double", C.c_int: "int", C.c_void_p: "ptr"}
    for name, S in ct.items():
        assert [(n, kind[t]) for n, t in S._fields_] == cs[name], name
    assert C.sizeof(pkg.Params) == 28 * 8 + 4 * 4 and C.sizeof(pkg.Fields) == 18 * 8


def test_bind_c_interfaces_match_the_header_and_the_ctypes_table(pkg, fmod):
    protos, ifaces = c_prototypes(), f_interfaces(fmod)
    lib = pkg.load_library()
    sig = pkg._abi._declare(lib)
    assert len(ifaces) >= 12
    cmap = {"unsigned long long": "long long"}
    for cname, (fres, fargs) in ifaces.items():
        assert cname in protos, f"{cname}: not declared in include/cloudsc2_b200.h"
        (cpass, cres), cargs = protos[cname]
        assert (fres == "int" and cpass == "val" and cres == "int") or (fres == "ptr" and cpass == "ptr"), cname
        assert len(fargs) == len(cargs), f"{cname}: {len(fargs)} Fortran dummies, {len(cargs)} C parameters"
        for i, ((fp, fk), (cp, ck)) in enumerate(zip(fargs, cargs)):
            ck = cmap.get(ck, ck)
            if fp == "val":
                assert cp == "val" and fk == ck, f"{cname} arg {i}: Fortran VALUE {fk} vs C {cp} {ck}"
            elif fk is None:                         # TYPE(C_PTR), VALUE: any pointer
                assert cp == "ptr", f"{cname} arg {i}: TYPE(C_PTR) vs C by-value {ck}"
            else:                                    # by reference
                assert cp == "ptr" and fk == ck, f"{cname} arg {i}: Fortran by-reference {fk} vs C {cp} {ck}"
        # the ctypes table (what the GPU tests actually call through)
        res, argtypes = sig[cname]
        assert len(argtypes) == len(cargs), cname
        for i, (a, (cp, ck)) in enumerate(zip(argtypes, cargs)):
            is_ptr = a in (C.c_void_p, C.c_char_p) or isinstance(a, type) and issubclass(a, C._Pointer)
            assert is_ptr == (cp == "ptr"), f"{cname} arg {i}: ctypes {a} vs C {cp} {ck}"
            if not is_ptr:
                want = {"int": C.c_int, "double": C.c_double, "long long": C.c_longlong,
                        "unsigned long long": C.c_ulonglong}[ck]
                assert a is want, f"{cname} arg {i}: ctypes {a} vs C {ck}"


def test_every_header_prototype_is_parsed():
    """The header parser sees every function the export test sees (so nothing is silently skipped)."""
    protos = c_prototypes()
    txt = _strip_c(HEADER.read_text())
    names = set(re.findall(r"\b(cloudsc2_[a-z0-9_]+)\s*\(", txt))
    assert names == set(protos), names ^ set(protos)


DRIVERS = {"cloudsc_driver_gpu_mod.F90": ("CLOUDSC_DRIVER_MOD", "CLOUDSC_DRIVER", "CLOUDSC2_GPU_NL"),
           "cloudsc_driver_tl_gpu_mod.F90": ("CLOUDSC_DRIVER_TL_MOD", "CLOUDSC_DRIVER_TL", "CLOUDSC2_GPU_TL_TAYLOR"),
           "cloudsc_driver_ad_gpu_mod.F90": ("CLOUDSC_DRIVER_AD_MOD", "CLOUDSC_DRIVER_AD", "CLOUDSC2_GPU_AD_TEST")}
# cloudsc2_nl/cloudsc_driver_mod.F90:22-30 (the TL / AD drivers have the same list)
DRIVER_ARGS = ("NUMOMP NPROMA NLEV NGPTOT NGPTOTG PTSPHY PT PQ TENDENCY_CML TENDENCY_LOC PAP PAPH PLU PLUDE PMFU PMFD "
               "PA PCLV PSUPSAT PCOVPTOT PFPLSL PFPLSN PFHPSL PFHPSN").split()


@pytest.mark.parametrize("fname", sorted(DRIVERS))
def test_driver_modules_keep_the_reference_interface(fname):
    mod, sub, entry = DRIVERS[fname]
    txt = _join_continuations((FDIR / fname).read_text())
    assert re.search(rf"^\s*MODULE {mod}\s*$", txt, flags=re.M) and re.search(rf"END MODULE {mod}", txt)
    m = re.search(rf"SUBROUTINE {sub}\s*\(([^)]*)\)", txt)
    assert [a.strip() for a in m.group(1).split(",")] == DRIVER_ARGS
    body = txt[m.end():]
    assert body.count(f"{entry}(") == 1                      # ONE library call replaces the block loop
    # the context outlives the call: created by the first CLOUDSC2_GPU_SETUP, never finalized per call
    assert "CLOUDSC2_GPU_SETUP(" in body and "CLOUDSC2_GPU_FINALIZE" not in body and "CLOUDSC2_GPU_INIT" not in body


def test_setup_is_once_only_and_multi_gpu(fmod):
    setup = fmod[fmod.index("SUBROUTINE CLOUDSC2_GPU_SETUP"):fmod.index("END SUBROUTINE CLOUDSC2_GPU_SETUP")]
    assert "IF (LGPU_INIT) THEN" in setup and "RETURN" in setup and "LGPU_INIT = .TRUE." in setup
    assert "CLOUDSC2_GPU_INIT_MULTI(" in setup                # one Fortran process drives all GPUs
    # the reference's reference-side check, when a compiler exists: gfortran -fsyntax-only needs the
    # reference's modules; recorded in INTEGRATION.md, not run here
    assert (ROOT / "INTEGRATION.md").exists()
