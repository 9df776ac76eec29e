#!/usr/bin/env python3
"""f90toc.py -- mechanical Fortran-90 -> C transliterator for the reference's CLOUDSC2 kernels.

TEST INFRASTRUCTURE (see oracle/cloudsc2_oracle.h).  There is no Fortran compiler in this image, so
the reference's own kernel sources cannot be compiled as they are.  This tool reads them WHERE THEY
LIE (under /root/reference/src, never copied into the repo) and emits C with the same statements in
the same order, the same expression trees (every Fortran operator node becomes one parenthesised C
operator node, so association and evaluation order are the Fortran parse's, not C's), the same
explicit-shape arrays (column-major index macros) and the same control flow.  `make -C oracle ref`
writes the generated C and the shared object to oracle/_ref/ (git-ignored).  The hand-written
restatement in oracle/*.c -- and through it the CUDA kernels -- is then checked against this
"reference compiled by other means" (tests/test_oracle_ref.py).

Supported subset = exactly what the kernel files use; anything else raises (no silent skipping):
  free-form source, '&' continuations, '!' comments, #include of *.func.h (statement functions) and
  *.intfb.h (interface blocks: ignored), SUBROUTINE with explicit-shape / scalar dummies, USE ... ONLY,
  REAL/INTEGER/LOGICAL declarations (INTENT, initialisers, lower bounds), ASSOCIATE of derived-type
  components, assignments (scalar, element, whole-array ':' broadcast of a scalar), IF/ELSEIF/ELSE/
  ENDIF, one-line IF, GO TO / labelled CONTINUE, DO/ENDDO (optional stride), CALL with whole arrays / leading-dimension sections
  / scalars by reference, RETURN, and in expressions + - * / ** (integer literal exponents become
  repeated multiplication like gfortran's powi expansion; real exponents become pow()), relational
  and logical operators in both spellings, MIN MAX EXP LOG SQRT TANH COSH ABS SIGN.
Modules (YOMCST, YOETHF, YOECLD, YOECLDP, YOEPHLI, YOPHNC, YOMNCL) are parsed for their variable and
derived-type declarations, emitted as C globals / structs; their HDF5 loader routines are not
translated (oracle/ref_glue.c fills the variables from a cloudsc2_params instead).

Usage: f90toc.py <reference src dir> <out dir>
"""
from __future__ import annotations

import re
import sys
from pathlib import Path

KERNELS = [  # (file relative to src, routine)
    ("cloudsc2_nl/satur.F90", "SATUR"),
    ("cloudsc2_nl/cuadjtqs.F90", "CUADJTQS"),
    ("cloudsc2_nl/cloudsc2.F90", "CLOUDSC2"),
    ("cloudsc2_tl/cuadjtqstl.F90", "CUADJTQSTL"),
    ("cloudsc2_tl/cloudsc2tl.F90", "CLOUDSC2TL"),
    ("cloudsc2_ad/cuadjtqsad.F90", "CUADJTQSAD"),
    ("cloudsc2_ad/cloudsc2ad.F90", "CLOUDSC2AD"),
]
MODULES = ["yomcst", "yoethf", "yoecld", "yoecldp", "yoephli", "yophnc", "yomncl"]
# files that exist once per program directory and must be identical copies (checked, not assumed)
DUPLICATES = {
    "cloudsc2_nl/satur.F90": ["cloudsc2_tl/satur.F90", "cloudsc2_ad/satur.F90"],
    "cloudsc2_nl/cuadjtqs.F90": ["cloudsc2_tl/cuadjtqs.F90", "cloudsc2_ad/cuadjtqs.F90"],
    "cloudsc2_tl/cuadjtqstl.F90": ["cloudsc2_ad/cuadjtqstl.F90"],
    "cloudsc2_tl/cloudsc2tl.F90": ["cloudsc2_ad/cloudsc2tl.F90"],
}


class F90Error(Exception):
    pass


# --------------------------------------------------------------------------------------------
# source -> logical lines
# --------------------------------------------------------------------------------------------
def strip_comment(line: str) -> str:
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out).rstrip()


def logical_lines(path: Path, incdir: Path, rel: str | None = None):
    """Yield (file, first line number, upper-cased statement text)."""
    rel = rel or path.name
    pending, start = "", 0
    for no, raw in enumerate(path.read_text().splitlines(), 1):
        if raw.startswith("#"):
            m = re.match(r'#include\s+"([^"]+)"', raw)
            if m:
                if m.group(1).endswith(".func.h"):
                    yield from logical_lines(incdir / m.group(1), incdir, "common/include/" + m.group(1))
                elif not m.group(1).endswith(".intfb.h"):
                    raise F90Error(f"{rel}:{no}: unsupported include {m.group(1)}")
                continue
            raise F90Error(f"{rel}:{no}: unsupported preprocessor line {raw!r}")
        s = strip_comment(raw).strip()
        if not s:
            continue
        if pending:
            if s.startswith("&"):
                s = s[1:].lstrip()
        else:
            start = no
        if s.endswith("&"):
            pending += s[:-1].rstrip() + " "
            continue
        text = (pending + s).upper()
        pending = ""
        for part in text.split(";"):
            if part.strip():
                yield rel, start, part.strip()
    if pending:
        raise F90Error(f"{rel}: dangling continuation")


# --------------------------------------------------------------------------------------------
# expressions
# --------------------------------------------------------------------------------------------
TOK = re.compile(r"""
    (?P<num>(?:\d+\.(?![A-Z]+\.)\d*|\.\d+|\d+)(?:[ED][+-]?\d+)?(?:_[A-Z0-9]+)?)
  | (?P<dotop>\.(?:AND|OR|NOT|EQV|NEQV|EQ|NE|LT|LE|GT|GE|TRUE|FALSE)\.)
  | (?P<name>[A-Z][A-Z0-9_]*)
  | (?P<op>\*\*|==|/=|<=|>=|=>|[-+*/<>=(),:%])
  | (?P<ws>\s+)
""", re.X)
RELOPS = {".EQ.": "==", ".NE.": "!=", ".LT.": "<", ".LE.": "<=", ".GT.": ">", ".GE.": ">=",
          "==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">="}


def tokenize(s: str):
    toks, pos = [], 0
    while pos < len(s):
        m = TOK.match(s, pos)
        if not m:
            raise F90Error(f"cannot tokenise {s[pos:pos + 20]!r} in {s!r}")
        pos = m.end()
        if m.lastgroup != "ws":
            toks.append((m.lastgroup, m.group()))
    return toks


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i][1] if self.i < len(self.t) else None

    def next(self):
        self.i += 1
        return self.t[self.i - 1]

    def expect(self, v):
        if self.peek() != v:
            raise F90Error(f"expected {v!r}, got {self.peek()!r} in {self.t}")
        self.i += 1

    def done(self):
        return self.i >= len(self.t)

    # Fortran precedence, lowest first
    def expr(self):
        n = self.p_or()
        while self.peek() in (".EQV.", ".NEQV."):
            op = self.next()[1]
            n = ("bin", op, n, self.p_or())
        return n

    def p_or(self):
        n = self.p_and()
        while self.peek() == ".OR.":
            self.next()
            n = ("bin", "||", n, self.p_and())
        return n

    def p_and(self):
        n = self.p_not()
        while self.peek() == ".AND.":
            self.next()
            n = ("bin", "&&", n, self.p_not())
        return n

    def p_not(self):
        if self.peek() == ".NOT.":
            self.next()
            return ("un", "!", self.p_not())
        return self.p_rel()

    def p_rel(self):
        n = self.p_add()
        if self.peek() in RELOPS:
            op = RELOPS[self.next()[1]]
            n = ("bin", op, n, self.p_add())
        return n

    def p_add(self):
        if self.peek() in ("+", "-"):
            op = self.next()[1]
            n = self.p_mul()
            n = ("un", op, n) if op == "-" else n
        else:
            n = self.p_mul()
        while self.peek() in ("+", "-"):
            op = self.next()[1]
            n = ("bin", op, n, self.p_mul())
        return n

    def p_mul(self):
        n = self.p_pow()
        while self.peek() in ("*", "/"):
            op = self.next()[1]
            n = ("bin", op, n, self.p_pow())
        return n

    def p_pow(self):
        n = self.p_primary()
        if self.peek() == "**":
            self.next()
            n = ("bin", "**", n, self.p_pow())      # right-associative
        return n

    def args(self):
        self.expect("(")
        out = []
        if self.peek() == ")":
            self.next()
            return out
        while True:
            if self.peek() == ":":
                self.next()
                out.append(("colon",))
            else:
                e = self.expr()
                if self.peek() == ":":
                    raise F90Error("bounded array sections are not supported")
                out.append(e)
            if self.peek() == ",":
                self.next()
                continue
            self.expect(")")
            return out

    def p_primary(self):
        kind, v = self.next()
        if kind == "num":
            return ("num", v)
        if v in (".TRUE.", ".FALSE."):
            return ("logical", v == ".TRUE.")
        if v == "(":
            e = self.expr()
            self.expect(")")
            return ("paren", e)
        if kind == "name":
            n = ("ref", v, self.args()) if self.peek() == "(" else ("name", v)
            while self.peek() == "%":
                self.next()
                k2, m = self.next()
                if k2 != "name":
                    raise F90Error("bad component reference")
                n = ("member", n, m, self.args() if self.peek() == "(" else None)
            return n
        raise F90Error(f"unexpected token {v!r} in {self.t}")


def parse_expr(s: str):
    p = Parser(tokenize(s))
    e = p.expr()
    if not p.done():
        raise F90Error(f"trailing tokens in expression {s!r}")
    return e


def split_top(s: str, sep: str = ","):
    out, depth, cur = [], 0, []
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        out.append("".join(cur).strip())
    return out


def match_paren(s: str, i: int) -> int:
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise F90Error(f"unbalanced parentheses in {s!r}")


def find_assign(s: str) -> int:
    depth = 0
    for j, ch in enumerate(s):
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "=" and depth == 0:
            prev, nxt = s[j - 1] if j else "", s[j + 1] if j + 1 < len(s) else ""
            if prev in "=/<>" or nxt in "=>":
                continue
            return j
    return -1


# --------------------------------------------------------------------------------------------
# symbols
# --------------------------------------------------------------------------------------------
class Sym:
    def __init__(self, name, typ, dims=None, intent=None, init=None, kind="var", struct=None):
        self.name, self.typ, self.dims, self.intent, self.init = name, typ, dims, intent, init
        self.kind, self.struct = kind, struct   # kind: var | module | stmtfunc | assoc
        self.dummy = False
        self.cexpr = None                       # for associate names / module components

    @property
    def ctype(self):
        return {"real": "double", "int": "int", "logical": "int"}[self.typ]


DECL = re.compile(r"^(REAL|INTEGER|LOGICAL|TYPE)\s*(\([^)]*\))?\s*((?:,\s*[A-Z]+(?:\([^)]*\))?\s*)*)::\s*(.*)$")


def parse_decl(text: str):
    """-> (typ, attrs, struct name or None, [(name, dims-src-list or None, init-src or None)])"""
    m = DECL.match(text)
    if not m:
        return None
    base, spec, attrs, ents = m.groups()
    typ = {"REAL": "real", "INTEGER": "int", "LOGICAL": "logical", "TYPE": "struct"}[base]
    struct = spec.strip("() ") if typ == "struct" else None
    attrs = [a.strip() for a in split_top(attrs.strip().lstrip(","))] if attrs.strip() else []
    out = []
    for e in split_top(ents):
        init = None
        k = e.find("=>")
        if k >= 0:
            e = e[:k].strip()
        else:
            k = find_assign(e)
            if k >= 0:
                e, init = e[:k].strip(), e[k + 1:].strip()
        mm = re.match(r"^([A-Z][A-Z0-9_]*)\s*(?:\((.*)\))?$", e)
        if not mm:
            raise F90Error(f"cannot parse entity {e!r}")
        out.append((mm.group(1), split_top(mm.group(2)) if mm.group(2) else None, init))
    return typ, attrs, struct, out


# --------------------------------------------------------------------------------------------
# modules
# --------------------------------------------------------------------------------------------
class Modules:
    def __init__(self):
        self.vars = {}      # module -> {name: Sym}
        self.types = {}     # type name -> [(member, typ, dims)]
        self.order = []     # emission order: ("var", module, name) / ("type", name)

    def load(self, path: Path, incdir: Path):
        mod, cur_type = None, None
        for rel, no, t in logical_lines(path, incdir, "common/module/" + path.name):
            if t.startswith("MODULE "):
                mod = t.split()[1]
                self.vars[mod] = {}
                continue
            if t == "CONTAINS" or t.startswith("END MODULE"):
                break
            if t.startswith("USE ") or t in ("IMPLICIT NONE", "SAVE"):
                continue
            m = re.match(r"^TYPE\s*(?:::)?\s*([A-Z][A-Z0-9_]*)$", t)
            if m:
                cur_type = m.group(1)
                self.types[cur_type] = []
                self.order.append(("type", cur_type))
                continue
            if t.startswith("END TYPE"):
                cur_type = None
                continue
            d = parse_decl(t)
            if not d:
                raise F90Error(f"{rel}:{no}: unsupported module statement {t!r}")
            typ, attrs, struct, ents = d
            for name, dims, init in ents:
                if cur_type:
                    self.types[cur_type].append((name, typ, dims, "ALLOCATABLE" in attrs))
                else:
                    s = Sym(name, typ, dims, init=init if "PARAMETER" in attrs else None, kind="module",
                            struct=struct)
                    self.vars[mod][name] = s
                    self.order.append(("var", mod, name))

    def emit(self):
        h = ["/* GENERATED by oracle/f90toc.py from the reference's common/module/*.F90 -- do not edit */",
             "#ifndef REF_MODULES_H", "#define REF_MODULES_H", "#define REF_HIDDEN __attribute__((visibility(\"hidden\")))"]
        c = ["/* GENERATED by oracle/f90toc.py -- do not edit */", '#include "ref_modules.h"']
        cty = {"real": "double", "int": "int", "logical": "int"}
        for item in self.order:
            if item[0] == "type":
                h.append(f"struct {item[1]} {{")
                for name, typ, dims, alloc in self.types[item[1]]:
                    if dims and (alloc or dims == [":"]):
                        h.append(f"  {cty[typ]} *{name};   /* {name}({','.join(dims)}) allocatable, lower bound 1 */")
                    elif dims:
                        n = []
                        for d in dims:
                            lo, hi = (d.split(":") + [None])[:2] if ":" in d else ("1", d)
                            n.append(f"(({hi})-({lo})+1)")
                        h.append(f"  {cty[typ]} {name}[{'*'.join(n)}];")
                    else:
                        h.append(f"  {cty[typ]} {name};")
                h.append("};")
            else:
                _, mod, name = item
                s = self.vars[mod][name]
                if s.typ == "struct":
                    h.append(f"extern REF_HIDDEN struct {s.struct} {name};   /* {mod}: TYPE({s.struct}) :: {name} */")
                    c.append(f"struct {s.struct} {name};")
                elif s.init is not None:
                    h.append(f"#define {name} ({s.init})   /* {mod}: PARAMETER */")
                elif s.dims:
                    raise F90Error(f"module array variable {name} not supported")
                else:
                    h.append(f"extern REF_HIDDEN {cty[s.typ]} {name};   /* {mod} */")
                    c.append(f"{cty[s.typ]} {name};")
        h.append("#endif")
        return "\n".join(h) + "\n", "\n".join(c) + "\n"


# --------------------------------------------------------------------------------------------
# routine translation
# --------------------------------------------------------------------------------------------
INTRINSIC_REAL1 = {"EXP": "exp", "LOG": "log", "SQRT": "sqrt", "TANH": "tanh", "COSH": "cosh"}


class Routine:
    def __init__(self, mods: Modules, relpath: str, protos: dict):
        self.mods, self.rel, self.protos = mods, relpath, protos
        self.sym: dict[str, Sym] = {}
        self.args: list[str] = []
        self.name = None
        self.body: list[str] = []
        self.stmtfuncs: dict[str, tuple] = {}   # name -> (args, expr ast, src loc)
        self.used_stmtfuncs: list[str] = []
        self.ind = 1
        self.calls: list[tuple] = []            # (callee, nargs, loc)
        self.block_stack: list[str] = []

    # ---- typing -----------------------------------------------------------------------------
    def typeof(self, n, local=None) -> str:
        k = n[0]
        if k == "num":
            return "real" if re.search(r"[.ED]", n[1].split("_")[0]) else "int"
        if k == "logical":
            return "logical"
        if k == "paren":
            return self.typeof(n[1], local)
        if k == "un":
            return "logical" if n[1] == "!" else self.typeof(n[2], local)
        if k == "bin":
            if n[1] in ("&&", "||", ".EQV.", ".NEQV.", "==", "!=", "<", "<=", ">", ">="):
                return "logical"
            a, b = self.typeof(n[2], local), self.typeof(n[3], local)
            return "real" if "real" in (a, b) else "int"
        if k == "name":
            if local and n[1] in local:
                return "real"
            return self.lookup(n[1]).typ
        if k == "ref":
            if n[1] in self.sym or n[1] in self.stmtfuncs:
                return "real" if n[1] in self.stmtfuncs else self.sym[n[1]].typ
            if n[1] in INTRINSIC_REAL1:
                return "real"
            if n[1] in ("MAX", "MIN", "ABS", "SIGN"):
                ts = [self.typeof(a, local) for a in n[2]]
                return "real" if "real" in ts else "int"
            raise F90Error(f"unknown function or array {n[1]}")
        if k == "member":
            return self.member(n)[1]
        raise F90Error(f"typeof: {n}")

    def lookup(self, name) -> Sym:
        if name not in self.sym:
            raise F90Error(f"{self.rel}: undeclared name {name} (IMPLICIT NONE)")
        return self.sym[name]

    def member(self, n):
        """('member', base, comp, args) -> (C expression, type, is_array)"""
        base, comp, args = n[1], n[2], n[3]
        if base[0] != "name":
            raise F90Error("nested component references are not supported")
        s = self.lookup(base[1])
        if s.typ != "struct":
            raise F90Error(f"{base[1]} is not a derived-type variable")
        for name, typ, dims, alloc in self.mods.types[s.struct]:
            if name == comp:
                return f"{base[1]}.{comp}", typ, bool(dims)
        raise F90Error(f"type {s.struct} has no component {comp}")

    # ---- C emission of expressions ------------------------------------------------------------
    def cnum(self, v: str) -> str:
        v = v.split("_")[0].replace("D", "E")
        return v

    def cx(self, n, local=None) -> str:
        k = n[0]
        if k == "num":
            return self.cnum(n[1])
        if k == "logical":
            return "1" if n[1] else "0"
        if k == "paren":
            return f"({self.cx(n[1], local)})"
        if k == "un":
            return f"({n[1]}{self.cx(n[2], local)})"
        if k == "bin":
            op = n[1]
            if op == "**":
                return self.cpow(n[2], n[3], local)
            if op in (".EQV.", ".NEQV."):
                op = "==" if op == ".EQV." else "!="
                return f"((!!{self.cx(n[2], local)}) {op} (!!{self.cx(n[3], local)}))"
            return f"({self.cx(n[2], local)} {op} {self.cx(n[3], local)})"
        if k == "name":
            if local and n[1] in local:
                return n[1]
            s = self.lookup(n[1])
            if s.dims:
                raise F90Error(f"whole-array reference {n[1]} in a scalar expression")
            return s.cexpr or n[1]
        if k == "ref":
            name, args = n[1], n[2]
            if name in self.sym and self.sym[name].dims:
                s = self.sym[name]
                if len(args) != len(s.dims):
                    raise F90Error(f"rank mismatch in {name}")
                if any(a[0] == "colon" for a in args):
                    raise F90Error(f"array section {name}(:) in a scalar expression")
                return f"{name}({', '.join(self.cx(a, local) for a in args)})"
            if name in self.stmtfuncs:
                if name not in self.used_stmtfuncs:
                    self.note_stmtfunc(name)
                if len(args) != len(self.stmtfuncs[name][0]):
                    raise F90Error(f"statement function {name}: argument count")
                return f"{name}({', '.join(self.cx(a, local) for a in args)})"
            if name in INTRINSIC_REAL1 and len(args) == 1:
                return f"{INTRINSIC_REAL1[name]}({self.cx(args[0], local)})"
            if name in ("MAX", "MIN") and len(args) >= 2:
                t = "d" if self.typeof(n, local) == "real" else "i"
                f = f"ref_{t}{name.lower()}"
                out = self.cx(args[0], local)
                for a in args[1:]:                       # MAX(a,b,c) = max(max(a,b),c)
                    out = f"{f}({out}, {self.cx(a, local)})"
                return out
            if name == "ABS" and len(args) == 1:
                return f"{'fabs' if self.typeof(n, local) == 'real' else 'abs'}({self.cx(args[0], local)})"
            if name == "SIGN" and len(args) == 2 and self.typeof(n, local) == "real":
                return f"copysign({self.cx(args[0], local)}, {self.cx(args[1], local)})"
            raise F90Error(f"unsupported function or undeclared array {name}")
        if k == "member":
            cexpr, typ, is_arr = self.member(n)
            if is_arr:
                if not n[3] or len(n[3]) != 1:
                    raise F90Error("component arrays: rank 1 only")
                return f"{cexpr}[({self.cx(n[3][0], local)}) - 1]"
            return cexpr
        raise F90Error(f"cx: {n}")

    def cpow(self, base, expo, local):
        b = self.cx(base, local)
        e = expo[1] if expo[0] == "paren" else expo
        if e[0] == "num" and self.typeof(e) == "int":
            k = int(e[1].split("_")[0])
            if self.typeof(base, local) == "real" and 1 <= k <= 4:
                return f"ref_pow{k}({b})"
            return f"__builtin_powi({b}, {k})"
        if self.typeof(expo, local) == "int":
            return f"__builtin_powi({b}, {self.cx(expo, local)})"
        return f"pow({b}, {self.cx(expo, local)})"

    def note_stmtfunc(self, name):
        args, ast, loc = self.stmtfuncs[name]
        self.used_stmtfuncs.append(name)          # reserve first to stop recursion
        self.cx(ast, set(args))                   # pulls in the functions it references
        self.used_stmtfuncs.remove(name)
        self.used_stmtfuncs.append(name)          # dependency order: after its callees

    # ---- statements -----------------------------------------------------------------------------
    def out(self, s, loc=None):
        tag = f"   /* {loc[0]}:{loc[1]} */" if loc else ""
        self.body.append("  " * self.ind + s + tag)

    def lhs(self, s: str):
        n = parse_expr(s)
        if n[0] == "name":
            sym = self.lookup(n[1])
            if sym.dims:
                raise F90Error(f"whole-array assignment to {n[1]} without (:)")
            if sym.kind in ("module", "assoc") or (sym.dummy and sym.intent == "IN"):
                raise F90Error(f"assignment to read-only {n[1]}")
            return ("scalar", sym.cexpr or n[1], sym)
        if n[0] == "ref" and n[1] in self.sym and self.sym[n[1]].dims:
            sym = self.sym[n[1]]
            if all(a[0] == "colon" for a in n[2]):
                if len(n[2]) != len(sym.dims):
                    raise F90Error("rank mismatch")
                return ("whole", n[1], sym)
            return ("scalar", self.cx(n), sym)
        raise F90Error(f"unsupported assignment target {s!r}")

    def assignment(self, t, loc):
        k = find_assign(t)
        kind, target, sym = self.lhs(t[:k].strip())
        rhs = parse_expr(t[k + 1:].strip())
        rt = self.typeof(rhs)
        if (sym.typ == "logical") != (rt == "logical"):
            raise F90Error(f"logical/numeric mismatch in {t!r}")
        r = self.cx(rhs)
        if kind == "scalar":
            self.out(f"{target} = {r};", loc)
        else:
            self.out(f"{{ const {sym.ctype} ref_v = {r}; for (long ref_i = 0; ref_i < {target}_size; ++ref_i) "
                     f"{target}_[ref_i] = ref_v; }}", loc)

    def call(self, t, loc):
        m = re.match(r"^CALL\s+([A-Z][A-Z0-9_]*)\s*\((.*)\)$", t)
        if not m:
            raise F90Error(f"unsupported CALL {t!r}")
        callee, args = m.group(1), split_top(m.group(2))
        cargs = []
        for a in args:
            n = parse_expr(a)
            if n[0] == "name":
                s = self.lookup(n[1])
                if s.dims:
                    cargs.append(f"{n[1]}_")
                elif s.kind in ("module", "assoc"):
                    cargs.append(f"&({s.ctype}){{{s.cexpr or n[1]}}}")
                else:
                    cargs.append(f"&{n[1]}")
            elif n[0] == "ref" and n[1] in self.sym and self.sym[n[1]].dims:
                s = self.sym[n[1]]
                idx, seen_scalar = [], False
                for d, a2 in zip(s.dims, n[2]):
                    if a2[0] == "colon":
                        if seen_scalar:
                            raise F90Error(f"non-contiguous section {a}")
                        idx.append(self.cx(d[0]))
                    else:
                        seen_scalar = True
                        idx.append(self.cx(a2))
                cargs.append(f"&{n[1]}({', '.join(idx)})")
            else:
                ty = {"real": "double", "int": "int", "logical": "int"}[self.typeof(n)]
                cargs.append(f"&({ty}){{{self.cx(n)}}}")
        self.calls.append((callee, len(cargs), loc))
        self.out(f"ref_{callee.lower()}({', '.join(cargs)});", loc)

    def statement(self, t, loc):
        if t.startswith("IF") and re.match(r"^IF\s*\(", t):
            i = t.index("(")
            j = match_paren(t, i)
            cond, rest = t[i + 1:j], t[j + 1:].strip()
            c = self.cx(parse_expr(cond))
            if self.typeof(parse_expr(cond)) != "logical":
                raise F90Error(f"IF condition is not logical: {cond}")
            if rest == "THEN":
                self.out(f"if ({c}) {{", loc)
                self.ind += 1
                self.block_stack.append("IF")
            else:
                self.out(f"if ({c}) {{", loc)
                self.ind += 1
                self.statement(rest, None)
                self.ind -= 1
                self.out("}")
            return
        m = re.match(r"^ELSE\s*IF\s*\((.*)\)\s*THEN$", t)
        if m:
            self.ind -= 1
            self.out(f"}} else if ({self.cx(parse_expr(m.group(1)))}) {{", loc)
            self.ind += 1
            return
        if t == "ELSE":
            self.ind -= 1
            self.out("} else {", loc)
            self.ind += 1
            return
        if t in ("ENDIF", "END IF"):
            if self.block_stack.pop() != "IF":
                raise F90Error(f"{loc}: ENDIF closes a DO")
            self.ind -= 1
            self.out("}")
            return
        m = re.match(r"^DO\s+([A-Z][A-Z0-9_]*)\s*=\s*(.*)$", t)
        if m:
            var, parts = m.group(1), split_top(m.group(2))
            if self.lookup(var).typ != "int" or len(parts) not in (2, 3):
                raise F90Error(f"unsupported DO {t!r}")
            lo, hi = self.cx(parse_expr(parts[0])), self.cx(parse_expr(parts[1]))
            step = parse_expr(parts[2]) if len(parts) == 3 else ("num", "1")
            neg = step[0] == "un" and step[1] == "-" and step[2][0] == "num"
            if not (neg or step[0] == "num"):
                raise F90Error(f"DO stride must be a literal: {t!r}")
            st = self.cx(step)
            self.out(f"{{ const int ref_hi = {hi}; for ({var} = {lo}; {var} {'>=' if neg else '<='} ref_hi; "
                     f"{var} += {st}) {{", loc)
            self.ind += 1
            self.block_stack.append("DO")
            return
        if t in ("ENDDO", "END DO"):
            if self.block_stack.pop() != "DO":
                raise F90Error(f"{loc}: ENDDO closes an IF")
            self.ind -= 1
            self.out("} }")
            return
        if t.startswith("CALL "):
            self.call(t, loc)
            return
        if t == "RETURN":
            self.out("goto ref_exit;", loc)
            return
        m = re.match(r"^GO\s*TO\s+(\d+)$", t)
        if m:
            self.out(f"goto L{m.group(1)};", loc)
            return
        m = re.match(r"^(\d+)\s+CONTINUE$", t)
        if m:
            self.out(f"L{m.group(1)}: ;", loc)
            return
        if t.startswith("ASSOCIATE"):
            i = t.index("(")
            for item in split_top(t[i + 1:match_paren(t, i)]):
                nm, sel = [x.strip() for x in item.split("=>")]
                n = parse_expr(sel)
                if n[0] != "member":
                    raise F90Error(f"ASSOCIATE selector {sel}")
                cexpr, typ, is_arr = self.member(n)
                s = Sym(nm, typ, dims=[(("num", "1"), None)] if is_arr else None, kind="assoc")
                if is_arr:
                    self.out(f"{s.ctype} *const {nm}_ = {cexpr};", loc)
                    self.out(f"#define {nm}(i1) {nm}_[(i1) - 1]")
                    self.undefs.append(nm)
                else:
                    self.out(f"const {s.ctype} {nm} = {cexpr};", loc)
                self.sym[nm] = s
            return
        if t.startswith("END ASSOCIATE") or t == "ENDASSOCIATE":
            return
        if find_assign(t) > 0:
            self.assignment(t, loc)
            return
        raise F90Error(f"{loc}: unsupported statement {t!r}")

    # ---- whole routine ----------------------------------------------------------------------------
    def translate(self, lines, want: str) -> str:
        self.undefs = []
        it = iter(lines)
        for rel, no, t in it:
            m = re.match(r"^SUBROUTINE\s+([A-Z][A-Z0-9_]*)\s*\((.*)\)$", t)
            if m:
                self.name = m.group(1)
                self.args = [a.strip() for a in split_top(m.group(2))]
                break
            raise F90Error(f"{rel}:{no}: expected SUBROUTINE, got {t!r}")
        if self.name != want:
            raise F90Error(f"{self.rel}: found {self.name}, wanted {want}")
        in_exec = False
        decl_order = []
        for rel, no, t in it:
            loc = (rel, no)
            if t.startswith("END SUBROUTINE"):
                break
            if not in_exec:
                if t.startswith("USE "):
                    m = re.match(r"^USE\s+([A-Z][A-Z0-9_]*)\s*(?:,\s*ONLY\s*:\s*(.*))?$", t)
                    mod = m.group(1)
                    if mod == "PARKIND1":
                        continue
                    if mod not in self.mods.vars:
                        raise F90Error(f"{rel}:{no}: module {mod} not loaded")
                    names = [x.strip() for x in m.group(2).split(",")] if m.group(2) else list(self.mods.vars[mod])
                    for nm in names:
                        if nm not in self.mods.vars[mod]:
                            raise F90Error(f"{rel}:{no}: {nm} is not in module {mod}")
                        self.sym[nm] = self.mods.vars[mod][nm]
                    continue
                if t == "IMPLICIT NONE":
                    continue
                d = parse_decl(t)
                if d:
                    typ, attrs, struct, ents = d
                    intent = next((a[7:-1].replace(" ", "") for a in attrs if a.startswith("INTENT(")), None)
                    for a in attrs:
                        if not a.startswith("INTENT("):
                            raise F90Error(f"{rel}:{no}: unsupported attribute {a}")
                    for name, dims, init in ents:
                        dd = None
                        if dims:
                            dd = []
                            for x in dims:
                                if ":" in x:
                                    lo, hi = x.split(":")
                                    dd.append((parse_expr(lo), parse_expr(hi)))
                                else:
                                    dd.append((("num", "1"), parse_expr(x)))
                        s = Sym(name, typ, dd, intent, init)
                        s.dummy = name in self.args
                        s.loc = loc
                        self.sym[name] = s
                        decl_order.append(name)
                    continue
                # statement function?  NAME(args) = expr with NAME a declared non-array scalar
                k = find_assign(t)
                m = re.match(r"^([A-Z][A-Z0-9_]*)\s*\(([^()]*)\)\s*$", t[:k]) if k > 0 else None
                if m and m.group(1) in self.sym and not self.sym[m.group(1)].dims \
                        and self.sym[m.group(1)].kind == "var" and not self.sym[m.group(1)].dummy:
                    fargs = [a.strip() for a in m.group(2).split(",")]
                    self.stmtfuncs[m.group(1)] = (fargs, parse_expr(t[k + 1:]), loc)
                    decl_order.remove(m.group(1))
                    del self.sym[m.group(1)]
                    continue
                in_exec = True
            self.statement(t, loc)
        if self.block_stack:
            raise F90Error(f"{self.rel}: unterminated {self.block_stack}")
        return self.assemble(decl_order)

    def assemble(self, decl_order) -> str:
        ctype = {"real": "double", "int": "int", "logical": "int"}
        o = [f"/* GENERATED by oracle/f90toc.py from {self.rel} ({self.name}) -- do not edit, do not commit */",
             '#include "ref_runtime.h"', '#include "ref_modules.h"', '#include "ref_protos.h"', ""]
        stmt_dummies = set()
        for fn in self.used_stmtfuncs:
            args, ast, loc = self.stmtfuncs[fn]
            stmt_dummies.update(args)
            o.append(f"/* statement function {loc[0]}:{loc[1]} */")
            o.append(f"static inline double {fn}({', '.join('double ' + a for a in args)}) "
                     f"{{ return {self.cx(ast, set(args))}; }}")
        for fn, (args, _, _) in self.stmtfuncs.items():
            stmt_dummies.update(args)
        params = []
        for a in self.args:
            s = self.lookup(a)
            params.append(f"{ctype[s.typ]} *{a}_" if s.dims else f"{ctype[s.typ]} *{a}_p")
        proto = f"void ref_{self.name.lower()}({', '.join(params)})"
        self.protos[self.name] = (proto, len(params))
        o += ["", proto, "{"]
        pre, post = [], []
        for a in self.args:                       # scalar dummies first: array bounds use them
            s = self.sym[a]
            if not s.dims:
                if s.intent != "IN":
                    raise F90Error(f"scalar dummy {a} must be INTENT(IN)")
                pre.append(f"  {ctype[s.typ]} {a} = *{a}_p;")
        for name in decl_order:
            s = self.sym[name]
            if s.dims:
                ext = [f"(({self.cx(hi)}) - ({self.cx(lo)}) + 1)" for lo, hi in s.dims]
                pre.append(f"  const long {name}_size = " + " * ".join(f"(long){e}" for e in ext) + ";")
                idx, stride = [], []
                for k, (lo, hi) in enumerate(s.dims):
                    term = f"((i{k + 1}) - ({self.cx(lo)}))"
                    if stride:
                        term = " * ".join(stride) + " * " + term
                    idx.append(term)
                    stride.append(f"(long){ext[k]}")
                margs = ", ".join(f"i{k + 1}" for k in range(len(s.dims)))
                if not s.dummy:
                    pre.append(f"  {ctype[s.typ]} *{name}_ = ({ctype[s.typ]} *)ref_alloc_{ctype[s.typ]}({name}_size);")
                    post.append(f"  ref_free({name}_);")
                pre.append(f"#define {name}({margs}) {name}_[{' + '.join(idx)}]")
                self.undefs.append(name)
            elif not s.dummy and name not in stmt_dummies:
                if s.init is not None:
                    pre.append(f"  static {ctype[s.typ]} {name} = {self.cx(parse_expr(s.init))};   /* initialised => SAVE */")
                elif s.typ == "real":
                    pre.append(f"  double {name} = REF_POISON;")
                else:
                    pre.append(f"  int {name} = 0;")
                pre.append(f"  (void){name};")
        o += pre + [""] + self.body + ["  goto ref_exit;", "ref_exit:"] + post
        o += [f"#undef {u}" for u in self.undefs] + ["}"]
        return "\n".join(o) + "\n"


RUNTIME_H = r"""/* GENERATED by oracle/f90toc.py -- run-time support of the transliterated reference kernels */
#ifndef REF_RUNTIME_H
#define REF_RUNTIME_H
#include <math.h>
#include <stdlib.h>
#include <string.h>
/* Fortran leaves locals undefined; poison them so that any use-before-set shows up as NaN */
#define REF_POISON (__builtin_nan(""))
static inline double ref_dmax(double a, double b) { return a > b ? a : b; }
static inline double ref_dmin(double a, double b) { return a < b ? a : b; }
static inline int ref_imax(int a, int b) { return a > b ? a : b; }
static inline int ref_imin(int a, int b) { return a < b ? a : b; }
/* x**k for literal k as gfortran expands __builtin_powi: repeated multiplication */
static inline double ref_pow1(double x) { return x; }
static inline double ref_pow2(double x) { return x * x; }
static inline double ref_pow3(double x) { return x * x * x; }
static inline double ref_pow4(double x) { double t = x * x; return t * t; }
static inline double *ref_alloc_double(long n) {
  double *p = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  if (!p) abort();
  for (long i = 0; i < n; ++i) p[i] = REF_POISON;
  return p;
}
static inline int *ref_alloc_int(long n) {
  int *p = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  if (!p) abort();
  memset(p, 0x7f, sizeof(int) * (size_t)(n > 0 ? n : 1));
  return p;
}
static inline void ref_free(void *p) { free(p); }
#endif
"""


def main(argv):
    if len(argv) != 3:
        print(__doc__)
        return 2
    src, out = Path(argv[1]), Path(argv[2])
    incdir = src / "common" / "include"
    out.mkdir(parents=True, exist_ok=True)
    for first, others in DUPLICATES.items():
        ref = (src / first).read_bytes()
        for o in others:
            if (src / o).read_bytes() != ref:
                raise F90Error(f"{o} differs from {first}: translate it separately")
    mods = Modules()
    for m in MODULES:
        mods.load(src / "common" / "module" / f"{m}.F90", incdir)
    h, c = mods.emit()
    (out / "ref_modules.h").write_text(h)
    (out / "ref_modules.c").write_text(c)
    (out / "ref_runtime.h").write_text(RUNTIME_H)
    protos: dict[str, tuple] = {}
    routines = []
    for rel, want in KERNELS:
        r = Routine(mods, rel, protos)
        code = r.translate(list(logical_lines(src / rel, incdir, rel)), want)
        (out / f"ref_{want.lower()}.c").write_text(code)
        routines.append(r)
        print(f"f90toc: {rel} -> ref_{want.lower()}.c ({len(r.body)} statements, "
              f"{len(r.used_stmtfuncs)} statement functions)")
    for r in routines:                              # interface check of every CALL
        for callee, n, loc in r.calls:
            if callee not in protos:
                raise F90Error(f"{loc[0]}:{loc[1]}: CALL of untranslated routine {callee}")
            if protos[callee][1] != n:
                raise F90Error(f"{loc[0]}:{loc[1]}: CALL {callee} with {n} arguments, routine has {protos[callee][1]}")
    ph = ["/* GENERATED by oracle/f90toc.py -- prototypes of the transliterated routines (all arguments by",
          " * reference, arrays as pointers to their first element: the gfortran calling convention) */",
          "#ifndef REF_PROTOS_H", "#define REF_PROTOS_H"]
    ph += [p[0] + ";" for p in protos.values()] + ["#endif"]
    (out / "ref_protos.h").write_text("\n".join(ph) + "\n")
    return 0


if __name__ == "__main__":
    try:
        sys.exit(main(sys.argv))
    except F90Error as e:
        print(f"f90toc: ERROR: {e}", file=sys.stderr)
        sys.exit(1)
