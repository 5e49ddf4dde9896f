"""Mechanical Fortran-77 (fixed form, the subset UVic ESCM 2.9 uses) -> C translator.

TEST INFRASTRUCTURE.  Used only by oracle/refgen/gen.py to turn the cpp-expanded reference
sources (read where they lie under /root/reference; nothing is copied into the repository)
into oracle/_ref/ref_gen.c, which is compiled into oracle/_ref/libref.so.  That library is
the *reference's own code* run through a compiler-like tool, with no hand editing; the
hand-written oracle/ restatement is pinned against it (tests/test_cpu_refpin.py).

Translation rules (all mechanical, none specific to a routine):
  * `real` is double (the reference builds with -r8), `integer`/`logical` are int;
  * every identifier X becomes X_ in C (no clashes with libc/libm);
  * COMMON members become file-scope globals (zero initialised, like static COMMON storage);
    all units that declare a block must declare the same members;
  * dummy arguments are pointers (Fortran passes by reference); actual arguments that are not
    variables are passed through C99 compound literals;
  * arrays keep the Fortran column-major layout and declared lower bounds;
  * expressions keep the Fortran parse tree: every binary operation is parenthesised in the
    order the Fortran grammar gives (left to right for + - * /, right to left for **);
    x**n with an integer n is repeated multiplication (what gfortran / ifort emit), with a real
    exponent it is pow();
  * local scalars are zero initialised and local arrays are static (zero initialised);
  * DO loops evaluate their bounds once (Fortran trip-count semantics);
  * statement functions become GCC nested functions;
  * array-section assignments / WHERE become element loops;
  * I/O statements (read/write/print/open/close/inquire/format/namelist), assignments to
    CHARACTER variables and calls to routines named in `skip_calls` are dropped (a comment
    marks each drop);  `stop` becomes abort().
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

# ----------------------------------------------------------------------------- fixed form


def read_fixed_form(text: str):
    """-> list of (label or None, statement text lower-cased outside strings)."""
    stmts = []
    cur = None
    for raw in text.split("\n"):
        line = raw.rstrip()
        if not line.strip():
            continue
        if line[0] in "cC*!":
            continue
        if line.lstrip().startswith("!") or line.lstrip().startswith("#"):
            continue
        line = line.replace("\t", "      ") if line.startswith("\t") else line
        line = _strip_comment(line)
        if not line.strip():
            continue
        cont = len(line) > 5 and line[:5].strip() == "" and line[5] not in " 0"
        if cont:
            if cur is None:
                raise ValueError("continuation without a statement: " + raw)
            cur[1] += line[6:]
            continue
        if cur is not None:
            stmts.append(cur)
        lab = line[:5].strip()
        body = line[6:] if len(line) > 6 else ""
        if lab and not lab.isdigit():
            # free-ish line starting before column 7 (does not occur after cpp in the files used)
            lab, body = "", line
        cur = [lab or None, body]
    if cur is not None:
        stmts.append(cur)
    out = []
    for lab, body in stmts:
        body = _lower_outside_strings(body).strip()
        # `a = b ; c = d` does not occur; keep one statement per entry
        if body:
            out.append((lab, body))
    return out


def _strip_comment(line):
    q = None
    for i, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!" and i != 5:
            return line[:i].rstrip()
    return line


def _lower_outside_strings(s):
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        else:
            if ch in "'\"":
                q = ch
                out.append(ch)
            else:
                out.append(ch.lower())
    return "".join(out)


# ----------------------------------------------------------------------------- tokens

_DOTOPS = ("eq", "ne", "lt", "le", "gt", "ge", "and", "or", "not", "true", "false", "eqv", "neqv")
_TOK = re.compile(r"""
    (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<dot>\.(?:eq|ne|lt|le|gt|ge|and|or|not|true|false|eqv|neqv)\.)
  | (?P<num>(?:\d+\.?\d*(?:[ed][+-]?\d+)?|\.\d+(?:[ed][+-]?\d+)?))
  | (?P<id>[a-z_][a-z0-9_]*)
  | (?P<op>\*\*|//|==|/=|<=|>=|::|[-+*/(),=:<>%])
""", re.X)


def tokenize(s):
    toks, i = [], 0
    while i < len(s):
        if s[i].isspace():
            i += 1
            continue
        m = _TOK.match(s, i)
        if not m:
            raise ValueError(f"cannot tokenize at {s[i:i+20]!r} in {s!r}")
        kind = m.lastgroup
        text = m.group()
        if kind == "num":
            # `1.eq.2` / `1.and.` : the dot belongs to the operator
            mm = re.match(r"(\d+)\.(?:eq|ne|lt|le|gt|ge|and|or|not|eqv|neqv)\.", s[i:])
            if mm:
                text = mm.group(1)
                toks.append(("num", text))
                i += len(text)
                continue
            # `1.d0` vs `1.e-3` handled by the regex; `2.e` followed by id chars (e.g. 2.eq) excluded above
        toks.append((kind, text))
        i = m.end()
    return toks


# ----------------------------------------------------------------------------- AST

@dataclass
class Num:
    text: str
    typ: str


@dataclass
class Str:
    text: str
    typ: str = "char"


@dataclass
class Var:
    name: str
    typ: str = "?"


@dataclass
class Ref:          # name(args): array element, section, function call or statement function
    name: str
    args: list
    typ: str = "?"
    kind: str = "?"  # array | intrinsic | stmtfunc | func


@dataclass
class Rng:          # lo:hi(:step) in a subscript
    lo: object
    hi: object
    typ: str = "int"


@dataclass
class Bin:
    op: str
    a: object
    b: object
    typ: str = "?"


@dataclass
class Un:
    op: str
    a: object
    typ: str = "?"


class Parser:
    def __init__(self, toks):
        self.t = toks
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else (None, None)

    def eat(self, text=None):
        k, v = self.peek()
        if text is not None and v != text:
            raise ValueError(f"expected {text!r}, got {v!r} in {self.t}")
        self.i += 1
        return k, v

    def done(self):
        return self.i >= len(self.t)

    # precedence climbing, Fortran 77 grammar
    def expr(self):
        return self.p_eqv()

    def p_eqv(self):
        a = self.p_or()
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.eat()[1]
            a = Bin(op, a, self.p_or())
        return a

    def p_or(self):
        a = self.p_and()
        while self.peek()[1] == ".or.":
            self.eat()
            a = Bin(".or.", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.peek()[1] == ".and.":
            self.eat()
            a = Bin(".and.", a, self.p_not())
        return a

    def p_not(self):
        if self.peek()[1] == ".not.":
            self.eat()
            return Un(".not.", self.p_not())
        return self.p_rel()

    _REL = {".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=",
            "==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">="}

    def p_rel(self):
        a = self.p_add()
        if self.peek()[1] in self._REL:
            op = self._REL[self.eat()[1]]
            a = Bin(op, a, self.p_add())
        return a

    def p_add(self):
        if self.peek()[1] in ("+", "-"):
            op = self.eat()[1]
            a = Un(op, self.p_mul())
        else:
            a = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.eat()[1]
            a = Bin(op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.eat()[1]
            a = Bin(op, a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_primary()
        if self.peek()[1] == "**":
            self.eat()
            # right associative; the exponent may carry a sign only inside parentheses in
            # standard Fortran, but `x**-2` is accepted by compilers: allow a unary sign
            if self.peek()[1] in ("+", "-"):
                op = self.eat()[1]
                b = Un(op, self.p_pow())
            else:
                b = self.p_pow()
            a = Bin("**", a, b)
        return a

    def p_primary(self):
        k, v = self.peek()
        if k == "num":
            self.eat()
            isreal = ("." in v) or ("e" in v) or ("d" in v)
            return Num(v, "real" if isreal else "int")
        if k == "str":
            self.eat()
            return Str(v)
        if v in (".true.", ".false."):
            self.eat()
            return Num("1" if v == ".true." else "0", "logical")
        if v == "(":
            self.eat()
            e = self.expr()
            self.eat(")")
            return Un("()", e)
        if k == "id":
            self.eat()
            if self.peek()[1] == "(":
                self.eat()
                args = []
                if self.peek()[1] != ")":
                    while True:
                        args.append(self.subscript())
                        if self.peek()[1] == ",":
                            self.eat()
                            continue
                        break
                self.eat(")")
                return Ref(v, args)
            return Var(v)
        raise ValueError(f"unexpected token {v!r} in {self.t}")

    def subscript(self):
        # expr | [expr] : [expr]
        if self.peek()[1] == ":":
            self.eat()
            hi = None
            if self.peek()[1] not in (",", ")"):
                hi = self.expr()
            return Rng(None, hi)
        e = self.expr()
        if self.peek()[1] == ":":
            self.eat()
            hi = None
            if self.peek()[1] not in (",", ")"):
                hi = self.expr()
            return Rng(e, hi)
        return e


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if not p.done():
        raise ValueError(f"trailing tokens in expression {s!r}")
    return e


# ----------------------------------------------------------------------------- symbols

@dataclass
class Sym:
    name: str
    typ: str = None              # int | real | logical | char
    dims: list = None            # list of (lo_ast or None, hi_ast or '*')
    kind: str = "local"          # local | arg | common | param | stmtfunc | func
    block: str = None
    value: object = None         # parameter AST
    sf_args: list = None
    sf_body: object = None


INTRINSICS = {
    # name: (C spelling for real, C spelling for int, result type rule)
    "abs": ("fabs", "abs", "same"), "dabs": ("fabs", None, "real"), "iabs": (None, "abs", "int"),
    "sqrt": ("sqrt", None, "real"), "dsqrt": ("sqrt", None, "real"),
    "exp": ("exp", None, "real"), "dexp": ("exp", None, "real"),
    "log": ("log", None, "real"), "alog": ("log", None, "real"), "dlog": ("log", None, "real"),
    "log10": ("log10", None, "real"), "alog10": ("log10", None, "real"),
    "sin": ("sin", None, "real"), "cos": ("cos", None, "real"), "tan": ("tan", None, "real"),
    "asin": ("asin", None, "real"), "acos": ("acos", None, "real"), "atan": ("atan", None, "real"),
    "atan2": ("atan2", None, "real"), "sinh": ("sinh", None, "real"), "cosh": ("cosh", None, "real"),
    "tanh": ("tanh", None, "real"),
    "max": ("f2c_dmax", "f2c_imax", "same"), "min": ("f2c_dmin", "f2c_imin", "same"),
    "amax1": ("f2c_dmax", None, "real"), "amin1": ("f2c_dmin", None, "real"),
    "dmax1": ("f2c_dmax", None, "real"), "dmin1": ("f2c_dmin", None, "real"),
    "max0": (None, "f2c_imax", "int"), "min0": (None, "f2c_imin", "int"),
    "sign": ("f2c_dsign", "f2c_isign", "same"), "dsign": ("f2c_dsign", None, "real"), "isign": (None, "f2c_isign", "int"),
    "mod": ("fmod", "f2c_imod", "same"), "amod": ("fmod", None, "real"),
    "int": ("(int)", "(int)", "int"), "ifix": ("(int)", "(int)", "int"), "nint": ("f2c_nint", "(int)", "int"),
    "float": ("(double)", "(double)", "real"), "real": ("(double)", "(double)", "real"),
    "dble": ("(double)", "(double)", "real"), "dfloat": ("(double)", "(double)", "real"),
    "aint": ("trunc", None, "real"), "anint": ("round", None, "real"),
}

PRELUDE = r"""/* GENERATED by oracle/refgen (f2c.py) from the cpp-expanded reference Fortran -- do not edit, do not commit. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline double f2c_dmax(double a, double b) { return a > b ? a : b; }
static inline double f2c_dmin(double a, double b) { return a < b ? a : b; }
static inline int f2c_imax(int a, int b) { return a > b ? a : b; }
static inline int f2c_imin(int a, int b) { return a < b ? a : b; }
static inline double f2c_dsign(double a, double b) { return copysign(fabs(a), b); }
static inline int f2c_isign(int a, int b) { return b >= 0 ? abs(a) : -abs(a); }
static inline int f2c_imod(int a, int b) { return a % b; }
static inline int f2c_nint(double a) { return (int)lround(a); }
/* x**n, integer n: repeated multiplication by the binary method (libgcc __powidf2, what gfortran calls / expands) */
static inline double f2c_powi(double x, int n) {
  unsigned m = n < 0 ? -(unsigned)n : (unsigned)n;
  double y = (m & 1) ? x : 1.0;
  while (m >>= 1) { x = x * x; if (m & 1) y = y * x; }
  return n < 0 ? 1.0 / y : y;
}
static inline int f2c_ipow(int x, int n) { int y = 1; if (n < 0) return (x == 1) ? 1 : (x == -1 ? ((n & 1) ? -1 : 1) : 0); while (n-- > 0) y *= x; return y; }
/* NAMELIST reads call this hook when the host has set it (group, n, member names, member addresses, 'r'/'i' per member) */
void (*f2c_namelist_hook)(const char *, int, const char **, void **, const char *) = 0;
static void f2c_stop(const char *where) { fprintf(stderr, "STOP %s\n", where); abort(); }
"""


class TranslateError(Exception):
    pass


IO_KEYWORDS = ("read", "write", "print", "open", "close", "inquire", "format", "namelist", "rewind", "backspace")
DECL_KEYWORDS = ("integer", "real", "double precision", "logical", "character", "dimension", "parameter", "common",
                 "save", "data", "external", "intrinsic", "implicit", "equivalence")


def _split_top(s, sep=","):
    """split on `sep` at parenthesis depth 0, outside strings"""
    parts, depth, q, cur = [], 0, None, []
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur))
    return [p.strip() for p in parts]


def _match_paren(s, i):
    """s[i] == '(' -> index of the matching ')'"""
    depth, q = 0, None
    for j in range(i, len(s)):
        ch = s[j]
        if q:
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses in " + s)


class Unit:
    """one subroutine / function"""

    def __init__(self, kind, name, args, rettype, tr):
        self.kind, self.name, self.args, self.rettype = kind, name, args, rettype
        self.tr = tr
        self.syms: dict[str, Sym] = {}
        self.decl_order = []
        self.body = []          # (label, text)
        self.entries = []       # (name, args, index into body)
        self.seen_exec = False
        self.tmp = 0
        self.common_blocks = {}  # block -> [names]
        self.implicit_none = False
        self.namelists = {}

    def lookup(self, name):
        """symbol with a type, applying the default implicit rule (i-n integer) unless `implicit none` was seen"""
        s = self.syms.get(name)
        if s is not None and s.typ is not None:
            return s
        if self.implicit_none:
            return None
        if s is None:
            s = self.sym(name)
        s.typ = "int" if name[0] in "ijklmn" else "real"
        return s

    def sym(self, name):
        s = self.syms.get(name)
        if s is None:
            s = Sym(name)
            self.syms[name] = s
            self.decl_order.append(name)
        return s


class Translator:
    def __init__(self, skip_calls=(), overrides=None, known_units=(), stub_calls=()):
        self.skip_calls = set(skip_calls)
        self.skip_why = dict(skip_calls) if isinstance(skip_calls, dict) else {}   # name -> reason recorded in the manifest
        self.stub_calls = set(stub_calls)   # diagnostics not translated: reaching one aborts
        self.overrides = dict(overrides or {})   # parameter name -> integer value (grid sizes)
        self.units: list[Unit] = []
        self.commons = {}        # block -> [(name, typ, dims numeric)]
        self.common_syms = {}    # name -> (block, typ, dims)
        self.known_units = set(known_units)
        self.dropped = []

    # ------------------------------------------------------------------ splitting into units
    def load(self, stmts, only=None):
        cur = None
        for lab, s in stmts:
            m = re.match(r"(?:(integer|real|logical|double precision)\s+)?(subroutine|function)\s+([a-z0-9_]+)\s*(?:\((.*)\))?\s*$", s)
            if m and cur is None:
                rettype = {"integer": "int", "real": "real", "double precision": "real", "logical": "logical", None: None}[m.group(1)]
                args = [a.strip() for a in m.group(4).split(",")] if m.group(4) and m.group(4).strip() else []
                cur = Unit(m.group(2), m.group(3), args, rettype, self)
                continue
            if cur is None:
                if s.startswith("program") or s.startswith("block data"):
                    cur = Unit("skip", s, [], None, self)
                continue
            if s == "end" or re.match(r"end\s+(subroutine|function|program)\b", s):
                if cur.kind != "skip" and (only is None or cur.name in only or any(e[0] in only for e in cur.entries)):
                    self.units.append(cur)
                cur = None
                continue
            cur.body.append((lab, s))
        return self

    # ------------------------------------------------------------------ declarations
    def _declare(self, u: Unit, s):
        """returns True when `s` was a declaration"""
        m = re.match(r"(integer|real|double precision|logical|character)\b\s*(\*\s*\d+|\*\s*\(\s*\*\s*\)|\(\s*[^)]*\))?\s*(::)?\s*(.*)$", s)
        if m and not re.match(r"(integer|real|logical|character)\s*(\(.*\))?\s*=", s) and not re.match(r"real\s*\(.*\)\s*[-+*/]", s):
            base = {"integer": "int", "real": "real", "double precision": "real", "logical": "logical", "character": "char"}[m.group(1)]
            rest = m.group(4)
            if base != "char" and m.group(2) and m.group(2).strip().startswith("("):
                # `real(kind=8) x`: double precision, what every real is in this build (mk passes -r8); `real(x)` as an
                # expression is excluded above
                if not (base == "real" and re.fullmatch(r"\(\s*kind\s*=\s*8\s*\)", m.group(2).strip())):
                    raise TranslateError("unsupported kind selector: " + s)
            if rest.startswith("function "):
                return False
            for item in _split_top(rest):
                if not item:
                    continue
                mm = re.match(r"([a-z0-9_]+)\s*(\((.*)\))?\s*(\*\s*\d+)?$", item)
                if not mm:
                    raise TranslateError(f"cannot parse declaration item {item!r} in {s!r}")
                sy = u.sym(mm.group(1))
                sy.typ = base
                if mm.group(3) is not None:
                    sy.dims = self._parse_dims(mm.group(3))
            return True
        if s.startswith("dimension"):
            for item in _split_top(s[len("dimension"):]):
                mm = re.match(r"([a-z0-9_]+)\s*\((.*)\)$", item)
                u.sym(mm.group(1)).dims = self._parse_dims(mm.group(2))
            return True
        if s.startswith("parameter"):
            inner = s[s.index("(") + 1:_match_paren(s, s.index("("))]
            for item in _split_top(inner):
                name, val = item.split("=", 1)
                sy = u.sym(name.strip())
                sy.kind = "param"
                sy.value = parse_expr(val.strip())
            return True
        if s.startswith("common"):
            rest = s[len("common"):].strip()
            for blk, items in _common_segments(rest):
                for item in _split_top(items):
                    if not item:
                        continue
                    mm = re.match(r"([a-z0-9_]+)\s*(\((.*)\))?$", item)
                    if not mm:
                        raise TranslateError(f"cannot parse common item {item!r} in {s!r}")
                    sy = u.sym(mm.group(1))
                    sy.kind = "common"
                    sy.block = blk
                    if mm.group(3) is not None:
                        sy.dims = self._parse_dims(mm.group(3))
                    u.common_blocks.setdefault(blk, []).append(sy.name)
            return True
        if re.match(r"(save|external|intrinsic|implicit)\b", s):
            if re.match(r"implicit\s+none", s):
                u.implicit_none = True
            elif s.startswith("implicit"):
                raise TranslateError("IMPLICIT other than NONE not supported: " + s)
            return True
        if s.startswith("namelist"):
            for grp, items in _common_segments(s[len("namelist"):].strip()):
                u.namelists.setdefault(grp, []).extend(x for x in _split_top(items) if x)
            return True
        if s.startswith("data ") or s.startswith("data("):
            raise TranslateError("DATA statement not supported: " + s)
        if s.startswith("equivalence"):
            raise TranslateError("EQUIVALENCE not supported: " + s)
        return False

    def _parse_dims(self, text):
        dims = []
        for d in _split_top(text):
            if d == "*":
                dims.append((None, "*"))
            elif ":" in _strip_parens_content(d):
                lo, hi = _split_top(d, ":")
                dims.append((parse_expr(lo), "*" if hi.strip() == "*" else parse_expr(hi)))
            else:
                dims.append((None, parse_expr(d)))
        return dims

    # ------------------------------------------------------------------ constant evaluation
    def const_int(self, u, e):
        v = self.const_eval(u, e)
        if v is None:
            return None
        return int(v)

    def const_eval(self, u, e):
        """integer constant folding of parameter expressions (dimension bounds); None if not constant"""
        if isinstance(e, Num):
            if e.typ == "int" or e.typ == "logical":
                return int(e.text)
            return float(e.text.replace("d", "e"))
        if isinstance(e, Var):
            if e.name in self.overrides:
                return self.overrides[e.name]
            sy = u.syms.get(e.name)
            if sy and sy.kind == "param":
                v = self.const_eval(u, sy.value)
                if v is not None and sy.typ == "int":
                    v = int(v)
                return v
            return None
        if isinstance(e, Un):
            v = self.const_eval(u, e.a)
            if v is None:
                return None
            return -v if e.op == "-" else v
        if isinstance(e, Bin):
            a, b = self.const_eval(u, e.a), self.const_eval(u, e.b)
            if a is None or b is None:
                return None
            if e.op == "+":
                return a + b
            if e.op == "-":
                return a - b
            if e.op == "*":
                return a * b
            if e.op == "/":
                if isinstance(a, int) and isinstance(b, int):
                    q = abs(a) // abs(b)
                    return q if (a >= 0) == (b >= 0) else -q
                return a / b
            if e.op == "**":
                return a ** b
            return None
        if isinstance(e, Ref) and e.name in ("max", "min") and not (e.name in u.syms and u.syms[e.name].dims):
            vals = [self.const_eval(u, a) for a in e.args]
            if any(v is None for v in vals):
                return None
            return max(vals) if e.name == "max" else min(vals)
        return None

    # ------------------------------------------------------------------ typing
    def typeof(self, u, e):
        if isinstance(e, (Num, Str)):
            return e.typ
        if isinstance(e, Var):
            sy = u.lookup(e.name)
            if sy is None:
                raise TranslateError(f"{u.name}: undeclared variable {e.name}")
            e.typ = sy.typ
            return sy.typ
        if isinstance(e, Un):
            t = self.typeof(u, e.a)
            e.typ = "logical" if e.op == ".not." else t
            return e.typ
        if isinstance(e, Bin):
            ta, tb = self.typeof(u, e.a), self.typeof(u, e.b)
            if e.op in ("==", "!=", "<", "<=", ">", ">=", ".and.", ".or.", ".eqv.", ".neqv."):
                e.typ = "logical"
            elif e.op == "**":
                e.typ = "real" if "real" in (ta, tb) else "int"
            else:
                e.typ = "real" if "real" in (ta, tb) else "int"
            return e.typ
        if isinstance(e, Ref):
            sy = u.syms.get(e.name)
            if sy is not None and sy.dims is not None:
                e.kind, e.typ = "array", sy.typ
                for a in e.args:
                    if isinstance(a, Rng):
                        if a.lo is not None:
                            self.typeof(u, a.lo)
                        if a.hi is not None:
                            self.typeof(u, a.hi)
                    else:
                        self.typeof(u, a)
                return e.typ
            if sy is not None and sy.kind == "stmtfunc":
                e.kind, e.typ = "stmtfunc", sy.typ
                for a in e.args:
                    self.typeof(u, a)
                return e.typ
            if e.name in INTRINSICS and (sy is None or sy.kind not in ("arg",)) and not (sy is not None and e.name in self.known_units):
                ts = [self.typeof(u, a) for a in e.args]
                rule = INTRINSICS[e.name][2]
                e.kind = "intrinsic"
                e.typ = ("real" if "real" in ts else "int") if rule == "same" else rule
                return e.typ
            if sy is not None and sy.typ is not None and (e.name in self.known_units or sy.kind in ("local", "func")):
                if e.name not in self.known_units:
                    raise TranslateError(f"{u.name}: call of untranslated function {e.name}")
                e.kind, e.typ = "func", sy.typ
                for a in e.args:
                    self.typeof(u, a)
                return e.typ
            raise TranslateError(f"{u.name}: cannot resolve {e.name}(...)")
        if isinstance(e, Rng):
            return "int"
        raise TranslateError(f"typeof: {e!r}")

    # ------------------------------------------------------------------ C expression
    def cexpr(self, u, e, sec=None):
        """sec: list of loop index names for the Rng subscripts / whole arrays of a section assignment"""
        if isinstance(e, Num):
            if e.typ == "real":
                t = e.text.replace("d", "e")
                if t.endswith("."):
                    t += "0"
                if t.startswith("."):
                    t = "0" + t
                t = re.sub(r"\.e", ".0e", t)
                if "." not in t and "e" in t:
                    t = t.replace("e", ".0e")
                return t
            return e.text
        if isinstance(e, Str):
            raise TranslateError("string in arithmetic expression")
        if isinstance(e, Var):
            return self.cvar(u, e.name, sec)
        if isinstance(e, Un):
            a = self.cexpr(u, e.a, sec)
            if e.op == "()":
                return f"({a})"
            if e.op == ".not.":
                return f"(!{a})"
            return f"({e.op}{a})"
        if isinstance(e, Bin):
            self.typeof(u, e)
            a, b = self.cexpr(u, e.a, sec), self.cexpr(u, e.b, sec)
            if e.op == "**":
                ta, tb = e.a.typ if hasattr(e.a, "typ") else self.typeof(u, e.a), self.typeof(u, e.b)
                ta = self.typeof(u, e.a)
                if tb == "int":
                    return f"f2c_powi({a}, {b})" if ta == "real" else f"f2c_ipow({a}, {b})"
                return f"pow({a}, {b})"
            op = {".and.": "&&", ".or.": "||", ".eqv.": "==", ".neqv.": "!="}.get(e.op, e.op)
            return f"({a} {op} {b})"
        if isinstance(e, Ref):
            self.typeof(u, e)
            if e.kind == "array":
                return self.carray(u, e, sec)
            if e.kind == "stmtfunc":
                return f"{e.name}_({', '.join(self.cexpr(u, a, sec) for a in e.args)})"
            if e.kind == "intrinsic":
                return self.cintrinsic(u, e, sec)
            if e.kind == "func":
                return f"{e.name}_({', '.join(self.cactual(u, a) for a in e.args)})"
        raise TranslateError(f"cexpr: {e!r}")

    def cintrinsic(self, u, e, sec):
        real_c, int_c, rule = INTRINSICS[e.name]
        ts = [self.typeof(u, a) for a in e.args]
        isreal = "real" in ts
        fn = real_c if isreal else int_c
        if fn is None:
            fn = real_c or int_c
        args = [self.cexpr(u, a, sec) for a in e.args]
        if fn.startswith("("):
            return f"({fn}({args[0]}))"
        if e.name in ("max", "min", "amax1", "amin1", "dmax1", "dmin1", "max0", "min0"):
            out = args[0]
            for a in args[1:]:
                out = f"{fn}({out}, {a})"
            return out
        return f"{fn}({', '.join(args)})"

    def cvar(self, u, name, sec=None):
        sy = u.lookup(name)
        if sy is None:
            raise TranslateError(f"{u.name}: undeclared variable {name}")
        if sy.dims is not None:
            if sec is None:
                raise TranslateError(f"{u.name}: whole array {name} in scalar context")
            return self.carray(u, Ref(name, [Rng(None, None) for _ in sy.dims]), sec)
        if sy.kind == "arg":
            return f"(*{name}_)"
        if u.kind == "function" and name == u.name:
            return f"{name}_result"
        return self.cname(u, name)

    def cname(self, u, name):
        sy = u.syms.get(name)
        if sy is not None and sy.kind == "common":
            return f"{sy.block}__{name}_"
        return f"{name}_"

    def dim_info(self, u, sy):
        """-> list of (lo C text, extent C text or None)"""
        out = []
        for lo, hi in sy.dims:
            lo_c = "1" if lo is None else self._cdim(u, lo)
            if hi == "*":
                out.append((lo_c, None))
                continue
            hi_c = self._cdim(u, hi)
            lo_v = 1 if lo is None else self.const_int(u, lo)
            hi_v = self.const_int(u, hi)
            if lo_v is not None and hi_v is not None:
                out.append((str(lo_v), str(max(hi_v - lo_v + 1, 0))))
            else:
                out.append((lo_c, f"(({hi_c}) - ({lo_c}) + 1)"))
        return out

    def _cdim(self, u, e):
        v = self.const_int(u, e)
        if v is not None:
            return str(v)
        return self.cexpr(u, e)

    def carray(self, u, e: Ref, sec=None):
        sy = u.syms[e.name]
        if len(e.args) != len(sy.dims):
            raise TranslateError(f"{u.name}: rank mismatch for {e.name}")
        info = self.dim_info(u, sy)
        idx = None
        si = 0
        terms = []
        for a, (lo_c, ext) in zip(e.args, info):
            if isinstance(a, Rng):
                if sec is None:
                    raise TranslateError(f"{u.name}: array section of {e.name} outside a section assignment")
                start = lo_c if a.lo is None else self.cexpr(u, a.lo)
                terms.append(f"(({start}) + {sec[si]} - ({lo_c}))")
                si += 1
            else:
                sub = self.cexpr(u, a, sec)
                if self.typeof(u, a) == "real":
                    sub = f"(int)({sub})"      # real subscript: ifort extension, truncated
                terms.append(f"(({sub}) - ({lo_c}))")
        for t, (lo_c, ext) in reversed(list(zip(terms, info))):
            if idx is None:
                idx = t
            else:
                idx = f"{t} + {ext}*({idx})"
        return f"{self.cname(u, e.name)}[{idx}]"

    def cactual(self, u, a):
        """actual argument -> pointer"""
        if isinstance(a, Var):
            sy = u.lookup(a.name)
            if sy is None:
                raise TranslateError(f"{u.name}: undeclared actual argument {a.name}")
            if sy.kind == "param":
                t = "int" if sy.typ in ("int", "logical") else "double"
                return f"&({t}){{{a.name}_}}"
            if sy.dims is not None:
                return self.cname(u, a.name)
            if sy.kind == "arg":
                return f"{a.name}_"
            if u.kind == "function" and a.name == u.name:
                return f"&{a.name}_result"
            return "&" + self.cname(u, a.name)
        if isinstance(a, Ref):
            self.typeof(u, a)
            if a.kind == "array":
                if any(isinstance(x, Rng) for x in a.args):
                    raise TranslateError(f"{u.name}: array section as actual argument")
                return "&" + self.carray(u, a)
        if isinstance(a, Str):
            return "0"       # character actual (names passed to skipped printing helpers)
        t = self.typeof(u, a)
        ct = "int" if t in ("int", "logical") else "double"
        return f"&({ct}){{{self.cexpr(u, a)}}}"

    # ------------------------------------------------------------------ statements
    def translate_unit(self, u: Unit):
        for a in u.args:
            u.sym(a).kind = "arg"
        if u.kind == "function":
            sy = u.sym(u.name)
            sy.kind = "func"
            if u.rettype:
                sy.typ = u.rettype
        # pass 1: declarations, statement functions, entries
        execs = []
        for lab, s in u.body:
            if not u.seen_exec:
                if self._declare(u, s):
                    continue
                if s.startswith("include"):
                    raise TranslateError("unexpanded include: " + s)
                m = re.match(r"([a-z0-9_]+)\s*\(([a-z0-9_,\s]*)\)\s*=(.*)$", s)
                if m and m.group(1) in u.syms and u.syms[m.group(1)].dims is None and u.syms[m.group(1)].kind == "local":
                    sy = u.syms[m.group(1)]
                    sy.kind = "stmtfunc"
                    sy.sf_args = [x.strip() for x in m.group(2).split(",") if x.strip()]
                    sy.sf_body = parse_expr(m.group(3).strip())
                    continue
            else:
                if self._declare(u, s) and not s.startswith("data"):
                    continue
            m = re.match(r"entry\s+([a-z0-9_]+)\s*(?:\((.*)\))?$", s)
            if m:
                eargs = [a.strip() for a in (m.group(2) or "").split(",") if a.strip()]
                for a in eargs:
                    u.sym(a).kind = "arg"
                u.entries.append((m.group(1), eargs, len(execs)))
                continue
            u.seen_exec = True
            execs.append((lab, s))
        # every declared symbol needs a type (implicit none everywhere in the files used)
        for name, sy in u.syms.items():
            if sy.typ is None:
                if sy.kind == "param":
                    # typed by a later declaration or implicit: infer from the value
                    sy.typ = self._infer_param_type(u, sy)
                elif not u.implicit_none:
                    u.lookup(name)
                else:
                    raise TranslateError(f"{u.name}: {name} has no type (implicit typing is not supported)")
        self._register_commons(u)
        funcs = [(u.name, u.args, 0)] + list(u.entries)
        out = []
        for fname, fargs, start in funcs:
            out.append(self._emit_function(u, fname, fargs, execs[start:]))
        return "\n".join(out)

    def _infer_param_type(self, u, sy):
        v = self.const_eval(u, sy.value)
        return "int" if isinstance(v, int) else "real"

    def _register_commons(self, u):
        for blk, names in u.common_blocks.items():
            desc = []
            for n in names:
                sy = u.syms[n]
                dims = None
                if sy.dims is not None:
                    dims = []
                    for lo, ext in self.dim_info(u, sy):
                        if ext is None or not ext.isdigit():
                            raise TranslateError(f"common /{blk}/ {n}: non-constant extent")
                        dims.append((int(lo), int(ext)))
                desc.append((n, sy.typ, dims))
            if blk in self.commons:
                old = self.commons[blk]
                # units may declare a prefix of the block or the same members; members are matched by name
                oldmap = {d[0]: d for d in old}
                for d in desc:
                    if d[0] in oldmap:
                        if oldmap[d[0]] != d:
                            raise TranslateError(f"common /{blk}/ member {d[0]} declared differently: {oldmap[d[0]]} vs {d}")
                    else:
                        old.append(d)
            else:
                self.commons[blk] = list(desc)
            for d in desc:
                self.common_syms[(blk, d[0])] = (d[1], d[2])

    def _ctype(self, typ):
        return {"int": "int", "logical": "int", "real": "double", "char": "char"}[typ]

    def _emit_function(self, u, fname, fargs, execs):
        lines = []
        ret = "void"
        if u.kind == "function":
            ret = self._ctype(u.syms[u.name].typ)
        params = []
        for a in fargs:
            sy = u.syms[a]
            if sy.typ is None:
                raise TranslateError(f"{u.name}: dummy {a} untyped")
            params.append(f"{self._ctype(sy.typ)} *{a}_")
        body_lines = []
        self.ind = 1
        self.block_stack = []
        for lab, s in execs:
            body_lines.extend(self._stmt(u, lab, s))
        if self.block_stack:
            raise TranslateError(f"{u.name}/{fname}: unterminated blocks {self.block_stack}")
        lines.append(f"{ret} {fname}_({', '.join(params) or 'void'})\n{{")
        # parameters (in declaration order: later ones may use earlier ones)
        for name in u.decl_order:
            sy = u.syms[name]
            if sy.kind == "param":
                if name in self.overrides:
                    lines.append(f"  const int {name}_ = {self.overrides[name]};")
                elif sy.typ in ("int", "logical"):
                    v = self.const_int(u, sy.value)
                    if v is None:
                        raise TranslateError(f"{u.name}: parameter {name} not constant")
                    lines.append(f"  const int {name}_ = {v};")
                else:
                    lines.append(f"  const double {name}_ = {self.cexpr(u, sy.value)};")
        # locals
        other_args = set(a for _, ea, _ in [(u.name, u.args, 0)] + u.entries for a in ea) - set(fargs)
        for name in u.decl_order:
            sy = u.syms[name]
            if sy.kind == "local" or (sy.kind == "arg" and name in other_args):
                if sy.typ == "char":
                    continue
                if sy.kind == "local" and sy.dims is None and name in self.known_units:
                    continue      # type declaration of an external function
                ct = self._ctype(sy.typ)
                if sy.kind == "arg":
                    # dummy of another entry point: never referenced on this path; a null pointer keeps the code compiling
                    lines.append(f"  {ct} *{name}_ = 0;")
                    continue
                if sy.dims is None:
                    lines.append(f"  {ct} {name}_ = 0;")
                else:
                    info = self.dim_info(u, sy)
                    if all(ext is not None and ext.isdigit() for _, ext in info):
                        n = 1
                        for _, ext in info:
                            n *= int(ext)
                        lines.append(f"  static {ct} {name}_[{max(n, 1)}];")
                    else:
                        n = " * ".join(f"(size_t)({ext})" for _, ext in info)
                        lines.append(f"  {ct} *{name}_ = ({ct} *)calloc({n}, sizeof({ct}));   /* automatic array (not freed: test infrastructure) */")
        if u.kind == "function":
            lines.append(f"  {ret} {u.name}_result = 0;")
        # statement functions as nested functions
        for name in u.decl_order:
            sy = u.syms[name]
            if sy.kind == "stmtfunc":
                ps = []
                saved = {}
                for a in sy.sf_args:
                    asy = u.syms.get(a)
                    if asy is None or asy.typ is None:
                        raise TranslateError(f"{u.name}: statement function {name} dummy {a} untyped")
                    ps.append(f"{self._ctype(asy.typ)} {a}_")
                    saved[a] = asy.kind
                # dummies are by value inside the nested function even when the host symbol is a dummy argument
                for a in sy.sf_args:
                    u.syms[a].kind = "local" if u.syms[a].kind == "arg" else u.syms[a].kind
                body = self.cexpr(u, sy.sf_body)
                for a, k in saved.items():
                    u.syms[a].kind = k
                lines.append(f"  auto {self._ctype(sy.typ)} {name}_({', '.join(ps)});")
                lines.append(f"  {self._ctype(sy.typ)} {name}_({', '.join(ps)}) {{ return {body}; }}")
        lines.extend(body_lines)
        lines.append(f"  return{' ' + u.name + '_result' if u.kind == 'function' else ''};")
        lines.append("}\n")
        return "\n".join(lines)

    def _pad(self):
        return "  " * self.ind

    def _stmt(self, u, lab, s):
        out = []
        pad = self._pad()
        # labelled DO terminator
        closes = 0
        if lab is not None:
            while self.block_stack and self.block_stack[-1] == ("dolabel", lab):
                closes += 1
                self.block_stack.pop()
        if lab is not None:
            out.append(self._pad() + f"L{lab}:;")
        body = self._stmt_inner(u, s)
        out.extend(body)
        for _ in range(closes):
            self.ind -= 1
            out.append(self._pad() + "}}")
        return out

    def _drop(self, u, s, why):
        self.dropped.append((u.name, why, s))
        txt = s.replace("*/", "* /")
        return [self._pad() + f"/* dropped ({why}): {txt[:100]} */;"]

    def _stmt_inner(self, u, s):
        pad = self._pad()
        kw = re.match(r"[a-z]+", s)
        kw = kw.group() if kw else ""
        # --- block closers / openers
        if re.match(r"end\s*if$", s):
            self._pop("if")
            return [self._pad() + "}"]
        if re.match(r"end\s*do$", s):
            k = self._pop("do")
            return [self._pad() + k]
        if re.match(r"end\s*where$", s):
            self._pop("where")
            return []
        if s == "else":
            self.ind -= 1
            r = [self._pad() + "} else {"]
            self.ind += 1
            return r
        m = re.match(r"else\s*if\s*\(", s)
        if m:
            j = _match_paren(s, s.index("("))
            cond = self.cexpr(u, parse_expr(s[s.index("(") + 1:j]))
            if s[j + 1:].strip() != "then":
                raise TranslateError("else if without then: " + s)
            self.ind -= 1
            r = [self._pad() + f"}} else if ({cond}) {{"]
            self.ind += 1
            return r
        if re.match(r"if\s*\(", s):
            j = _match_paren(s, s.index("("))
            cond_e = parse_expr(s[s.index("(") + 1:j])
            rest = s[j + 1:].strip()
            cond = self.cexpr(u, cond_e)
            if rest == "then":
                self.block_stack.append(("if", None))
                self.ind += 1
                return [pad + f"if ({cond}) {{"]
            if re.match(r"\d+\s*,\s*\d+\s*,\s*\d+$", rest):
                raise TranslateError("arithmetic IF not supported: " + s)
            self.ind += 1
            inner = self._stmt_inner(u, rest)
            self.ind -= 1
            return [pad + f"if ({cond}) {{"] + inner + [pad + "}"]
        m = re.match(r"do\s+while\s*\(", s)
        if m:
            j = _match_paren(s, s.index("("))
            cond = self.cexpr(u, parse_expr(s[s.index("(") + 1:j]))
            self.block_stack.append(("do", "}"))
            self.ind += 1
            return [pad + f"while ({cond}) {{"]
        m = re.match(r"do\s+(\d+\s+)?([a-z0-9_]+)\s*=\s*(.*)$", s)
        if m and "," in m.group(3):
            label = m.group(1).strip() if m.group(1) else None
            var = m.group(2)
            parts = _split_top(m.group(3))
            if len(parts) not in (2, 3):
                raise TranslateError("bad DO: " + s)
            v = self.cvar(u, var)
            lo = self.cexpr(u, parse_expr(parts[0]))
            hi = self.cexpr(u, parse_expr(parts[1]))
            u.tmp += 1
            t = u.tmp
            if len(parts) == 3:
                st_e = parse_expr(parts[2])
                st_v = self.const_int(u, st_e)
                st = self.cexpr(u, st_e)
            else:
                st_v, st = 1, "1"
            if st_v is not None and st_v > 0:
                head = f"{{ const int hi{t} = {hi}; for ({v} = {lo}; {v} <= hi{t}; {v} += {st_v}) {{"
            elif st_v is not None and st_v < 0:
                head = f"{{ const int hi{t} = {hi}; for ({v} = {lo}; {v} >= hi{t}; {v} += ({st_v})) {{"
            else:
                head = (f"{{ const int hi{t} = {hi}, st{t} = {st}; for ({v} = {lo}; st{t} > 0 ? {v} <= hi{t} : {v} >= hi{t}; "
                        f"{v} += st{t}) {{")
            if label:
                self.block_stack.append(("dolabel", label))
            else:
                self.block_stack.append(("do", "}}"))
            self.ind += 1
            return [pad + head]
        m = re.match(r"where\s*\(", s)
        if m:
            j = _match_paren(s, s.index("("))
            rest = s[j + 1:].strip()
            mask = parse_expr(s[s.index("(") + 1:j])
            if rest:
                return self._section_assign(u, rest, mask)
            self.block_stack.append(("where", mask))
            return []
        # --- simple statements
        if s == "continue":
            return [pad + ";"]
        if s == "return":
            return [pad + (f"return {u.name}_result;" if u.kind == "function" else "return;")]
        if s == "exit":
            return [pad + "break;"]
        if s == "cycle":
            return [pad + "continue;"]
        if kw == "stop" and re.match(r"stop\b", s):
            msg = s[4:].strip().strip("'\"").replace('"', "'")
            return [pad + f'f2c_stop("{msg} ({u.name})");']
        m = re.match(r"read\s*\(\s*[a-z0-9_]+\s*,\s*([a-z0-9_]+)\s*(,[^)]*)?\)$", s)
        if m and m.group(1) in u.namelists:
            # NAMELIST read: handed to the host through a hook (names, addresses, types); the pin tests feed it the
            # values of run/control.in that gen.py recorded in the manifest
            names = u.namelists[m.group(1)]
            syms = [u.lookup(n) for n in names]
            nm = ", ".join(f'"{n}"' for n in names)
            pp = ", ".join(("" if (sy.dims is not None or sy.kind == "arg") else "&") + (f"{n}_" if sy.kind == "arg" else self.cname(u, n)) for n, sy in zip(names, syms))
            ty = "".join("r" if sy.typ == "real" else "i" for sy in syms)
            return [pad + f'{{ static const char *nm[] = {{{nm}}}; void *pp[] = {{{pp}}}; '
                          f'if (f2c_namelist_hook) f2c_namelist_hook("{m.group(1)}", {len(names)}, nm, pp, "{ty}"); }}']
        if kw in IO_KEYWORDS and re.match(kw + r"\s*[\(\*'\"0-9]|" + kw + r"\s+[a-z(]", s) and not re.match(kw + r"\s*(\(.*\))?\s*=[^=]", s):
            return self._drop(u, s, "I/O")
        m = re.match(r"go\s*to\s*(\d+)$", s)
        if m:
            return [pad + f"goto L{m.group(1)};"]
        if re.match(r"go\s*to\b", s):
            raise TranslateError("computed / assigned GOTO not supported: " + s)
        m = re.match(r"call\s+([a-z0-9_]+)\s*(\((.*)\))?$", s)
        if m:
            name = m.group(1)
            if name in self.stub_calls and name not in self.known_units:
                return [pad + f'f2c_stop("untranslated diagnostics routine {name} reached from {u.name}");']
            if name in self.skip_calls or name not in self.known_units:
                if name not in self.skip_calls:
                    raise TranslateError(f"{u.name}: call of untranslated routine {name}")
                return self._drop(u, s, self.skip_why.get(name, "I/O helper"))
            args = []
            if m.group(3) and m.group(3).strip():
                p = Parser(tokenize(m.group(3)))
                while True:
                    args.append(p.subscript())
                    if p.peek()[1] == ",":
                        p.eat()
                        continue
                    break
            # array-section actuals (explicit-shape dummy): copy-in / copy-out through a contiguous temporary
            pre, post, cargs = [], [], []
            for a in args:
                if isinstance(a, Ref) and u.syms.get(a.name) is not None and u.syms[a.name].dims is not None \
                        and any(isinstance(x, Rng) for x in a.args):
                    sy = u.syms[a.name]
                    rngs = [(x, inf) for x, inf in zip(a.args, self.dim_info(u, sy)) if isinstance(x, Rng)]
                    u.tmp += 1
                    tid = u.tmp
                    t = f"sec{tid}"
                    ct = self._ctype(sy.typ)
                    ns, qs = [], []
                    for d, (x, (lo_c, ext)) in enumerate(rngs):
                        start = lo_c if x.lo is None else self.cexpr(u, x.lo)
                        ns.append(f"(({lo_c}) + ({ext}) - ({start}))" if x.hi is None else f"(({self.cexpr(u, x.hi)}) - ({start}) + 1)")
                        qs.append(f"q{tid}_{d}")
                    el = self.carray(u, a, qs)
                    total = " * ".join(ns)
                    flat, mul = [], []
                    for d in range(len(qs)):
                        flat.append(" * ".join([qs[d]] + mul) if mul else qs[d])
                        mul.append(ns[d])
                    flat = " + ".join(flat)
                    loops = "".join(f"for (int {qs[d]} = 0; {qs[d]} < {ns[d]}; {qs[d]}++) " for d in reversed(range(len(qs))))
                    pre.append(pad + f"  {ct} {t}[{total}]; {loops}{t}[{flat}] = {el};")
                    post.append(pad + f"  {loops}{el} = {t}[{flat}];")
                    cargs.append(t)
                else:
                    cargs.append(self.cactual(u, a))
            call = pad + ("  " if pre else "") + f"{name}_({', '.join(cargs)});"
            if pre:
                return [pad + "{"] + pre + [call] + post + [pad + "}"]
            return [call]
        # --- assignment
        eq = _find_assign(s)
        if eq is None:
            raise TranslateError(f"{u.name}: unsupported statement: {s}")
        where_mask = None
        for b in reversed(self.block_stack):
            if b[0] == "where":
                where_mask = b[1]
                break
        lhs = parse_expr(s[:eq].strip())
        lname = lhs.name
        sy = u.lookup(lname)
        if sy is None:
            raise TranslateError(f"{u.name}: assignment to undeclared {lname}")
        if sy.typ == "char":
            return self._drop(u, s, "CHARACTER assignment")
        is_section = (isinstance(lhs, Var) and sy.dims is not None) or (isinstance(lhs, Ref) and any(isinstance(a, Rng) for a in lhs.args))
        if is_section or where_mask is not None:
            return self._section_assign(u, s, where_mask)
        rhs = parse_expr(s[eq + 1:].strip())
        if _has_char(u, rhs):
            return self._drop(u, s, "CHARACTER expression")
        self.typeof(u, rhs)
        return [pad + f"{self.cexpr(u, lhs)} = {self.cexpr(u, rhs)};"]

    def _section_assign(self, u, s, mask):
        pad = self._pad()
        eq = _find_assign(s)
        lhs = parse_expr(s[:eq].strip())
        rhs = parse_expr(s[eq + 1:].strip())
        sy = u.syms[lhs.name]
        if isinstance(lhs, Var):
            lhs = Ref(lhs.name, [Rng(None, None) for _ in sy.dims])
        info = self.dim_info(u, sy)
        loops, sec = [], []
        for a, (lo_c, ext) in zip(lhs.args, info):
            if isinstance(a, Rng):
                u.tmp += 1
                v = f"s{u.tmp}"
                start = lo_c if a.lo is None else self.cexpr(u, a.lo)
                if a.hi is None:
                    if ext is None:
                        raise TranslateError("section of an assumed-size dimension")
                    n = f"(({lo_c}) + ({ext}) - ({start}))"
                else:
                    n = f"(({self.cexpr(u, a.hi)}) - ({start}) + 1)"
                loops.append((v, n))
                sec.append(v)
        body = f"{self.cexpr(u, lhs, sec)} = {self.cexpr(u, rhs, sec)};"
        if mask is not None:
            body = f"if ({self.cexpr(u, mask, sec)}) {body}"
        out = []
        # first section dimension innermost (array element order)
        for v, n in reversed(loops):
            out.append(pad + f"for (int {v} = 0; {v} < {n}; {v}++)")
        out.append(pad + "  " + body)
        return out

    def _pop(self, kind):
        if not self.block_stack or self.block_stack[-1][0] != kind:
            raise TranslateError(f"block mismatch: closing {kind}, stack {self.block_stack}")
        k = self.block_stack.pop()
        if kind != "where":
            self.ind -= 1
        return k[1]

    # ------------------------------------------------------------------ whole file
    def emit(self):
        self.known_units |= {u.name for u in self.units} | {e[0] for u in self.units for e in u.entries}
        bodies = []
        protos = []
        for u in self.units:
            bodies.append(self.translate_unit(u))
        out = [PRELUDE]
        out.append("/* COMMON blocks (zero-initialised static storage) */")
        for blk, members in self.commons.items():
            out.append(f"/* common /{blk}/ */")
            for name, typ, dims in members:
                ct = self._ctype(typ)
                if dims is None:
                    out.append(f"{ct} {blk}__{name}_ = 0;")
                else:
                    n = 1
                    for _, ext in dims:
                        n *= ext
                    out.append(f"{ct} {blk}__{name}_[{max(n, 1)}];")
        # prototypes
        for b in bodies:
            for m in re.finditer(r"^(void|double|int) ([a-z0-9_]+_)\(([^)]*)\)\n\{", b, re.M):
                protos.append(f"{m.group(1)} {m.group(2)}({m.group(3)});")
        out.append("\n".join(protos))
        out.extend(bodies)
        return "\n".join(out)

    def manifest(self):
        man = {"commons": {}, "routines": {}}
        for blk, members in self.commons.items():
            for name, typ, dims in members:
                man["commons"].setdefault(name, []).append({"block": blk, "symbol": f"{blk}__{name}_", "type": typ, "dims": dims})
        for u in self.units:
            for fname, fargs, _ in [(u.name, u.args, 0)] + list(u.entries):
                man["routines"][fname] = {"args": [{"name": a, "type": u.syms[a].typ, "array": u.syms[a].dims is not None} for a in fargs],
                                          "returns": u.syms[u.name].typ if u.kind == "function" else None}
        man["dropped"] = [{"unit": a, "why": b, "stmt": c} for a, b, c in self.dropped]
        return man


def _common_segments(rest):
    """`/a/ x, y(n/2) /b/ z` -> [(a, 'x, y(n/2)'), (b, 'z')]; a '/' inside parentheses is a division"""
    segs, depth, i, blk, cur = [], 0, 0, "", []
    if not rest.startswith("/"):
        blk = ""          # blank common
    while i < len(rest):
        ch = rest[i]
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "/" and depth == 0:
            j = rest.index("/", i + 1)
            if "".join(cur).strip():
                segs.append((blk, "".join(cur).strip().rstrip(",")))
            blk, cur = rest[i + 1:j].strip(), []
            i = j + 1
            continue
        cur.append(ch)
        i += 1
    if "".join(cur).strip():
        segs.append((blk, "".join(cur).strip().rstrip(",")))
    return segs


def _strip_parens_content(s):
    """remove everything inside parentheses (to look for a top-level ':')"""
    out, depth = [], 0
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif depth == 0:
            out.append(ch)
    return "".join(out)


def _find_assign(s):
    """index of the top-level '=' of an assignment (not ==, <=, >=, /=)"""
    depth, q = 0, None
    for i, ch in enumerate(s):
        if q:
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "=" and depth == 0:
            if s[i + 1:i + 2] == "=" or s[i - 1:i] in ("=", "<", ">", "/"):
                continue
            return i
    return None


def _has_char(u, e):
    if isinstance(e, Str):
        return True
    if isinstance(e, Var):
        sy = u.syms.get(e.name)
        return sy is not None and sy.typ == "char"
    if isinstance(e, Ref):
        sy = u.syms.get(e.name)
        if sy is not None and sy.typ == "char":
            return True
        if e.name in ("trim", "len", "len_trim", "adjustl", "char", "ichar"):
            return True
        return any(_has_char(u, a) for a in e.args if not isinstance(a, Rng))
    if isinstance(e, Un):
        return _has_char(u, e.a)
    if isinstance(e, Bin):
        return e.op == "//" or _has_char(u, e.a) or _has_char(u, e.b)
    return False
