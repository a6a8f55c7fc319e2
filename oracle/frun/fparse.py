"""fparse.py -- free-form Fortran 90 source -> syntax trees, for oracle/frun (TEST INFRASTRUCTURE ONLY).

Why this exists: the reference (CABLE, /root/reference) is Fortran and neither this container nor the GPU box has a
Fortran compiler (profiles/r02_fortran_compiler_probe.txt), so the reference cannot be built into oracle/_ref.  Instead
of trusting only hand restatements, oracle/frun EXECUTES the reference's own source files, unmodified, where they lie
under /root/reference: this module parses them, finterp.py interprets them with the declared kinds (REAL = binary32,
REAL(r_2) = binary64), per-operator promotion and IEEE arithmetic.  tests/golden/make_fortran_golden.py drives the
reference SUBROUTINE cbm that way and commits its outputs as golden vectors; the C++ oracle and the CUDA path are then
checked against those.  Nothing here knows anything about CABLE: it is language semantics only.

Subset: what the hot-path files use -- modules, USE (ONLY / renames), derived types with default initialisation,
PARAMETERs, generic interfaces (MODULE PROCEDURE), subroutines / functions (RESULT, ELEMENTAL, OPTIONAL, internal
procedures), explicit-shape / assumed-shape / allocatable / pointer arrays with lower bounds, whole-array expressions
and sections, WHERE / ELSEWHERE (nested), IF, DO, DO WHILE, SELECT CASE, EXIT / CYCLE / RETURN / STOP, ALLOCATE /
DEALLOCATE, array constructors.  I/O statements parse and are ignored (PRINT / WRITE) or fail when executed.
"""
from __future__ import annotations

import re

# ------------------------------------------------------------------------------------------------------------------
# preprocessor (the few #ifdef blocks of cable_define_types.F90 etc.; no macro is defined in an offline build)


def cpp(text: str, defines=()) -> str:
    out, stack = [], []          # stack of (taking, taken_already)
    defs = set(defines)
    for line in text.split("\n"):
        s = line.strip()
        if s.startswith("#"):
            d = s[1:].strip()
            m = re.match(r"(ifdef|ifndef|if|elif|else|endif|define|undef|include)\b\s*(.*)", d)
            if not m:
                out.append("")
                continue
            kw, rest = m.group(1), m.group(2).strip()
            if kw in ("ifdef", "ifndef", "if"):
                if kw == "if":
                    mm = re.match(r"defined\s*\(?\s*(\w+)\s*\)?", rest)
                    val = (mm.group(1) in defs) if mm else False
                else:
                    val = (rest.split()[0] in defs) == (kw == "ifdef")
                parent = all(t for t, _ in stack)
                stack.append((parent and val, parent and val))
            elif kw == "elif":
                t, done = stack.pop()
                parent = all(t2 for t2, _ in stack)
                mm = re.match(r"defined\s*\(?\s*(\w+)\s*\)?", rest)
                val = (mm.group(1) in defs) if mm else False
                take = parent and (not done) and val
                stack.append((take, done or take))
            elif kw == "else":
                t, done = stack.pop()
                parent = all(t2 for t2, _ in stack)
                stack.append((parent and not done, True))
            elif kw == "endif":
                stack.pop()
            elif kw == "define" and all(t for t, _ in stack):
                defs.add(rest.split()[0])
            out.append("")
            continue
        out.append(line if all(t for t, _ in stack) else "")
    return "\n".join(out)


# ------------------------------------------------------------------------------------------------------------------
# physical lines -> logical statements (lower-cased outside character literals)

def _strip(line: str):
    out, q, i, n = [], None, 0, len(line)
    while i < n:
        c = line[i]
        if q:
            out.append(c)
            if c == q:
                if i + 1 < n and line[i + 1] == q:
                    out.append(q); i += 1
                else:
                    q = None
        elif c in "'\"":
            q = c; out.append(c)
        elif c == "!":
            break
        else:
            out.append(c.lower())
        i += 1
    code = "".join(out).rstrip()
    cont = code.endswith("&")
    if cont:
        code = code[:-1]
    return code, cont


def _split_semicolons(s: str):
    parts, cur, q = [], [], None
    for c in s:
        if q:
            cur.append(c)
            if c == q:
                q = None
        elif c in "'\"":
            q = c; cur.append(c)
        elif c == ";":
            parts.append("".join(cur)); cur = []
        else:
            cur.append(c)
    parts.append("".join(cur))
    return [p.strip() for p in parts if p.strip()]


def logical_lines(text: str):
    """-> [(first line number, statement text)]"""
    res, buf, start, pending = [], "", 0, False
    for ln, raw in enumerate(text.split("\n"), 1):
        code, cont = _strip(raw)
        if pending:
            s = code.lstrip()
            if not s and not cont:
                if raw.strip() == "" or raw.strip().startswith("!"):
                    continue                      # comment / blank line inside a continued statement
            if s.startswith("&"):
                s = s[1:]
            buf += " " + s if not buf.endswith(("'", '"')) or True else s
        else:
            if not code.strip():
                continue
            buf, start = code, ln
        pending = cont
        if not pending:
            for part in _split_semicolons(buf):
                res.append((start, part))
            buf = ""
    if buf.strip():
        res.append((start, buf.strip()))
    return res


# ------------------------------------------------------------------------------------------------------------------
# tokens

_TOK = re.compile(r"""\s*(?:
   (?P<dotop>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.(?:_\w+)?)
  |(?P<real>(?:\d+\.(?![a-z]+\.)\d*|\.\d+)(?:[ed][+-]?\d+)?(?:_\w+)?|\d+[ed][+-]?\d+(?:_\w+)?)
  |(?P<int>\d+(?:_\w+)?)
  |(?P<name>[a-z_$]\w*)
  |(?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  |(?P<op>\*\*|//|==|/=|<=|>=|=>|::|\(/|/\)|[-+*/<>=(),:%\[\]])
 )""", re.X)


class ParseError(Exception):
    pass


def tokenize(s: str):
    toks, pos, n = [], 0, len(s)
    while pos < n:
        if s[pos:].strip() == "":
            break
        m = _TOK.match(s, pos)
        if not m or m.end() == pos:
            raise ParseError(f"cannot tokenize at {s[pos:pos + 20]!r} in {s!r}")
        kind = m.lastgroup
        toks.append((kind, m.group(kind)))
        pos = m.end()
    return toks


# ------------------------------------------------------------------------------------------------------------------
# expressions.  Nodes are tuples:
#   ('num', text, kind)           kind: 'i' | 'r'          text incl. exponent / kind suffix
#   ('log', bool)  ('str', text)
#   ('des', [(name, args|None), (member, args|None), ...])      designator / function reference
#   ('bin', op, l, r)  ('un', op, e)  ('arr', [items])  ('ido', [items], var, lo, hi, step)
#   args: list of expr | ('kw', name, expr) | ('sec', lo|None, hi|None, step|None)

_REL = {"==": "==", "/=": "/=", "<": "<", "<=": "<=", ">": ">", ">=": ">=",
        ".eq.": "==", ".ne.": "/=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}


class ExprParser:
    def __init__(self, toks, text=""):
        self.t, self.i, self.text = toks, 0, text

    def peek(self, k=0):
        j = self.i + k
        return self.t[j] if j < len(self.t) else (None, None)

    def next(self):
        tok = self.peek(); self.i += 1
        return tok

    def at(self, v):
        return self.peek()[1] == v and self.peek()[0] in ("op", "dotop")

    def accept(self, v):
        if self.at(v):
            self.i += 1
            return True
        return False

    def expect(self, v):
        if not self.accept(v):
            raise ParseError(f"expected {v!r} at token {self.i} ({self.peek()}) in {self.text!r}")

    def done(self):
        return self.i >= len(self.t)

    # precedence levels, lowest first
    def expr(self):
        left = self.p_or()
        while self.at(".eqv.") or self.at(".neqv."):
            op = self.next()[1]
            left = ("bin", op, left, self.p_or())
        return left

    def p_or(self):
        left = self.p_and()
        while self.accept(".or."):
            left = ("bin", ".or.", left, self.p_and())
        return left

    def p_and(self):
        left = self.p_not()
        while self.accept(".and."):
            left = ("bin", ".and.", left, self.p_not())
        return left

    def p_not(self):
        if self.accept(".not."):
            return ("un", ".not.", self.p_not())
        return self.p_rel()

    def p_rel(self):
        left = self.p_cat()
        k, v = self.peek()
        if k in ("op", "dotop") and v in _REL:
            self.next()
            return ("bin", _REL[v], left, self.p_cat())
        return left

    def p_cat(self):
        left = self.p_add()
        while self.accept("//"):
            left = ("bin", "//", left, self.p_add())
        return left

    def p_add(self):
        if self.at("+") or self.at("-"):
            op = self.next()[1]
            left = ("un", op, self.p_mul())
        else:
            left = self.p_mul()
        while self.at("+") or self.at("-"):
            op = self.next()[1]
            left = ("bin", op, left, self.p_mul())
        return left

    def p_mul(self):
        left = self.p_pow()
        while self.at("*") or self.at("/"):
            op = self.next()[1]
            if self.at("+") or self.at("-"):            # a * -b (common extension)
                uop = self.next()[1]
                right = ("un", uop, self.p_pow())
            else:
                right = self.p_pow()
            left = ("bin", op, left, right)
        return left

    def p_pow(self):
        base = self.p_primary()
        if self.accept("**"):
            if self.at("+") or self.at("-"):
                uop = self.next()[1]
                return ("bin", "**", base, ("un", uop, self.p_pow()))
            return ("bin", "**", base, self.p_pow())        # right associative
        return base

    def p_primary(self):
        k, v = self.peek()
        if k is None:
            raise ParseError(f"unexpected end of expression in {self.text!r}")
        if k == "int":
            self.next(); return ("num", v, "i")
        if k == "real":
            self.next(); return ("num", v, "r")
        if k == "str":
            self.next()
            q = v[0]
            return ("str", v[1:-1].replace(q + q, q))
        if k == "dotop" and v.startswith((".true.", ".false.")):
            self.next(); return ("log", v.startswith(".true."))
        if k == "op" and v == "(":
            self.next()
            e = self.expr()
            self.expect(")")
            return ("par", e)
        if k == "op" and v in ("(/", "["):
            self.next()
            close = "/)" if v == "(/" else "]"
            items = []
            if not self.at(close):
                while True:
                    items.append(self.p_ac_item())
                    if not self.accept(","):
                        break
            self.expect(close)
            return ("arr", items)
        if k == "name":
            return self.p_designator()
        raise ParseError(f"unexpected token {v!r} in {self.text!r}")

    def p_ac_item(self):
        # implied do: ( items , i = lo , hi [, step] )
        if self.at("("):
            save = self.i
            try:
                self.next()
                items = [self.expr()]
                while self.accept(","):
                    if self.peek()[0] == "name" and self.peek(1) == ("op", "="):
                        var = self.next()[1]; self.next()
                        lo = self.expr(); self.expect(","); hi = self.expr()
                        step = self.expr() if self.accept(",") else None
                        self.expect(")")
                        return ("ido", items, var, lo, hi, step)
                    items.append(self.expr())
                raise ParseError("not an implied do")
            except ParseError:
                self.i = save
        return self.expr()

    def p_designator(self):
        parts = []
        while True:
            k, v = self.next()
            if k != "name":
                raise ParseError(f"expected a name, got {v!r} in {self.text!r}")
            args = None
            if self.at("("):
                args = self.p_args()
                if self.at("("):                     # substring a(i)(1:3): keep as second arg list (strings only)
                    sub = self.p_args()
                    parts.append((v, args)); parts.append(("$substr", sub))
                    if self.accept("%"):
                        continue
                    break
            parts.append((v, args))
            if not self.accept("%"):
                break
        return ("des", parts)

    def p_args(self):
        self.expect("(")
        args = []
        if self.accept(")"):
            return args
        while True:
            args.append(self.p_arg())
            if not self.accept(","):
                break
        self.expect(")")
        return args

    def p_arg(self):
        if self.peek()[0] == "name" and self.peek(1) == ("op", "=") and self.peek(2) != ("op", "="):
            name = self.next()[1]; self.next()
            return ("kw", name, self.expr())
        lo = None
        if not self.at(":"):
            lo = self.expr()
            if not self.at(":"):
                return lo
        self.expect(":")
        hi = step = None
        if not (self.at(",") or self.at(")") or self.at(":")):
            hi = self.expr()
        if self.accept(":"):
            step = self.expr()
        return ("sec", lo, hi, step)


def parse_expr(text: str):
    p = ExprParser(tokenize(text), text)
    e = p.expr()
    if not p.done():
        raise ParseError(f"trailing tokens in expression {text!r}")
    return e


# ------------------------------------------------------------------------------------------------------------------
# helpers on raw statement text

def match_paren(s: str, i: int) -> int:
    """index of the ')' matching the '(' at s[i] (character literals respected)."""
    depth, q = 0, None
    for j in range(i, len(s)):
        c = s[j]
        if q:
            if c == q:
                q = None
        elif c in "'\"":
            q = c
        elif c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ParseError(f"unbalanced parentheses in {s!r}")


def split_top(s: str, sep: str = ","):
    """split at top-level separators (outside parentheses, brackets, (/ /) and literals)."""
    parts, cur, depth, q, i, n = [], [], 0, None, 0, len(s)
    while i < n:
        c = s[i]
        if q:
            cur.append(c)
            if c == q:
                q = None
        elif c in "'\"":
            q = c; cur.append(c)
        elif c in "([":
            depth += 1; cur.append(c)
        elif c in ")]":
            depth -= 1; cur.append(c)
        elif c == sep and depth == 0:
            parts.append("".join(cur)); cur = []
        else:
            cur.append(c)
        i += 1
    parts.append("".join(cur))
    return [p.strip() for p in parts]


def find_top(s: str, target: str, start: int = 0) -> int:
    """first top-level occurrence of `target` (not inside parentheses / literals), or -1."""
    depth, q, i, n, L = 0, None, start, len(s), len(target)
    while i < n:
        c = s[i]
        if q:
            if c == q:
                q = None
        elif c in "'\"":
            q = c
        elif c in "([":
            depth += 1
        elif c in ")]":
            depth -= 1
        elif depth == 0 and s.startswith(target, i):
            return i
        i += 1
    return -1


# ------------------------------------------------------------------------------------------------------------------
# declarations

_TYPE_START = re.compile(r"^(real|integer|logical|character|complex|double\s*precision|type\s*\(|class\s*\()")


class Decl:
    """one declared entity"""
    __slots__ = ("name", "base", "kind", "tname", "dims", "attrs", "init", "charlen", "intent")

    def __init__(self):
        self.name = None; self.base = None; self.kind = None; self.tname = None
        self.dims = None          # None (scalar) or list of (lo_expr|None, hi_expr|None|'*'|':')
        self.attrs = set(); self.init = None; self.charlen = None; self.intent = None

    def __repr__(self):
        return f"Decl({self.name}:{self.base}{'*' + str(self.kind) if self.kind else ''}{self.dims or ''} {sorted(self.attrs)})"


def parse_dims(text: str):
    dims = []
    for d in split_top(text):
        if d == ":":
            dims.append((None, ":"))
        elif d == "*":
            dims.append((None, "*"))
        else:
            k = find_top(d, ":")
            if k >= 0:
                lo, hi = d[:k].strip(), d[k + 1:].strip()
                if hi == "":
                    dims.append((parse_expr(lo), ":"))
                elif hi == "*":
                    dims.append((parse_expr(lo), "*"))
                else:
                    dims.append((parse_expr(lo) if lo else None, parse_expr(hi)))
            else:
                dims.append((None, parse_expr(d)))
    return dims


def parse_declaration(stmt: str):
    """'real(r_2), dimension(mp,ms), intent(in) :: a, b(3) = 0.0'  ->  [Decl, ...]   (None if not a declaration)"""
    m = _TYPE_START.match(stmt)
    if not m:
        return None
    s = stmt
    base = m.group(1).replace(" ", "")
    pos = m.end()
    kind = None; tname = None; charlen = None
    if base.startswith(("type(", "class(")):
        j = match_paren(s, pos - 1)
        tname = s[pos:j].strip()
        base = "type"
        pos = j + 1
    else:
        if base == "doubleprecision":
            base, kind = "real", ("num", "8", "i")
        rest = s[pos:].lstrip()
        off = len(s) - len(rest)
        if rest.startswith("*") and base != "character":           # real*8
            mm = re.match(r"\*\s*(\d+)", rest)
            kind = ("num", mm.group(1), "i"); pos = off + mm.end()
        elif rest.startswith("("):
            j = match_paren(s, off)
            inner = s[off + 1:j].strip()
            if base == "character":
                charlen = inner
            else:
                inner = re.sub(r"^kind\s*=\s*", "", inner)
                kind = parse_expr(inner)
            pos = j + 1
        elif rest.startswith("*") and base == "character":
            mm = re.match(r"\*\s*(\(\s*\*\s*\)|\d+)", rest)
            charlen = mm.group(1); pos = off + mm.end()
    rest = s[pos:].strip()
    k = find_top(rest, "::")
    attrs_text, ents_text = (rest[:k], rest[k + 2:]) if k >= 0 else ("", rest)
    if k < 0 and rest.startswith(","):
        raise ParseError(f"declaration with attributes but no '::': {stmt!r}")
    if k < 0:
        # 'real x, y' old style -- but not 'real(x) = ...' (never valid) ; reject things that look like assignments
        if find_top(rest, "=") >= 0 and not re.match(r"^[a-z_]\w*\s*(\(.*\))?\s*=", rest) is None and base != "type":
            pass
    attrs, dims_attr, intent = set(), None, None
    for a in split_top(attrs_text.strip().lstrip(",")):
        if not a:
            continue
        am = re.match(r"(\w+)\s*(\((.*)\))?$", a, re.S)
        if not am:
            continue
        an = am.group(1)
        if an == "dimension":
            dims_attr = parse_dims(am.group(3))
        elif an == "intent":
            intent = am.group(3).replace(" ", "")
        else:
            attrs.add(an)
    out = []
    for e in split_top(ents_text):
        if not e:
            continue
        d = Decl()
        d.base, d.kind, d.tname, d.charlen, d.attrs, d.intent = base, kind, tname, charlen, set(attrs), intent
        init = None
        k2 = find_top(e, "=")
        if k2 >= 0 and not e[k2:k2 + 2] == "=>" and not e[k2:k2 + 2] == "==":
            init = e[k2 + 1:].strip(); e = e[:k2].strip()
        elif k2 >= 0 and e[k2:k2 + 2] == "=>":
            e = e[:k2].strip()                      # => null()
        mm = re.match(r"([a-z_]\w*)\s*(\((.*)\))?\s*(\*\s*\d+)?$", e, re.S)
        if not mm:
            raise ParseError(f"cannot parse entity {e!r} in {stmt!r}")
        d.name = mm.group(1)
        d.dims = parse_dims(mm.group(3)) if mm.group(2) else dims_attr
        d.init = parse_expr(init) if init is not None else None
        out.append(d)
    return out


# ------------------------------------------------------------------------------------------------------------------
# program units

class TypeDef:
    def __init__(self, name):
        self.name = name; self.comps = []          # [Decl]
        self.module = None


class Procedure:
    def __init__(self, kind, name, args, result=None, prefix=""):
        self.kind, self.name, self.args, self.result, self.prefix = kind, name, args, result, prefix
        self.decls = {}            # name -> Decl
        self.decl_order = []
        self.uses = []             # [(module, only: None | {local: remote})]
        self.body = []
        self.internal = {}         # name -> Procedure
        self.host = None           # Module or Procedure
        self.module = None
        self.file = None; self.line = 0
        self.saved = {}            # persistent entities (SAVE / initialised locals / parameters)
        self.elemental = "elemental" in prefix


class Module:
    def __init__(self, name):
        self.name = name
        self.uses = []
        self.decls = {}; self.decl_order = []
        self.types = {}; self.procs = {}; self.generics = {}
        self.private_default = False
        self.public, self.private = set(), set()
        self.file = None
        self.names = None          # instantiated entities (finterp)


_RE = {
    "module": re.compile(r"^module\s+([a-z_]\w*)$"),
    "endunit": re.compile(r"^end\s*(module|subroutine|function|program)?\b\s*([a-z_]\w*)?$"),
    "use": re.compile(r"^use\b\s*(?:,\s*intrinsic\s*)?(?:::)?\s*([a-z_]\w*)\s*(?:,\s*(only\s*:)?\s*(.*))?$", re.S),
    "sub": re.compile(r"^((?:(?:recursive|pure|elemental|impure)\s+)*)subroutine\s+([a-z_]\w*)\s*(\((.*)\))?\s*(bind\s*\(.*\))?$", re.S),
    "fun": re.compile(r"^((?:(?:recursive|pure|elemental|impure|real|integer|logical|double\s*precision|real\s*\([^)]*\)|integer\s*\([^)]*\)|type\s*\([^)]*\))\s+)*)function\s+([a-z_]\w*)\s*\((.*?)\)\s*(?:result\s*\(\s*([a-z_]\w*)\s*\))?\s*(bind\s*\(.*\))?$", re.S),
    "typedef": re.compile(r"^type\s*(?:,\s*[^:]*)?(?:::)?\s*([a-z_]\w*)$"),
    "typedef2": re.compile(r"^type\s+([a-z_]\w*)$"),
    "endtype": re.compile(r"^end\s*type\b"),
    "interface": re.compile(r"^(abstract\s+)?interface\b\s*(.*)$"),
    "endinterface": re.compile(r"^end\s*interface\b"),
    "modproc": re.compile(r"^module\s+procedure\s+(.*)$"),
}


class SourceParser:
    """one file -> Modules (and stand-alone procedures)"""

    def __init__(self, path: str, defines=()):
        self.path = path
        with open(path, errors="replace") as fh:
            self.lines = logical_lines(cpp(fh.read(), defines))
        self.i = 0
        self.modules = {}
        self.procs = {}

    def parse(self):
        while self.i < len(self.lines):
            ln, s = self.lines[self.i]
            m = _RE["module"].match(s)
            if m and not s.startswith("module procedure"):
                self.i += 1
                mod = self.parse_module(m.group(1))
                self.modules[mod.name] = mod
                continue
            p = self.try_procedure_header(s)
            if p:
                self.i += 1
                self.parse_procedure_body(p, None)
                self.procs[p.name] = p
                continue
            if s.startswith("program"):
                # skip a main program
                self.i += 1
                while self.i < len(self.lines) and not re.match(r"^end\s*program\b", self.lines[self.i][1]):
                    self.i += 1
                self.i += 1
                continue
            self.i += 1
        return self

    # -- specification part shared by modules and procedures ------------------------------------------------------
    def parse_spec_stmt(self, unit, s, ln) -> bool:
        """returns True if s was a specification statement (consumed)."""
        m = _RE["use"].match(s)
        if m:
            only = None
            rest = (m.group(3) or "").strip()
            if m.group(2) or rest:
                only_flag = bool(m.group(2))
                mapping = {}
                for item in split_top(rest):
                    if not item:
                        continue
                    if "=>" in item:
                        loc, rem = [x.strip() for x in item.split("=>")]
                    else:
                        loc = rem = item
                    mapping[loc] = rem
                only = ("only", mapping) if only_flag else ("rename", mapping)
            unit.uses.append((m.group(1), only))
            return True
        if s.startswith("implicit") or s == "save" or s.startswith(("external", "intrinsic", "namelist", "common", "equivalence", "include", "import")):
            return True
        if s == "private" or s == "public":
            if isinstance(unit, Module):
                unit.private_default = (s == "private")
            return True
        mm = re.match(r"^(public|private|save|optional|target|allocatable|pointer|contiguous|volatile|protected)\b\s*(?:::)?\s*(.+)$", s)
        if mm and "=" not in mm.group(2):
            names = [x.strip() for x in split_top(mm.group(2))]
            if isinstance(unit, Module) and mm.group(1) in ("public", "private"):
                (unit.public if mm.group(1) == "public" else unit.private).update(names)
            elif mm.group(1) in ("optional", "save", "target", "allocatable", "pointer"):
                for nme in names:
                    nme = re.sub(r"\(.*\)$", "", nme).strip()
                    if nme in unit.decls:
                        unit.decls[nme].attrs.add(mm.group(1))
                    else:
                        unit.__dict__.setdefault("late_attrs", []).append((nme, mm.group(1)))
            return True
        mm = re.match(r"^(dimension|parameter|data|intent)\b", s)
        if mm:
            if mm.group(1) == "parameter":
                inner = s[s.index("(") + 1:match_paren(s, s.index("("))]
                for item in split_top(inner):
                    k = find_top(item, "=")
                    nme = item[:k].strip()
                    if nme in unit.decls:
                        unit.decls[nme].attrs.add("parameter"); unit.decls[nme].init = parse_expr(item[k + 1:].strip())
                return True
            if mm.group(1) == "data":
                unit.__dict__.setdefault("data_stmts", []).append(s)
                return True
            return True
        # derived type definition
        m = _RE["typedef"].match(s) if not re.match(r"^type\s*\(", s) else None
        if m and not s.startswith("type is"):
            td = TypeDef(m.group(1))
            self.i += 1
            private_comps = False
            while not _RE["endtype"].match(self.lines[self.i][1]):
                ln2, s2 = self.lines[self.i]
                if s2 in ("private", "sequence", "public") or s2.startswith(("contains", "procedure", "generic", "final")):
                    if s2.startswith("contains"):
                        # type-bound procedures: skip to end type
                        while not _RE["endtype"].match(self.lines[self.i][1]):
                            self.i += 1
                        break
                    self.i += 1
                    continue
                try:
                    ds = parse_declaration(s2)
                except ParseError:
                    ds = None
                if ds:
                    td.comps.extend(ds)
                self.i += 1
            unit.types[td.name] = td if hasattr(unit, "types") else None
            if not hasattr(unit, "types"):
                unit.__dict__.setdefault("local_types", {})[td.name] = td
            return True
        m = _RE["interface"].match(s)
        if m:
            gname = m.group(2).strip()
            self.i += 1
            specifics = []
            depth = 0
            while True:
                ln2, s2 = self.lines[self.i]
                if _RE["endinterface"].match(s2):
                    break
                mp_ = _RE["modproc"].match(s2)
                if mp_:
                    specifics.extend(x.strip() for x in split_top(mp_.group(1)))
                elif re.match(r"^procedure\b", s2):
                    specifics.extend(x.strip() for x in split_top(re.sub(r"^procedure\s*(::)?", "", s2)))
                self.i += 1
            if gname and not gname.startswith(("operator", "assignment")) and hasattr(unit, "generics"):
                unit.generics.setdefault(gname, []).extend(specifics)
            return True
        try:
            ds = parse_declaration(s)
        except ParseError as e:
            if isinstance(unit, Module):
                unit.__dict__.setdefault("skipped", []).append((ln, s, str(e)))
                return True
            raise
        if ds is not None:
            # 'real function f(x)' is a procedure header, not a declaration
            for d in ds:
                unit.decls[d.name] = d
                unit.decl_order.append(d.name)
            return True
        return False

    def parse_module(self, name):
        mod = Module(name); mod.file = self.path
        while self.i < len(self.lines):
            ln, s = self.lines[self.i]
            if s == "contains":
                self.i += 1
                break
            m = _RE["endunit"].match(s)
            if m and (m.group(1) in (None, "module")) and not _RE["endtype"].match(s):
                self.i += 1
                return mod
            if self.try_procedure_header(s):
                break
            if not self.parse_spec_stmt(mod, s, ln):
                mod.__dict__.setdefault("skipped", []).append((ln, s, "not a specification statement"))
            self.i += 1
        # module procedures
        while self.i < len(self.lines):
            ln, s = self.lines[self.i]
            m = _RE["endunit"].match(s)
            if m and m.group(1) in (None, "module"):
                self.i += 1
                break
            p = self.try_procedure_header(s)
            if p:
                self.i += 1
                p.line = ln
                self.parse_procedure_body(p, mod)
                mod.procs[p.name] = p
                continue
            self.i += 1
        return mod

    def try_procedure_header(self, s):
        m = _RE["sub"].match(s)
        if m:
            args = [a.strip() for a in split_top(m.group(4) or "") if a.strip()]
            return Procedure("subroutine", m.group(2), args, prefix=m.group(1) or "")
        if "function" in s and not s.startswith("end"):
            m = _RE["fun"].match(s)
            if m:
                args = [a.strip() for a in split_top(m.group(3) or "") if a.strip()]
                p = Procedure("function", m.group(2), args, result=m.group(4) or m.group(2), prefix=m.group(1) or "")
                # a type prefix declares the result
                pre = re.sub(r"\b(recursive|pure|elemental|impure)\b", "", p.prefix).strip()
                if pre:
                    ds = parse_declaration(pre + " :: " + p.result)
                    if ds:
                        p.decls[p.result] = ds[0]; p.decl_order.append(p.result)
                return p
        return None

    def parse_procedure_body(self, p: Procedure, host):
        p.host = host; p.file = self.path
        p.module = host if isinstance(host, Module) else (host.module if host is not None else None)
        p.types = {}
        # specification part
        while self.i < len(self.lines):
            ln, s = self.lines[self.i]
            if self.is_end_of(p, s) or s == "contains":
                break
            try:
                is_spec = self.parse_spec_stmt(p, s, ln)
            except ParseError:
                is_spec = False
            if not is_spec:
                break
            self.i += 1
        for nme, attr in p.__dict__.get("late_attrs", []):
            if nme in p.decls:
                p.decls[nme].attrs.add(attr)
        # execution part
        p.body = self.parse_block(lambda s: self.is_end_of(p, s) or s == "contains")
        ln, s = self.lines[self.i]
        if s == "contains":
            self.i += 1
            while self.i < len(self.lines):
                ln, s = self.lines[self.i]
                if self.is_end_of(p, s):
                    break
                q = self.try_procedure_header(s)
                if q:
                    self.i += 1
                    q.line = ln
                    self.parse_procedure_body(q, p)
                    p.internal[q.name] = q
                    continue
                self.i += 1
        self.i += 1           # the END statement

    @staticmethod
    def is_end_of(p, s):
        m = _RE["endunit"].match(s)
        return bool(m) and m.group(1) in (None, p.kind)

    # -- executable statements --------------------------------------------------------------------------------------
    def parse_block(self, stop):
        """parse statements until stop(s) is true for the current line (not consumed)."""
        out = []
        while self.i < len(self.lines):
            ln, s = self.lines[self.i]
            if stop(s):
                return out
            self.i += 1
            st = self.parse_stmt(s, ln)
            if st is not None:
                out.append(st)
        return out

    def parse_stmt(self, s, ln):
        # strip a statement label / construct name
        m = re.match(r"^(\d+)\s+(.*)$", s)
        if m:
            s = m.group(2)
        m = re.match(r"^([a-z_]\w*)\s*:\s*(do|if|where|select)\b(.*)$", s)
        if m and not s.startswith(("else", "case")):
            s = m.group(2) + m.group(3)
        try:
            return self._parse_stmt(s, ln)
        except ParseError as e:
            return ("unsupported", ln, s, str(e))

    def _parse_stmt(self, s, ln):
        if s.startswith("if") and re.match(r"^if\s*\(", s):
            i0 = s.index("(")
            j = match_paren(s, i0)
            cond = parse_expr(s[i0 + 1:j])
            rest = s[j + 1:].strip()
            if rest == "then":
                arms, else_body = [], None
                body = self.parse_block(lambda t: bool(re.match(r"^(else\b|elseif\b|end\s*if\b|endif\b)", t)))
                arms.append((cond, body))
                while True:
                    ln2, t = self.lines[self.i]
                    self.i += 1
                    if re.match(r"^(end\s*if|endif)\b", t):
                        break
                    mm = re.match(r"^else\s*if\s*\(", t)
                    if mm:
                        i1 = t.index("(")
                        j1 = match_paren(t, i1)
                        c2 = parse_expr(t[i1 + 1:j1])
                        b2 = self.parse_block(lambda u: bool(re.match(r"^(else\b|elseif\b|end\s*if\b|endif\b)", u)))
                        arms.append((c2, b2))
                    else:   # else
                        else_body = self.parse_block(lambda u: bool(re.match(r"^(end\s*if|endif)\b", u)))
                return ("if", ln, arms, else_body)
            inner = self.parse_stmt(rest, ln)
            return ("if", ln, [(cond, [inner] if inner else [])], None)
        if re.match(r"^where\s*\(", s):
            i0 = s.index("(")
            j = match_paren(s, i0)
            mask = parse_expr(s[i0 + 1:j])
            rest = s[j + 1:].strip()
            if rest:
                inner = self.parse_stmt(rest, ln)
                return ("where", ln, [(mask, [inner])])
            arms = []
            body = self.parse_block(lambda t: bool(re.match(r"^(else\s*where|end\s*where)\b", t)))
            arms.append((mask, body))
            while True:
                ln2, t = self.lines[self.i]
                self.i += 1
                if re.match(r"^end\s*where\b", t):
                    break
                mm = re.match(r"^else\s*where\s*(\((.*)\))?\s*([a-z_]\w*)?$", t, re.S)
                m2 = None
                if mm and mm.group(1):
                    i1 = t.index("(")
                    m2 = parse_expr(t[i1 + 1:match_paren(t, i1)])
                b2 = self.parse_block(lambda u: bool(re.match(r"^(else\s*where|end\s*where)\b", u)))
                arms.append((m2, b2))
            return ("where", ln, arms)
        m = re.match(r"^do\s+while\s*\(", s)
        if m:
            i0 = s.index("(")
            cond = parse_expr(s[i0 + 1:match_paren(s, i0)])
            body = self.parse_block(lambda t: bool(re.match(r"^(end\s*do|enddo)\b", t)))
            self.i += 1
            return ("dowhile", ln, cond, body)
        if s == "do":
            body = self.parse_block(lambda t: bool(re.match(r"^(end\s*do|enddo)\b", t)))
            self.i += 1
            return ("doforever", ln, body)
        m = re.match(r"^do\s+(?:\d+\s+)?([a-z_]\w*)\s*=\s*(.*)$", s, re.S)
        if m:
            parts = split_top(m.group(2))
            lo, hi = parse_expr(parts[0]), parse_expr(parts[1])
            step = parse_expr(parts[2]) if len(parts) > 2 else None
            body = self.parse_block(lambda t: bool(re.match(r"^(end\s*do|enddo)\b", t)))
            self.i += 1
            return ("do", ln, m.group(1), lo, hi, step, body)
        m = re.match(r"^select\s*case\s*\(", s)
        if m:
            i0 = s.index("(")
            sel = parse_expr(s[i0 + 1:match_paren(s, i0)])
            cases = []
            # skip to first case
            while not re.match(r"^(case\b|end\s*select)", self.lines[self.i][1]):
                self.i += 1
            while True:
                ln2, t = self.lines[self.i]
                self.i += 1
                if re.match(r"^end\s*select\b", t):
                    break
                if re.match(r"^case\s+default", t):
                    vals = None
                else:
                    i1 = t.index("(")
                    vals = []
                    for item in split_top(t[i1 + 1:match_paren(t, i1)]):
                        k = find_top(item, ":")
                        if k >= 0:
                            lo_, hi_ = item[:k].strip(), item[k + 1:].strip()
                            vals.append(("range", parse_expr(lo_) if lo_ else None, parse_expr(hi_) if hi_ else None))
                        else:
                            vals.append(parse_expr(item))
                body = self.parse_block(lambda u: bool(re.match(r"^(case\b|end\s*select)", u)))
                cases.append((vals, body))
            return ("select", ln, sel, cases)
        m = re.match(r"^call\s+([a-z_]\w*(?:\s*%\s*[a-z_]\w*)*)\s*(\(.*\))?$", s, re.S)
        if m:
            args = ExprParser(tokenize(m.group(2)), s).p_args() if m.group(2) else []
            return ("call", ln, m.group(1).replace(" ", ""), args)
        if s == "return":
            return ("return", ln)
        if s == "exit" or re.match(r"^exit\s+\w+$", s):
            return ("exit", ln)
        if s == "cycle" or re.match(r"^cycle\s+\w+$", s):
            return ("cycle", ln)
        if s == "continue":
            return None
        if re.match(r"^(error\s+)?stop\b", s):
            return ("stop", ln, s)
        if re.match(r"^(print\b|write\s*\(|format\s*\(|flush\b)", s):
            return ("io_ignored", ln, s)
        if re.match(r"^(open|close|read|rewind|backspace|inquire)\s*\(", s) or s.startswith("read "):
            return ("io_fail", ln, s)
        m = re.match(r"^(allocate|deallocate|nullify)\s*\(", s)
        if m:
            i0 = s.index("(")
            inner = s[i0 + 1:match_paren(s, i0)]
            items = []
            for item in split_top(inner):
                if re.match(r"^(stat|errmsg|source|mold)\s*=", item):
                    continue
                items.append(parse_expr(item))
            return (m.group(1), ln, items)
        # assignment / pointer assignment
        k = find_top(s, "=")
        while k >= 0 and (s[k:k + 2] == "==" or s[k - 1] in "<>/=" or s[k:k + 2] == "=>"):
            if s[k:k + 2] == "=>":
                lhs = parse_expr(s[:k].strip()); rhs = parse_expr(s[k + 2:].strip())
                return ("ptrassign", ln, lhs, rhs)
            k = find_top(s, "=", k + 2 if s[k:k + 2] == "==" else k + 1)
        if k > 0:
            lhs = parse_expr(s[:k].strip())
            rhs = parse_expr(s[k + 1:].strip())
            if lhs[0] != "des":
                raise ParseError(f"bad assignment target in {s!r}")
            return ("assign", ln, lhs, rhs)
        raise ParseError(f"unrecognised statement {s!r}")
