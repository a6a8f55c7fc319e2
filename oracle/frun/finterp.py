"""finterp.py -- tree-walking interpreter for the Fortran subset parsed by fparse.py (TEST INFRASTRUCTURE ONLY).

Runs the reference's unmodified source files (see fparse.py for why).  Semantics implemented the way a conforming
compiler without value-changing optimisation would (the reference builds with -fp-model precise / plain -O3,
CMakeLists.txt:44-57):

  * kinds: default REAL = binary32, REAL(r_2) / DOUBLE PRECISION = binary64, INTEGER = int32, LOGICAL;
    un-suffixed real literals are binary32, `d` exponents and `_r_2` suffixes binary64;
  * every operator promotes its two operands only (int -> real32 -> real64), so sub-expressions of REAL operands are
    evaluated in binary32 even inside a REAL(r_2) statement; assignment converts once;
  * x**n with integer n is repeated multiplication (binary powering, as gfortran's powi expansion and ifort do);
    x**y with real y, EXP, LOG, LOG10, SIN, COS, TAN, ATAN on binary32 are evaluated in binary64 and rounded once
    (= a correctly rounded libm; SQRT, ABS, +, -, *, / are IEEE exact in both kinds);
  * SUM / dot products run left to right; integer division truncates; NINT rounds half away from zero;
    SIGN takes the sign bit; MAX / MIN fold left to right;
  * WHERE evaluates its mask once on entry; nested WHERE / ELSEWHERE masks combine; masked assignment only touches
    selected elements;
  * arguments are passed by reference (array sections and elements as views), explicit-shape dummies re-map bounds,
    locals without SAVE are fresh on every call (zero-filled instead of undefined), initialised locals are SAVEd.
"""
from __future__ import annotations

import glob
import os
import re

import numpy as np

from . import fparse
from .fparse import Decl, Module, Procedure, TypeDef

F32, F64, I32, B = np.float32, np.float64, np.int32, np.bool_
_RANK = {np.dtype(B): 0, np.dtype(I32): 1, np.dtype(np.int64): 1, np.dtype(F32): 2, np.dtype(F64): 3}
np.seterr(all="ignore")


class FortranStop(Exception):
    pass


class FortranError(Exception):
    pass


class StubError(FortranError):
    """control reached something that lives in a stubbed module"""


class Arr:
    """a variable: data (0-d ndarray for scalars) + lower bounds"""
    __slots__ = ("a", "lb", "decl")

    def __init__(self, a, lb=(), decl=None):
        self.a, self.lb, self.decl = a, lb, decl


class Str:
    __slots__ = ("s", "n")

    def __init__(self, s="", n=None):
        self.s, self.n = s, n


class Struct:
    __slots__ = ("tdef", "f")

    def __init__(self, tdef):
        self.tdef, self.f = tdef, {}


class StructArr:
    """array of derived type (rare: only module-level tables)"""
    __slots__ = ("items", "lb")

    def __init__(self, items, lb):
        self.items, self.lb = items, lb


ABSENT = object()


class _Ctl(Exception):
    pass


class Frame:
    __slots__ = ("proc", "vars", "host", "interp")

    def __init__(self, proc, host, interp):
        self.proc, self.vars, self.host, self.interp = proc, {}, host, interp


def _dtype_of(v):
    if isinstance(v, np.ndarray) or isinstance(v, np.generic):
        return v.dtype
    if isinstance(v, bool):
        return np.dtype(B)
    if isinstance(v, int):
        return np.dtype(I32)
    if isinstance(v, float):
        return np.dtype(F64)
    raise FortranError(f"not a numeric value: {v!r}")


def _promote(a, b):
    ta, tb = _dtype_of(a), _dtype_of(b)
    if ta == tb:
        return a, b, ta
    ra, rb = _RANK[ta], _RANK[tb]
    if ra >= rb:
        t = ta
        b = np.asarray(b).astype(t) if isinstance(b, np.ndarray) and b.ndim else t.type(b)
    else:
        t = tb
        a = np.asarray(a).astype(t) if isinstance(a, np.ndarray) and a.ndim else t.type(a)
    return a, b, t


# binary64 transcendentals through the C library (glibc: correctly rounded in all but a vanishing fraction of cases), not
# NumPy's SIMD kernels, which are allowed 1-4 ulp
import math as _math


def _libm(fn):
    uf = np.frompyfunc(lambda v: fn(float(v)), 1, 1)

    def f(x):
        if isinstance(x, np.ndarray) and x.ndim:
            with np.errstate(all="ignore"):
                return _guard(uf, x).astype(F64)
        return F64(_guard1(fn, float(x)))
    return f


def _guard1(fn, v):
    try:
        return fn(v)
    except (ValueError, OverflowError):
        if fn is _math.log or fn is _math.log10:
            return float("-inf") if v == 0.0 else float("nan")
        if fn is _math.exp:
            return float("inf")
        return float("nan")


def _guard(uf, x):
    try:
        return uf(x)
    except (ValueError, OverflowError):
        fn = uf.__dict__.get("fn") if hasattr(uf, "__dict__") else None
        out = np.empty(x.shape, object)
        for ix in np.ndindex(*x.shape):
            try:
                out[ix] = uf(x[ix])
            except (ValueError, OverflowError):
                v = float(x[ix])
                out[ix] = float("nan") if not (v == 0.0) else float("-inf")
        return out


def _pow64(a, b):
    def one(x, y):
        try:
            return _math.pow(float(x), float(y))
        except (ValueError, OverflowError):
            return float(np.power(F64(x), F64(y)))
    if (isinstance(a, np.ndarray) and a.ndim) or (isinstance(b, np.ndarray) and b.ndim):
        return np.frompyfunc(one, 2, 1)(a, b).astype(F64)
    return F64(one(a, b))


_exp64, _log64, _log1064 = _libm(_math.exp), _libm(_math.log), _libm(_math.log10)
_sin64, _cos64, _tan64, _atan64 = _libm(_math.sin), _libm(_math.cos), _libm(_math.tan), _libm(_math.atan)
_asin64, _acos64, _tanh64, _sinh64, _cosh64 = _libm(_math.asin), _libm(_math.acos), _libm(_math.tanh), _libm(_math.sinh), _libm(_math.cosh)


def _cr32(fn, x):
    """binary32 intrinsic evaluated in binary64, rounded once"""
    if isinstance(x, np.ndarray) and x.ndim:
        return fn(x.astype(F64)).astype(F32)
    return F32(fn(F64(x)))


def _elem(fn64):
    def f(x):
        t = _dtype_of(x)
        if t == np.dtype(F32):
            return _cr32(fn64, x)
        if t == np.dtype(F64):
            return fn64(x)
        raise FortranError(f"real intrinsic on {t}")
    return f


def _ipow(x, n: int):
    """x**n, integer n: binary powering (multiplications only; reciprocal for n < 0)"""
    if n == 0:
        return (x * 0 + 1) if not isinstance(x, np.ndarray) else np.ones_like(x)
    m = abs(n)
    result, base = None, x
    while m:
        if m & 1:
            result = base if result is None else result * base
        m >>= 1
        if m:
            base = base * base
    if n < 0:
        one = _dtype_of(x).type(1)
        return one / result
    return result


def _seq_sum(a, axis=None, mask=None):
    """left-to-right sum (Fortran SUM as a plain loop)"""
    if mask is not None:
        a = np.where(mask, a, a.dtype.type(0))
    if axis is None:
        flat = np.ravel(a, order="F")
        acc = a.dtype.type(0)
        for v in flat:
            acc = acc + v
        return acc
    n = a.shape[axis]
    acc = np.zeros(np.delete(a.shape, axis), a.dtype)
    for k in range(n):
        acc = acc + np.take(a, k, axis=axis)
    return acc


class Interp:
    def __init__(self, src_root: str, exclude=("coupled",), stub_modules=()):
        self.src_root = src_root
        self.index = {}                 # module name -> file
        for f in sorted(glob.glob(os.path.join(src_root, "**", "*.[Ff]90"), recursive=True)):
            rel = os.path.relpath(f, src_root)
            if any(rel.startswith(e + os.sep) or (os.sep + e + os.sep) in rel for e in exclude):
                continue
            with open(f, errors="replace") as fh:
                for m in re.finditer(r"^\s*module\s+([a-z_]\w*)\s*$", fh.read().lower(), re.M):
                    if m.group(1) != "procedure":
                        self.index.setdefault(m.group(1), f)
        self.modules = {}
        self.stub_modules = set(stub_modules)
        self.files_parsed = {}
        self.trace = None
        self.nstmt = 0
        self.tolerate_stubs_in = set()      # procedure names in which a statement that needs a stubbed module is skipped
        self.skipped = []                   # ... and recorded here

    # ---- modules ---------------------------------------------------------------------------------------------------
    def module(self, name: str) -> Module:
        m = self.modules.get(name)
        if m is not None:
            return m
        if name in self.stub_modules or name in ("iso_c_binding", "iso_fortran_env", "mpi", "netcdf", "ieee_arithmetic"):
            m = Module(name); m.names = {}
            if name == "iso_fortran_env":       # the kind constants (byte sizes, as gfortran / ifort / nvfortran define them)
                m.names = {k: Arr(np.array(v, np.int32)) for k, v in (("int8", 1), ("int16", 2), ("int32", 4), ("int64", 8),
                                                                      ("real32", 4), ("real64", 8))}
            self.modules[name] = m
            return m
        f = self.index.get(name)
        if f is None:
            raise FortranError(f"module {name} not found under {self.src_root}")
        if f not in self.files_parsed:
            self.files_parsed[f] = fparse.SourceParser(f).parse()
        m = self.files_parsed[f].modules[name]
        self.modules[name] = m
        self._instantiate_module(m)
        return m

    def _instantiate_module(self, m: Module):
        m.names = {}
        for t in m.types.values():
            t.module = m
        for p in m.procs.values():
            m.names[p.name] = p
        for g, specs in m.generics.items():
            m.names[g] = ("generic", m, specs)
        for t in m.types.values():
            m.names.setdefault("$type:" + t.name, t)
        fr = Frame(None, m, self)
        for name in m.decl_order:
            d = m.decls[name]
            if name in m.procs:
                continue
            try:
                m.names[name] = self._make_entity(d, fr, module_level=True)
            except Exception as e:          # entities of unrelated subsystems (I/O tables ...) may need things we do not model
                m.names[name] = ("broken", f"{m.name}::{name}: {e}")

    def lookup_in_module(self, m: Module, name: str, seen=None):
        if m.names is None:
            self._instantiate_module(m)
        if name in m.names:
            return m.names[name]
        seen = seen or set()
        if m.name in seen:
            return None
        seen.add(m.name)
        for uname, only in m.uses:
            r = self._lookup_use(uname, only, name, seen)
            if r is not None:
                m.names[name] = r
                return r
        return None

    def _lookup_use(self, uname, only, name, seen=None):
        remote = name
        if only is not None:
            kind, mapping = only
            pre, bare = ("$type:", name[6:]) if name.startswith("$type:") else ("", name)
            if bare in mapping:
                remote = pre + mapping[bare]
            elif kind == "only":
                return None
            elif bare in mapping.values():
                return None                    # renamed away
        if uname in self.stub_modules:
            # only names a USE statement imports explicitly resolve to a stub; a blanket USE of a stubbed module exports nothing
            return ("stub", uname, remote) if (only is not None and only[0] == "only") else None
        um = self.module(uname)
        return self.lookup_in_module(um, remote, seen)

    def lookup(self, fr: Frame, name: str):
        f = fr
        while f is not None:
            if isinstance(f, Frame):
                v = f.vars.get(name)
                if v is not None:
                    return v
                p = f.proc
                if p is not None:
                    if name in p.internal:
                        return ("internal", p.internal[name], f)
                    lt = p.__dict__.get("types") or {}
                    if name.startswith("$type:") and name[6:] in lt:
                        return lt[name[6:]]
                    for uname, only in p.uses:
                        r = self._lookup_use(uname, only, name)
                        if r is not None:
                            return r
                f = f.host
            else:                                # a Module
                return self.lookup_in_module(f, name)
        return None

    # ---- entities --------------------------------------------------------------------------------------------------
    def _kind_dtype(self, d: Decl, fr):
        if d.base == "real":
            if d.kind is None:
                return F32
            k = int(self.eval(d.kind, fr))
            return F64 if k == 8 else F32
        if d.base == "integer":
            return I32
        if d.base == "logical":
            return B
        raise FortranError(f"no dtype for {d}")

    def _bounds(self, d: Decl, fr):
        lbs, shape = [], []
        for lo, hi in d.dims:
            if hi in (":", "*"):
                return None
            l = int(self.eval(lo, fr)) if lo is not None else 1
            h = int(self.eval(hi, fr))
            lbs.append(l); shape.append(max(0, h - l + 1))
        return tuple(lbs), tuple(shape)

    def find_type(self, fr, tname):
        t = self.lookup(fr, "$type:" + tname)
        if not isinstance(t, TypeDef):
            raise FortranError(f"derived type {tname} not found")
        return t

    def new_struct(self, tdef: TypeDef):
        s = Struct(tdef)
        fr = Frame(None, tdef.module, self)
        for c in tdef.comps:
            s.f[c.name] = self._make_entity(c, fr, component=True)
        return s

    def _make_entity(self, d: Decl, fr, module_level=False, component=False):
        if d.base == "character":
            init = self.eval(d.init, fr) if d.init is not None else ""
            n = None
            if d.charlen and re.match(r"^(len\s*=\s*)?\d+$", d.charlen):
                n = int(re.sub(r"\D", "", d.charlen))
            if d.dims is not None:
                b = self._bounds(d, fr)
                if b is None:
                    return Arr(None, (), d)
                return StructArr([Str("", n) for _ in range(int(np.prod(b[1])))], b[0])
            return Str(init if isinstance(init, str) else "", n)
        if d.base == "type":
            t_ = self.lookup(fr, "$type:" + d.tname)
            if isinstance(t_, tuple) and t_[0] == "stub":        # a type of a stubbed module: never touched
                return Arr(None, (), d)
            tdef = self.find_type(fr, d.tname)
            if d.dims is not None:
                b = self._bounds(d, fr)
                if b is None:
                    return Arr(None, (), d)
                return StructArr([self.new_struct(tdef) for _ in range(int(np.prod(b[1])))], b[0])
            if "pointer" in d.attrs or "allocatable" in d.attrs:
                return Arr(None, (), d)
            return self.new_struct(tdef)
        dt = self._kind_dtype(d, fr)
        if d.dims is None:
            a = np.zeros((), dt)
            if d.init is not None:
                a[...] = self._convert(self.eval(d.init, fr), dt)
            return Arr(a, (), d)
        b = self._bounds(d, fr)
        if b is None or "allocatable" in d.attrs or ("pointer" in d.attrs and b is None):
            if b is None and d.init is not None and "parameter" in d.attrs:      # implied-shape parameter (*)
                v = np.asarray(self.eval(d.init, fr))
                return Arr(np.asfortranarray(v.astype(dt)), (1,) * v.ndim, d)
            return Arr(None, (), d)
        lbs, shape = b
        a = np.zeros(shape, dt, order="F")
        if d.init is not None:
            v = self.eval(d.init, fr)
            a[...] = self._convert(v, dt) if not (isinstance(v, np.ndarray) and v.ndim) else np.reshape(self._convert(v, dt), shape, order="F")
        return Arr(a, lbs, d)

    @staticmethod
    def _convert(v, dt):
        """assignment conversion to dtype dt (real -> integer truncates toward zero)"""
        if isinstance(v, (Str, str)):
            raise FortranError("character value in numeric context")
        t = _dtype_of(v)
        if t == np.dtype(dt):
            return v
        if np.dtype(dt) == np.dtype(I32) and _RANK[t] >= 2:
            return np.trunc(v).astype(I32) if isinstance(v, np.ndarray) and v.ndim else I32(np.trunc(v))
        if isinstance(v, np.ndarray) and v.ndim:
            return v.astype(dt)
        return np.dtype(dt).type(v)

    # ---- expressions -----------------------------------------------------------------------------------------------
    def eval(self, e, fr):
        k = e[0]
        if k == "num":
            return self._literal(e, fr)
        if k == "des":
            return self.eval_des(e, fr)
        if k == "bin":
            return self.binop(e[1], self.eval(e[2], fr), self.eval(e[3], fr))
        if k == "par":
            v = self.eval(e[1], fr)
            return v.copy() if isinstance(v, np.ndarray) else v
        if k == "un":
            v = self.eval(e[2], fr)
            if e[1] == "-":
                return -v
            if e[1] == "+":
                return v
            return np.logical_not(v)
        if k == "log":
            return B(e[1])
        if k == "str":
            return e[1]
        if k == "arr":
            vals = []
            for it in e[1]:
                self._ac_items(it, fr, vals)
            if vals and isinstance(vals[0], str):
                return vals
            t = vals[0]
            for v in vals[1:]:
                _, _, tt = _promote(t, v)
                t = tt.type(0)
            return np.array(vals, dtype=_dtype_of(t))
        raise FortranError(f"cannot evaluate {e!r}")

    def _ac_items(self, it, fr, out):
        if it[0] == "ido":
            _, items, var, lo, hi, step = it
            lo, hi = int(self.eval(lo, fr)), int(self.eval(hi, fr))
            st = int(self.eval(step, fr)) if step is not None else 1
            saved = fr.vars.get(var)
            cell = Arr(np.zeros((), I32))
            fr.vars[var] = cell
            i = lo
            while (st > 0 and i <= hi) or (st < 0 and i >= hi):
                cell.a[...] = i
                for sub in items:
                    self._ac_items(sub, fr, out)
                i += st
            if saved is not None:
                fr.vars[var] = saved
            else:
                del fr.vars[var]
            return
        v = self.eval(it, fr)
        if isinstance(v, np.ndarray) and v.ndim:
            out.extend(np.ravel(v, order="F"))
        elif isinstance(v, np.ndarray):
            out.append(v[()])
        else:
            out.append(v)

    _lit_cache = {}

    def _literal(self, e, fr):
        text, kind = e[1], e[2]
        key = (text, kind)
        c = self._lit_cache.get(key)
        if c is not None:
            return c
        suffix = None
        if "_" in text:
            text, suffix = text.split("_", 1)
        if kind == "i":
            v = I32(int(text))
        else:
            if "d" in text:
                v = F64(float(text.replace("d", "e")))
            elif suffix is not None:
                kv = self.lookup(fr, suffix) if not suffix.isdigit() else None
                kk = int(suffix) if suffix.isdigit() else int(kv.a)
                v = F64(float(text)) if kk == 8 else F32(float(text))
                if not suffix.isdigit():
                    return v              # kind parameter looked up in scope: do not cache across scopes
            else:
                v = F32(float(text))
        self._lit_cache[key] = v
        return v

    def binop(self, op, a, b):
        if isinstance(a, str) or isinstance(b, str) or isinstance(a, Str) or isinstance(b, Str):
            sa = a.s if isinstance(a, Str) else a
            sb = b.s if isinstance(b, Str) else b
            if op == "//":
                return sa + sb
            if op == "==":
                return B(sa.rstrip() == sb.rstrip())
            if op == "/=":
                return B(sa.rstrip() != sb.rstrip())
            raise FortranError(f"character operator {op}")
        if op in (".and.", ".or.", ".eqv.", ".neqv."):
            if op == ".and.":
                return np.logical_and(a, b)
            if op == ".or.":
                return np.logical_or(a, b)
            if op == ".eqv.":
                return np.equal(a, b)
            return np.not_equal(a, b)
        if op == "**":
            tb = _dtype_of(b)
            if _RANK[tb] == 1:                          # integer exponent
                if isinstance(b, np.ndarray) and b.ndim:
                    raise FortranError("array-valued integer exponent")
                n = int(b)
                if _RANK[_dtype_of(a)] == 1:
                    return I32(int(a) ** n) if n >= 0 and not (isinstance(a, np.ndarray) and a.ndim) else _ipow(a, n)
                return _ipow(a, n)
            a, b, t = _promote(a, b)
            if t == np.dtype(F32):
                if isinstance(a, np.ndarray) and a.ndim or isinstance(b, np.ndarray) and b.ndim:
                    return _pow64(np.asarray(a, F64), np.asarray(b, F64)).astype(F32)
                return F32(_pow64(F64(a), F64(b)))
            return _pow64(a, b)
        a, b, t = _promote(a, b)
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            if _RANK[t] == 1:
                q = np.abs(a) // np.abs(b)
                return (q * (np.sign(a) * np.sign(b))).astype(I32) if isinstance(q, np.ndarray) and q.ndim else I32(q * np.sign(a) * np.sign(b))
            return a / b
        if op == "==":
            return a == b
        if op == "/=":
            return a != b
        if op == "<":
            return a < b
        if op == "<=":
            return a <= b
        if op == ">":
            return a > b
        if op == ">=":
            return a >= b
        raise FortranError(f"operator {op}")

    # ---- designators -----------------------------------------------------------------------------------------------
    def _index(self, arr: Arr, args, fr, want_view=False):
        """subscript / section of an array variable -> ndarray view (or scalar value when not want_view)"""
        a, lb = arr.a, arr.lb
        if a is None:
            raise FortranError("reference to an unallocated array")
        if len(args) != a.ndim:
            raise FortranError(f"rank mismatch: {len(args)} subscripts for rank {a.ndim}")
        idx, squeeze, all_scalar = [], [], True
        for d, s in enumerate(args):
            l = lb[d] if lb else 1
            if s[0] == "sec":
                all_scalar = False
                lo = int(self.eval(s[1], fr)) - l if s[1] is not None else 0
                hi = int(self.eval(s[2], fr)) - l + 1 if s[2] is not None else a.shape[d]
                st = int(self.eval(s[3], fr)) if s[3] is not None else 1
                if st < 0:
                    hi2 = hi - 2
                    idx.append(slice(lo, hi2 if hi2 >= 0 else None, st))
                else:
                    idx.append(slice(lo, hi, st))
            else:
                v = self.eval(s, fr)
                if isinstance(v, np.ndarray) and v.ndim:
                    all_scalar = False
                    if want_view:
                        raise FortranError("vector subscript in a variable-definition context")
                    idx.append(v.astype(np.int64) - l)           # vector subscript (copy)
                else:
                    i = int(v) - l
                    if i < 0 or i >= a.shape[d]:
                        raise FortranError(f"subscript {int(v)} out of bounds [{l}, {l + a.shape[d] - 1}] in dimension {d + 1}")
                    if want_view:
                        idx.append(slice(i, i + 1)); squeeze.append(d)
                    else:
                        idx.append(i)
        r = a[tuple(idx)]
        if want_view and squeeze:
            r = np.squeeze(r, axis=tuple(squeeze))
        return r

    def resolve(self, e, fr, want_view):
        """designator -> ('val', value/view) | ('struct', Struct) | ('str', Str) | ('call', value)"""
        parts = e[1]
        name, args = parts[0]
        ent = self.lookup(fr, name)
        if ent is None:
            fn = INTRINSICS.get(name)
            if fn is not None and args is not None:
                return ("call", fn(self, fr, args))
            raise FortranError(f"unknown name {name!r} in {fr.proc.name if fr.proc else fr.host}")
        cur = ent
        i = 0
        while True:
            nm, ar = parts[i]
            if isinstance(cur, Arr):
                if cur.a is None and cur.decl is not None and cur.decl.base == "type" and ar is None and i + 1 < len(parts):
                    raise FortranError(f"{nm}: unallocated derived-type pointer")
                if ar is not None:
                    val = self._index(cur, ar, fr, want_view)
                else:
                    val = cur.a
                    if val is None:
                        if want_view == "alloc":
                            return ("arr", cur)
                        raise FortranError(f"{nm}: not allocated")
                if i + 1 < len(parts):
                    raise FortranError(f"component of a numeric value: {parts}")
                return ("val", val)
            if isinstance(cur, Struct):
                if i + 1 == len(parts):
                    return ("struct", cur)
                i += 1
                nm2, ar2 = parts[i]
                nxt = cur.f.get(nm2)
                if nxt is None:
                    raise FortranError(f"type {cur.tdef.name} has no component {nm2}")
                if want_view == "alloc" and i + 1 == len(parts) and isinstance(nxt, Arr):
                    return ("arr", nxt, ar2)
                cur = nxt
                continue
            if isinstance(cur, StructArr):
                if ar is None:
                    raise FortranError("whole array of derived type")
                flat, mul = 0, 1
                # column-major position
                dims = cur.__dict__ if False else None
                idxs = [int(self.eval(a_, fr)) for a_ in ar]
                if len(idxs) != 1:
                    raise FortranError("multi-dimensional arrays of derived type are not supported")
                cur = cur.items[idxs[0] - cur.lb[0]]
                parts = parts[:i] + [(nm, None)] + parts[i + 1:]
                continue
            if isinstance(cur, Str):
                if i + 1 < len(parts) and parts[i + 1][0] == "$substr":
                    sec = parts[i + 1][1][0]
                    lo = int(self.eval(sec[1], fr)) if sec[1] is not None else 1
                    hi = int(self.eval(sec[2], fr)) if sec[2] is not None else len(cur.s)
                    return ("val", cur.s[lo - 1:hi])
                return ("str", cur)
            if isinstance(cur, Procedure) or (isinstance(cur, tuple) and cur[0] in ("generic", "internal", "stub", "broken")):
                if isinstance(cur, tuple) and cur[0] == "broken":
                    raise FortranError(cur[1])
                if ar is None and not isinstance(cur, Procedure):
                    raise FortranError(f"procedure {nm} used as a value")
                return ("call", self.call_function(cur, ar or [], fr))
            raise FortranError(f"cannot resolve {parts} (entity {type(cur).__name__})")

    def eval_des(self, e, fr):
        r = self.resolve(e, fr, False)
        k = r[0]
        if k in ("val", "call"):
            return r[1]
        if k == "str":
            return r[1].s
        return r[1]           # struct

    # ---- calls -----------------------------------------------------------------------------------------------------
    def _pick_specific(self, gen, args, fr):
        _, mod, specs = gen
        actuals = None
        for sname in specs:
            p = self.lookup_in_module(mod, sname)
            if not isinstance(p, Procedure):
                continue
            if actuals is None:
                actuals = []
                for a in args:
                    ex = a[2] if a[0] == "kw" else a
                    try:
                        actuals.append(self.eval(ex, fr) if ex[0] != "des" else self.resolve(ex, fr, True)[1])
                    except FortranError:
                        actuals.append(None)
            ok = len(args) <= len(p.args)
            for dn, av in zip(p.args, actuals):
                d = p.decls.get(dn)
                if d is None:
                    continue
                if d.base == "type":
                    if not (isinstance(av, Struct) and av.tdef.name == d.tname):
                        ok = False
                elif isinstance(av, Struct):
                    ok = False
                elif isinstance(av, (np.ndarray, np.generic)):
                    want = {"real": (2, 3), "integer": (1,), "logical": (0,)}.get(d.base, ())
                    if _RANK[av.dtype] not in want:
                        ok = False
                    if d.base == "real" and ok:
                        dt = self._kind_dtype(d, Frame(p, p.host, self))
                        if np.dtype(dt) != av.dtype:
                            ok = False
                    nd = av.ndim if isinstance(av, np.ndarray) else 0
                    if (len(d.dims) if d.dims else 0) != nd:
                        ok = False
            if ok:
                return p
        raise FortranError(f"no specific procedure of generic matches the call ({specs})")

    def call_function(self, ent, args, fr):
        if isinstance(ent, tuple):
            if ent[0] == "generic":
                ent = self._pick_specific(ent, args, fr)
            elif ent[0] == "internal":
                return self.invoke(ent[1], args, fr, host_frame=ent[2])
            elif ent[0] == "stub":
                raise StubError(f"call into stubbed module {ent[1]}::{ent[2]}")
        return self.invoke(ent, args, fr)

    def invoke(self, p: Procedure, args, caller: Frame, host_frame=None):
        """call procedure p with actual-argument syntax trees `args` evaluated in `caller`"""
        # ---- associate actuals
        bound = {}
        pos = 0
        for a in args:
            if a[0] == "kw":
                dn, ex = a[1], a[2]
            else:
                if pos >= len(p.args):
                    raise FortranError(f"too many arguments in call to {p.name}")
                dn, ex = p.args[pos], a
                pos += 1
            bound[dn] = ex
        actual = {}
        for dn in p.args:
            ex = bound.get(dn)
            if ex is None:
                actual[dn] = ABSENT
                continue
            if ex[0] == "des":
                r = self.resolve(ex, caller, True)
                if r[0] == "val":
                    v = r[1]
                    actual[dn] = v if isinstance(v, np.ndarray) else np.array(v)
                elif r[0] == "call":
                    v = r[1]
                    actual[dn] = v if isinstance(v, (np.ndarray, Struct, str, list)) else np.array(v)
                elif r[0] == "str":
                    actual[dn] = r[1]
                else:
                    actual[dn] = r[1]
            else:
                v = self.eval(ex, caller)
                actual[dn] = v if isinstance(v, (np.ndarray, str, list)) else np.array(v)
        # elemental procedure referenced with array actuals
        if p.elemental and any(isinstance(v, np.ndarray) and v.ndim for v in actual.values()):
            return self._invoke_elemental(p, actual, caller, host_frame)
        return self._run(p, actual, host_frame)

    def _invoke_elemental(self, p, actual, caller, host_frame):
        arrs = [v for v in actual.values() if isinstance(v, np.ndarray) and v.ndim]
        shape = np.broadcast(*arrs).shape
        out = None
        for ix in np.ndindex(*shape):
            one = {}
            for dn, v in actual.items():
                if isinstance(v, np.ndarray) and v.ndim:
                    vv = np.broadcast_to(v, shape)
                    one[dn] = np.array(vv[ix])
                else:
                    one[dn] = v
            r = self._run(p, one, host_frame)
            if p.kind == "function":
                if out is None:
                    out = np.zeros(shape, _dtype_of(r), order="F")
                out[ix] = r
            else:
                for dn, v in actual.items():       # write back intent(out) scalars
                    if isinstance(v, np.ndarray) and v.ndim and v.flags.writeable and isinstance(one[dn], np.ndarray):
                        v[ix] = one[dn]
        return out

    def _run(self, p: Procedure, actual: dict, host_frame=None):
        fr = Frame(p, host_frame if host_frame is not None else p.host, self)
        V = fr.vars
        # dummies first (scalars, then arrays whose bounds may use them)
        pending = []
        for dn in p.args:
            av = actual[dn]
            d = p.decls.get(dn)
            if av is ABSENT:
                V[dn] = ABSENT
                continue
            if d is None:
                raise FortranError(f"{p.name}: dummy {dn} has no declaration (implicit typing is not supported)")
            if d.base == "type":
                if not isinstance(av, (Struct, StructArr)):
                    raise FortranError(f"{p.name}: dummy {dn} wants TYPE({d.tname}), got {type(av).__name__}")
                V[dn] = av
            elif d.base == "character":
                V[dn] = av if isinstance(av, Str) else Str(av if isinstance(av, str) else str(av))
            elif d.dims is None:
                dt = self._kind_dtype(d, fr)
                if not isinstance(av, np.ndarray) or av.ndim != 0:
                    raise FortranError(f"{p.name}: scalar dummy {dn} associated with {type(av).__name__} rank {getattr(av, 'ndim', '?')}")
                if av.dtype != np.dtype(dt):
                    raise FortranError(f"{p.name}: dummy {dn} is {np.dtype(dt)}, actual is {av.dtype}")
                V[dn] = Arr(av, (), d)
            else:
                pending.append((dn, d, av))
        for dn, d, av in pending:
            dt = self._kind_dtype(d, fr)
            if not isinstance(av, np.ndarray):
                raise FortranError(f"{p.name}: array dummy {dn} associated with {type(av).__name__}")
            if av.dtype != np.dtype(dt):
                raise FortranError(f"{p.name}: dummy {dn} is {np.dtype(dt)}, actual is {av.dtype}")
            b = self._bounds(d, fr)
            if b is None:                             # assumed shape / size
                lbs = tuple(int(self.eval(lo, fr)) if lo is not None else 1 for lo, _ in d.dims)
                if len(lbs) != av.ndim:
                    if d.dims[-1][1] == "*":
                        av = np.ravel(av, order="F"); lbs = lbs[:1]
                    else:
                        raise FortranError(f"{p.name}: dummy {dn} rank {len(lbs)} vs actual rank {av.ndim}")
                V[dn] = Arr(av, lbs, d)
            else:
                lbs, shape = b
                if tuple(av.shape) != tuple(shape) and av.ndim == len(shape) and av.shape[:-1] == tuple(shape[:-1]) \
                        and av.shape[-1] < shape[-1]:
                    # actual shorter than the dummy in its last dimension (e.g. veg%taul(mp,2) passed to VegTaul(mp,nrb),
                    # cbl_init_radiation.F90): storage association with the part that exists; a reference beyond it
                    # is caught by the subscript check
                    V[dn] = Arr(av, lbs, d)
                    continue
                if tuple(av.shape) != tuple(shape):
                    if av.size < int(np.prod(shape)):
                        raise FortranError(f"{p.name}: dummy {dn}{shape} larger than actual {av.shape}")
                    flat = np.ravel(av, order="F")
                    if not np.shares_memory(flat, av) and av.size:
                        raise FortranError(f"{p.name}: dummy {dn}{shape} needs sequence association with a non-contiguous actual {av.shape}")
                    av = np.reshape(flat[:int(np.prod(shape))], shape, order="F")
                V[dn] = Arr(av, lbs, d)
        # locals
        for name in p.decl_order:
            if name in V or name in p.args:
                continue
            d = p.decls[name]
            if name in p.saved:
                V[name] = p.saved[name]
                continue
            if p.kind == "function" and name == p.name and p.result != p.name:
                continue
            ent = self._make_entity(d, fr)
            V[name] = ent
            if d.init is not None or "save" in d.attrs or "parameter" in d.attrs:
                p.saved[name] = ent
        if p.kind == "function" and p.result not in V:
            raise FortranError(f"function {p.name}: result {p.result} not declared")
        # body
        try:
            self.exec_block(p.body, fr, None)
        except _Return:
            pass
        if p.kind == "function":
            r = V[p.result]
            if isinstance(r, Arr):
                return r.a if r.a.ndim else r.a[()]
            return r
        return None

    def call(self, module: str, name: str, *actuals):
        """host entry: call module procedure with Python-side objects (Arr / ndarray / Struct / numpy scalars)"""
        m = self.module(module)
        p = self.lookup_in_module(m, name)
        if isinstance(p, tuple) and p[0] == "generic":
            raise FortranError("call a specific procedure from the host")
        act = {}
        for dn, v in zip(p.args, actuals):
            if isinstance(v, Arr):
                v = v.a
            if isinstance(v, (np.generic, int, float, bool)):
                v = np.array(v)
            act[dn] = v
        for dn in p.args[len(actuals):]:
            act[dn] = ABSENT
        return self._run(p, act)

    # ---- statements ------------------------------------------------------------------------------------------------
    def exec_block(self, stmts, fr, mask):
        for st in stmts:
            self.exec(st, fr, mask)

    def exec(self, st, fr, mask):
        k = st[0]
        self.nstmt += 1
        try:
            if k == "assign":
                self.assign(st[2], st[3], fr, mask)
            elif k == "if":
                for cond, body in st[2]:
                    c = self.eval(cond, fr)
                    if isinstance(c, np.ndarray) and c.ndim:
                        raise FortranError("array-valued IF condition")
                    if bool(c):
                        self.exec_block(body, fr, mask)
                        break
                else:
                    if st[3] is not None:
                        self.exec_block(st[3], fr, mask)
            elif k == "do":
                _, ln, var, lo, hi, step, body = st
                lo, hi = int(self.eval(lo, fr)), int(self.eval(hi, fr))
                stp = int(self.eval(step, fr)) if step is not None else 1
                cell = self.lookup(fr, var)
                if not isinstance(cell, Arr):
                    raise FortranError(f"DO variable {var} is not an integer variable")
                i = lo
                while (stp > 0 and i <= hi) or (stp < 0 and i >= hi):
                    cell.a[...] = i
                    try:
                        self.exec_block(body, fr, mask)
                    except _Cycle:
                        pass
                    except _Exit:
                        break
                    i += stp
                else:
                    cell.a[...] = i
            elif k == "dowhile":
                while bool(self.eval(st[2], fr)):
                    try:
                        self.exec_block(st[3], fr, mask)
                    except _Cycle:
                        pass
                    except _Exit:
                        break
            elif k == "doforever":
                while True:
                    try:
                        self.exec_block(st[2], fr, mask)
                    except _Cycle:
                        pass
                    except _Exit:
                        break
            elif k == "where":
                self.where(st, fr, mask)
            elif k == "call":
                self.exec_call(st, fr)
            elif k == "return":
                raise _Return()
            elif k == "exit":
                raise _Exit()
            elif k == "cycle":
                raise _Cycle()
            elif k == "select":
                v = self.eval(st[2], fr)
                if isinstance(v, Str):
                    v = v.s
                chosen = None
                for vals, body in st[3]:
                    if vals is None:
                        if chosen is None:
                            chosen = body
                        continue
                    hit = False
                    for c in vals:
                        if c[0] == "range":
                            lo = self.eval(c[1], fr) if c[1] is not None else None
                            hi = self.eval(c[2], fr) if c[2] is not None else None
                            hit = (lo is None or v >= lo) and (hi is None or v <= hi)
                        else:
                            cv = self.eval(c, fr)
                            hit = (v.rstrip() == cv.rstrip()) if isinstance(v, str) else bool(v == cv)
                        if hit:
                            break
                    if hit:
                        chosen = body
                        break
                if chosen is not None:
                    self.exec_block(chosen, fr, mask)
            elif k == "allocate":
                for item in st[2]:
                    r = self.resolve_alloc(item, fr)
                    arr, dims = r
                    d = arr.decl
                    lbs, shape = [], []
                    for s in dims:
                        if s[0] == "sec":
                            l, h = int(self.eval(s[1], fr)), int(self.eval(s[2], fr))
                        else:
                            l, h = 1, int(self.eval(s, fr))
                        lbs.append(l); shape.append(max(0, h - l + 1))
                    if d.base == "type":
                        raise FortranError("ALLOCATE of derived-type arrays is not supported")
                    arr.a = np.zeros(tuple(shape), self._kind_dtype(d, fr), order="F")
                    arr.lb = tuple(lbs)
            elif k == "deallocate":
                for item in st[2]:
                    r = self.resolve(item, fr, "alloc")
                    if r[0] == "arr":
                        r[1].a = None
                    else:
                        # allocated: find the Arr again
                        arr, _ = self.resolve_alloc(("des", item[1][:-1] + [(item[1][-1][0], [])]), fr)
                        arr.a = None
            elif k == "nullify" or k == "io_ignored":
                pass
            elif k == "stop":
                raise FortranStop(st[2])
            elif k == "io_fail":
                raise FortranError(f"I/O statement executed: {st[2]}")
            elif k == "unsupported":
                raise FortranError(f"unsupported statement executed: {st[2]} ({st[3]})")
            elif k == "ptrassign":
                raise FortranError("pointer assignment is not supported")
            else:
                raise FortranError(f"statement kind {k}")
        except (FortranError, FortranStop) as e:
            if isinstance(e, StubError) and fr.proc is not None and fr.proc.name in self.tolerate_stubs_in:
                self.skipped.append((fr.proc.name, st[1], str(e)))
                return
            if not getattr(e, "_located", False):
                where_ = f"{os.path.relpath(fr.proc.file, self.src_root) if fr.proc and fr.proc.file else '?'}:{st[1]} in {fr.proc.name if fr.proc else '?'}"
                e.args = (f"{e.args[0] if e.args else ''}  [at {where_}]",)
                e._located = True
            raise

    def resolve_alloc(self, item, fr):
        """ALLOCATE( a(n) ) / ALLOCATE( v%x(n,m) ): -> (Arr, dims)"""
        parts = item[1]
        dims = parts[-1][1]
        base = ("des", parts[:-1] + [(parts[-1][0], None)])
        name0 = parts[0][0]
        if len(parts) == 1:
            ent = self.lookup(fr, name0)
            if not isinstance(ent, Arr):
                raise FortranError(f"ALLOCATE of {name0}: not an allocatable array")
            return ent, dims
        r = self.resolve(base, fr, "alloc")
        if r[0] != "arr":
            # already allocated: walk to the Arr
            cur = self.lookup(fr, name0)
            for nm, ar in parts[1:]:
                cur = cur.f[nm]
            return cur, dims
        return r[1], dims

    def exec_call(self, st, fr):
        _, ln, name, args = st
        if "%" in name:
            raise StubError(f"type-bound call {name}")
        ent = self.lookup(fr, name)
        if ent is None:
            raise FortranError(f"CALL of unknown procedure {name}")
        if isinstance(ent, tuple):
            if ent[0] == "generic":
                ent = self._pick_specific(ent, args, fr)
            elif ent[0] == "internal":
                self.invoke(ent[1], args, fr, host_frame=ent[2])
                return
            elif ent[0] == "stub":
                raise StubError(f"CALL into stubbed module {ent[1]}::{ent[2]}")
            elif ent[0] == "broken":
                raise FortranError(ent[1])
        if self.trace is not None:
            self.trace(ent.name)
        self.invoke(ent, args, fr)

    def assign(self, lhs, rhs, fr, mask):
        v = self.eval(rhs, fr)
        r = self.resolve(lhs, fr, True)
        if r[0] == "str":
            s = v.s if isinstance(v, Str) else v
            if not isinstance(s, str):
                raise FortranError("numeric value assigned to a character variable")
            r[1].s = s if r[1].n is None else s[:r[1].n]
            return
        if r[0] == "struct":
            if isinstance(v, Struct):
                self._copy_struct(r[1], v)
                return
            raise FortranError("assignment to a derived-type object")
        if r[0] != "val":
            raise FortranError("assignment to a function reference")
        tgt = r[1]
        if not isinstance(tgt, np.ndarray):
            raise FortranError("assignment target is not a variable")
        if isinstance(v, list):
            raise FortranError("character array assigned to numeric variable")
        val = self._convert(v, tgt.dtype)
        if mask is None:
            if isinstance(val, np.ndarray) and val.ndim and tgt.ndim == 0:
                raise FortranError("array assigned to a scalar")
            if isinstance(val, np.ndarray) and val.ndim and val.shape != tgt.shape:
                raise FortranError(f"shape mismatch in assignment: {tgt.shape} = {val.shape}")
            tgt[...] = val
        else:
            if tgt.shape != mask.shape:
                raise FortranError(f"WHERE: assignment target shape {tgt.shape} does not conform with the mask {mask.shape}")
            if isinstance(val, np.ndarray) and val.ndim and val.shape != tgt.shape:
                raise FortranError(f"WHERE: shape mismatch {tgt.shape} = {val.shape}")
            np.copyto(tgt, val, where=mask)

    def _copy_struct(self, dst: Struct, src: Struct):
        for k, v in src.f.items():
            d = dst.f[k]
            if isinstance(v, Arr):
                if v.a is None:
                    d.a = None
                elif d.a is not None and d.a.shape == v.a.shape:
                    d.a[...] = v.a
                else:
                    d.a, d.lb = v.a.copy(order="F"), v.lb
            elif isinstance(v, Str):
                d.s = v.s
            elif isinstance(v, Struct):
                self._copy_struct(d, v)

    def where(self, st, fr, outer):
        arms = st[2]
        pending = None            # elements not yet claimed by an earlier arm
        for m_expr, body in arms:
            if m_expr is not None:
                m = np.asarray(self.eval(m_expr, fr), dtype=bool)
                if m.ndim == 0:
                    raise FortranError("scalar WHERE mask")
                cur = m if pending is None else (pending & m)
                pending = (~m) if pending is None else (pending & ~m)
            else:
                cur = pending
            eff = cur if outer is None else (outer & cur)
            for s in body:
                if s[0] == "assign":
                    self.nstmt += 1
                    try:
                        self.assign(s[2], s[3], fr, eff)
                    except FortranError as e:
                        if not getattr(e, "_located", False):
                            e.args = (f"{e.args[0]}  [at {os.path.relpath(fr.proc.file, self.src_root)}:{s[1]} in {fr.proc.name}]",)
                            e._located = True
                        raise
                elif s[0] == "where":
                    self.where(s, fr, eff)
                else:
                    raise FortranError(f"statement {s[0]} inside WHERE")


class _Return(_Ctl):
    pass


class _Exit(_Ctl):
    pass


class _Cycle(_Ctl):
    pass


# ------------------------------------------------------------------------------------------------------------------
# intrinsics: fn(interp, frame, arg trees) -> value

def _vals(I, fr, args):
    pos, kw = [], {}
    for a in args:
        if a[0] == "kw":
            kw[a[1]] = I.eval(a[2], fr)
        else:
            pos.append(I.eval(a, fr))
    return pos, kw


def _fold(op):
    def f(I, fr, args):
        pos, _ = _vals(I, fr, args)
        acc = pos[0]
        for v in pos[1:]:
            a, b, t = _promote(acc, v)
            acc = op(a, b)
        return acc
    return f


def _unary(fn):
    def f(I, fr, args):
        pos, _ = _vals(I, fr, args)
        return fn(pos[0])
    return f


def _kind_arg(I, fr, pos, kw, at=1):
    k = kw.get("kind", pos[at] if len(pos) > at else None)
    return None if k is None else int(k)


def _i_real(I, fr, args):
    pos, kw = _vals(I, fr, args)
    k = _kind_arg(I, fr, pos, kw)
    x = pos[0]
    dt = F64 if k == 8 else F32
    if k is None and _dtype_of(x) == np.dtype(F64):
        dt = F32                                       # REAL(x) of a double is default real
    return x.astype(dt) if isinstance(x, np.ndarray) and x.ndim else dt(x)


def _i_dble(I, fr, args):
    pos, _ = _vals(I, fr, args)
    x = pos[0]
    return x.astype(F64) if isinstance(x, np.ndarray) and x.ndim else F64(x)


def _i_int(I, fr, args):
    pos, _ = _vals(I, fr, args)
    x = pos[0]
    return np.trunc(x).astype(I32) if isinstance(x, np.ndarray) and x.ndim else I32(np.trunc(x))


def _i_nint(I, fr, args):
    pos, _ = _vals(I, fr, args)
    x = pos[0]
    r = np.sign(x) * np.floor(np.abs(x) + 0.5)
    return r.astype(I32) if isinstance(r, np.ndarray) and r.ndim else I32(r)


def _i_floor(I, fr, args):
    pos, _ = _vals(I, fr, args)
    r = np.floor(pos[0])
    return r.astype(I32) if isinstance(r, np.ndarray) and r.ndim else I32(r)


def _i_ceiling(I, fr, args):
    pos, _ = _vals(I, fr, args)
    r = np.ceil(pos[0])
    return r.astype(I32) if isinstance(r, np.ndarray) and r.ndim else I32(r)


def _i_sign(I, fr, args):
    pos, _ = _vals(I, fr, args)
    a, b = pos
    t = _dtype_of(a)
    if _RANK[t] == 1:
        return np.where(np.asarray(b) >= 0, np.abs(a), -np.abs(a)).astype(I32)[()]
    mag = np.abs(a)
    r = np.where(np.signbit(b), -mag, mag)
    return r.astype(t) if isinstance(r, np.ndarray) and r.ndim else t.type(r)


def _i_mod(I, fr, args):
    pos, _ = _vals(I, fr, args)
    a, b, t = _promote(pos[0], pos[1])
    return np.fmod(a, b)


def _i_modulo(I, fr, args):
    pos, _ = _vals(I, fr, args)
    a, b, t = _promote(pos[0], pos[1])
    return np.mod(a, b)


def _i_sum(I, fr, args):
    pos, kw = _vals(I, fr, args)
    a = np.asarray(pos[0])
    dim = kw.get("dim"); mask = kw.get("mask")
    for v in pos[1:]:
        if _dtype_of(v) == np.dtype(B):
            mask = v
        else:
            dim = v
    return _seq_sum(a, None if dim is None else int(dim) - 1, mask)


def _i_product(I, fr, args):
    pos, kw = _vals(I, fr, args)
    a = np.asarray(pos[0])
    acc = a.dtype.type(1)
    for v in np.ravel(a, order="F"):
        acc = acc * v
    return acc


def _i_spread(I, fr, args):
    pos, kw = _vals(I, fr, args)
    src = pos[0]
    dim = int(kw.get("dim", pos[1] if len(pos) > 1 else 1))
    n = int(kw.get("ncopies", pos[2] if len(pos) > 2 else 1))
    a = np.asarray(src)
    return np.repeat(np.expand_dims(a, dim - 1), n, axis=dim - 1)


def _i_merge(I, fr, args):
    pos, kw = _vals(I, fr, args)
    t, f, m = pos[0], pos[1], (pos[2] if len(pos) > 2 else kw["mask"])
    if isinstance(t, str):
        return t if bool(m) else f
    t, f, dt = _promote(t, f)
    r = np.where(m, t, f)
    return r if isinstance(r, np.ndarray) and r.ndim else dt.type(r)


def _i_size(I, fr, args):
    pos, kw = _vals(I, fr, args)
    a = np.asarray(pos[0])
    dim = kw.get("dim", pos[1] if len(pos) > 1 else None)
    return I32(a.size if dim is None else a.shape[int(dim) - 1])


def _bound(which):
    def f(I, fr, args):
        ex = args[0]
        r = I.lookup(fr, ex[1][0][0]) if ex[0] == "des" and len(ex[1]) == 1 and ex[1][0][1] is None else None
        pos, kw = _vals(I, fr, args[1:])
        dim = kw.get("dim", pos[0] if pos else None)
        if isinstance(r, Arr):
            lbs = r.lb or (1,) * r.a.ndim
            b = [l if which == "l" else l + s - 1 for l, s in zip(lbs, r.a.shape)]
        else:
            a = np.asarray(I.eval(ex, fr))
            b = [1 if which == "l" else s for s in a.shape]
        return I32(b[int(dim) - 1]) if dim is not None else np.array(b, I32)
    return f


def _minmaxval(fn):
    def f(I, fr, args):
        pos, kw = _vals(I, fr, args)
        a = np.asarray(pos[0])
        dim = kw.get("dim"); mask = kw.get("mask")
        for v in pos[1:]:
            if _dtype_of(v) == np.dtype(B):
                mask = v
            else:
                dim = v
        if mask is not None:
            fill = (np.finfo(a.dtype).max if a.dtype.kind == "f" else np.iinfo(a.dtype).max)
            a = np.where(mask, a, -fill if fn is np.max else fill)
        return fn(a) if dim is None else fn(a, axis=int(dim) - 1)
    return f


def _i_any(I, fr, args):
    pos, kw = _vals(I, fr, args)
    dim = kw.get("dim", pos[1] if len(pos) > 1 else None)
    return B(np.any(pos[0])) if dim is None else np.any(pos[0], axis=int(dim) - 1)


def _i_all(I, fr, args):
    pos, kw = _vals(I, fr, args)
    dim = kw.get("dim", pos[1] if len(pos) > 1 else None)
    return B(np.all(pos[0])) if dim is None else np.all(pos[0], axis=int(dim) - 1)


def _i_count(I, fr, args):
    pos, kw = _vals(I, fr, args)
    dim = kw.get("dim", pos[1] if len(pos) > 1 else None)
    return I32(np.count_nonzero(pos[0])) if dim is None else np.count_nonzero(pos[0], axis=int(dim) - 1).astype(I32)


def _i_present(I, fr, args):
    name = args[0][1][0][0]
    return B(fr.vars.get(name, ABSENT) is not ABSENT)


def _i_allocated(I, fr, args):
    r = I.resolve(args[0], fr, "alloc")
    return B(r[0] != "arr")


def _i_kind(I, fr, args):
    pos, _ = _vals(I, fr, args)
    t = _dtype_of(pos[0])
    return I32(8 if t == np.dtype(F64) else 4)


def _i_srk(I, fr, args):
    pos, kw = _vals(I, fr, args)
    p = int(kw.get("p", pos[0] if pos else 6))
    return I32(8 if p > 6 else 4)


def _i_sik(I, fr, args):
    return I32(4)


def _num_inquiry(attr):
    def f(I, fr, args):
        pos, _ = _vals(I, fr, args)
        t = _dtype_of(pos[0])
        if t.kind == "f":
            fi = np.finfo(t)
            return t.type({"tiny": fi.tiny, "huge": fi.max, "epsilon": fi.eps}[attr])
        return I32(np.iinfo(np.int32).max) if attr == "huge" else I32(0)
    return f


def _i_reshape(I, fr, args):
    pos, kw = _vals(I, fr, args)
    src, shape = np.asarray(pos[0]), kw.get("shape", pos[1] if len(pos) > 1 else None)
    return np.reshape(np.ravel(src, order="F"), tuple(int(x) for x in np.asarray(shape)), order="F")


def _i_trim(I, fr, args):
    pos, _ = _vals(I, fr, args)
    s = pos[0].s if isinstance(pos[0], Str) else pos[0]
    return s.rstrip()


def _i_len_trim(I, fr, args):
    pos, _ = _vals(I, fr, args)
    s = pos[0].s if isinstance(pos[0], Str) else pos[0]
    return I32(len(s.rstrip()))


def _i_abs(I, fr, args):
    pos, _ = _vals(I, fr, args)
    return np.abs(pos[0])


def _i_sqrt(I, fr, args):
    pos, _ = _vals(I, fr, args)
    return np.sqrt(pos[0])


def _i_atan2(I, fr, args):
    pos, _ = _vals(I, fr, args)
    a, b, t = _promote(pos[0], pos[1])
    if t == np.dtype(F32):
        r = np.arctan2(np.asarray(a, F64), np.asarray(b, F64))
        return r.astype(F32) if r.ndim else F32(r)
    return np.arctan2(a, b)


def _i_isnan(I, fr, args):
    pos, _ = _vals(I, fr, args)
    return np.isnan(pos[0])


def _i_pack(I, fr, args):
    pos, kw = _vals(I, fr, args)
    a, m = np.asarray(pos[0]), np.asarray(pos[1])
    return np.ravel(a, order="F")[np.ravel(np.broadcast_to(m, a.shape), order="F")]


def _i_dot(I, fr, args):
    pos, _ = _vals(I, fr, args)
    a, b, t = _promote(pos[0], pos[1])
    return _seq_sum(a * b)


def _i_null(I, fr, args):
    return None


INTRINSICS = {
    "max": _fold(np.maximum), "min": _fold(np.minimum), "amax1": _fold(np.maximum), "amin1": _fold(np.minimum),
    "max0": _fold(np.maximum), "min0": _fold(np.minimum),
    "abs": _i_abs, "sqrt": _i_sqrt,
    "exp": _unary(_elem(_exp64)), "log": _unary(_elem(_log64)), "alog": _unary(_elem(_log64)),
    "log10": _unary(_elem(_log1064)), "alog10": _unary(_elem(_log1064)),
    "sin": _unary(_elem(_sin64)), "cos": _unary(_elem(_cos64)), "tan": _unary(_elem(_tan64)),
    "atan": _unary(_elem(_atan64)), "asin": _unary(_elem(_asin64)), "acos": _unary(_elem(_acos64)),
    "tanh": _unary(_elem(_tanh64)), "sinh": _unary(_elem(_sinh64)), "cosh": _unary(_elem(_cosh64)),
    "atan2": _i_atan2,
    "real": _i_real, "float": _i_real, "sngl": _i_real, "dble": _i_dble, "int": _i_int, "nint": _i_nint,
    "floor": _i_floor, "ceiling": _i_ceiling, "sign": _i_sign, "mod": _i_mod, "modulo": _i_modulo,
    "sum": _i_sum, "product": _i_product, "spread": _i_spread, "merge": _i_merge, "size": _i_size,
    "lbound": _bound("l"), "ubound": _bound("u"),
    "maxval": _minmaxval(np.max), "minval": _minmaxval(np.min), "any": _i_any, "all": _i_all, "count": _i_count,
    "present": _i_present, "allocated": _i_allocated, "associated": _i_allocated,
    "kind": _i_kind, "selected_real_kind": _i_srk, "selected_int_kind": _i_sik,
    "tiny": _num_inquiry("tiny"), "huge": _num_inquiry("huge"), "epsilon": _num_inquiry("epsilon"),
    "reshape": _i_reshape, "trim": _i_trim, "len_trim": _i_len_trim, "adjustl": _i_trim,
    "isnan": _i_isnan, "ieee_is_nan": _i_isnan, "pack": _i_pack, "dot_product": _i_dot, "null": _i_null,
}
