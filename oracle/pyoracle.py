"""ctypes driver of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
LIB_CR_PATH = os.path.join(_HERE, "liboracle_cr.so")
_libs = {}


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", _HERE, "-j8"])


def load(cr_math: bool = False) -> C.CDLL:
    """cr_math=False: host libm float intrinsics (a gfortran build's semantics);
    cr_math=True: correctly rounded fp32 intrinsics (isolates logic from libm differences)."""
    path = LIB_CR_PATH if cr_math else LIB_PATH
    _lib = _libs.get(path)
    if _lib is None:
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.oracle_nfields.restype = C.c_int
        lib.oracle_field_name.restype = C.c_char_p
        lib.oracle_field_name.argtypes = [C.c_int]
        lib.oracle_create.restype = C.c_void_p
        lib.oracle_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        lib.oracle_cbm.argtypes = [C.c_void_p, C.c_int, C.c_float]
        lib.oracle_dryleaf_warnings.restype = C.c_longlong
        lib.oracle_dryleaf_warnings.argtypes = [C.c_void_p]
        lib.oracle_destroy.argtypes = [C.c_void_p]
        lib.oracle_destroy.restype = None
        lib.oracle_trimb.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int]
        lib.oracle_trimb.restype = None
        for fn in (lib.oracle_psim, lib.oracle_psis):
            fn.restype = C.c_float
            fn.argtypes = [C.c_float]
        lib.oracle_qsat.restype = C.c_float
        lib.oracle_qsat.argtypes = [C.c_float, C.c_float]
        _libs[path] = _lib = lib
    return _lib


class Oracle:
    """Runs the restated cbm() in place on a dict of (ncomp, mp) arrays (registry layout)."""

    def __init__(self, tiles: dict[str, np.ndarray], cfg, cr_math: bool = False):
        lib = load(cr_math)
        n = lib.oracle_nfields()
        names = [lib.oracle_field_name(i).decode() for i in range(n)]
        self.tiles = tiles
        mp = tiles["met_tk"].shape[-1]
        ptrs = (C.c_void_p * n)()
        for i, nm in enumerate(names):
            a = tiles[nm]
            assert a.flags["C_CONTIGUOUS"], nm
            ptrs[i] = a.ctypes.data
        self._cfg = cfg
        self._h = lib.oracle_create(mp, C.addressof(cfg), ptrs)
        if not self._h:
            raise RuntimeError("oracle_create failed")
        self._lib = lib

    def cbm(self, ktau: int, dels: float) -> None:
        rc = self._lib.oracle_cbm(self._h, int(ktau), float(dels))
        if rc:
            raise RuntimeError("oracle_cbm failed")

    def warnings(self) -> int:
        return int(self._lib.oracle_dryleaf_warnings(self._h))

    def close(self) -> None:
        if self._h:
            self._lib.oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
