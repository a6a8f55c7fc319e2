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


# ---- driver stages either side of cbm (oracle/o_driver.cpp) -----------------------------------------------------
DRIVER_ARRAYS = ("tscrn_max_daily", "tscrn_min_daily", "sumpn", "sumrp", "sumrpw", "sumrpr", "sumrs", "sumrd", "dsumpn",
                 "dsumrp", "dsumrd", "owb", "wbal", "wbal_tot", "precip_tot", "rnoff_tot", "evap_tot", "radbal", "ebalsoil",
                 "ebalveg", "ebal", "ebal_tot", "radbalsum")
# names of the same arrays at the C ABI of the product (cable_b200_driver_download)
DRIVER_ABI_NAMES = dict(zip(DRIVER_ARRAYS, (
    "canopy_tscrn_max_daily", "canopy_tscrn_min_daily", "sum_flux_sumpn", "sum_flux_sumrp", "sum_flux_sumrpw",
    "sum_flux_sumrpr", "sum_flux_sumrs", "sum_flux_sumrd", "sum_flux_dsumpn", "sum_flux_dsumrp", "sum_flux_dsumrd",
    "bal_owb", "bal_wbal", "bal_wbal_tot", "bal_precip_tot", "bal_rnoff_tot", "bal_evap_tot", "bal_Radbal",
    "bal_EbalSoil", "bal_Ebalveg", "bal_ebal", "bal_ebal_tot", "bal_Radbalsum")))


class _DriverArrays(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in DRIVER_ARRAYS]


def sinbet(doy: float, lat: float, hod: float, cr_math: bool = True) -> float:
    lib = load(cr_math)
    lib.oracle_sinbet.restype = C.c_float
    lib.oracle_sinbet.argtypes = [C.c_float] * 3
    return float(lib.oracle_sinbet(doy, lat, hod))


def met_expand(tiles: dict, land: np.ndarray, cstart: np.ndarray, cend: np.ndarray, latitude: np.ndarray,
               tair_offset: float, psurf_scale: float, rainf_scale: float, co2_scale: float, snowf_from_tair: bool,
               cr_math: bool = True) -> None:
    """get_met_data's tile expansion into tiles['met_*'] (in place)."""
    lib = load(cr_math)
    mp = tiles["met_tk"].shape[-1]
    land = np.ascontiguousarray(land, np.float32)
    cs, ce = np.ascontiguousarray(cstart, np.int32), np.ascontiguousarray(cend, np.int32)
    lat = np.ascontiguousarray(latitude, np.float32)
    lib.oracle_met_expand.restype = None
    lib.oracle_met_expand.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_float] * 4 + [C.c_int] + [C.c_void_p] * 11
    names = ("met_fsd", "met_tk", "met_pmb", "met_qv", "met_ua", "met_precip", "met_precip_sn", "met_fld", "met_ca",
             "met_coszen", "met_doy")
    lib.oracle_met_expand(mp, cs.size, land.ctypes.data, cs.ctypes.data, ce.ctypes.data, lat.ctypes.data,
                          tair_offset, psurf_scale, rainf_scale, co2_scale, int(snowf_from_tair),
                          *[tiles[n].ctypes.data for n in names])


class OracleDriver:
    """Driver-owned arrays (bal%*, sum_flux%*, daily tscrn extremes) + the post-step statements on an Oracle."""

    def __init__(self, oracle: Oracle):
        self.o = oracle
        mp = oracle.tiles["met_tk"].shape[-1]
        self.arrays = {n: np.zeros(mp, np.float64 if n == "owb" else np.float32) for n in DRIVER_ARRAYS}
        self.arrays["tscrn_max_daily"][:] = -np.finfo(np.float32).max      # aggregator.F90:1068-1119 / :1015-1066
        self.arrays["tscrn_min_daily"][:] = np.finfo(np.float32).max
        self._c = _DriverArrays(*[self.arrays[n].ctypes.data for n in DRIVER_ARRAYS])
        lib = oracle._lib
        lib.oracle_post_step.restype = None
        lib.oracle_post_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]

    def post_step(self, ktau: int, kstart: int, dels: float, mass_bal: bool = True, energy_bal: bool = True) -> None:
        self.o._lib.oracle_post_step(self.o._h, C.addressof(self._c), int(ktau), int(kstart), float(dels), int(mass_bal), int(energy_bal))


def aggregate(src: np.ndarray, method: int, agg: np.ndarray, counter: int, scale=1.0, div=1.0, offset=0.0, cr_math=True) -> None:
    lib = load(cr_math)
    lib.oracle_aggregate.restype = None
    lib.oracle_aggregate.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int]
    dt = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2}[src.dtype]
    assert agg.dtype == np.float64 and src.flags["C_CONTIGUOUS"] and agg.flags["C_CONTIGUOUS"]
    lib.oracle_aggregate(src.size, src.ctypes.data, dt, method, scale, div, offset, agg.ctypes.data, counter)


def grid_reduce(x: np.ndarray, patchfrac: np.ndarray, cstart: np.ndarray, cend: np.ndarray, cr_math=True) -> np.ndarray:
    lib = load(cr_math)
    lib.oracle_grid_reduce.restype = None
    lib.oracle_grid_reduce.argtypes = [C.c_int] + [C.c_void_p] * 5
    cs, ce = np.ascontiguousarray(cstart, np.int32), np.ascontiguousarray(cend, np.int32)
    x, pf = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(patchfrac, np.float32)
    out = np.empty(cs.size, np.float32)
    lib.oracle_grid_reduce(cs.size, cs.ctypes.data, ce.ctypes.data, x.ctypes.data, pf.ctypes.data, out.ctypes.data)
    return out
