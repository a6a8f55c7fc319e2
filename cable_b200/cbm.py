"""Host-side mirror of the reference interface of the hot path.

Reference: MODULE cable_cbm_module, SUBROUTINE cbm(ktau, dels, air, bgc, canopy, met, bal, rad,
rough, soil, ssnow, sum_flux, veg, climate, xk, c1, rhoch)
(src/offline/cbl_model_driver_offline.F90:38-40; callers cable_serial.F90:594, cable_mpiworker.F90:503).

`CableB200` is the handle the Fortran shim keeps behind that signature: it binds the
caller's column-major arrays once, and `cbm()` then behaves like `CALL cbm(...)`.  The derived types
are presented as namespaces of NumPy views (`DerivedTypes`), so tests read like the reference call.
Everything numerical happens in libcable_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import numpy as np

from . import lib as _lib
from .registry import FIELDS, BY_NAME, ROLE, FLAG, alloc_tiles

# reference type names (cable_define_types.F90) for the registry prefixes
TYPE_NAMES = {
    "met": "met_type", "air": "air_type", "veg": "veg_parameter_type", "soil": "soil_parameter_type",
    "ssnow": "soil_snow_type", "canopy": "canopy_type", "rad": "radiation_type", "rough": "roughness_type",
    "bal": "balances_type", "bgc": "bgc_pool_type", "scr": "(xk, c1, rhoch scratch)",
}


def derived_types(tiles: dict[str, np.ndarray]) -> SimpleNamespace:
    """Group the flat field dict into the reference's derived types: types.ssnow.tgg is (6, mp)."""
    groups: dict[str, SimpleNamespace] = {}
    for f in FIELDS:
        groups.setdefault(f.type, SimpleNamespace())
        a = tiles[f.name]
        setattr(groups[f.type], f.member, a[0] if f.ncomp == 1 else a)
    return SimpleNamespace(**groups)


class CableB200:
    """One handle per (process, GPU); not thread-safe; calls are stream-ordered."""

    def __init__(self, mp: int, cfg: _lib.CableCfg | None = None, device: int = -1):
        self._lib = _lib.load()
        self.cfg = cfg if cfg is not None else _lib.default_cfg()
        self.mp = int(mp)
        self._h = C.c_void_p()
        _lib.check(self._lib.cable_b200_create(self.mp, C.byref(self.cfg), device, C.byref(self._h)))
        self._bound: dict[str, np.ndarray] = {}

    # -- life cycle ---------------------------------------------------------------------------
    def close(self) -> None:
        if self._h:
            self._lib.cable_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- binding ------------------------------------------------------------------------------
    def bind(self, tiles: dict[str, np.ndarray]) -> None:
        """Bind caller-owned arrays (what the shim does with C_LOC of every derived-type member)."""
        for f in FIELDS:
            a = tiles.get(f.name)
            if a is None:
                continue
            if a.dtype != f.dtype or a.size != f.ncomp * self.mp or not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"{f.name}: need C-contiguous {f.dtype.__name__} of {f.ncomp}x{self.mp}")
            _lib.check(self._lib.cable_b200_bind_field(self._h, f.id, a.ctypes.data_as(C.c_void_p)))
            self._bound[f.name] = a

    def upload_params(self) -> None:
        _lib.check(self._lib.cable_b200_upload(self._h, ROLE["PARAM"]))

    def upload_state(self) -> None:
        _lib.check(self._lib.cable_b200_upload(self._h, ROLE["STATE"]))

    def download_state(self) -> None:
        _lib.check(self._lib.cable_b200_download(self._h, ROLE["STATE"], 0))

    def download_diag(self, star_only: bool = True) -> None:
        _lib.check(self._lib.cable_b200_download(self._h, ROLE["DIAG"], FLAG["STAR"] if star_only else 0))

    # -- stepping -----------------------------------------------------------------------------
    def set_forcing_async(self, slot: int = 0) -> None:
        _lib.check(self._lib.cable_b200_set_forcing_async(self._h, slot))

    def step(self, ktau: int, dels: float, slot: int = 0) -> None:
        _lib.check(self._lib.cable_b200_step(self._h, int(ktau), float(dels), slot))

    def cbm(self, ktau: int, dels: float) -> None:
        """Drop-in call: forcing up, one step, outputs (cfg.output_level) back, synchronised."""
        _lib.check(self._lib.cable_b200_cbm(self._h, int(ktau), float(dels)))

    def sync(self) -> None:
        _lib.check(self._lib.cable_b200_sync(self._h))

    # -- device access / measurement ----------------------------------------------------------
    def device_ptr(self, name: str, slot: int = 0) -> int:
        p = self._lib.cable_b200_device_ptr(self._h, BY_NAME[name].id, slot)
        if not p:
            raise KeyError(name)
        return int(p)

    def compute_stream(self) -> int:
        return int(self._lib.cable_b200_compute_stream(self._h) or 0)

    def profile(self, enable: bool) -> None:
        _lib.check(self._lib.cable_b200_profile(self._h, int(enable)))

    def counters(self) -> _lib.Counters:
        c = _lib.Counters()
        _lib.check(self._lib.cable_b200_get_counters(self._h, C.byref(c)))
        return c

    def reset_counters(self) -> None:
        _lib.check(self._lib.cable_b200_reset_counters(self._h))

    def grid_reduce(self, name: str, comp: int, d_patchfrac: int, d_cstart: int, d_cend: int, nland: int, d_out: int) -> None:
        _lib.check(self._lib.cable_b200_grid_reduce(self._h, BY_NAME[name].id, comp, d_patchfrac, d_cstart, d_cend, nland, d_out))


_HANDLES: dict[int, CableB200] = {}


def cbm(ktau, dels, air, bgc, canopy, met, bal, rad, rough, soil, ssnow, sum_flux, veg, climate, xk, c1, rhoch,
        cfg: _lib.CableCfg | None = None):
    """Same argument list as the reference `cbm` (cbl_model_driver_offline.F90:38-40).

    The derived-type arguments are namespaces of (ncomp, mp) NumPy arrays (see `derived_types`).
    On first call for a given `ssnow` object a handle is created and every member is bound; parameters
    and state are uploaded once; afterwards each call is cable_b200_cbm().  `sum_flux` and `climate`
    are accepted and ignored exactly as the reference ignores them on the default path.
    """
    key = id(ssnow)
    h = _HANDLES.get(key)
    if h is None:
        groups = {"air": air, "bgc": bgc, "canopy": canopy, "met": met, "bal": bal, "rad": rad, "rough": rough,
                  "soil": soil, "ssnow": ssnow, "veg": veg, "scr": SimpleNamespace(xk=xk, c1=c1, rhoch=rhoch)}
        mp = np.asarray(met.tk).shape[-1]
        tiles = {}
        for f in FIELDS:
            a = getattr(groups[f.type], f.member, None)
            if a is not None:
                tiles[f.name] = np.asarray(a).reshape(f.ncomp, mp)
        h = CableB200(mp, cfg)
        h.bind(tiles)
        h.upload_params()
        h.upload_state()
        _HANDLES[key] = h
    h.cbm(ktau, dels)
    return h


def release(ssnow) -> None:
    h = _HANDLES.pop(id(ssnow), None)
    if h is not None:
        h.close()
