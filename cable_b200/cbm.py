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
    "bal": "balances_type", "bgc": "bgc_pool_type", "scr": "(xk, c1, rhoch scratch)", "climate": "climate_type",
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

    def param_table_classes(self) -> int:
        """bit 0: veg%* served from per-PFT tables in shared memory, bit 1: soil%* from per-soil-type tables"""
        return int(self._lib.cable_b200_param_table_classes(self._h))

    def mark_dirty(self, *names: str) -> None:
        """The host wrote these resident (PARAM / STATE) arrays: the next step uploads them first."""
        for n in names:
            _lib.check(self._lib.cable_b200_mark_dirty(self._h, BY_NAME[n].id))

    def set_output_mask(self, names) -> None:
        """Restrict what cbm() mirrors to the host every step to these fields (empty: the output_level default)."""
        ids = np.asarray([BY_NAME[n].id for n in names], np.int32)
        _lib.check(self._lib.cable_b200_set_output_mask(self._h, ids.ctypes.data if ids.size else None, int(ids.size)))

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


    # -- driver stages either side of cbm ---------------------------------------------
    def driver_init(self, cstart: np.ndarray, cend: np.ndarray, patchfrac: np.ndarray, latitude: np.ndarray) -> None:
        """landpt(:)%cstart-1 / %cend-1 (0-based, inclusive), patch(:)%frac, rad%latitude (per tile)."""
        cs, ce = np.ascontiguousarray(cstart, np.int32), np.ascontiguousarray(cend, np.int32)
        pf, la = np.ascontiguousarray(patchfrac, np.float32).ravel(), np.ascontiguousarray(latitude, np.float32).ravel()
        self.nland = int(cs.size)
        _lib.check(self._lib.cable_b200_driver_init(self._h, self.nland, cs.ctypes.data, ce.ctypes.data, pf.ctypes.data, la.ctypes.data))

    def set_met_async(self, slot: int, met_land: np.ndarray, convert: _lib.MetConvert) -> None:
        """met_land: float32 [len(MET_ROWS), nland], C-contiguous; must stay alive until the copy has run."""
        assert met_land.dtype == np.float32 and met_land.shape == (len(_lib.MET_ROWS), self.nland) and met_land.flags["C_CONTIGUOUS"]
        _lib.check(self._lib.cable_b200_set_met_async(self._h, slot, met_land.ctypes.data, C.byref(convert)))

    def upload_lai(self) -> None:
        _lib.check(self._lib.cable_b200_upload_lai(self._h))

    def post_step(self, ktau: int, kstart: int, dels: float, mass_bal: bool = True, energy_bal: bool = True) -> None:
        _lib.check(self._lib.cable_b200_post_step(self._h, int(ktau), int(kstart), float(dels), int(mass_bal), int(energy_bal)))

    def output_plan(self, rows) -> None:
        """rows: iterable of (name, comp, method[, scale, div, offset]); name is a registry field or a driver array."""
        ids, comps, meth, sc, dv, off = [], [], [], [], [], []
        for r in rows:
            name, comp, method = r[0], r[1], r[2]
            if name in BY_NAME:
                ids.append(BY_NAME[name].id)
            else:
                k = self._lib.cable_b200_driver_field_id(name.encode())
                if k < 0:
                    raise KeyError(name)
                ids.append(-(1 + k))
            comps.append(comp); meth.append(_lib.AGG[method] if isinstance(method, str) else int(method))
            sc.append(r[3] if len(r) > 3 else 1.0); dv.append(r[4] if len(r) > 4 else 1.0); off.append(r[5] if len(r) > 5 else 0.0)
        a = [np.asarray(ids, np.int32), np.asarray(comps, np.int32), np.asarray(meth, np.int32),
             np.asarray(sc, np.float32), np.asarray(dv, np.float32), np.asarray(off, np.float32)]
        self.nrows = len(ids)
        _lib.check(self._lib.cable_b200_output_plan(self._h, self.nrows, *[x.ctypes.data for x in a]))

    def output_accumulate(self) -> None:
        _lib.check(self._lib.cable_b200_output_accumulate(self._h))

    def output_fetch_async(self, host_out: np.ndarray) -> None:
        assert host_out.dtype == np.float32 and host_out.size == self.nrows * self.nland and host_out.flags["C_CONTIGUOUS"]
        _lib.check(self._lib.cable_b200_output_fetch_async(self._h, host_out.ctypes.data))

    # -- multi-GPU gather of the output block (NCCL inside the library; no torch involved) ----------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """rank 0 creates it; the caller broadcasts the 128 bytes with whatever it has (MPI_Bcast in the Fortran driver)."""
        buf = C.create_string_buffer(128)
        _lib.check(_lib.load().cable_b200_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, nranks: int) -> None:
        assert len(unique_id) == 128
        self.rank, self.nranks = int(rank), int(nranks)
        _lib.check(self._lib.cable_b200_comm_init(self._h, C.create_string_buffer(unique_id, 128), self.rank, self.nranks))

    def output_gather_async(self, root: int, host_out, nland_of_rank) -> None:
        """host_out: float32 [nrows, sum(nland_of_rank)] on the root (pinned memory recommended), None elsewhere."""
        counts = np.ascontiguousarray(nland_of_rank, np.int32)
        if host_out is not None:
            assert host_out.dtype == np.float32 and host_out.size == self.nrows * int(counts.sum()) and host_out.flags["C_CONTIGUOUS"]
        _lib.check(self._lib.cable_b200_output_gather_async(self._h, int(root), host_out.ctypes.data if host_out is not None else None,
                                                            counts.ctypes.data))

    def output_wait(self) -> None:
        _lib.check(self._lib.cable_b200_output_wait(self._h))

    def driver_download(self, name: str) -> np.ndarray:
        out = np.empty(self.mp, np.float64 if name == "bal_owb" else np.float32)
        _lib.check(self._lib.cable_b200_driver_download(self._h, name.encode(), out.ctypes.data))
        return out


_HANDLES: dict[int, CableB200] = {}


def cbm(ktau, dels, air, bgc, canopy, met, bal, rad, rough, soil, ssnow, sum_flux, veg, climate, xk, c1, rhoch,
        cfg: _lib.CableCfg | None = None):
    """Same argument list as the reference `cbm` (cbl_model_driver_offline.F90:38-40).

    The derived-type arguments are namespaces of (ncomp, mp) NumPy arrays (see `derived_types`).
    On first call for a given `ssnow` object a handle is created and every member is bound; parameters
    and state are uploaded once; afterwards each call is cable_b200_cbm().  `sum_flux` is accepted and ignored
    exactly as the reference ignores it; `climate` (may be None) is read only under cable_user%call_climate.
    """
    key = id(ssnow)
    h = _HANDLES.get(key)
    if h is None:
        groups = {"air": air, "bgc": bgc, "canopy": canopy, "met": met, "bal": bal, "rad": rad, "rough": rough,
                  "soil": soil, "ssnow": ssnow, "veg": veg, "scr": SimpleNamespace(xk=xk, c1=c1, rhoch=rhoch),
                  "climate": climate if climate is not None else SimpleNamespace()}
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
