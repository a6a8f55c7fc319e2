"""run_cbm.py -- drive the reference's own SUBROUTINE cbm (src/offline/cbl_model_driver_offline.F90:38) through the
Fortran interpreter of this package, on registry-layout tile arrays (TEST INFRASTRUCTURE ONLY; needs /root/reference).

What the harness does is what the offline driver does around the call and nothing else:
  * cable_def_types_mod::mp, mvtype, mstype are set; every derived type is allocated by the reference's own
    alloc_*_type routines (cable_define_types.F90:760-1420) -- i.e. ALL members exist, zero-filled;
  * members that have a registry row (include/cable_b200_fields.def) are filled from / read back into the caller's
    (ncomp, mp) arrays; soil%zse, soil%zshh, bgc%ratecp, bgc%ratecs come from cable_cfg;
  * module switches: cable_user%*, cable_runtime%offline = .TRUE., icycle, the snow / soil tunables of
    cable_common_module, redistrb, wiltParam, satuParam;
  * the caller's duties before CALL cbm: canopy%oldcansto = canopy%cansto (cable_serial.F90:573) and
    met%tvair = met%tvrad = met%tk (cable_input.F90:2679-2680) when the configuration says the caller does them.
"""
from __future__ import annotations

import numpy as np

from .finterp import Arr, Interp, Str, Struct, FortranError

REF_SRC = "/root/reference/src"
# modules that are USEd by the hot-path files but never executed in this configuration (SLI, groundwater, the output
# aggregators): referenced names resolve to stubs that fail loudly if control ever reaches them
STUBS = ("sli_main_mod", "cable_gw_hydro_module", "gwstempv_mod", "aggregator_mod", "cable_iovars", "cable_io_vars_module",
         "netcdf", "mpi", "cable_abort_module")

TYPE_OF = {"met": "met_type", "air": "air_type", "veg": "veg_parameter_type", "soil": "soil_parameter_type",
           "ssnow": "soil_snow_type", "canopy": "canopy_type", "rad": "radiation_type", "rough": "roughness_type",
           "bal": "balances_type", "bgc": "bgc_pool_type", "sum_flux": "sum_flux_type", "climate": "climate_type"}
ALLOC_OF = {"met": "alloc_met_type", "air": "alloc_air_type", "veg": "alloc_veg_parameter_type", "soil": "alloc_soil_parameter_type",
            "ssnow": "alloc_soil_snow_type", "canopy": "alloc_canopy_type", "rad": "alloc_radiation_type",
            "rough": "alloc_roughness_type", "bal": "alloc_balances_type", "bgc": "alloc_bgc_pool_type",
            "sum_flux": "alloc_sum_flux_type"}

GS = {0: "leuning", 1: "medlyn"}
FWSOIL = {0: "standard", 1: "non-linear extrapolation", 2: "Lai and Ktaul 2000", 3: "Haverd2013"}


def _to_fortran(a: np.ndarray, n1: int, n2: int) -> np.ndarray:
    """registry (n1*n2, mp), component k = a + n1*b  ->  Fortran (mp[, n1[, n2]])"""
    mp = a.shape[-1]
    if n1 == 1 and n2 == 1:
        return a[0]
    if n2 == 1:
        return a.T
    return a.reshape(n2, n1, mp).transpose(2, 1, 0)


class FortranCbm:
    """The reference cbm() executed from its source, stepping registry-layout tiles in place (same interface as
    oracle.pyoracle.Oracle)."""

    def __init__(self, tiles: dict, cfg, fields, src_root: str = REF_SRC):
        self.I = I = Interp(src_root, stub_modules=STUBS)
        # alloc_canopy_type ends with two statements that set up output aggregators of the I/O layer (type-bound
        # procedures of aggregator_mod, cable_define_types.F90:1200-1201); they have no part in cbm and are skipped
        I.tolerate_stubs_in.add("alloc_canopy_type")
        self.tiles, self.cfg, self.fields = tiles, cfg, fields
        mp = self.mp = int(tiles["met_tk"].shape[-1])
        dt = I.module("cable_def_types_mod")
        self._set(dt, "mp", mp); self._set(dt, "mvtype", int(cfg.mvtype)); self._set(dt, "mstype", 9)
        self._set(dt, "mland", mp); self._set(dt, "mp_global", mp); self._set(dt, "mland_global", mp)
        # derived types, allocated by the reference's own routines
        self.S = {}
        for short, tname in TYPE_OF.items():
            tdef = I.lookup_in_module(dt, "$type:" + tname)
            s = I.new_struct(tdef)
            if short in ALLOC_OF:
                I.call("cable_def_types_mod", ALLOC_OF[short], s, np.int32(mp))
            else:           # climate_type: only qtemp_max_last_year is read (call_climate)
                for c in tdef.comps:
                    if c.name == "qtemp_max_last_year":
                        s.f[c.name].a = np.zeros(mp, np.float32, order="F"); s.f[c.name].lb = (1,)
            self.S[short] = s
        self.scr = {n: np.zeros((mp, 3), np.float32, order="F") for n in ("xk", "c1", "rhoch")}
        # module-scope inputs
        cm = I.module("cable_common_module")
        user = I.lookup_in_module(cm, "cable_user")
        self._sets(user, "gs_switch", GS[int(cfg.gs_switch)])
        self._sets(user, "fwsoil_switch", FWSOIL[int(cfg.fwsoil_switch)])
        self._sets(user, "ssnow_potev", "P-M" if cfg.ssnow_potev else "HDM")
        self._sets(user, "diag_soil_resp", "ON " if cfg.diag_soil_resp_on else "off")
        self._sets(user, "soil_struc", "default")
        for name in ("l_new_runoff_speed", "l_new_reduce_soilevp", "litter", "or_evap", "gw_model", "l_rev_corr",
                     "soil_thermal_fix", "l_new_roughness_soil", "call_climate"):
            self._setl(user, name, bool(getattr(cfg, name)))
        rt = I.lookup_in_module(cm, "cable_runtime")
        for name in ("um", "um_explicit", "um_implicit", "um_radiation", "um_hydrology", "esm15", "mk3l"):
            if name in rt.f:
                self._setl(rt, name, False)
        self._setl(rt, "offline", True)
        for name, val in (("snmin", cfg.snmin), ("max_glacier_snowd", cfg.max_glacier_snowd), ("snow_ccnsw", cfg.snow_ccnsw),
                          ("max_ssdn", cfg.max_ssdn), ("max_sconds", cfg.max_sconds), ("frozen_limit", cfg.frozen_limit),
                          ("wiltparam", cfg.wiltParam), ("satuparam", cfg.satuParam)):
            self._set(cm, name, np.float32(val))
        self._set(cm, "redistrb", bool(cfg.redistrb))
        self._set(cm, "ktau_gl", 0); self._set(cm, "kend_gl", 10 ** 6); self._set(cm, "knode_gl", 0)
        self._set(I.module("casadimension"), "icycle", int(cfg.icycle))
        # non-per-tile members of the derived types
        self.S["soil"].f["zse"].a[...] = np.asarray(cfg.zse, np.float32)
        self.S["soil"].f["zshh"].a[...] = np.asarray(cfg.zshh, np.float32)
        self.S["bgc"].f["ratecp"].a[...] = np.asarray(cfg.ratecp, np.float32)
        self.S["bgc"].f["ratecs"].a[...] = np.asarray(cfg.ratecs, np.float32)
        self.push(all_fields=True)

    # ---- helpers -------------------------------------------------------------------------------------------------
    def _set(self, mod, name, value):
        ent = self.I.lookup_in_module(mod, name)
        if not isinstance(ent, Arr):
            raise FortranError(f"{mod.name}::{name} is not a variable ({ent!r})")
        ent.a[...] = value

    @staticmethod
    def _sets(struct: Struct, name: str, value: str):
        ent = struct.f[name]
        assert isinstance(ent, Str), name
        ent.s = value

    @staticmethod
    def _setl(struct: Struct, name: str, value: bool):
        struct.f[name].a[...] = value

    def member(self, f):
        """ndarray of the derived-type member behind registry field f (None if it has no counterpart)"""
        if f.type == "scr":
            return self.scr[f.member]
        s = self.S.get(f.type)
        ent = s.f.get(f.member.lower()) if s is not None else None
        return ent.a if isinstance(ent, Arr) else None

    def push(self, all_fields=False, roles=(1,)):
        """tiles -> derived types (forcing every step; everything at start-up)"""
        for f in self.fields:
            if not all_fields and f.role not in roles:
                continue
            dst = self.member(f)
            if dst is None:
                raise FortranError(f"registry field {f.name} has no member in the reference type {TYPE_OF.get(f.type)}")
            src = _to_fortran(self.tiles[f.name], f.n1, f.n2)
            if dst.shape != src.shape:
                raise FortranError(f"{f.name}: reference member has shape {dst.shape}, registry says {src.shape}")
            if dst.dtype != src.dtype:
                raise FortranError(f"{f.name}: reference member is {dst.dtype}, registry says {src.dtype}")
            dst[...] = src

    def pull(self):
        """derived types -> tiles (every state and diagnostic field)"""
        for f in self.fields:
            if f.role in (1, 2):
                continue
            src = self.member(f)
            _to_fortran(self.tiles[f.name], f.n1, f.n2)[...] = src

    def cbm(self, ktau: int, dels: float):
        S, T = self.S, self.tiles
        self.push(roles=(1,))
        if self.cfg.caller_duties:
            S["canopy"].f["oldcansto"].a[...] = S["canopy"].f["cansto"].a            # cable_serial.F90:573
        else:
            S["canopy"].f["oldcansto"].a[...] = T["canopy_oldcansto"][0]
        if self.cfg.met_tv_is_tk:
            S["met"].f["tvair"].a[...] = S["met"].f["tk"].a                           # cable_input.F90:2679-2680
            S["met"].f["tvrad"].a[...] = S["met"].f["tk"].a
        else:
            S["met"].f["tvair"].a[...] = T["met_tvair"][0]
        self._set(self.I.module("cable_common_module"), "ktau_gl", int(ktau))
        self.I.call("cable_cbm_module", "cbm", np.int32(ktau), np.float32(dels), S["air"], S["bgc"], S["canopy"], S["met"],
                    S["bal"], S["rad"], S["rough"], S["soil"], S["ssnow"], S["sum_flux"], S["veg"], S["climate"],
                    self.scr["xk"], self.scr["c1"], self.scr["rhoch"])
        self.pull()
