"""run_casa.py -- drive the reference's own SUBROUTINE bgcdriver / biogeochem (src/science/casa-cnp/bgcdriver.F90:7,
biogeochem_casa.F90:7) through the Fortran interpreter of this package, on registry-layout arrays
(TEST INFRASTRUCTURE ONLY; needs /root/reference).  The CASA derived types are allocated by the reference's own
alloc_casavariable / alloc_phenvariable; POP is absent (cable_user%CALL_POP = .FALSE.), its modules are stubs."""
from __future__ import annotations

import numpy as np

from .finterp import Arr, Interp, Struct
from .run_cbm import STUBS, _to_fortran

REF_SRC = "/root/reference/src"
CASA_STUBS = STUBS + ("popmodule", "pop_types", "cable_phenology_module", "landuse_constant")
TYPE_OF = {"casabiome": "casa_biome", "casapool": "casa_pool", "casaflux": "casa_flux", "casamet": "casa_met", "casabal": "casa_balance"}


class FortranCasa:
    def __init__(self, tiles: dict, casa: dict, casa_fields, cfg, silt, clay, src_root=REF_SRC):
        self.I = I = Interp(src_root, stub_modules=CASA_STUBS)
        I.tolerate_stubs_in.add("alloc_canopy_type")
        self.tiles, self.casa, self.fields, self.cfg = tiles, casa, casa_fields, cfg
        mp = self.mp = int(tiles["met_tk"].shape[-1])
        dt = I.module("cable_def_types_mod")
        for n, v in (("mp", mp), ("mvtype", int(cfg.mvtype)), ("mstype", 9), ("mland", mp)):
            I.lookup_in_module(dt, n).a[...] = v
        cv = I.module("casavariable")
        self.S = S = {}
        for short, t in TYPE_OF.items():
            S[short] = I.new_struct(I.lookup_in_module(cv, "$type:" + t))
        I.call("casavariable", "alloc_casavariable", S["casabiome"], S["casapool"], S["casaflux"], S["casamet"], S["casabal"], np.int32(mp))
        S["phen"] = I.new_struct(I.lookup_in_module(I.module("phenvariable"), "$type:phen_variable"))
        I.call("phenvariable", "alloc_phenvariable", S["phen"], np.int32(mp))
        for short, tn, al in (("veg", "veg_parameter_type", "alloc_veg_parameter_type"), ("soil", "soil_parameter_type", "alloc_soil_parameter_type"),
                              ("met", "met_type", "alloc_met_type"), ("ssnow", "soil_snow_type", "alloc_soil_snow_type"),
                              ("canopy", "canopy_type", "alloc_canopy_type"), ("bgc", "bgc_pool_type", "alloc_bgc_pool_type"),
                              ("sum_flux", "sum_flux_type", "alloc_sum_flux_type")):
            s = I.new_struct(I.lookup_in_module(dt, "$type:" + tn))
            I.call("cable_def_types_mod", al, s, np.int32(mp))
            S[short] = s
        S["climate"] = I.new_struct(I.lookup_in_module(I.module("cable_climate_type_mod"), "$type:climate_type"))
        q = S["climate"].f["qtemp_max_last_year"]; q.a = np.zeros(mp, np.float32, order="F"); q.lb = (1,)
        S["pop"] = Struct(None)
        cm = I.module("cable_common_module")
        user = I.lookup_in_module(cm, "cable_user")
        user.f["call_climate"].a[...] = bool(cfg.call_climate); user.f["call_pop"].a[...] = False
        user.f["phenology_switch"].s = "MODIS"; user.f["l_limit_labile"].a[...] = bool(cfg.l_limit_labile)
        user.f["srf"].a[...] = False; user.f["mettype"].s = "site"
        I.lookup_in_module(I.module("casadimension"), "icycle").a[...] = int(cfg.icycle)
        S["veg"].f["iveg"].a[...] = tiles["veg_iveg"][0]
        S["veg"].f["froot"].a[...] = tiles["veg_froot"].T
        for n in ("sfc", "swilt", "ssat"):
            S["soil"].f[n].a[...] = tiles["soil_" + n][0]
        S["soil"].f["silt"].a[...] = silt; S["soil"].f["clay"].a[...] = clay
        if "climate_qtemp_max_last_year" in tiles:
            q.a[...] = tiles["climate_qtemp_max_last_year"][0]
        self.push()

    def member(self, f):
        ent = self.S[f.type].f.get(f.member)
        return ent.a if isinstance(ent, Arr) else None

    def push(self):
        for f in self.fields:
            dst = self.member(f)
            src = _to_fortran(self.casa[f.name], f.n1, f.n2)
            assert dst is not None and dst.shape == src.shape and dst.dtype == src.dtype, (f.name, None if dst is None else (dst.shape, dst.dtype), src.shape, src.dtype)
            dst[...] = src

    def pull(self):
        for f in self.fields:
            if f.key == 0:
                _to_fortran(self.casa[f.name], f.n1, f.n2)[...] = self.member(f)

    def bgcdriver(self, ktau, kstart, kend, dels, ktauday, idoy, loy=365):
        """the cbm-side inputs are taken from self.tiles (met_tk, ssnow_tgg, ssnow_wb, canopy_fpn, canopy_frday)"""
        S, T = self.S, self.tiles
        S["met"].f["tk"].a[...] = T["met_tk"][0]; S["ssnow"].f["tgg"].a[...] = T["ssnow_tgg"].T; S["ssnow"].f["wb"].a[...] = T["ssnow_wb"].T
        S["canopy"].f["fpn"].a[...] = T["canopy_fpn"][0]; S["canopy"].f["frday"].a[...] = T["canopy_frday"][0]
        self.I.call("bgcdriver_mod", "bgcdriver", np.int32(ktau), np.int32(kstart), np.int32(kend), np.float32(dels), S["met"], S["ssnow"],
                    S["canopy"], S["veg"], S["soil"], S["climate"], S["casabiome"], S["casapool"], S["casaflux"], S["casamet"], S["casabal"],
                    S["phen"], S["pop"], np.bool_(False), np.bool_(False), np.int32(ktauday), np.int32(idoy), np.int32(loy),
                    np.bool_(False), np.bool_(False), np.int32(self.cfg.lalloc))
        self.pull()

    SUMCFLUX_CANOPY = ("frp", "frs", "frpw", "frpr", "fnpp", "fgpp", "fra", "fnee")
    SUMCFLUX_SUMS = ("sumpn", "sumrp", "sumrpw", "sumrpr", "sumrs", "sumrd", "dsumpn", "dsumrp", "dsumrd")

    def sumcflux(self, ktau, kstart, kend, dels):
        """CALL sumcflux (casa_sumcflux.F90:37; call site cable_serial.F90:713) after bgcdriver: canopy%fpn / frday are the ones
        bgcdriver was given; canopy%frp / frs / frpw / frpr come from self.tiles when icycle == 0 (cbm's simple carbon)"""
        S, T = self.S, self.tiles
        for n in ("frp", "frs", "frpw", "frpr"):
            S["canopy"].f[n].a[...] = T["canopy_" + n][0]
        self.I.call("sumcflux_mod", "sumcflux", np.int32(ktau), np.int32(kstart), np.int32(kend), np.float32(dels), S["bgc"], S["canopy"],
                    S["soil"], S["ssnow"], S["sum_flux"], S["veg"], S["met"], S["casaflux"], np.bool_(False))
        out = {"canopy_" + n: S["canopy"].f[n].a.copy() for n in self.SUMCFLUX_CANOPY}
        out.update({"sum_flux_" + n: S["sum_flux"].f[n].a.copy() for n in self.SUMCFLUX_SUMS})
        return out
