"""Host-side plumbing of the CASA-CNP daily step (include/cable_b200.h, second half): registry of the casa_* / phen members
(include/cable_b200_casa_fields.def, generated from the reference types), ctypes binding, and a synthetic biome / pool
generator for tests and the config-5 bench leg (the reference's pftlookup.csv and pool files are external CABLE-AUX data).

Reference interface: CALL bgcdriver(...) src/science/casa-cnp/bgcdriver.F90:7, call site src/offline/cable_serial.F90:621."""
from __future__ import annotations

import ctypes as C
import os
import re
from dataclasses import dataclass

import numpy as np

from . import lib as _lib

_DEF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "cable_b200_casa_fields.def")
DT = {"double": np.float64, "float": np.float32, "int": np.int32}
MSO = 12


@dataclass(frozen=True)
class CasaField:
    id: int
    name: str
    type: str
    member: str
    dtype: type
    n1: int
    n2: int
    key: int      # 0 per tile, 1 per vegetation type, 2 per soil order

    @property
    def ncomp(self):
        return self.n1 * self.n2


def load_fields(path=_DEF):
    rx = re.compile(r"^CASA_FA\(\s*(\w+)\s*,\s*(\w+)\s*,\s*(\w+)\s*,\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*\)")
    out = []
    with open(path) as fh:
        for line in fh:
            m = rx.match(line)
            if m:
                t, mem, ct, n1, n2, key = m.groups()
                out.append(CasaField(len(out), f"{t}_{mem}", t, mem, DT[ct], int(n1), int(n2), int(key)))
    return out


FIELDS = load_fields()
BY_NAME = {f.name: f for f in FIELDS}


class CasaCfg(C.Structure):
    _fields_ = [("struct_bytes", C.c_int), ("icycle", C.c_int), ("lalloc", C.c_int), ("call_climate", C.c_int),
                ("l_limit_labile", C.c_int), ("mvtype", C.c_int), ("call_pop", C.c_int), ("srf", C.c_int),
                ("phenology_climate", C.c_int), ("l_landuse", C.c_int)]


EXPORTS = ["cable_b200_casa_default_cfg", "cable_b200_casa_nfields", "cable_b200_casa_field_id", "cable_b200_casa_field_info",
           "cable_b200_casa_init", "cable_b200_casa_bind", "cable_b200_casa_upload", "cable_b200_casa_download",
           "cable_b200_bgcdriver", "cable_b200_casa_biogeochem", "cable_b200_casa_feedback"]


def _bind_lib():
    L = _lib.load()
    if getattr(L, "_casa_bound", False):
        return L
    H = C.c_void_p
    L.cable_b200_casa_default_cfg.argtypes = [C.POINTER(CasaCfg)]; L.cable_b200_casa_default_cfg.restype = None
    L.cable_b200_casa_field_id.argtypes = [C.c_char_p]
    L.cable_b200_casa_field_info.argtypes = [C.c_int, C.POINTER(_lib.FieldInfo), C.POINTER(C.c_int)]
    L.cable_b200_casa_init.argtypes = [H, C.POINTER(CasaCfg)]
    L.cable_b200_casa_bind.argtypes = [H, C.c_char_p, C.c_void_p]
    L.cable_b200_casa_upload.argtypes = [H]; L.cable_b200_casa_download.argtypes = [H]
    L.cable_b200_bgcdriver.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    L.cable_b200_casa_biogeochem.argtypes = [H, C.c_int]
    L.cable_b200_casa_feedback.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int]
    L._casa_bound = True
    return L


def default_cfg() -> CasaCfg:
    cfg = CasaCfg()
    _bind_lib().cable_b200_casa_default_cfg(C.byref(cfg))
    return cfg


def alloc(mp: int, mvtype: int = 17) -> dict:
    """zero arrays in registry layout: (ncomp, lead) C-order == Fortran (lead, n1, n2)"""
    out = {}
    for f in FIELDS:
        lead = mp if f.key == 0 else mvtype if f.key == 1 else MSO
        out[f.name] = np.zeros((f.ncomp, lead), f.dtype)
    return out


class Casa:
    """CASA-CNP on an existing CableB200 handle (device-resident next to the cbm state)."""

    def __init__(self, handle, cfg: CasaCfg):
        self.L = _bind_lib(); self.h = handle; self.cfg = cfg
        _lib.check(self.L.cable_b200_casa_init(handle._h, C.byref(cfg)))
        self._keep = {}

    def bind(self, arrays: dict, silt=None, clay=None):
        for f in FIELDS:
            a = arrays.get(f.name)
            if a is None:
                continue
            assert a.dtype == f.dtype and a.flags["C_CONTIGUOUS"], f.name
            _lib.check(self.L.cable_b200_casa_bind(self.h._h, f.name.encode(), a.ctypes.data_as(C.c_void_p)))
            self._keep[f.name] = a
        for nm, a in (("soil_silt", silt), ("soil_clay", clay)):
            if a is not None:
                a = np.ascontiguousarray(a, np.float32); self._keep[nm] = a
                _lib.check(self.L.cable_b200_casa_bind(self.h._h, nm.encode(), a.ctypes.data_as(C.c_void_p)))

    def upload(self):
        _lib.check(self.L.cable_b200_casa_upload(self.h._h))

    def download(self):
        _lib.check(self.L.cable_b200_casa_download(self.h._h))

    def bgcdriver(self, ktau, kstart, kend, dels, ktauday, idoy, loy=365):
        _lib.check(self.L.cable_b200_bgcdriver(self.h._h, int(ktau), int(kstart), int(kend), float(dels), int(ktauday), int(idoy), int(loy)))

    def feedback(self, slot=0, vcmax=True, lai=False, walker2014=False):
        """casa_feedback / l_laiFeedbk before the step that reads forcing slot `slot` (cable_serial.F90:587-590)"""
        _lib.check(self.L.cable_b200_casa_feedback(self.h._h, int(slot), int(vcmax), int(lai), int(walker2014)))

    def biogeochem(self, idoy):
        _lib.check(self.L.cable_b200_casa_biogeochem(self.h._h, int(idoy)))


# ------------------------------------------------------------------------------------------------------------------
# synthetic biome parameters and initial pools (plausible magnitudes; one row per CABLE vegetation type)
WOODY = np.array([1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0], bool)        # types 1-5, 12, 13 carry wood


def synth_casa(grid, tiles, cfg: CasaCfg, seed: int = 11) -> dict:
    mp, mv = grid.mp, cfg.mvtype
    rng = np.random.default_rng([seed, 5])
    A = alloc(mp, mv)
    iv = tiles["veg_iveg"][0] - 1
    t = np.arange(mv, dtype=np.float64)
    woody = WOODY[:mv]

    def tab(name, vals):
        A["casabiome_" + name][...] = np.asarray(vals, np.float64).reshape(A["casabiome_" + name].shape)
    tab("ivt2", np.where(woody, 3, 1).astype(np.int32)); A["casabiome_ivt2"] = A["casabiome_ivt2"].astype(np.int32)
    tab("xkleafcoldmax", 0.2 + 0.02 * t); tab("xkleafcoldexp", 3.0 + 0 * t); tab("xkleafdrymax", 0.05 + 0.005 * t); tab("xkleafdryexp", 3.0 + 0 * t)
    tab("glaimax", np.where(woody, 7.0, 5.0) - 0.05 * t); tab("glaimin", 0.1 + 0.01 * t); tab("sla", 0.010 + 0.0012 * t)
    tab("ratiofrootleaf", 1.0 + 0.1 * t); tab("kroot", 5.5 + 0 * t); tab("krootlen", 14.9e3 + 0 * t); tab("rootdepth", 1.5 + 0 * t); tab("kuptake", 2.0 + 0 * t)
    tab("kminn", 2.0 + 0.05 * t); tab("kuplabp", 0.5 + 0.01 * t); tab("kclabrate", 0.05 / (0.15 + 0.01 * t) / 365.0); tab("xnpmax", 1.5 + 0.01 * t)
    tab("q10soil", 1.72 + 0.01 * t); tab("xkoptlitter", 0.4 + 0.01 * t); tab("xkoptsoil", 0.33 + 0.01 * t)
    so = np.arange(MSO, dtype=np.float64)
    tab("xkplab", 0.5e-4 * (1 + so)); tab("xkpsorb", 1e-4 / (1 + so)); tab("xkpocc", 1e-5 * (1 + 0.1 * so))
    tab("prodptase", 0.5 + 0.1 * t); tab("costnpup", 25.0 + 2.0 * t); tab("maxfinelitter", 1500.0 + 10 * t); tab("maxcwd", 1500.0 + 20 * t)
    tab("nintercept", 6.3 + 0 * t); tab("nslope", 18.2 + 0 * t)
    age = np.stack([1.0 + 0.2 * t, np.where(woody, 40.0 + t, 1.0), 3.0 + 0.3 * t])                      # leaf, wood, froot (years)
    tab("plantrate", 1.0 / (age * 365.0)); tab("rmplant", np.stack([0.1 + 0 * t, 0.3 + 0.01 * t, 0.8 + 0.02 * t]) / 365.0)
    fa = np.stack([0.35 - 0.005 * t, np.where(woody, 0.30, 0.0), 0.35 + 0.005 * t]); tab("fracnpptop", fa / fa.sum(0))
    tab("fraclignin", np.stack([0.2 + 0 * t, 0.4 + 0 * t, 0.2 + 0 * t])); tab("fraclabile", np.stack([0.6 + 0 * t, 0.0 * t, 0.6 + 0 * t]))
    cn = np.stack([40.0 + t, 200.0 + 5 * t, 60.0 + t])
    tab("rationcplantmin", 1.0 / (cn * 1.2)); tab("rationcplantmax", 1.0 / (cn * 0.8))
    tab("rationpplantmin", np.stack([10.0 + 0 * t, 10.0 + 0 * t, 10.0 + 0 * t])); tab("rationpplantmax", np.stack([20.0 + 0 * t, 20.0 + 0 * t, 20.0 + 0 * t]))
    tab("fracligninplant", np.stack([0.25 - 0.003 * t, 0.4 + 0 * t, 0.25 + 0 * t])); tab("ftransnptol", np.stack([0.5 + 0 * t, 0.95 + 0 * t, 0.9 + 0 * t]))
    tab("ftranspptol", np.stack([0.5 + 0 * t, 0.95 + 0 * t, 0.9 + 0 * t]))
    tab("litterrate", 1.0 / (np.stack([0.04 + 0 * t, 0.23 + 0.01 * t, 0.82 + 0.02 * t]) * 365.0))
    tab("ratiopcplantmin", A["casabiome_rationcplantmin"] / 20.0); tab("ratiopcplantmax", A["casabiome_rationcplantmax"] / 10.0)
    tab("soilrate", 1.0 / (np.stack([0.07 + 0 * t, 3.0 + 0.1 * t, 100.0 + t]) * 365.0))
    A["phen_tkshed"][0] = 268.0 + 0.3 * t
    # per tile
    ice = tiles["veg_iveg"][0] >= 16
    A["casamet_iveg2"][0] = np.where(ice, 0, np.where(woody[np.minimum(iv, mv - 1)], 3, 1))
    A["casamet_lnonwood"][0] = np.where(woody[np.minimum(iv, mv - 1)] & ~ice, 0, 1)
    A["casamet_isorder"][0] = rng.integers(1, MSO + 1, mp)
    A["casamet_glai"][0] = np.where(ice, 0.0, rng.uniform(0.3, 4.0, mp))
    A["phen_phase"][0] = rng.integers(0, 4, mp)
    base = rng.integers(60, 140, mp)
    for k, off in enumerate((0, 30, 150, 200)):
        A["phen_doyphase"][k] = (base + off - 1) % 365 + 1
    cpl = rng.uniform(40, 300, (3, mp)) * np.array([[1.0], [20.0], [1.0]])
    cpl[1] *= (A["casamet_lnonwood"][0] == 0)
    cpl[:, ice] = 0.0
    A["casapool_cplant"][...] = cpl
    A["casapool_clitter"][...] = rng.uniform(20, 400, (3, mp)) * ~ice
    A["casapool_csoil"][...] = rng.uniform(100, 5000, (3, mp)) * np.array([[0.1], [1.0], [1.0]]) * ~ice
    A["casapool_clabile"][0] = rng.uniform(0, 5, mp) * ~ice
    ncp = 0.5 * (A["casabiome_rationcplantmin"][:, np.minimum(iv, mv - 1)] + A["casabiome_rationcplantmax"][:, np.minimum(iv, mv - 1)])
    A["casapool_rationcplant"][...] = ncp; A["casapool_nplant"][...] = ncp * cpl
    A["casapool_rationclitter"][...] = 1.0 / rng.uniform(40, 120, (3, mp)); A["casapool_nlitter"][...] = A["casapool_rationclitter"] * A["casapool_clitter"]
    A["casapool_rationcsoil"][...] = 1.0 / np.array([[8.0], [16.0], [16.0]]) * np.ones((3, mp))
    A["casapool_rationcsoilmin"][...] = A["casapool_rationcsoil"] * 0.8; A["casapool_rationcsoilmax"][...] = A["casapool_rationcsoil"] * 1.3
    A["casapool_rationcsoilnew"][...] = A["casapool_rationcsoilmax"]; A["casapool_nsoil"][...] = A["casapool_rationcsoil"] * A["casapool_csoil"]
    A["casapool_nsoilmin"][0] = rng.uniform(0.3, 4.0, mp)
    A["casapool_ratiopcplant"][...] = ncp / 15.0; A["casapool_pplant"][...] = A["casapool_ratiopcplant"] * cpl
    A["casapool_ratiopclitter"][...] = A["casapool_rationclitter"] / 20.0; A["casapool_plitter"][...] = A["casapool_ratiopclitter"] * A["casapool_clitter"]
    A["casapool_ratiopcsoil"][...] = A["casapool_rationcsoil"] / 12.0; A["casapool_psoil"][...] = A["casapool_ratiopcsoil"] * A["casapool_csoil"]
    A["casapool_rationpplant"][...] = 15.0; A["casapool_rationplitter"][...] = 20.0; A["casapool_rationpsoil"][...] = 12.0
    A["casapool_psoillab"][0] = rng.uniform(0.2, 3.0, mp); A["casapool_psoilsorb"][0] = rng.uniform(10, 60, mp); A["casapool_psoilocc"][0] = rng.uniform(10, 60, mp)
    A["casaflux_frac_sapwood"][0] = 1.0; A["casaflux_sapwood_area"][0] = rng.uniform(0, 2e-3, mp)
    A["casaflux_nmindep"][0] = rng.uniform(1e-4, 3e-3, mp); A["casaflux_nminfix"][0] = rng.uniform(1e-4, 2e-3, mp)
    A["casaflux_fnminloss"][0] = 0.05; A["casaflux_fnminleach"][0] = 0.001
    A["casaflux_pdep"][0] = rng.uniform(1e-5, 1e-4, mp); A["casaflux_pwea"][0] = rng.uniform(1e-5, 1e-4, mp); A["casaflux_fpleach"][0] = 0.0005
    A["casaflux_psorbmax"][0] = rng.uniform(50, 150, mp); A["casaflux_kmlabp"][0] = rng.uniform(20, 80, mp)
    return A


def soil_texture(tiles):
    """soil%silt / soil%clay per tile from the soil type (cable_soilparm.nml); casa_coeffsoil reads them"""
    from .synth import SOIL
    ist = tiles["soil_isoilm"][0] - 1
    return SOIL["silt"][ist].astype(np.float32), SOIL["clay"][ist].astype(np.float32)
