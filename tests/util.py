"""Shared helpers of the test-suite: field comparison under the north-star tolerances and the reference's
own closure invariants (src/offline/cable_checks.F90:472-618)."""
from __future__ import annotations

import numpy as np

from cable_b200 import lib, synth
from cable_b200.registry import FIELDS, ROLE, FLAG

RTOL_F32 = 1e-4      # BASELINE.json north_star: per-step fluxes/states within 1e-4 relative for fp32 fields
RTOL_F64 = 1e-6      #                            and 1e-6 relative for fp64 fields
DELS = 10800.0


def output_fields():
    return [f for f in FIELDS if not (f.flags & FLAG["HOSTONLY"]) and f.role in (ROLE["STATE"], ROLE["DIAG"])]


def field_errors(a: np.ndarray, b: np.ndarray, dtype) -> tuple[float, float, np.ndarray]:
    """Relative error per element with a field-scale floor: |a-b| / max(|a|, |b|, 1e-3*max|a|).
    Fluxes that cancel to ~0 (balances, night-time assimilation) are judged against the field's own scale,
    which is how a per-field relative tolerance is meaningful for them."""
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    floor = 1e-3 * max(float(np.abs(a64).max()), 1e-30)
    rel = np.abs(a64 - b64) / np.maximum(np.maximum(np.abs(a64), np.abs(b64)), floor)
    tol = RTOL_F32 if dtype == np.float32 else RTOL_F64
    return float(rel.max()), tol, rel


def compare_tiles(ref: dict, got: dict, fields=None):
    """-> {name: (max_rel, tol, fraction_of_elements_within_tol)}; raises on non-finite output."""
    out = {}
    for f in (fields or output_fields()):
        b = got[f.name]
        if not np.all(np.isfinite(b)):
            raise AssertionError(f"{f.name}: non-finite values on the device path")
        mx, tol, rel = field_errors(ref[f.name], b, f.dtype)
        out[f.name] = (mx, tol, float(np.mean(rel <= tol)))
    return out


def make_case(nland=200, nap=5, seed=synth.SEED, cfg=None, start_doy=120, site_lat=None, single_pft=None):
    cfg = cfg or lib.default_cfg()
    grid = synth.make_grid(nland, nap, seed=seed, site_lat=site_lat)
    tiles = synth.make_tiles(grid, cfg, single_pft=single_pft)
    forcing = synth.Forcing(grid, tiles, DELS, start_doy=start_doy)
    return cfg, grid, tiles, forcing


def water_balance(T, dels, wbtot_prev):
    """bal%wbal of mass_balance (cable_checks.F90:510-523) in mm per step; cbm's runoff is per second
    (the driver multiplies by dels afterwards, cable_serial.F90:602-605)."""
    delwb = T["ssnow_wbtot"][0] - wbtot_prev
    evap = (T["canopy_fevw"][0] + T["canopy_fevc"][0] + T["canopy_fes"][0] / T["ssnow_cls"][0]) * dels / T["air_rlam"][0]
    return (T["met_precip"][0] - T["canopy_delwc"][0] - T["ssnow_snowd"][0] + T["ssnow_osnowd"][0]
            - T["ssnow_runoff"][0] * dels - evap - delwb)


def energy_balances(T):
    """Radbal, EbalSoil, Ebalveg, Ebal of energy_balance (cable_checks.F90:585-604), W/m2."""
    sb, fsd = 5.67e-8, T["met_fsd"]
    q = T["rad_qcan"]
    radbal = (fsd[0] + fsd[1] + T["met_fld"][0] - T["rad_albedo"][0] * fsd[0] - T["rad_albedo"][1] * fsd[1]
              - sb * T["rad_transd"][0] * T["ssnow_otss"][0].astype(np.float64) ** 4
              - sb * (1 - T["rad_transd"][0]) * T["canopy_tv"][0].astype(np.float64) ** 4 - T["canopy_fnv"][0] - T["canopy_fns"][0])
    ebalsoil = T["canopy_fns"][0] - T["canopy_fes"][0] - T["canopy_fhs"][0] - T["canopy_ga"][0]
    ebalveg = T["canopy_fnv"][0] - T["canopy_fev"][0] - T["canopy_fhv"][0]
    ebal = (q[0] + q[1] + q[2] + q[3] + T["rad_qssabs"][0] + T["met_fld"][0]
            - sb * T["canopy_tv"][0].astype(np.float64) ** 4 * (1 - T["rad_transd"][0]) - T["rad_flws"][0] * T["rad_transd"][0]
            - T["canopy_fev"][0] - T["canopy_fes"][0] - T["canopy_fh"][0] - T["canopy_ga"][0])
    return radbal, ebalsoil, ebalveg, ebal


def ragged_case(nland=200, seed=7, **kw):
    """A case whose land points keep 1..5 active patches, as landpt(:)%cstart/cend are in the reference (nap = number of
    active patches of a point, src/offline/cable_input.F90:158-160): tiles dropped from the tail of each point of a uniform
    case, patch fractions renormalised per point.  -> cfg, grid, tiles, forcing, idx (kept tiles of the uniform case, for the
    per-tile LAI the forcing generator returns)."""
    import dataclasses
    cfg, grid, T, F = make_case(nland, **kw)
    rng = np.random.default_rng(seed)
    keep_n = rng.integers(1, grid.nap + 1, grid.nland)
    idx = np.flatnonzero((np.arange(grid.mp) - grid.cstart[grid.tile2land]) < keep_n[grid.tile2land])
    T2 = {n: np.ascontiguousarray(a[:, idx]) for n, a in T.items()}
    cend = (np.cumsum(keep_n) - 1).astype(np.int32)
    cstart = (cend - keep_n + 1).astype(np.int32)
    tl = grid.tile2land[idx].astype(np.int32)
    pf = grid.patchfrac[idx].astype(np.float64)
    tot = np.zeros(grid.nland)
    np.add.at(tot, tl, pf)
    g2 = dataclasses.replace(grid, mp=int(idx.size), tile2land=tl, cstart=cstart, cend=cend, patchfrac=(pf / tot[tl]).astype(np.float32))
    return cfg, g2, T2, F, idx
