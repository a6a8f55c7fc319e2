"""Seeded synthetic grid, parameters, initial state and met forcing for the cbm() path.

There is no forcing / gridinfo data in the reference tree (large blobs are absent), so the
benchmark shapes of BASELINE.json are reproduced synthetically:
  * PFT and soil-type parameter tables: src/offline/pft_params.nml, src/offline/cable_soilparm.nml
    (data, transcribed), expanded per tile the way `init_veg_from_vegin` does
    (src/offline/cable_parameters.F90:3277-3345) plus the derived soil parameters
    (:2236-2241, :1685-1691, :1828-1831).
  * default initial state: `write_default_params` (cable_parameters.F90:1187-1262) and the
    frozen-soil / glacier special initialisation (cbl_soilsnow_init_special.F90:34-77).
  * forcing per land point, replicated to its tiles; coszen by `sinbet`
    (src/science/radiation/cbl_sinbet.F90:12-28); snow/rain split at tfrz
    (src/offline/cable_input.F90:2666-2671); SW split 50/50 VIS/NIR (:1880-1883).
All values stay inside the reference's input ranges (src/offline/cable_checks.F90:53-201).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .registry import FIELDS, alloc_tiles

SEED = 20261017
TFRZ = np.float32(273.16)
ZSE = np.array([.022, .058, .154, .409, 1.085, 2.872], dtype=np.float32)   # cable_parameters.F90:1241


def _rep(spec):
    """Expand namelist repeat syntax already turned into python: list of (count, value) or values."""
    out = []
    for s in spec:
        if isinstance(s, tuple):
            out += [s[1]] * s[0]
        else:
            out.append(s)
    assert len(out) == 17, len(out)
    return np.array(out, dtype=np.float32)


# ---- src/offline/pft_params.nml (17 PFTs) -------------------------------------------------------
PFT = {
    "a1gs": _rep([(6, 9.0), 4.0, 9.0, 9.0, 4.0, (7, 9.0)]),
    "alpha": _rep([(6, 0.2), 0.05, 0.2, 0.2, 0.05, (7, 0.2)]),
    "canst1": _rep([(17, 0.1)]),
    "clitt": _rep([20., 6., 10., 13., 2., 2., 0.3, 0.3, 0., 0., 2., 2., (5, 0.)]),
    "cfrd": _rep([(6, 0.015), 0.025, 0.015, 0.015, 0.025, (7, 0.015)]),
    "conkc0": _rep([(17, 0.000302)]),
    "conko0": _rep([(17, 0.256)]),
    "convex": _rep([(6, 0.01), 0.8, 0.01, 0.01, 0.8, (7, 0.01)]),
    "cplant1": _rep([200., 300., 200., 300., 159., 250., 250., 250., 150., 150., 250., 1., 0.1, 0., 1., 1., 0.]),
    "cplant2": _rep([10217., 16833., 5967., 12000., 5000., (12, 0.)]),
    "cplant3": _rep([876., 1443., 511., 1029., 500., 500., 500., 500., 607., 607., 500., 1., 0.1, 0., 1., 1., 0.]),
    "csoil1": _rep([184., 303., 107., 216., 100., 275., 275., 275., 149., 149., 275., 1., 0.1, 1., 1., 1., 1.]),
    "csoil2": _rep([367., 606., 214., 432., 250., 314., 314., 314., 300., 300., 314., 1., 0.1, 1., 1., 1., 1.]),
    "d0gs": _rep([(17, 1500.)]),
    "ekc": _rep([(17, 59430.)]),
    "eko": _rep([(17, 36000.)]),
    "extkn": _rep([(17, 0.001)]),
    "frac4": _rep([(6, 0.), 1., 0., 0., 1., (7, 0.)]),
    "g0": _rep([(17, 0.)]),
    "g1": _rep([2.346064, 4.114762, 2.346064, 4.447321, 4.694803, 5.2485, 1.616178, 2.222156, 5.789377,
                1.616178, 5.2485, 5.2485, 0., 5.2485, 5.2485, 5.2485, 5.2485]),
    "gswmin": _rep([(6, 0.01), 0.04, 0.01, 0.01, 0.04, (7, 0.01)]),
    "hc": _rep([17., 35., 15.5, 20., 0.6, 0.567, 0.567, 0.567, 0.55, 0.55, 0.567, 0.2, 6.017, 0.2, 0.2, 0.2, 0.2]),
    "length": _rep([0.055, 0.1, 0.04, 0.15, 0.1, (6, 0.3), 0.03, 0.242, 0.03, 0.03, 0.03, 0.03]),
    "refl1": _rep([0.09, 0.09, 0.075, 0.09, 0.09, 0.11, 0.11, 0.075, 0.11, 0.11, 0.108, 0.055, 0.091, 0.238, 0.143, 0.143, 0.159]),
    "refl2": _rep([0.3, 0.29, 0.3, 0.29, 0.3, 0.34, 0.34, 0.32, 0.34, 0.34, 0.343, 0.19, 0.31, 0.457, 0.275, 0.275, 0.305]),
    "refl3": _rep([(17, 0.01)]),
    "taul1": _rep([0.09, 0.09, 0.075, 0.09, 0.09, 0.11, 0.11, 0.075, 0.11, 0.11, 0.075, 0.023, 0.059, 0.039, 0.023, 0.023, 0.026]),
    "taul2": _rep([0.3, 0.29, 0.3, 0.29, 0.3, 0.34, 0.34, 0.32, 0.34, 0.34, 0.146, 0.198, 0.163, 0.189, 0.113, 0.113, 0.113]),
    "taul3": _rep([(17, 0.01)]),
    "rootbeta": _rep([0.943, 0.962, 0.966, 0.961, 0.964, 0.943, 0.943, 0.943, 0.961, 0.961, 0.943, 0.975, (5, 0.961)]),
    "rp20": _rep([3., 0.6, 3., 2.2, 1., 1.5, 2.8, 2.5, 1.5, 1., 1.5, (6, 1.)]),
    "rs20": _rep([(11, 1.), 0., 1., 0., 0., 0., 0.]),
    "shelrb": _rep([(17, 2.)]),
    "vbeta": _rep([2., 2., 2., 2., 4., 4., 4., 4., 2., 2., 4., 4., 2., 4., 4., 4., 4.]),
    "vcmax": _rep([0.00004, 0.000055, 0.00004, 0.00006, 0.00004, 0.00006, 0.00001, 0.00004, 0.00008, 0.00008,
                   0.00006, 0.000017, 0.000001, 0.000017, 0.000017, 0.000017, 0.000017]),
    "vegcf": _rep([9., 14., 9., 8., 5., 7., 7., 5., 7., 1., 7., (6, 1.)]),
    "width": _rep([0.001, 0.05, 0.001, 0.08, 0.005, (6, 0.01), 0.003, 0.015, 0.001, 0.001, 0.001, 0.001]),
    "xfang": _rep([0.01, 0.1, 0.01, 0.25, 0.01, (6, -0.3), 0.1, (5, 0.)]),
}
# not in the namelist: a plausible peak LAI per PFT for the synthetic seasonal cycle
LAIMAX = np.array([4.5, 5.5, 3.5, 4.5, 1.5, 2.0, 2.5, 1.0, 3.0, 3.5, 2.5, 1.0, 1.0, 0.3, 0.3, 0., 0.], dtype=np.float32)

# ---- src/offline/cable_soilparm.nml (9 soil types) -------------------------------------------------
SOIL = {
    "bch": [4.2, 7.1, 11.4, 5.15, 10.4, 10.4, 7.12, 5.83, 7.1],
    "clay": [0.09, 0.3, 0.67, 0.2, 0.42, 0.48, 0.27, 0.17, 0.3],
    "css": [850] * 7 + [1920, 2100],
    "hyds": [0.000166, 0.000004, 0.000001, 0.000021, 0.000002, 0.000001, 0.000006, 0.0008, 0.000001],
    "rhosoil": [1600, 1600, 1381, 1373, 1476, 1521, 1373, 1537, 917],
    "sand": [0.83, 0.37, 0.16, 0.6, 0.52, 0.27, 0.58, 0.13, 0.37],
    "sfc": [0.143, 0.301, 0.367, 0.218, 0.31, 0.37, 0.255, 0.45, 0.301],
    "silt": [0.08, 0.33, 0.17, 0.2, 0.06, 0.25, 0.15, 0.7, 0.33],
    "ssat": [0.398, 0.479, 0.482, 0.443, 0.426, 0.482, 0.42, 0.451, 0.479],
    "sucs": [-0.106, -0.591, -0.405, -0.348, -0.153, -0.49, -0.299, -0.356, -0.153],
    "swilt": [0.072, 0.216, 0.286, 0.135, 0.219, 0.283, 0.175, 0.395, 0.216],
}
SOIL = {k: np.array(v, dtype=np.float32) for k, v in SOIL.items()}


@dataclass
class Grid:
    nland: int
    nap: int
    mp: int
    lat: np.ndarray        # (nland,) degrees north
    lon: np.ndarray
    elev: np.ndarray       # (nland,) m
    tile2land: np.ndarray  # (mp,) int32
    cstart: np.ndarray     # (nland,) first tile of land point (0-based)
    cend: np.ndarray       # (nland,) last tile (inclusive)
    patchfrac: np.ndarray  # (mp,) float32
    tmean: np.ndarray      # (nland,) annual-mean air temperature (K)
    tamp: np.ndarray       # (nland,) seasonal amplitude (K)
    seed: int


def make_grid(nland: int, nap: int = 5, seed: int = SEED, site_lat: float | None = None) -> Grid:
    rng = np.random.default_rng([seed, 1])
    if site_lat is not None:
        lat = np.full(nland, site_lat, dtype=np.float32)
    else:   # area-weighted latitude in [-56, 84]
        s0, s1 = np.sin(np.deg2rad(-56.0)), np.sin(np.deg2rad(84.0))
        lat = np.rad2deg(np.arcsin(rng.uniform(s0, s1, nland))).astype(np.float32)
    lon = rng.uniform(-180, 180, nland).astype(np.float32)
    # land points in raster order of the land mask (rows of `dlat` degrees from the north, west -> east inside
    # a row), the order in which the reference numbers land points from its gridinfo mask
    # (src/offline/cable_parameters.F90 countLandPoints / landpt): neighbours in memory are neighbours in space.
    if site_lat is None and nland > 1:
        dlat = 0.5
        order = np.lexsort((lon, -np.floor((lat + 90.0) / dlat)))
        lat, lon = lat[order], lon[order]
    # terrain height: smooth in space plus local roughness
    elev = (1250.0 + 900.0 * np.sin(np.deg2rad(3.0 * lon)) * np.cos(np.deg2rad(2.0 * lat))
            + rng.uniform(-350, 350, nland)).clip(0, 2500).astype(np.float32)
    mp = nland * nap
    tile2land = np.repeat(np.arange(nland, dtype=np.int32), nap)
    cstart = (np.arange(nland, dtype=np.int32) * nap)
    cend = cstart + nap - 1
    pf = rng.exponential(1.0, (nland, nap)).astype(np.float64)      # Dirichlet(1)
    pf = (pf / pf.sum(axis=1, keepdims=True)).astype(np.float32).reshape(mp)
    alat = np.abs(lat)
    tmean = (273.15 + 27.0 - 0.55 * alat - 0.0065 * elev * 0.3).astype(np.float32)
    tamp = (0.25 * alat).astype(np.float32)
    return Grid(nland, nap, mp, lat, lon, elev, tile2land, cstart, cend, pf, tmean, tamp, seed)


def _pick_pft(rng, lat_tile: np.ndarray) -> np.ndarray:
    """PFT per tile from {1..11,14} with a latitude-dependent prior; lakes/ice are added by the caller."""
    alat = np.abs(lat_tile)
    cands = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 14])
    w = np.ones((lat_tile.size, cands.size))
    tropic, temperate, boreal, polar = alat < 23, (alat >= 23) & (alat < 45), (alat >= 45) & (alat < 65), alat >= 65
    w[tropic] *= np.array([0.2, 4, 0.1, 1.5, 1.5, 1, 3, 0.05, 1, 2, 0.7, 1.5])
    w[temperate] *= np.array([2, 1, 0.5, 3, 2, 3, 1.5, 0.2, 3, 1, 0.7, 1.5])
    w[boreal] *= np.array([4, 0.05, 3, 1.5, 1.5, 2, 0.1, 2, 1, 0.05, 1.5, 0.7])
    w[polar] *= np.array([0.7, 0.01, 1, 0.2, 1.5, 1, 0.01, 5, 0.05, 0.01, 1, 2])
    w /= w.sum(axis=1, keepdims=True)
    u = rng.random(lat_tile.size)
    idx = (np.cumsum(w, axis=1) < u[:, None]).sum(axis=1).clip(0, cands.size - 1)
    return cands[idx].astype(np.int32)


def make_tiles(grid: Grid, cfg=None, single_pft: int | None = None) -> dict[str, np.ndarray]:
    """Parameters + initial prognostic state for every tile (all registry fields allocated)."""
    rng = np.random.default_rng([grid.seed, 2])
    mp, nland = grid.mp, grid.nland
    zse = np.array(cfg.zse[:], dtype=np.float32) if cfg is not None else ZSE
    frozen_limit = np.float32(cfg.frozen_limit if cfg is not None else 0.85)
    max_glacier = np.float32(cfg.max_glacier_snowd if cfg is not None else 1100.0)
    T = alloc_tiles(mp)
    lat_t = grid.lat[grid.tile2land]
    # surface types
    if single_pft is not None:
        iveg = np.full(mp, single_pft, dtype=np.int32)
        isoilm = np.full(mp, 2, dtype=np.int32)
    else:
        iveg = _pick_pft(rng, lat_t)
        iveg[rng.random(mp) < 0.01] = 16                                     # lakes
        isoilm = rng.integers(1, 9, mp).astype(np.int32)
        polar = np.abs(grid.lat) > 65
        ice_pt = polar & (rng.random(nland) < 0.08)
        if polar.any() and not ice_pt.any():                                  # always cover the glacier path
            ice_pt[np.flatnonzero(polar)[0]] = True
        ice_t = ice_pt[grid.tile2land]
        iveg[ice_t] = 17
        isoilm[ice_t] = 9
    iv = iveg - 1
    T["veg_iveg"][0] = iveg
    T["veg_meth"][0] = 1                                                     # cable_parameters.F90:1262
    for name in ("a1gs", "alpha", "canst1", "cfrd", "conkc0", "conko0", "convex", "d0gs", "ekc", "eko", "extkn",
                 "frac4", "g0", "g1", "gswmin", "hc", "rp20", "rs20", "shelrb", "vbeta", "vcmax", "vegcf", "xfang"):
        T["veg_" + name][0] = PFT[name][iv]
    T["veg_dleaf"][0] = np.sqrt(PFT["width"] * PFT["length"])[iv]            # cable_pft_params.F90
    T["veg_ejmax"][0] = np.float32(2.0) * T["veg_vcmax"][0]                   # cable_parameters.F90:1543
    for b, (r, tl) in enumerate((("refl1", "taul1"), ("refl2", "taul2"))):        # (mp,2) offline, cable_define_types.F90:1083
        T["veg_refl"][b] = PFT[r][iv]
        T["veg_taul"][b] = PFT[tl][iv]
    # froot from rootbeta (cable_parameters.F90:3335-3343), float32 like the reference
    rootbeta = PFT["rootbeta"][iv]
    froot = np.zeros((6, mp), dtype=np.float32)
    totdepth = np.float32(0.0)
    for k in range(5):
        totdepth = np.float32(totdepth + zse[k] * np.float32(100.0))
        froot[k] = np.minimum(np.float32(1.0), np.float32(1.0) - np.power(rootbeta, totdepth, dtype=np.float32))
    froot[5] = np.float32(1.0) - froot[4]
    for k in range(4, 0, -1):
        froot[k] = froot[k] - froot[k - 1]
    T["veg_froot"][:] = froot
    # soil
    T["soil_isoilm"][0] = isoilm
    st = isoilm - 1
    for name in ("css", "hyds", "rhosoil", "sfc", "ssat", "swilt"):
        T["soil_" + name][0] = SOIL[name][st]
    bch, sucs = SOIL["bch"][st], SOIL["sucs"][st]
    T["soil_hsbh"][0] = T["soil_hyds"][0] * np.abs(sucs) * bch                # cable_parameters.F90:2236
    nint_bch = np.floor(bch + np.float32(0.5)).astype(np.float32)            # NINT
    T["soil_ibp2"][0] = nint_bch + 2
    T["soil_i2bp3"][0] = 2 * nint_bch + 3
    T["soil_pwb_min"][0] = np.power(T["soil_swilt"][0] / T["soil_ssat"][0], T["soil_ibp2"][0], dtype=np.float32).astype(np.float64)
    T["soil_cnsd"][0] = (SOIL["sand"][st] * np.float32(0.3) + SOIL["clay"][st] * np.float32(0.25)
                         + SOIL["silt"][st] * np.float32(0.265)).astype(np.float64)
    vis = rng.uniform(0.08, 0.35, nland).astype(np.float32)[grid.tile2land]
    T["soil_albsoil"][0] = vis
    T["soil_albsoil"][1] = np.minimum(np.float32(2.0) * vis, np.float32(0.9))
    T["soil_albsoil"][2] = 0.05
    for k in range(6):                                                       # cable_parameters.F90:1685-1691
        T["soil_swilt_vec"][k] = T["soil_swilt"][0].astype(np.float64)
        T["soil_sfc_vec"][k] = T["soil_sfc"][0].astype(np.float64)
        T["soil_ssat_vec"][k] = T["soil_ssat"][0].astype(np.float64)
        T["soil_zse_vec"][k] = np.float64(zse[k])
        # soil-type-table path of the reference (cable_parameters.F90:1526-1529); cnsd_vec as the spread of cnsd
        T["soil_sand_vec"][k] = SOIL["sand"][st].astype(np.float64)
        T["soil_cnsd_vec"][k] = T["soil_cnsd"][0]
        T["soil_watr"][k] = np.float64(np.float32(0.01))
    T["veg_clitt"][0] = PFT["clitt"][iv].astype(np.float64)                  # pft_params.nml vegin%clitt
    T["canopy_us"][0] = 0.1                                                  # cable_parameters.F90 write_default_params
    # climate%qtemp_max_last_year (deg C, mean of the warmest quarter): only read under call_climate
    T["climate_qtemp_max_last_year"][0] = (grid.tmean + np.float32(0.7) * grid.tamp - np.float32(273.15))[grid.tile2land]
    T["rough_za_uv"][0] = 40.0
    T["rough_za_tq"][0] = 40.0
    # ---- initial state (write_default_params) ----
    tg0 = grid.tmean[grid.tile2land]
    T["ssnow_tgg"][:] = tg0[None, :]
    T["ssnow_tggsn"][:] = TFRZ
    T["ssnow_ssdn"][:] = 120.0
    T["ssnow_ssdnn"][0] = 120.0
    T["ssnow_sconds"][:] = 0.06
    T["ssnow_rtsoil"][0] = 100.0
    T["ssnow_t_snwlr"][0] = 0.05
    wb0 = (np.float32(0.5) * (T["soil_sfc"][0] + T["soil_swilt"][0])).astype(np.float64)
    T["ssnow_wb"][:] = wb0[None, :]
    # spec_init_soil_snow (cbl_soilsnow_init_special.F90:40-72)
    ssat64 = T["soil_ssat"][0].astype(np.float64)
    wb = T["ssnow_wb"]
    for k in range(6):
        wb[k] = np.minimum(ssat64, np.maximum(wb[k].astype(np.float32), T["soil_swilt"][0]).astype(np.float64))
    wb[3] = np.minimum(ssat64, np.maximum(wb[3].astype(np.float32), np.float32(0.5) * (T["soil_sfc"][0] + T["soil_swilt"][0])))
    wb[4] = np.minimum(ssat64, np.maximum(wb[4].astype(np.float32), np.float32(0.8) * T["soil_sfc"][0]))
    wb[5] = np.minimum(ssat64, np.maximum(wb[5].astype(np.float32), T["soil_sfc"][0]))
    wbice = T["ssnow_wbice"]
    for k in range(6):
        cold = T["ssnow_tgg"][k] <= TFRZ
        wbice[k][cold] = 0.5 * wb[k][cold]
        colder = T["ssnow_tgg"][k] < TFRZ
        wbice[k][colder] = np.float64(frozen_limit) * wb[k][colder]
    ice = isoilm == 9
    T["ssnow_snowd"][0][ice] = max_glacier
    T["ssnow_tgg"][0][ice] -= np.float32(1.0)
    for k in range(6):
        wb[k][ice] = (np.float32(0.95) * T["soil_ssat"][0][ice]).astype(np.float64)
        wbice[k][ice] = np.float64(frozen_limit) * wb[k][ice]
    T["ssnow_wbliq"][:] = wb - wbice
    T["ssnow_tss"][0] = T["ssnow_tgg"][0]
    T["ssnow_otss"][0] = T["ssnow_tss"][0]
    # owetfac (cable_parameters.F90:2245-2254)
    ow = np.clip((wb[0].astype(np.float32) - T["soil_swilt"][0]) / (T["soil_sfc"][0] - T["soil_swilt"][0]), 0.0, 1.0).astype(np.float32)
    has_ice = wbice[0] > 0
    tmp2 = (wbice[0] / np.where(wb[0] > 0, wb[0], 1.0)).astype(np.float32)
    ow[has_ice] = ow[has_ice] * (np.float32(1.0) - tmp2[has_ice]) ** 2
    T["ssnow_owetfac"][0] = ow
    T["bgc_cplant"][0], T["bgc_cplant"][1], T["bgc_cplant"][2] = PFT["cplant1"][iv], PFT["cplant2"][iv], PFT["cplant3"][iv]
    T["bgc_csoil"][0], T["bgc_csoil"][1] = PFT["csoil1"][iv], PFT["csoil2"][iv]
    return T


def sinbet(doy, xslat, hod):
    """cbl_sinbet.F90:22-26 in float32."""
    f = np.float32
    pi, pi180 = f(3.1415927), f(3.1415927) / f(180.0)
    sindec = -np.sin(f(23.45) * pi180, dtype=np.float32) * np.cos(f(2.) * pi * (f(doy) + f(10.0)) / f(365.0), dtype=np.float32)
    z = (np.sin(pi180 * xslat, dtype=np.float32) * sindec
         + np.cos(pi180 * xslat, dtype=np.float32) * np.sqrt(f(1.) - sindec * sindec, dtype=np.float32)
         * np.cos(pi * (f(hod) - f(12.0)) / f(12.0), dtype=np.float32))
    return np.maximum(z, f(1e-8)).astype(np.float32)


class Forcing:
    """Met forcing generator: `fill(T, step)` writes the FORCING fields of step `step` (0-based) into T."""

    def __init__(self, grid: Grid, tiles: dict[str, np.ndarray], dels: float, start_doy: int = 1):
        self.g, self.dels, self.start_doy = grid, float(dels), start_doy
        self.steps_per_day = int(round(86400.0 / dels))
        self.iveg = tiles["veg_iveg"][0].copy()
        self.laimax = LAIMAX[self.iveg - 1]
        self.peak = np.where(grid.lat >= 0, 200.0, 20.0).astype(np.float32)[grid.tile2land]
        self.lon_shift = (grid.lon / 15.0).astype(np.float32)               # local solar time offset (h)

    def time_of(self, step: int) -> tuple[int, float]:
        day, sub = divmod(step, self.steps_per_day)
        doy = (self.start_doy - 1 + day) % 365 + 1
        hod = (sub + 0.5) * self.dels / 3600.0
        return doy, hod

    def _land(self, step: int) -> dict[str, np.ndarray]:
        """Per-land-point forcing of step `step` (what one time slice of a gridded met file holds)."""
        g, f = self.g, np.float32
        doy, hod = self.time_of(step)
        day = step // self.steps_per_day
        rd = np.random.default_rng([g.seed, 3, day])        # per-day draws
        rs = np.random.default_rng([g.seed, 4, step])       # per-step draws
        n = g.nland
        # synoptic-scale weather: a few random planetary waves per day (spatially coherent, as real forcing is),
        # plus small per-point noise
        ph = rd.uniform(0, 2 * np.pi, 6)
        lonr, latr = np.deg2rad(g.lon.astype(np.float64)), np.deg2rad(g.lat.astype(np.float64))
        wave = (np.sin(3 * lonr + ph[0]) * np.cos(2 * latr + ph[1]) + np.sin(5 * lonr + ph[2]) * np.cos(4 * latr + ph[3])
                + 0.5 * np.sin(9 * lonr + ph[4]) * np.cos(7 * latr + ph[5])) / 2.5        # in [-1, 1]
        cloud = np.clip(0.5 + 0.5 * wave + rd.normal(0, 0.05, n), 0.0, 1.0)
        tau = (0.75 - 0.5 * cloud).astype(np.float32)                                   # 0.25 .. 0.75
        emis = (0.70 + 0.25 * cloud).astype(np.float32)                                 # 0.70 .. 0.95
        rh = np.clip(0.3 + 0.65 * cloud + rs.normal(0, 0.03, n), 0.2, 0.98).astype(np.float32)
        lst = (f(hod) + self.lon_shift) % f(24.0)
        coszen = sinbet(doy, g.lat, lst)
        sw = f(1370.0) * coszen * tau
        sw[coszen <= f(1e-4)] = 0.0
        sgn = np.where(g.lat >= 0, 0.0, 182.0).astype(np.float32)
        tair = (g.tmean + g.tamp * np.cos(f(2 * np.pi) * (f(doy) - f(200.0) - sgn) / f(365.0))
                + f(5.0) * np.cos(f(2 * np.pi) * (lst - f(15.0)) / f(24.0))).astype(np.float32)
        pmb = (f(1000.0) * np.exp(-g.elev / f(8000.0))).astype(np.float32)
        tc = tair - TFRZ
        qsat = (f(0.018016 / 0.02897) * f(6.106) * np.exp(f(17.27) * tc / (f(237.3) + tc)) / pmb).astype(np.float32)
        qv = (rh * qsat).astype(np.float32)
        ua = np.clip(rs.lognormal(np.log(3.0), 0.5, n), 0.1, 20.0).astype(np.float32)
        fld = (emis * f(5.67e-8) * tair ** 4).astype(np.float32)
        wet = rs.random(n) < 0.30 * cloud ** 2                                          # ~0.12 on average, under cloud
        precip = np.where(wet, rs.exponential(0.5 * self.dels / 1800.0, n), 0.0).astype(np.float32)
        precip_sn = np.where(tair <= TFRZ, precip, f(0.0)).astype(np.float32)    # cable_input.F90:2666-2671
        return dict(sw=sw, tair=tair, pmb=pmb, qv=qv, ua=ua, precip=precip, precip_sn=precip_sn, fld=fld, coszen=coszen,
                    doy=doy, lst=lst)

    def land_slice(self, step: int, out: np.ndarray | None = None) -> np.ndarray:
        """One time slice in met-file (ALMA) units, rows = cable_b200.lib.MET_ROWS: the input of
        cable_b200_set_met_async.  PSurf in Pa, Rainf in kg/m2/s, no Snowf variable (derived from Tair), CO2 in ppm."""
        L = self._land(step)
        f = np.float32
        if out is None:
            out = np.empty((11, self.g.nland), np.float32)
        out[0] = L["sw"]; out[1] = L["tair"]; out[2] = L["qv"]; out[3] = L["pmb"] * f(100.0); out[4] = L["ua"]
        out[5] = L["precip"] / f(self.dels); out[6] = 0.0; out[7] = L["fld"]; out[8] = f(350.0); out[9] = L["lst"]
        out[10] = f(L["doy"])
        return out

    def lai(self, step: int) -> np.ndarray:
        f = np.float32
        doy, _ = self.time_of(step)
        lai = self.laimax * (f(0.55) + f(0.45) * np.cos(f(2 * np.pi) * (f(doy) - self.peak) / f(365.0)))
        lai[self.iveg >= 14] = 0.0                                                # cable_serial.F90:575
        return lai.astype(np.float32)

    def fill(self, T: dict[str, np.ndarray], step: int) -> None:
        g, f = self.g, np.float32
        L = self._land(step)
        sw, tair, pmb, qv, ua, precip, precip_sn, fld, coszen, doy = (L[k] for k in
            ("sw", "tair", "pmb", "qv", "ua", "precip", "precip_sn", "fld", "coszen", "doy"))
        t2l = g.tile2land
        T["met_fsd"][0] = (f(0.5) * sw)[t2l]                                      # cable_input.F90:1880-1883
        T["met_fsd"][1] = (f(0.5) * sw)[t2l]
        T["met_tk"][0] = tair[t2l]
        T["met_pmb"][0] = pmb[t2l]
        T["met_qv"][0] = qv[t2l]
        T["met_ua"][0] = ua[t2l]
        T["met_precip"][0] = precip[t2l]
        T["met_precip_sn"][0] = precip_sn[t2l]
        T["met_fld"][0] = fld[t2l]
        T["met_ca"][0] = f(350.0e-6)                                              # cable.nml:32 fixedCO2
        T["met_coszen"][0] = coszen[t2l]
        T["met_doy"][0] = f(doy)
        T["met_tvrad"][0] = T["met_tk"][0]
        T["veg_vlai"][0] = self.lai(step)


FORCING_FIELDS = [f.name for f in FIELDS if f.role == 1 and not (f.flags & 8)]
BYTES_FORCING_PER_TILE = sum(4 * f.ncomp for f in FIELDS if f.role == 1 and not (f.flags & 8))
