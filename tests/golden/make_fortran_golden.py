"""Golden vectors of the REFERENCE ITSELF: /root/reference's SUBROUTINE cbm executed from its unmodified Fortran source
by the interpreter in oracle/frun (this image has no Fortran compiler, profiles/r02_fortran_compiler_probe.txt), on the
seeded synthetic cases below.  Run in the build container (needs /root/reference):

    python tests/golden/make_fortran_golden.py            # writes tests/golden/fortran_cbm_v1.npz  (~10 min, 8 processes)

The inputs are regenerated from the seeds by the tests (cable_b200.synth), so the fixture holds outputs only: every state
and diagnostic field of the registry after the last step of each case, plus a per-step trace of a few fluxes and stores.
tests/test_fortran_golden.py pins the C++ oracle against it on the CPU; tests/test_gpu_fortran_golden.py the CUDA path.
"""
from __future__ import annotations

import multiprocessing as mp_
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

OUT = os.path.join(HERE, "fortran_cbm_v1.npz")
TRACE = ("canopy_fe", "canopy_fh", "canopy_fpn", "canopy_fes", "ssnow_runoff", "ssnow_snowd", "ssnow_tss", "canopy_cansto",
         "ssnow_wbtot", "canopy_ga", "ssnow_isflag")

# name -> (nland, nsteps, start_doy, dels, site_lat, switches)
CASES = {
    "leuning_spring":   (60, 12, 120, 10800.0, None, dict()),
    "medlyn_winter":    (60, 12, 10, 10800.0, None, dict(gs_switch=1)),
    "leuning_july":     (40, 10, 200, 10800.0, None, dict()),
    "site_half_hourly": (1, 60, 340, 1800.0, -35.6, dict()),
    "fwsoil_nonlinear_pm": (24, 6, 200, 10800.0, None, dict(fwsoil_switch=1, ssnow_potev=1)),
    "fwsoil_lai_ktaul":    (24, 6, 200, 10800.0, None, dict(fwsoil_switch=2, gs_switch=1)),
    "carbon_runoff_opts":  (24, 6, 30, 10800.0, None, dict(diag_soil_resp_on=0, l_new_runoff_speed=1, l_new_reduce_soilevp=1)),
    "xsw_litter_revcorr":  (24, 6, 200, 10800.0, None, dict(litter=1, l_rev_corr=1, ssnow_potev=1)),
    "xsw_thermal_rough":   (24, 6, 30, 10800.0, None, dict(soil_thermal_fix=1, l_new_roughness_soil=1)),
    "xsw_redistrb_climate": (24, 6, 200, 10800.0, None, dict(redistrb=1, call_climate=1, gs_switch=1)),
    "caller_inputs":       (24, 8, 200, 10800.0, None, dict(caller_duties=0, met_tv_is_tk=0)),
    # ten model days through the spring melt: state carried over 80 steps (snow packs building up and melting, soil thawing,
    # canopy storage filling and draining, snow age), so a difference that only shows after many steps has room to show
    "leuning_ten_days":    (20, 80, 85, 10800.0, None, dict()),
    # ... and ten days of deep winter with the other stomatal model: multi-layer snow packs ageing, densifying and re-layering
    "medlyn_winter_ten_days": (20, 80, 5, 10800.0, None, dict(gs_switch=1)),
}


def case_inputs(name):
    """-> cfg, grid, tiles, forcing for a named case (shared with the tests)"""
    from cable_b200 import lib, synth
    nland, nsteps, doy, dels, site_lat, sw = CASES[name]
    cfg = lib.default_cfg()
    for k, v in sw.items():
        setattr(cfg, k, v)
    cfg.output_level = 2
    grid = synth.make_grid(nland, 5, seed=synth.SEED + 7, site_lat=site_lat)
    tiles = synth.make_tiles(grid, cfg)
    forcing = synth.Forcing(grid, tiles, dels, start_doy=doy)
    return cfg, grid, tiles, forcing


def caller_step(name, T, forcing, k):
    """what the caller does before step k of a case (forcing; the caller-set inputs of the `caller_inputs` case)"""
    forcing.fill(T, k)
    if name == "caller_inputs":
        T["met_tvair"][0] = T["met_tk"][0] + np.float32(0.25)
        T["met_tvrad"][0] = T["met_tk"][0] - np.float32(0.5)
        T["canopy_oldcansto"][...] = T["canopy_cansto"]          # cable_serial.F90:573, done by the caller


def run_case(name):
    from cable_b200.registry import FIELDS
    from oracle.frun.run_cbm import FortranCbm
    nland, nsteps, doy, dels, site_lat, sw = CASES[name]
    cfg, grid, T, F = case_inputs(name)
    t0 = time.time()
    fc = FortranCbm(T, cfg, FIELDS)
    trace = {n: [] for n in TRACE}
    for k in range(nsteps):
        caller_step(name, T, F, k)
        fc.cbm(k + 1, dels)
        for n in TRACE:
            trace[n].append(T[n].copy())
    out = {}
    for f in FIELDS:
        if f.role in (4, 8) and not (f.flags & 4):
            out[f"{name}/final/{f.name}"] = T[f.name].copy()
    for n in TRACE:
        out[f"{name}/trace/{n}"] = np.stack(trace[n])
    cover = dict(tiles=grid.mp, snow=int((T["ssnow_snowd"] > 0).sum()), three_layer=int((T["ssnow_isflag"] == 1).sum()),
                 ice=int((T["soil_isoilm"] == 9).sum()), lakes=int((T["veg_iveg"] == 16).sum()),
                 veg=int((T["canopy_vlaiw"] > 0.001).sum()), frozen=int((T["ssnow_wbice"] > 0).any(axis=0).sum()),
                 statements=int(fc.I.nstmt), seconds=round(time.time() - t0, 1), skipped=[s[:2] for s in fc.I.skipped])
    print(name, cover, flush=True)
    return out


def main():
    names = sys.argv[1:] or list(CASES)
    with mp_.get_context("fork").Pool(min(8, len(names))) as pool:
        parts = pool.map(run_case, names)
    merged = {}
    if os.path.exists(OUT) and sys.argv[1:]:
        with np.load(OUT) as z:
            merged.update({k: z[k] for k in z.files})
    for p in parts:
        merged.update(p)
    np.savez_compressed(OUT, **merged)
    print("wrote", OUT, f"{os.path.getsize(OUT) / 1e6:.2f} MB,", len(merged), "arrays")


if __name__ == "__main__":
    main()
