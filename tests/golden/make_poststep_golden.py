"""Golden vectors of the driver statements serialdrv runs right after CALL cbm, from the reference's own Fortran source executed by
oracle/frun: the dels scaling of smelt / rnof1 / rnof2 / runoff (src/offline/cable_serial.F90:602-605, restated here: four array
statements), then CALL sumcflux (src/science/casa-cnp/casa_sumcflux.F90:37, icycle = 0), CALL mass_balance and CALL energy_balance
(src/offline/cable_checks.F90:472, 565) -- the reference's routines, unmodified.  Run in the build container:

    python tests/golden/make_poststep_golden.py        # writes tests/golden/fortran_poststep_v1.npz  (~3 min)

The fixture holds bal%* / sum_flux%* / canopy%fnee after every step (14 steps: the ktau == 1 initialisations and the ktau > 10
accumulation branch of mass_balance).  tests/test_fortran_golden.py pins the C++ oracle's post-step on it (CPU, bit for bit);
tests/test_gpu_driver.py the device's cable_b200_post_step."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, HERE)
OUT = os.path.join(HERE, "fortran_poststep_v1.npz")
NLAND, NSTEPS, DOY, DELS = 16, 14, 100, 10800.0
BAL = ("wbal", "wbal_tot", "precip_tot", "rnoff_tot", "evap_tot", "radbal", "ebalsoil", "ebalveg", "ebal", "ebal_tot", "radbalsum")
SUMS = ("sumpn", "sumrp", "sumrpw", "sumrpr", "sumrs", "sumrd", "dsumpn", "dsumrp", "dsumrd")
SCALED = ("smelt", "rnof1", "rnof2", "runoff")


def case_inputs():
    from cable_b200 import lib, synth
    cfg = lib.default_cfg(); cfg.output_level = 2
    grid = synth.make_grid(NLAND, 5, seed=synth.SEED + 11)
    tiles = synth.make_tiles(grid, cfg)
    forcing = synth.Forcing(grid, tiles, DELS, start_doy=DOY)
    return cfg, grid, tiles, forcing


def main():
    from cable_b200.registry import FIELDS
    from oracle.frun.run_cbm import FortranCbm
    cfg, grid, T, F = case_inputs()
    fc = FortranCbm(T, cfg, FIELDS)
    S, I = fc.S, fc.I
    consts = I.module("cable_phys_constants_mod") if "cable_phys_constants_mod" in getattr(I, "modules", {}) else None
    out = {}
    for k in range(NSTEPS):
        F.fill(T, k); fc.cbm(k + 1, DELS)
        for n in SCALED:                                                          # cable_serial.F90:602-605
            S["ssnow"].f[n].a[...] = S["ssnow"].f[n].a * np.float32(DELS)
        I.call("sumcflux_mod", "sumcflux", np.int32(k + 1), np.int32(1), np.int32(NSTEPS), np.float32(DELS), S["bgc"], S["canopy"],
               S["soil"], S["ssnow"], S["sum_flux"], S["veg"], S["met"], fc_casaflux(fc), np.bool_(False))
        I.call("cable_checks_module", "mass_balance", np.float32(DELS), np.int32(k + 1), S["ssnow"], S["soil"], S["canopy"], S["met"],
               S["air"], S["bal"])
        I.call("cable_checks_module", "energy_balance", np.float32(DELS), np.int32(k + 1), S["met"], S["rad"], S["canopy"], S["bal"],
               S["ssnow"], np.float32(5.67e-8), np.float32(1.0), np.float32(1.0))        # CSBOLTZ, CEMLEAF, CEMSOIL
        for n in BAL:
            out[f"step{k}/bal_{n}"] = S["bal"].f[n].a.copy()
        for n in SUMS:
            out[f"step{k}/sum_flux_{n}"] = S["sum_flux"].f[n].a.copy()
        out[f"step{k}/canopy_fnee"] = S["canopy"].f["fnee"].a.copy()
        for n in SCALED:
            out[f"step{k}/ssnow_{n}"] = S["ssnow"].f[n].a.copy()
        print("step", k + 1, "wbal max", float(np.abs(S["bal"].f["wbal"].a).max()), "ebal max", float(np.abs(S["bal"].f["ebal"].a).max()), flush=True)
        # the scaled rates live on in ssnow until the next cbm overwrites them: the reference does not undo the scaling
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out), "arrays, statements", I.nstmt)


_CASAFLUX = {}


def fc_casaflux(fc):
    """sumcflux takes casaflux even when icycle = 0 (never touched): an allocated casa_flux of the reference's own type"""
    if "x" not in _CASAFLUX:
        I = fc.I
        cv = I.module("casavariable")
        mp = int(fc.tiles["met_tk"].shape[-1])
        S = {t: I.new_struct(I.lookup_in_module(cv, "$type:" + t)) for t in ("casa_biome", "casa_pool", "casa_flux", "casa_met", "casa_balance")}
        I.call("casavariable", "alloc_casavariable", S["casa_biome"], S["casa_pool"], S["casa_flux"], S["casa_met"], S["casa_balance"], np.int32(mp))
        _CASAFLUX["x"] = S["casa_flux"]
    return _CASAFLUX["x"]


if __name__ == "__main__":
    main()
