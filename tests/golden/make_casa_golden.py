"""Golden vectors of the reference's CASA-CNP daily step: /root/reference's biogeochem (biogeochem_casa.F90:7) and bgcdriver
(bgcdriver.F90:7) executed from their unmodified Fortran source by oracle/frun.  Run in the build container:

    python tests/golden/make_casa_golden.py        # writes tests/golden/fortran_casa_v1.npz

Two kinds of case:
  bio/<name>   biogeochem alone, NDAYS consecutive days on a synthetic casa state with fixed daily met: the fixture holds the
               per-tile casa arrays after every day.  Inputs are regenerated from the seeds by the test -> exact comparison.
  drv/<name>   bgcdriver called after every cbm step for two model days (cbm = the pinned C++ oracle): the fixture holds the
               casa arrays at the end, for the device's cbm + bgcdriver pipeline.
"""
from __future__ import annotations

import multiprocessing as mp_
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(HERE, "fortran_casa_v1.npz")
NDAYS = 4
TRACE = ("casaflux_cnpp", "casapool_cplant", "casapool_nsoilmin", "casapool_psoillab", "casamet_glai", "casabal_cbalance", "phen_phase",
         "casaflux_crsoil", "casaflux_fraccalloc", "casaflux_kplant")      # stored after every day; every field after the last
# name -> (icycle, lalloc, call_climate, l_limit_labile)
BIO = {"c_fixed": (1, 0, 0, 0), "cn_dynamic": (2, 1, 0, 0), "cnp_fixed": (3, 0, 0, 0), "cnp_dynamic": (3, 1, 0, 1),
       "cnp_lasa": (3, 3, 0, 0), "cn_climate": (2, 0, 1, 0), "cnp_climate_dyn": (3, 1, 1, 0),
       "cnp_year_sweep": (3, 1, 0, 0)}
# day of year of every biogeochem call: the default wraps a year end; the sweep visits every month, so that every tile walks
# through all four phenology phases (casa_cnp.F90:2305-2360) with its leaf onset / fall rates
IDOY = {"cnp_year_sweep": [1] + [15 + 30 * m for m in range(12)]}


def idoys(name):
    return IDOY.get(name, [1 if day == 0 else 364 + day for day in range(NDAYS)])
DRV = {"drv_cnp": (3, 1, 0, 0), "drv_c": (1, 0, 0, 0)}
# casa_feedback: name -> (icycle, cable_user%vcmax)
FB = {"fb_cn_standard": (2, "standard"), "fb_cnp_standard": (3, "standard"), "fb_cnp_walker": (3, "Walker2014")}
NLAND, DOY = 24, 200


def casa_cfg(spec):
    from cable_b200 import casa
    c = casa.CasaCfg()
    c.struct_bytes = 40
    c.icycle, c.lalloc, c.call_climate, c.l_limit_labile = spec
    c.mvtype = 17
    return c


def bio_inputs(name):
    """-> cfg, grid, tiles, casa arrays, silt, clay, first idoy: synthetic state with a day's met already in casamet/casaflux"""
    from cable_b200 import casa
    from util import make_case
    cfg, grid, T, F = make_case(NLAND, start_doy=DOY)
    ccfg = casa_cfg(BIO[name])
    A = casa.synth_casa(grid, T, ccfg, seed=31)
    rng = np.random.default_rng([31, 9])
    mp = grid.mp
    A["casamet_tairk"][0] = rng.uniform(255.0, 303.0, mp)
    A["casamet_tsoil"][...] = rng.uniform(262.0, 300.0, (6, mp))
    A["casamet_tsoil"][:, ::17] = 249.0                       # frozen columns: the tsoilavg > 250 tests
    A["casamet_moist"][...] = rng.uniform(0.05, 0.45, (6, mp))
    A["casaflux_cgpp"][0] = rng.uniform(0.0, 9.0, mp) * (A["casamet_iveg2"][0] != 0)
    A["casaflux_cgpp"][0][::5] *= 0.002                       # starving tiles: the NPP < 0 branches of casa_rplant / casa_allocation / casa_delplant
    A["casaflux_crmplant"][0] = 0.12 * A["casaflux_cgpp"][0]
    T["climate_qtemp_max_last_year"][0] = rng.uniform(285.0, 305.0, mp).astype(np.float32)
    silt, clay = casa.soil_texture(T)
    return cfg, grid, T, A, silt, clay, ccfg


def run_bio(name):
    from cable_b200 import casa
    from oracle.frun.run_casa import FortranCasa
    cfg, grid, T, A, silt, clay, ccfg = bio_inputs(name)
    fc = FortranCasa(T, A, casa.FIELDS, ccfg, silt, clay)
    S = fc.S
    out = {}
    days = idoys(name)
    for day, idoy in enumerate(days):                          # day 0: idoy == 1 resets the annual sums
        xs = [np.zeros(grid.mp, np.float64) for _ in range(7)]; ys = [np.zeros(grid.mp, np.float32) for _ in range(15)]
        # the leaf maintenance respiration is an input of every day (bgcdriver sets it from the day's mean)
        S["casaflux"].f["crmplant"].a[:, 0] = 0.12 * S["casaflux"].f["cgpp"].a
        fc.I.call("biogeochem_mod", "biogeochem", np.int32(8 * (day + 1)), np.float32(10800.0), np.int32(idoy), np.int32(ccfg.lalloc),
                  S["veg"], S["soil"], S["casabiome"], S["casapool"], S["casaflux"], S["casamet"], S["casabal"], S["phen"], S["pop"],
                  S["climate"], *xs, *ys)
        fc.pull()
        for f in casa.FIELDS:
            if f.key == 0 and (day == len(days) - 1 or f.name in TRACE):
                out[f"bio/{name}/day{day}/{f.name}"] = A[f.name].copy()
    neg = int((A["casaflux_cnpp"][0] < 0).sum()); pos = int((A["casaflux_cnpp"][0] > 0).sum())
    if name in IDOY:
        print(name, "phenology phases seen at the end:", np.bincount(A["phen_phase"][0], minlength=4).tolist(), flush=True)
    print(name, "tiles", grid.mp, "NPP>0", pos, "NPP<0", neg, "statements", fc.I.nstmt, flush=True)
    return out


def run_drv(name):
    from cable_b200 import casa
    from oracle.frun.run_casa import FortranCasa
    from oracle.pyoracle import Oracle
    from util import make_case, DELS
    cfg, grid, T, F = make_case(NLAND, start_doy=DOY)
    ccfg = casa_cfg(DRV[name])
    cfg.icycle = ccfg.icycle                                   # cbm skips its simple carbon model, as in a CASA run (cbm:214)
    A = casa.synth_casa(grid, T, ccfg, seed=31)
    silt, clay = casa.soil_texture(T)
    o = Oracle(T, cfg, cr_math=True)
    fc = FortranCasa(T, A, casa.FIELDS, ccfg, silt, clay)
    for k in range(16):
        F.fill(T, k); o.cbm(k + 1, DELS)
        fc.bgcdriver(k + 1, 1, 10000, DELS, 8, DOY + k // 8)
        post = fc.sumcflux(k + 1, 1, 10000, DELS)              # cable_serial.F90:713, right after bgcdriver
    out = {f"drv/{name}/{f.name}": A[f.name].copy() for f in casa.FIELDS if f.key == 0}
    out.update({f"drv/{name}/post/{n}": v for n, v in post.items()})
    print(name, "done; NPP>0:", int((A["casaflux_cnpp"][0] > 0).sum()), flush=True)
    return out


def fb_inputs(name):
    """synthetic pools with every branch of casa_feedback present: bare leaf pools, LAI below glaimin, zero leaf P"""
    cfg, grid, T, A, silt, clay, ccfg = bio_inputs("cnp_fixed")
    ccfg.icycle = FB[name][0]
    A["casapool_cplant"][0][::11] = 0.0
    A["casamet_glai"][0][::7] = 0.05
    A["casapool_pplant"][0][::13] = 0.0
    A["casapool_nplant"][0][5::23] *= 4.0                      # N:C above ratioNCplantmax, N:P above 30
    A["casapool_pplant"][0][3::19] *= 6.0                      # N:P below 8
    return cfg, grid, T, A, silt, clay, ccfg


def run_fb(name):
    from cable_b200 import casa
    from oracle.frun.run_casa import FortranCasa
    cfg, grid, T, A, silt, clay, ccfg = fb_inputs(name)
    fc = FortranCasa(T, A, casa.FIELDS, ccfg, silt, clay)
    S = fc.S
    fc.I.lookup_in_module(fc.I.module("cable_common_module"), "cable_user").f["vcmax"].s = FB[name][1]
    S["veg"].f["vcmax"].a[...] = T["veg_vcmax"][0]; S["veg"].f["ejmax"].a[...] = T["veg_ejmax"][0]
    fc.I.call("feedback_mod", "casa_feedback", np.int32(1), S["veg"], S["casabiome"], S["casapool"], S["casamet"])
    out = {f"fb/{name}/veg_vcmax": S["veg"].f["vcmax"].a.copy(), f"fb/{name}/veg_ejmax": S["veg"].f["ejmax"].a.copy()}
    print(name, "vcmax changed on", int((out[f"fb/{name}/veg_vcmax"] != T["veg_vcmax"][0]).sum()), "of", grid.mp, "tiles", flush=True)
    return out


def _job(j):
    return {"bio": run_bio, "drv": run_drv, "fb": run_fb}[j[0]](j[1])


def main():
    jobs = [("bio", n) for n in BIO] + [("drv", n) for n in DRV] + [("fb", n) for n in FB]
    with mp_.get_context("fork").Pool(8) as pool:
        parts = pool.map(_job, jobs)
    merged = {}
    for p in parts:
        merged.update(p)
    np.savez_compressed(OUT, **merged)
    print("wrote", OUT, f"{os.path.getsize(OUT) / 1e6:.2f} MB,", len(merged), "arrays")


if __name__ == "__main__":
    main()
