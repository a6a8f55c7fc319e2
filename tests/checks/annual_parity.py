"""One model YEAR (2 920 three-hourly steps) of the CUDA path against the host-libm oracle -- the stand-in for a gfortran
build of the reference, whose float intrinsics (expf / logf / powf ...) may differ from the correctly rounded ones by 1 ulp
-- on a grid of `nland` land points x 5 tiles.  (The correctly rounded oracle is pinned bit for bit on the reference's own
Fortran source, tests/test_fortran_golden.py; the device matches that build to 1e-7 on every element.)

Reports  (a) annual totals of the headline fluxes and end-of-year stores: per-tile and grid-mean relative differences
         (north star: annual totals within 1e-5);
         (b) per field, over ALL steps and elements, the exact fraction outside the per-step tolerance (1e-4 binary32,
         1e-6 binary64) and the worst relative difference.

usage: python tests/checks/annual_parity.py [nland=2000] [nsteps=2920] [out=gpurun_out/annual_parity.txt]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cable_b200 import lib, synth
from cable_b200.cbm import CableB200
from oracle.pyoracle import Oracle
from util import output_fields

nland = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2920
out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "annual_parity.txt")
dels = 10800.0
TOTALS = ("canopy_fe", "canopy_fh", "canopy_fpn", "canopy_fes", "canopy_fev", "ssnow_runoff", "canopy_fns", "canopy_ga",
          "canopy_frday", "ssnow_smelt", "canopy_through")
STORES = ("ssnow_wb", "ssnow_tgg", "ssnow_snowd", "bgc_cplant", "bgc_csoil", "ssnow_wbice", "canopy_cansto")

cfg = lib.default_cfg(); cfg.output_level = 2
g = synth.make_grid(nland, 5)
To = synth.make_tiles(g, cfg)
Tg = {k: v.copy() for k, v in To.items()}
F = synth.Forcing(g, To, dels, start_doy=1)
o = Oracle(To, cfg, cr_math=False)
fields = output_fields()
tot_o = {n: np.zeros(g.mp) for n in TOTALS}; tot_g = {n: np.zeros(g.mp) for n in TOTALS}
n_out = {f.name: 0 for f in fields}; n_all = {f.name: 0 for f in fields}; worst = {f.name: 0.0 for f in fields}
t0 = time.time()
with CableB200(g.mp, cfg) as h:
    h.bind(Tg); h.upload_params(); h.upload_state()
    for k in range(nsteps):
        F.fill(To, k)
        for n in synth.FORCING_FIELDS:
            Tg[n][...] = To[n]
        o.cbm(k + 1, dels)
        h.cbm(k + 1, dels)
        for n in TOTALS:
            tot_o[n] += To[n][0]; tot_g[n] += Tg[n][0]
        for f in fields:
            a, b = To[f.name], Tg[f.name]
            if f.dtype != np.float64:
                a = a.astype(np.float64); b = b.astype(np.float64)
            d = np.abs(a - b)
            if not d.any():
                n_all[f.name] += a.size
                continue
            floor = 1e-3 * max(float(np.abs(a).max()), 1e-30)
            rel = d / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
            tol = 1e-4 if f.dtype == np.float32 else 1e-6
            n_out[f.name] += int(np.count_nonzero(rel > tol)); n_all[f.name] += a.size
            worst[f.name] = max(worst[f.name], float(rel.max()))
        if k % 365 == 0:
            print(f"step {k}: {time.time() - t0:.0f} s", flush=True)
lines = [f"annual parity, device (correctly rounded intrinsics) vs host-libm oracle: {g.mp} tiles x {nsteps} steps of {int(dels)} s, "
         f"{time.time() - t0:.0f} s; oracle dryLeaf warnings {o.warnings()}", ""]
lines.append("(a) annual totals: relative difference per tile |sum_dev - sum_ref| / max(|sum_ref|, 1e-3 max|sum_ref|), and of the grid mean")
lines.append(f"{'field':18s} {'grid-mean rel':>14s} {'median tile':>12s} {'p99 tile':>12s} {'max tile':>12s} {'tiles > 1e-5':>13s}")
summary = {"tiles": g.mp, "steps": nsteps, "totals": {}, "fields": {}}
for n in TOTALS:
    scale = np.maximum(np.abs(tot_o[n]), 1e-3 * np.abs(tot_o[n]).max() + 1e-30)
    rel = np.abs(tot_g[n] - tot_o[n]) / scale
    gm = abs(tot_g[n].mean() - tot_o[n].mean()) / max(abs(tot_o[n].mean()), 1e-30)
    lines.append(f"{n:18s} {gm:14.3e} {np.median(rel):12.3e} {np.percentile(rel, 99):12.3e} {rel.max():12.3e} {int((rel > 1e-5).sum()):13d}")
    summary["totals"][n] = {"grid_mean_rel": gm, "median": float(np.median(rel)), "p99": float(np.percentile(rel, 99)),
                            "max": float(rel.max()), "tiles_over_1e-5": int((rel > 1e-5).sum())}
lines.append("")
lines.append("end-of-year stores: max relative difference (field-scale floor)")
for n in STORES:
    a, b = To[n].astype(np.float64), Tg[n].astype(np.float64)
    floor = 1e-3 * max(float(np.abs(a).max()), 1e-30)
    rel = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    lines.append(f"{n:18s} max {rel.max():10.3e}   fraction > 1e-5: {np.mean(rel > 1e-5):.3e}")
lines.append("")
lines.append("(b) per field over all steps: fraction of elements outside the per-step tolerance (1e-4 binary32 / 1e-6 binary64)")
lines.append(f"{'field':26s} {'dtype':>6s} {'fraction outside':>17s} {'worst rel':>11s}")
for f in sorted(fields, key=lambda f: -n_out[f.name] / max(n_all[f.name], 1)):
    fr = n_out[f.name] / max(n_all[f.name], 1)
    summary["fields"][f.name] = {"fraction_outside": fr, "worst": worst[f.name]}
    if fr > 0 or worst[f.name] > 1e-7:
        lines.append(f"{f.name:26s} {'f64' if f.dtype == np.float64 else 'f32' if f.dtype == np.float32 else 'i32':>6s} {fr:17.3e} {worst[f.name]:11.3e}")
clean = sum(1 for f in fields if n_out[f.name] == 0)
lines.append(f"{clean} of {len(fields)} fields never leave the tolerance; overall fraction outside: "
             f"{sum(n_out.values()) / max(sum(n_all.values()), 1):.3e}")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
open(out_path, "w").write("\n".join(lines) + "\n")
json.dump(summary, open(out_path.replace(".txt", ".json"), "w"), indent=1)
print("\n".join(lines))
