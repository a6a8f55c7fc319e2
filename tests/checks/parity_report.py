"""Run oracle and CUDA path side by side on the same seeded inputs and print a per-field report.
Usage: python tests/checks/parity_report.py [nland] [nsteps] [gs_switch]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from cable_b200 import lib, synth
from cable_b200.cbm import CableB200
from cable_b200.registry import FIELDS, ROLE, FLAG
from oracle.pyoracle import Oracle

nland = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
gs = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cr = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dels = 10800.0
cfg = lib.default_cfg(); cfg.gs_switch = gs; cfg.output_level = 2
g = synth.make_grid(nland, 5)
To = synth.make_tiles(g, cfg)
Tg = {k: v.copy() for k, v in To.items()}
F = synth.Forcing(g, To, dels, start_doy=100)
o = Oracle(To, cfg, cr_math=bool(cr))
h = CableB200(g.mp, cfg)
h.bind(Tg); h.upload_params(); h.upload_state()
worst = {}
for k in range(nsteps):
    F.fill(To, k)
    for n in synth.FORCING_FIELDS: Tg[n][...] = To[n]
    o.cbm(k + 1, dels)
    h.cbm(k + 1, dels)
    for f in FIELDS:
        if f.flags & FLAG["HOSTONLY"] or f.role in (ROLE["FORCING"], ROLE["PARAM"]): continue
        a, b = To[f.name].astype(np.float64), Tg[f.name].astype(np.float64)
        scale = np.maximum(np.abs(a), np.abs(b))
        tol = 1e-4 if f.dtype == np.float32 else 1e-6
        floor = 1e-3 * max(np.abs(a).max(), 1e-30)      # field-scale absolute floor
        err = np.abs(a - b)
        badmask = err > tol * scale + tol * floor
        nb = int(np.count_nonzero(badmask)) if np.all(np.isfinite(b)) else -1
        rel = float((err / np.maximum(scale, floor)).max()) if np.all(np.isfinite(b)) else float("inf")
        w = worst.get(f.name, (0, 0.0, 0))
        worst[f.name] = (max(w[0], nb) if nb >= 0 else -1, max(w[1], rel), k if rel > w[1] else w[2])
print(f"mp={g.mp} steps={nsteps} gs={gs} cr_oracle={cr}; oracle warns={o.warnings()} gpu warns={h.counters().n_dryleaf_warn}")
print(f"{'field':28s} {'max bad tiles':>14s} {'max rel err':>12s} step")
for n, (nb, rel, k) in sorted(worst.items(), key=lambda kv: -kv[1][1]):
    if rel > 1e-7 or nb != 0:
        print(f"{n:28s} {nb:14d} {rel:12.3e} {k}")
nclean = sum(1 for v in worst.values() if v[1] <= 1e-7 and v[0] == 0)
print(f"{nclean} of {len(worst)} fields identical to 1e-7")
