"""GPU debugging aid: run one switch case for a few steps and describe the tiles whose results differ from the oracle.
usage: python tests/checks/debug_switch.py soil_thermal_fix=1 [nsteps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
from cable_b200 import lib, synth
from cable_b200.cbm import CableB200
from oracle.pyoracle import Oracle
from util import DELS, make_case, compare_tiles
cfg = lib.default_cfg()
nsteps = 1
for a in sys.argv[1:]:
    if "=" in a:
        k, v = a.split("="); setattr(cfg, k, int(v))
    else:
        nsteps = int(a)
cfg.output_level = 2
cfg, grid, T, F = make_case(300, cfg=cfg, start_doy=200)
G = {k: v.copy() for k, v in T.items()}
o = Oracle(T, cfg, cr_math=True)
with CableB200(grid.mp, cfg) as h:
    h.bind(G); h.upload_params(); h.upload_state()
    for k in range(nsteps):
        F.fill(T, k)
        for n in synth.FORCING_FIELDS: G[n][...] = T[n]
        pre = {n: T[n].copy() for n in ("ssnow_tgg", "ssnow_wb", "ssnow_wbice", "ssnow_wbliq", "ssnow_snowd", "ssnow_isflag")}
        o.cbm(k + 1, DELS); h.cbm(k + 1, DELS)
        res = compare_tiles(T, G)
        bad = {n: r for n, r in res.items() if r[0] > r[1]}
        print("step", k + 1, "fields outside tolerance:", len(bad), sorted(bad))
        if bad:
            name = "ssnow_tgg" if "ssnow_tgg" in bad else sorted(bad)[0]
            a, b = T[name].astype(np.float64), G[name].astype(np.float64)
            rel = np.abs(a - b) / np.maximum(np.abs(a), 1e-30)
            tiles = np.unique(np.nonzero(rel > 1e-5)[-1])
            print(name, "differs on", tiles.size, "tiles of", grid.mp)
            for i in tiles[:12]:
                print(" tile", i, "iveg", T["veg_iveg"][0][i], "isoilm", T["soil_isoilm"][0][i], "isflag", pre["ssnow_isflag"][0][i],
                      "snowd", pre["ssnow_snowd"][0][i], "tgg", pre["ssnow_tgg"][:, i], "wbice", pre["ssnow_wbice"][:, i],
                      "wb", pre["ssnow_wb"][:, i], "wbliq", pre["ssnow_wbliq"][:, i], "\n   ref", a[..., i].ravel(), "\n   gpu", b[..., i].ravel())
            break
