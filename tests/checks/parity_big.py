"""Parity at a size that takes kernel A's 768-thread path (>= 148*768 tiles): device vs CR oracle, every output field.
usage: python tests/checks/parity_big.py [nland] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
from cable_b200 import lib, synth
from cable_b200.cbm import CableB200
from oracle.pyoracle import Oracle
from util import compare_tiles, DELS
nland = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = lib.default_cfg(); cfg.output_level = 2
g = synth.make_grid(nland, 5); T = synth.make_tiles(g, cfg); F = synth.Forcing(g, T, DELS, start_doy=172)
Tg = {k: v.copy() for k, v in T.items()}
o = Oracle(T, cfg, cr_math=True)
with CableB200(g.mp, cfg) as h:
    h.bind(Tg); h.upload_params(); h.upload_state()
    for k in range(steps):
        F.fill(T, k)
        for n in synth.FORCING_FIELDS: Tg[n][...] = T[n]
        o.cbm(k + 1, DELS); h.cbm(k + 1, DELS)
    warn = h.counters().n_dryleaf_warn
res = compare_tiles(T, Tg)
bad = {n: r for n, r in res.items() if r[0] > r[1]}
worst = max(r[0] for r in res.values())
print(f"parity_big: mp={g.mp} steps={steps} lib={os.path.basename(lib.LIB_PATH)} worst rel {worst:.2e} bad fields {len(bad)} {list(bad.items())[:4]} warns {warn}/{o.warnings()}")
sys.exit(1 if bad else 0)
