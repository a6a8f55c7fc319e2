"""Bit-identity check between builds / launch settings: run the resident step loop and print one SHA-256 over every
prognostic and driver-visible diagnostic array.  Pure scheduling changes (tile order, block shape, barriers) and
same-value math rewrites must leave the digest unchanged.
usage: [CABLE_B200_LIB=...] [CABLE_B200_PIPE_STREAMS=0] python tools/state_hash.py [nland] [steps]"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cable_b200 import lib, synth
from cable_b200.cbm import CableB200
from cable_b200.registry import FIELDS, ROLE, FLAG
nland = int(sys.argv[1]) if len(sys.argv) > 1 else 62000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cfg = lib.default_cfg(); cfg.n_forcing_slots = 8; cfg.output_level = 1
g = synth.make_grid(nland, 5); T = synth.make_tiles(g, cfg); F = synth.Forcing(g, T, 10800.0, start_doy=172)
with CableB200(g.mp, cfg) as h:
    h.bind(T); h.upload_params(); h.upload_state()
    for k in range(steps):
        F.fill(T, k); h.set_forcing_async(k % 8); h.sync()
        h.step(k + 1, 10800.0, k % 8)
    h.sync(); h.download_state(); h.download_diag(True)
dig = hashlib.sha256()
for f in FIELDS:
    if f.flags & FLAG["HOSTONLY"]: continue
    if f.role == ROLE["STATE"] or (f.role == ROLE["DIAG"] and f.flags & FLAG["STAR"]):
        dig.update(np.ascontiguousarray(T[f.name]).tobytes())
print(f"state_hash mp={g.mp} steps={steps} lib={os.path.basename(lib.LIB_PATH)} pipe_streams={os.environ.get('CABLE_B200_PIPE_STREAMS', 'default')}: {dig.hexdigest()[:32]}")
