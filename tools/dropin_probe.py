"""Unchanged-caller drop-in (cable_b200_cbm with the driver's output mask): time per call vs the pipeline's chunk count.
python tools/dropin_probe.py [nland]     (CABLE_B200_CHUNKS / CABLE_B200_GRAPH from the environment)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cable_b200 import lib, synth
from cable_b200.cbm import CableB200
from cable_b200.registry import FIELDS, ROLE, FLAG
import bench
nland = int(sys.argv[1]) if len(sys.argv) > 1 else 62000
DELS = 10800.0
cfg = lib.default_cfg(); cfg.output_level = 1
g = synth.make_grid(nland, 5); T = synth.make_tiles(g, cfg); F = synth.Forcing(g, T, DELS, start_doy=172)
keep = []
def pin(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True); t.numpy()[...] = a; keep.append(t); return t.numpy()
for f in FIELDS:
    if not (f.flags & FLAG["HOSTONLY"]) and (f.role == ROLE["STATE"] or (f.role == ROLE["DIAG"] and f.flags & FLAG["STAR"])):
        T[f.name] = pin(T[f.name])
fs = []
for k in range(4):
    F.fill(T, k); fs.append({n: pin(T[n].copy()) for n in synth.FORCING_FIELDS})
h = CableB200(g.mp, cfg); h.bind(T); h.upload_params(); h.upload_state()
h.set_output_mask(bench.DRIVER_READS)
for k in range(3):
    h.bind(fs[k % 4]); h.cbm(k + 1, DELS)
h.reset_counters()
n = 12
t0 = time.perf_counter()
for k in range(3, 3 + n):
    h.bind(fs[k % 4]); h.cbm(k + 1, DELS)
dt = (time.perf_counter() - t0) / n
c = h.counters()
print(f"chunks={os.environ.get('CABLE_B200_CHUNKS', 'auto')} graph={os.environ.get('CABLE_B200_GRAPH', '1')}: {dt * 1e3:.3f} ms/call, {g.mp / dt / 1e6:.1f} M tile-steps/s, "
      f"H2D {c.h2d_bytes / n / 1e6:.1f} MB, D2H {c.d2h_bytes / n / 1e6:.1f} MB per call -> {(c.h2d_bytes + c.d2h_bytes) / n / dt / 1e9:.1f} GB/s", flush=True)
