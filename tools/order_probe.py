"""Probe: does the ORDER of tiles in the device arrays matter for the step time?  (tiles are independent, so the library is
free to keep them in any order on the device)  python tools/order_probe.py [nland] [steps] order...
orders: none | patch:<window land points> | iveg:<window land points> | iveg (global, stable) | random"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cable_b200 import lib, synth
from cable_b200.cbm import CableB200
nland = int(sys.argv[1]); steps = int(sys.argv[2]); orders = sys.argv[3:]
cfg = lib.default_cfg(); cfg.n_forcing_slots = 8; cfg.output_level = 1
g = synth.make_grid(nland, 5); T0 = synth.make_tiles(g, cfg); F = synth.Forcing(g, T0, 10800.0, start_doy=172)
mp = g.mp
FS = []
for k in range(8):
    F.fill(T0, k); FS.append({n: T0[n].copy() for n in synth.FORCING_FIELDS})
def perm_of(order):
    i = np.arange(mp); land = g.tile2land; patch = i - g.cstart[land]; iveg = T0["veg_iveg"][0]
    if order == "none": return i
    if order == "random": return np.random.default_rng(1).permutation(mp)
    if order == "iveg": return np.argsort(iveg, kind="stable")
    kind, w = order.split(":"); w = int(w)
    key = patch if kind == "patch" else iveg
    return np.lexsort((i, key, land // w))
for order in orders:
    p = perm_of(order)
    T = {n: np.ascontiguousarray(a[..., p]) if a.shape[-1] == mp else a.copy() for n, a in T0.items()}
    h = CableB200(mp, cfg); h.bind(T); h.upload_params(); h.upload_state()
    for k in range(8):
        h.bind({n: np.ascontiguousarray(a[..., p]) for n, a in FS[k].items()}); h.set_forcing_async(k); h.sync()
    for k in range(8): h.step(k + 1, 10800.0, k % 8)
    h.sync(); h.reset_counters(); h.profile(True)
    t0 = time.perf_counter()
    for k in range(8, 8 + steps): h.step(k + 1, 10800.0, k % 8)
    h.sync(); dt = time.perf_counter() - t0
    c = h.counters()
    print(f"order={order}: mp={mp} kernel {c.kernel_ms / c.kernel_ms_count:.3f} ms/step, wall {dt / steps * 1e3:.3f} ms/step, "
          f"{mp * steps / dt / 1e6:.1f} M tile-steps/s redo={c.n_fastdiv_redo_blocks}", flush=True)
    h.close()
