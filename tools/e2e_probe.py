"""Where the offline-driver loop loses time against the bare resident step: python tools/e2e_probe.py [nland] [steps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cable_b200 import lib, synth
from cable_b200.cbm import CableB200
import bench
nland = int(sys.argv[1]) if len(sys.argv) > 1 else 62000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
DELS = 10800.0
cfg = lib.default_cfg(); cfg.n_forcing_slots = 8; cfg.output_level = 1
g = synth.make_grid(nland, 5); T = synth.make_tiles(g, cfg); F = synth.Forcing(g, T, DELS, start_doy=172)
conv = lib.MetConvert(tair_offset=0.0, psurf_scale=0.01, rainf_scale=DELS, co2_scale=1.0e-6, snowf_from_tair=1)
slices = []
for k in range(8):
    t = torch.empty((len(lib.MET_ROWS), nland), dtype=torch.float32, pin_memory=True); t.numpy()[...] = F.land_slice(k); slices.append(t)
outs = [torch.zeros((len(bench.OUTPUT_ROWS), nland), dtype=torch.float32, pin_memory=True) for _ in range(2)]
def run(mode):
    Tm = {k: v.copy() for k, v in T.items()}
    h = CableB200(g.mp, cfg); h.bind(Tm); h.upload_params(); h.upload_state()
    h.driver_init(g.cstart, g.cend, g.patchfrac, g.lat[g.tile2land]); h.output_plan(bench.OUTPUT_ROWS)
    Tm["veg_vlai"][0] = F.lai(0); h.upload_lai()
    def one(k):
        if "met" in mode: h.set_met_async(k % 8, slices[k % 8].numpy(), conv)
        h.step(k + 1, DELS, k % 8)
        if "post" in mode: h.post_step(k + 1, 1, DELS)
        if "out" in mode:
            h.output_wait(); h.output_fetch_async(outs[k % 2].numpy())
    if "met" not in mode:
        for k in range(8): h.set_met_async(k, slices[k].numpy(), conv)
    for k in range(8): one(k)
    h.output_wait(); h.sync()
    t0 = time.perf_counter()
    for k in range(8, 8 + steps): one(k)
    h.output_wait(); h.sync()
    dt = (time.perf_counter() - t0) / steps * 1e3
    print(f"{mode:18s} {dt:.3f} ms/step  {g.mp / dt / 1e3:.1f} M tile-steps/s", flush=True)
    h.close()
for mode in ("step", "met", "post", "met+post", "out", "met+post+out"):
    run(mode)
