"""Kernel-only timing for quick iteration: python tools/quick_perf.py [nland] [steps] [block]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cable_b200 import lib, synth
from cable_b200.cbm import CableB200
nland = int(sys.argv[1]) if len(sys.argv) > 1 else 62000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
block = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg = lib.default_cfg(); cfg.n_forcing_slots = 8; cfg.output_level = 1; cfg.threads_per_block = block
g = synth.make_grid(nland, 5); T = synth.make_tiles(g, cfg); F = synth.Forcing(g, T, 10800.0, start_doy=172)
h = CableB200(g.mp, cfg); h.bind(T); h.upload_params(); h.upload_state()
for k in range(8):
    F.fill(T, k); h.set_forcing_async(k); h.sync()
for k in range(8): h.step(k + 1, 10800.0, k % 8)
h.sync(); h.reset_counters(); h.profile(True)
t0 = time.perf_counter()
for k in range(8, 8 + steps): h.step(k + 1, 10800.0, k % 8)
h.sync(); dt = time.perf_counter() - t0
c = h.counters()
print(f"mp={g.mp} steps={steps} block={block or 128}: kernel {c.kernel_ms / c.kernel_ms_count:.3f} ms/step, wall {dt / steps * 1e3:.3f} ms/step, "
      f"{g.mp * steps / dt / 1e6:.1f} M tile-steps/s, warns={c.n_dryleaf_warn}, fastdiv redo blocks={c.n_fastdiv_redo_blocks}")
