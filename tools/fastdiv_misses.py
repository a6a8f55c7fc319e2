"""How many kernel-A blocks the CBL_FASTDIV build hands back over a long run, and (with a -DCBL_FASTDIV_DEBUG library in
CABLE_B200_LIB and DUMP=1) which dv() sites and operands raise the flag:  python tools/fastdiv_misses.py [nland] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cable_b200 import lib, synth
from cable_b200.cbm import CableB200
nland = int(sys.argv[1]) if len(sys.argv) > 1 else 62000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 240
cfg = lib.default_cfg(); cfg.n_forcing_slots = 8; cfg.output_level = 1
g = synth.make_grid(nland, 5); T = synth.make_tiles(g, cfg); F = synth.Forcing(g, T, 10800.0, start_doy=172)
h = CableB200(g.mp, cfg); h.bind(T); h.upload_params(); h.upload_state()
for k in range(8):
    F.fill(T, k); h.set_forcing_async(k); h.sync()
last = 0
for k in range(steps):
    if os.environ.get("DUMP") and k == steps - 2: os.environ["CABLE_B200_FASTDIV_DEBUG"] = "1"
    h.step(k + 1, 10800.0, k % 8)
    if k % 40 == 39 or k == steps - 1:
        h.sync(); c = h.counters()
        print(f"steps {k + 1}: fastdiv redo blocks so far {c.n_fastdiv_redo_blocks} (+{c.n_fastdiv_redo_blocks - last})", flush=True); last = c.n_fastdiv_redo_blocks
