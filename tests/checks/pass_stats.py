"""Distribution of dryLeaf passes per tile and stability iteration (oracle + the NumPy restatement's pass counter): what the
warps of kernel A wait for.  python tests/checks/pass_stats.py [nland] [steps]"""
import os, sys
import ctypes as C
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cable_b200 import lib
from oracle.pyoracle import Oracle
from util import make_case, DELS
from np_dryleaf import dryleaf
import test_oracle_numpy_xcheck as X

nland = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = lib.default_cfg()
cfg, grid, T, F = make_case(nland, cfg=cfg, start_doy=172)
o = Oracle(T, cfg, cr_math=True)
mp = grid.mp
cap = []
def view(ptr, dtype, ncol):
    n = mp * ncol
    buf = (C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    a = np.frombuffer(buf, dtype=dtype, count=n)
    return a.reshape(mp, ncol).copy() if ncol > 1 else a.copy()
HOOK = C.CFUNCTYPE(None, C.c_int, C.c_int, C.POINTER(C.c_void_p))
def hook(when, it, work):
    if when == 0:
        snap = {n: T[n].copy() for n in X.DRYLEAF_FIELDS_IN}
        for j, (n, dt, nc) in enumerate(X.DRYLEAF_WORK):
            snap["w_" + n] = view(work[j], dt, nc)
        cap.append((it, snap))
cb = HOOK(hook)
o._lib.oracle_set_dryleaf_hook.argtypes = [C.c_void_p, HOOK]; o._lib.oracle_set_dryleaf_hook.restype = None
o._lib.oracle_set_dryleaf_hook(o._h, cb)
NP = []
for k in range(steps):
    F.fill(T, k); cap.clear(); o.cbm(k + 1, DELS)
    NP.append(np.stack([dryleaf(DELS, it, False, inp, False, 0)["npass"] for it, inp in cap]))
NP = np.stack(NP)            # steps, 4, mp
np.save("/tmp/npass.npy", NP)
veg = NP.max((0, 1)) > 0
print("tiles", mp, "vegetated", int(veg.sum()))
h = np.bincount(NP.ravel(), minlength=21)
print("passes histogram (tile-calls):", {k: int(v) for k, v in enumerate(h) if v})
tot = NP.sum()
for B in (32, 640):
    n = mp // B * B
    M = NP[:, :, :n].reshape(steps, 4, n // B, B)
    print(f"group {B}: sum of max over group x size = {int(M.max(-1).sum() * B)}, tile-passes {int(NP[:, :, :n].sum())}, efficiency {NP[:, :, :n].sum() / (M.max(-1).sum() * B):.3f}")
for cut in (2, 3, 4, 5):
    # main loop runs min(passes, cut) per warp; stragglers (passes > cut) of a 640-block are re-packed
    n = mp // 640 * 640
    A = NP[:, :, :n]
    W = A.reshape(steps, 4, n // 32, 32)
    main = np.minimum(W.max(-1), cut).sum() * 32
    S = A.reshape(steps, 4, n // 640, 640)
    extra = 0
    for k in range(cut + 1, 21):
        extra += (np.ceil((S >= k).sum(-1) / 32) * 32).sum()
    print(f"cut {cut}: lane-passes main {int(main)} + repacked stragglers {int(extra)} = {int(main + extra)} vs now {int(W.max(-1).sum() * 32)}; stragglers/block mean {(S > cut).sum(-1).mean():.1f} max {(S > cut).sum(-1).max()}")
