"""Per-step deviation of the device from the Fortran run of the ten-day golden case (tests/golden/fortran_cbm_v1.npz,
leuning_ten_days): which traced fields differ at all, where they first differ, where most.  usage (GPU box): python
tests/checks/tenday_trace.py   [CABLE_B200_LIB=... for a variant build]"""
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tests/golden')
import numpy as np
import make_fortran_golden as G
from cable_b200.cbm import CableB200
from util import field_errors
z = np.load('tests/golden/fortran_cbm_v1.npz')
case = 'leuning_ten_days'
nland, nsteps, doy, dels, site_lat, sw = G.CASES[case]
cfg, grid, T, F = G.case_inputs(case)
worst = {}
with CableB200(grid.mp, cfg) as h:
    h.bind(T); h.upload_params(); h.upload_state()
    for k in range(nsteps):
        G.caller_step(case, T, F, k); h.cbm(k + 1, dels)
        for n in G.TRACE:
            want = z[f"{case}/trace/{n}"][k]
            mx, tol, rel = field_errors(want, T[n], want.dtype.type)
            if mx > 0: worst.setdefault(n, []).append((k + 1, float(mx), int(np.argmax(rel.max(0) if rel.ndim > 1 else rel)), float((rel > tol).mean())))
for n, v in worst.items():
    print(n, 'first', v[0], 'max', max(v, key=lambda t: t[1]), 'steps differing', len(v), 'steps over tol', sum(1 for t in v if t[3] > 0))
