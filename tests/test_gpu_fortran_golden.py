"""GPU: the CUDA path, through the C ABI, against golden vectors produced by the reference's own Fortran source
(tests/golden/make_fortran_golden.py: /root/reference's cbm executed by the interpreter in oracle/frun).  No oracle in
between: this is device vs reference.  Tolerances are the north star's: 1e-4 relative for binary32 fields, 1e-6 for
binary64 fields, per element against max(|a|, |b|, 1e-3 x field scale)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_fortran_golden as G            # noqa: E402
from cable_b200.cbm import CableB200        # noqa: E402
from cable_b200.registry import FIELDS      # noqa: E402
from util import field_errors               # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(HERE, "golden", "fortran_cbm_v1.npz")


@pytest.mark.parametrize("case", list(G.CASES))
def test_device_matches_the_fortran_run(case):
    z = np.load(GOLD)
    nland, nsteps, doy, dels, site_lat, sw = G.CASES[case]
    cfg, grid, T, F = G.case_inputs(case)
    worst, nbit, ntot = 0.0, 0, 0
    with CableB200(grid.mp, cfg) as h:
        h.bind(T); h.upload_params(); h.upload_state()
        for k in range(nsteps):
            G.caller_step(case, T, F, k)
            h.cbm(k + 1, dels)
            for n in G.TRACE:
                want = z[f"{case}/trace/{n}"][k]
                mx, tol, _ = field_errors(want, T[n], want.dtype.type)
                assert mx <= tol, (case, k + 1, n, mx)
    for f in FIELDS:
        key = f"{case}/final/{f.name}"
        if key not in z.files or f.name in ("met_tvair", "canopy_oldcansto") and case == "caller_inputs":
            continue
        want = z[key]
        assert np.all(np.isfinite(T[f.name])), f.name
        mx, tol, _ = field_errors(want, T[f.name], f.dtype)
        assert mx <= tol, (case, "final", f.name, mx)
        worst = max(worst, mx)
        ntot += 1; nbit += int(np.array_equal(want, T[f.name]))
    print(f"{case}: worst relative difference vs the Fortran run {worst:.2e}; {nbit}/{ntot} fields bit-identical")
    assert ntot >= 170 and worst < 1e-5
