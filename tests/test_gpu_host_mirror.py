"""GPU: the C++ host-side mirror (cable_b200/csrc/host_mirror.hpp -- same argument list as the reference cbm,
same call sequence as the Fortran shim) driven by a C++ test double of the Fortran caller."""
import os
import subprocess

import numpy as np
import pytest

from cable_b200 import lib, synth
from cable_b200.registry import FIELDS, ROLE, FLAG
from oracle.pyoracle import Oracle
from util import DELS, compare_tiles, make_case, output_fields

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_mirror")


def test_cpp_caller_matches_oracle(tmp_path):
    if not os.path.exists(EXE):
        import __graft_entry__
        __graft_entry__.build()
    cfg, grid, T, F = make_case(150)
    nsteps = 4
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as fh:
        fh.write(np.array([grid.mp, nsteps], dtype=np.int32).tobytes())
        for f in FIELDS:
            fh.write(T[f.name].tobytes())
        Tf = {k: v.copy() for k, v in T.items()}
        for k in range(nsteps):
            F.fill(Tf, k)
            for f in FIELDS:
                if f.role == ROLE["FORCING"] and not (f.flags & FLAG["OPTIN"]):
                    fh.write(Tf[f.name].tobytes())
    r = subprocess.run([EXE, str(fin), str(fout)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "host_mirror ok" in r.stdout
    got = {}
    with open(fout, "rb") as fh:
        for f in FIELDS:
            got[f.name] = np.frombuffer(fh.read(T[f.name].nbytes), dtype=f.dtype).reshape(T[f.name].shape)
    o = Oracle(T, cfg, cr_math=True)
    for k in range(nsteps):
        F.fill(T, k)
        o.cbm(k + 1, DELS)
    star = [f for f in output_fields() if f.role == ROLE["STATE"] or f.flags & FLAG["STAR"]]
    res = compare_tiles(T, got, star)
    bad = {n: r for n, r in res.items() if r[0] > r[1]}
    assert not bad, bad


def test_cpp_caller_with_casa_matches_the_fortran_run(tmp_path):
    """The same C++ test double playing serialdrv's CASA-CNP part too: after every cbm() it calls
    bgcdriver_mod::bgc_device::bgcdriver(...) -- the reference's 25 arguments, C++ mirrors of casa_biome ... phen_variable
    generated from the CASA registry -- for two model days, against the golden vectors of the reference's Fortran bgcdriver
    (tests/golden/make_casa_golden.py, case drv_cnp: pinned-oracle cbm + Fortran bgcdriver)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_casa_golden as G
    from cable_b200 import casa
    if not os.path.exists(EXE):
        import __graft_entry__
        __graft_entry__.build()
    z = np.load(os.path.join(ROOT, "tests", "golden", "fortran_casa_v1.npz"))
    case = "drv_cnp"
    cfg, grid, T, F = make_case(G.NLAND, start_doy=G.DOY)
    ccfg = G.casa_cfg(G.DRV[case])
    A = casa.synth_casa(grid, T, ccfg, seed=31)
    silt, clay = casa.soil_texture(T)
    nsteps = 16
    fin, fout, cin, cout = (tmp_path / n for n in ("in.bin", "out.bin", "casa_in.bin", "casa_out.bin"))
    with open(fin, "wb") as fh:
        fh.write(np.array([grid.mp, nsteps], dtype=np.int32).tobytes())
        for f in FIELDS:
            fh.write(T[f.name].tobytes())
        Tf = {k: v.copy() for k, v in T.items()}
        for k in range(nsteps):
            F.fill(Tf, k)
            for f in FIELDS:
                if f.role == ROLE["FORCING"] and not (f.flags & FLAG["OPTIN"]):
                    fh.write(Tf[f.name].tobytes())
    with open(cin, "wb") as fh:
        fh.write(np.array([ccfg.icycle, ccfg.lalloc, ccfg.mvtype, 8, G.DOY], dtype=np.int32).tobytes())
        for f in casa.FIELDS:
            fh.write(A[f.name].tobytes())
        fh.write(np.ascontiguousarray(silt, np.float32).tobytes()); fh.write(np.ascontiguousarray(clay, np.float32).tobytes())
    r = subprocess.run([EXE, str(fin), str(fout), str(cin), str(cout)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    checked = 0
    with open(cout, "rb") as fh:
        for f in casa.FIELDS:
            got = np.frombuffer(fh.read(A[f.name].nbytes), dtype=f.dtype).reshape(A[f.name].shape)
            key = f"drv/{case}/{f.name}"
            if key not in z.files or f.dtype == np.int32:
                continue
            want = z[key]
            assert np.array_equal(np.isnan(got), np.isnan(want)), f.name
            a, b = np.nan_to_num(want.astype(np.float64)), np.nan_to_num(got.astype(np.float64))
            if f.name.startswith("casabal_") and "balance" in f.name or f.name in ("casabal_sumcbal", "casabal_sumnbal", "casabal_sumpbal"):
                assert float(np.abs(a - b).max()) < 1e-6, f.name
            else:
                floor = 1e-6 * max(float(np.abs(a).max()), 1e-300)
                rel = float((np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)).max())
                assert rel <= 2e-5, (f.name, rel)
            checked += 1
    assert checked > 150
