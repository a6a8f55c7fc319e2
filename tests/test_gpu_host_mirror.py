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
