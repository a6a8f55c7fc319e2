"""Regenerate tests/golden/oracle_cr_v1.npz:  python tests/golden/make_golden.py

The reference ships no golden vectors for cbm() and cannot be built here (no Fortran compiler), so these
vectors do NOT come from the reference binary -- they freeze the outputs of the correctly-rounded oracle
build (oracle/liboracle_cr.so; bit-reproducible across hosts because every fp32 intrinsic is an fp64
evaluation rounded once) on a small seeded case, so that any later change to the oracle, the synthetic
generator or the device path is caught.  Two cable_user configurations: cable.nml (leuning) and code
defaults (medlyn)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from cable_b200 import lib  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402
from util import make_case, output_fields, DELS  # noqa: E402

NLAND, NSTEPS = 24, 16


def run(gs_switch: int):
    cfg = lib.default_cfg()
    cfg.gs_switch = gs_switch
    cfg, grid, tiles, forcing = make_case(NLAND, cfg=cfg, start_doy=100)
    o = Oracle(tiles, cfg, cr_math=True)
    acc = {}
    for k in range(NSTEPS):
        forcing.fill(tiles, k)
        o.cbm(k + 1, DELS)
        for n in ("canopy_fe", "canopy_fh", "canopy_fpn", "ssnow_runoff", "canopy_fes", "rad_swnet"):
            acc[n] = acc.get(n, 0.0) + tiles[n].astype(np.float64)
    out = {f"final/{f.name}": tiles[f.name].copy() for f in output_fields()}
    out.update({f"sum/{n}": a for n, a in acc.items()})
    return out


if __name__ == "__main__":
    data = {}
    for gs, tag in ((0, "leuning"), (1, "medlyn")):
        for k, v in run(gs).items():
            data[f"{tag}/{k}"] = v
    path = os.path.join(HERE, "oracle_cr_v1.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes,", len(data), "arrays")
