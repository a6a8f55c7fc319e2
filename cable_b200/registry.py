"""Field registry of the cbm() hot path, parsed from include/cable_b200_fields.def.

The .def file is the single source of truth shared by the C ABI, the CUDA kernel
and the oracle; it mirrors the members of the reference derived types
(reference: src/offline/cable_define_types.F90:79-717).
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass

import numpy as np

ROLE = {"FORCING": 1, "PARAM": 2, "STATE": 4, "DIAG": 8}
FLAG = {"STAR": 1, "COND": 2, "HOSTONLY": 4, "OPTIN": 8, "XCH": 16, "PHB": 32, "STA": 64}
DTYPE = {"float": np.float32, "double": np.float64, "int": np.int32}

_DEF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "cable_b200_fields.def")


@dataclass(frozen=True)
class Field:
    id: int
    name: str          # "<type>_<member>", e.g. "ssnow_tgg"
    type: str          # reference derived type: met, air, veg, soil, ssnow, canopy, rad, rough, bal, bgc, scr
    member: str
    dtype: type
    n1: int
    n2: int
    role: int
    flags: int

    @property
    def ncomp(self) -> int:
        return self.n1 * self.n2

    def star(self) -> bool:
        return bool(self.flags & FLAG["STAR"])


def _parse_flags(txt: str) -> int:
    v = 0
    for tok in txt.split("|"):
        tok = tok.strip()
        if tok and tok != "0":
            v |= FLAG[tok]
    return v


def load_fields(path: str = _DEF) -> list[Field]:
    rx1 = re.compile(r"^CABLE_F1\(\s*(\w+)\s*,\s*(\w+)\s*,\s*(\w+)\s*,\s*(\w+)\s*,\s*([\w|]+)\s*\)")
    rxa = re.compile(r"^CABLE_FA\(\s*(\w+)\s*,\s*(\w+)\s*,\s*(\w+)\s*,\s*(\d+)\s*,\s*(\d+)\s*,\s*(\w+)\s*,\s*([\w|]+)\s*\)")
    out: list[Field] = []
    with open(path) as fh:
        for line in fh:
            m = rx1.match(line)
            if m:
                t, mem, ct, role, fl = m.groups()
                out.append(Field(len(out), f"{t}_{mem}", t, mem, DTYPE[ct], 1, 1, ROLE[role], _parse_flags(fl)))
                continue
            m = rxa.match(line)
            if m:
                t, mem, ct, n1, n2, role, fl = m.groups()
                out.append(Field(len(out), f"{t}_{mem}", t, mem, DTYPE[ct], int(n1), int(n2), ROLE[role], _parse_flags(fl)))
    return out


FIELDS: list[Field] = load_fields()
BY_NAME: dict[str, Field] = {f.name: f for f in FIELDS}


def alloc_tiles(mp: int, fill_nan_diag: bool = False) -> dict[str, np.ndarray]:
    """One zero-initialised array per field, shaped (ncomp, mp) C-order == Fortran (mp,n1,n2)."""
    arrs = {}
    for f in FIELDS:
        a = np.zeros((f.ncomp, mp), dtype=f.dtype)
        arrs[f.name] = a
    return arrs
