"""Land-point block decomposition across ranks / GPUs.

Rule of the reference: contiguous blocks of land points whose sizes differ by at most one,
the first `mland mod nranks` blocks being the larger ones
(`array_partition`, src/util/cable_array_utils.F90:48-75; legacy `master_decomp`,
src/offline/cable_mpimaster.F90:1428-1445).  Tiles of one land point stay on one rank so the
patch -> grid-cell reduction is local (cable_mpimaster.F90:1454-1463).
"""
from __future__ import annotations

import numpy as np


def array_partition(n: int, k: int, i: int) -> tuple[int, int]:
    """(start, count) of block i (0-based) when n items are split into k contiguous blocks."""
    if k <= 0 or not (0 <= i < k) or n < 0:
        raise ValueError("bad partition request")
    base, rem = divmod(n, k)
    count = base + (1 if i < rem else 0)
    start = i * base + min(i, rem)
    return start, count


def land_to_tile_range(cstart: np.ndarray, cend: np.ndarray, l0: int, nl: int) -> tuple[int, int]:
    """Tile range [t0, t1) owned by land points l0 .. l0+nl-1 (cstart/cend are 0-based, inclusive)."""
    if nl == 0:
        return 0, 0
    return int(cstart[l0]), int(cend[l0 + nl - 1]) + 1
