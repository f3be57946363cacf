"""Synthetic force tables of the reference's shape.

The reference's real table (``./data/fit_force.txt``, README.md:53-55) is not
in the repository.  It is a list of ``force_grid_length`` floats indexed by the
SQUARED pair distance (reference simulation.c:217-218, main.c:150):
entry ``i`` belongs to ``d_i = force_grid_start + force_grid * i``.

``T-ref`` is attractive, ``F(d) = -A (sqrt(d)/6)^-p``
with ``A = 1e-6``, ``p = 4``.  For even ``p`` this is ``-A (36/d)^(p/2)``, which
is evaluated here with plain IEEE double multiplications and one division so
that every implementation (numpy here, C in host/heroam_table.c) produces the
same float32 bits without depending on a libm ``pow``.
"""
from __future__ import annotations

import numpy as np

TREF_LENGTH = 1_596_000
TREF_START = 4.0
TREF_GRID = 0.001


def synthetic_table(
    length: int = TREF_LENGTH,
    start: float = TREF_START,
    grid: float = TREF_GRID,
    amplitude: float = 1.0e-6,
    half_power: int = 2,
) -> np.ndarray:
    """float32 table ``-amplitude * (36/d_i)^half_power``; T-ref by default."""
    d = np.float64(start) + np.float64(grid) * np.arange(length, dtype=np.float64)
    ratio = np.float64(36.0) / d
    v = np.full(length, np.float64(amplitude))
    for _ in range(int(half_power)):
        v = v * ratio
    return (-v).astype(np.float32)


def write_force_file(path: str, table: np.ndarray, start: float = TREF_START, grid: float = TREF_GRID) -> None:
    """The reference's text format: 4 tab-separated columns per row, the table
    value in column 2 (reference main.c:136).  T-ref in this format is ~102 MB,
    the "exceeds 100MB" file of README.md:55."""
    with open(path, "w") as f:
        for i, v in enumerate(table):
            f.write("%.9e\t%.9e\t%.9e\t%.9e\n" % (start + grid * i, v, 0.0, 0.0))
