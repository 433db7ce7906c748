"""Linear algebra over GF(2) on plain uint8 arrays (the reference leans on the `galois` package for
this: mentpy/calculator/linalg.py, mbqc/flow.py:216-229)."""
from typing import Optional

import numpy as np


def _reduce(aug: np.ndarray, n_cols: int):
    """Row-reduce `aug` in place over its first n_cols columns; returns the pivot column of every pivot row."""
    pivots, row = [], 0
    for col in range(n_cols):
        hits = np.nonzero(aug[row:, col])[0]
        if hits.size == 0:
            continue
        p = row + hits[0]
        if p != row:
            aug[[row, p]] = aug[[p, row]]
        others = np.nonzero(aug[:, col])[0]
        others = others[others != row]
        aug[others] ^= aug[row]
        pivots.append(col)
        row += 1
        if row == aug.shape[0]:
            break
    return pivots


def gf2_solve(a, b) -> Optional[np.ndarray]:
    """One solution x of a x = b over GF(2) (free variables 0), or None when there is none."""
    a = np.asarray(a, dtype=np.uint8) & 1
    b = (np.asarray(b, dtype=np.uint8) & 1).reshape(-1)
    aug = np.concatenate([a, b[:, None]], axis=1)
    pivots = _reduce(aug, a.shape[1])
    if np.any(aug[len(pivots):, -1]):
        return None
    x = np.zeros(a.shape[1], dtype=np.uint8)
    for r, c in enumerate(pivots):
        x[c] = aug[r, -1]
    return x


def gf2_rank(a) -> int:
    a = (np.asarray(a, dtype=np.uint8) & 1).copy()
    return len(_reduce(a, a.shape[1])) if a.size else 0
