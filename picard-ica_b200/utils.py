"""Evaluation utilities of the reference (utils.rs:16-103): `permute` and `amari_distance`.

These are N x N host-side metrics in the reference too (out of scope for the GPU); they are
restated here with numpy because the Amari distance is the end-to-end parity metric.
"""
from __future__ import annotations

import numpy as np

__all__ = ["permute", "amari_distance"]


def permute(a, scale: bool = True) -> np.ndarray:
    """utils.rs:16-68: swap rows until the diagonal dominates, optionally scale rows to a unit diagonal,
    then order rows/columns by absolute column sum."""
    a = np.array(a, dtype=np.float64, copy=True)
    n = a.shape[0]
    done = False
    while not done:
        done = True
        for i in range(n):
            for j in range(i):
                if a[i, i] ** 2 + a[j, j] ** 2 < a[i, j] ** 2 + a[j, i] ** 2:
                    a[[i, j], :] = a[[j, i], :]
                    done = False
    if scale:
        for i in range(n):
            d = a[i, i]
            if abs(d) > 1e-10:
                a[i, :] /= d
    col_sums = np.abs(a[:n, :n]).sum(axis=0)
    order = np.argsort(col_sums, kind="stable")  # slice::sort_by is stable
    return a[np.ix_(order, order)]


def amari_distance(w, a) -> float:
    """utils.rs:82-103: 0 when `w @ a` is a scaled permutation."""
    p = np.abs(np.asarray(w, dtype=np.float64) @ np.asarray(a, dtype=np.float64))
    n = p.shape[0]

    def s(r):
        sq = r * r
        mx = sq.max(axis=1)
        ok = mx > 1e-15
        return float(np.sum(sq.sum(axis=1)[ok] / mx[ok] - 1.0))

    return (s(p) + s(p.T)) / (2.0 * n)
