"""Seeded synthetic inputs shared by the tests, the smoke test and bench.py."""
from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1


def lcg_bench_data(n_features: int, n_samples: int, seed: int = 42) -> np.ndarray:
    """The reference's own bench generator, bit for bit (benches/benchmarks.rs:8-35): a 64-bit LCG
    `state = state * 6364136223846793005 + 1`, u = (state >> 33) / 2^31, Laplace sources by inverse CDF,
    mixing entries u - 0.5, X = mixing @ sources.  Pure integer arithmetic, vectorised by jump-ahead."""
    a, c = 6364136223846793005, 1
    total = n_features * n_samples + n_features * n_features
    # state_k = A_k * seed + C_k with A_k = a^k, C_k = c (a^k - 1)/(a - 1) mod 2^64, built by doubling
    A = np.empty(total, dtype=np.uint64)
    Cc = np.empty(total, dtype=np.uint64)
    A[0], Cc[0] = a, c
    filled = 1
    with np.errstate(over="ignore"):
        while filled < total:
            m = min(filled, total - filled)
            # (A_j, C_j) o (A_filled, C_filled): state_{filled + j} = A_j * state_filled + C_j
            Af, Cf = A[filled - 1], Cc[filled - 1]
            A[filled:filled + m] = A[:m] * Af
            Cc[filled:filled + m] = A[:m] * Cf + Cc[:m]
            filled += m
        states = A * np.uint64(seed & _M64) + Cc
    u = (states >> np.uint64(33)).astype(np.float64) / float(1 << 31)
    us = u[: n_features * n_samples].reshape(n_features, n_samples)
    with np.errstate(divide="ignore"):
        src = np.where(us < 0.5, np.log(2.0 * us), -np.log(2.0 * (1.0 - us)))
    mixing = (u[n_features * n_samples:] - 0.5).reshape(n_features, n_features)
    return mixing @ src


def sources(n: int, t: int, seed: int, kind: str = "mixed") -> np.ndarray:
    """Unit-variance sources: Laplace(b=1/sqrt 2), uniform[-sqrt 3, sqrt 3], or mixed (first ceil(n/2) Laplace)."""
    rng = np.random.default_rng(seed)
    s = np.empty((n, t))
    n_lap = {"laplace": n, "uniform": 0, "mixed": (n + 1) // 2}[kind]
    if n_lap:
        s[:n_lap] = rng.laplace(scale=1.0 / np.sqrt(2.0), size=(n_lap, t))
    if n_lap < n:
        s[n_lap:] = rng.uniform(-np.sqrt(3.0), np.sqrt(3.0), size=(n - n_lap, t))
    return s


def mixture(n: int, t: int, seed: int = 0, kind: str = "mixed"):
    """(X = A S, A, S) with A i.i.d. N(0,1)."""
    s = sources(n, t, seed, kind)
    a = np.random.default_rng(seed + 1000003).standard_normal((n, n))
    return a @ s, a, s


def orthogonal(n: int, seed: int) -> np.ndarray:
    """QR-orthogonalised N(0,1) matrix (the explicit w_init of every parity run)."""
    q, r = np.linalg.qr(np.random.default_rng(seed).standard_normal((n, n)))
    return q * np.sign(np.diag(r))


def whitened(n: int, t: int, seed: int = 0, kind: str = "mixed") -> np.ndarray:
    """A centred, exactly whitened mixture rotated by a random orthogonal matrix: what core::run sees."""
    x, _, _ = mixture(n, t, seed, kind)
    x = x - x.mean(axis=1, keepdims=True)
    d, e = np.linalg.eigh(x @ x.T / t)
    k = (e / np.sqrt(d)).T
    return orthogonal(n, seed + 7) @ (k @ x)


def rel_err(a, b) -> float:
    """max|a - b| / max|b|  (the tolerance definition of SURVEY.md §7 / BASELINE north_star)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = float(np.max(np.abs(b))) if b.size else 0.0
    num = float(np.max(np.abs(a - b))) if b.size else 0.0
    return num / den if den > 0 else num
