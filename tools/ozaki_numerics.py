"""CPU experiment for DESIGN.md section 7: how many signed 7-bit slices does an error-free (Ozaki-type) INT8 evaluation of the
pass contraction Y = W X (K = N = 128) need to meet the parity bar?  Rows of W and columns (samples) of X are scaled by their
own power of two, cut into s int8 slices, the slice products with p + q <= s - 1 are accumulated exactly in int32 per level
p + q (what an INT8 tensor-core accumulator does) and combined in f64.  Reference: exact rational arithmetic via fractions on a
sample of entries.  Run: python tools/ozaki_numerics.py"""
import fractions

import numpy as np


def slices(a, axis, s):
    """a scaled per row (axis=1) or per column (axis=0) to (-1, 1), cut into s slices of 7 bits (truncation: exact residuals)."""
    m = np.max(np.abs(a), axis=axis, keepdims=True)
    e = np.where(m > 0, np.ceil(np.log2(np.where(m > 0, m, 1.0))) + 1, 0.0)
    r = a / np.exp2(e)
    out = []
    for _ in range(s):
        r = r * 128.0
        q = np.trunc(r)
        out.append(q.astype(np.int8).astype(np.int32))
        r = r - q
    return out, e


def ozaki_matmul(w, x, s):
    ws, ew = slices(w, 1, s)
    xs, ex = slices(x, 0, s)
    y = np.zeros((w.shape[0], x.shape[1]))
    for level in range(s - 1, -1, -1):  # small terms first
        acc = np.zeros((w.shape[0], x.shape[1]), dtype=np.int64)
        for p in range(level + 1):
            acc += ws[p] @ xs[level - p]   # |acc| <= (level + 1) K 127^2 < 2^31 for K = 128, s <= 8: exact in int32
        assert np.max(np.abs(acc)) < 2 ** 31
        y += acc.astype(np.float64) * 2.0 ** (-7 * (level + 2))
    return y * np.exp2(ew) * np.exp2(ex)


def exact_entries(w, x, idx):
    out = []
    for i, t in idx:
        out.append(float(sum(fractions.Fraction(float(w[i, k])) * fractions.Fraction(float(x[k, t])) for k in range(w.shape[1]))))
    return np.array(out)


rng = np.random.default_rng(0)
n, t = 128, 4096
lap = rng.laplace(size=(n // 2, t)) / np.sqrt(2.0)
uni = rng.uniform(-np.sqrt(3.0), np.sqrt(3.0), size=(n - n // 2, t))
x = rng.standard_normal((n, n)) @ np.vstack([lap, uni]) / np.sqrt(n)
w = np.linalg.qr(rng.standard_normal((n, n)))[0]
idx = [(int(rng.integers(n)), int(rng.integers(t))) for _ in range(400)]
ref = exact_entries(w, x, idx)
scale = np.max(np.abs(w @ x))
f64 = np.array([(w @ x)[i, tt] for i, tt in idx])
print(f"f64 dgemm      max|err|/max|y| = {np.max(np.abs(f64 - ref)) / scale:.2e}")
for s in (4, 5, 6, 7, 8):
    y = ozaki_matmul(w, x, s)
    got = np.array([y[i, tt] for i, tt in idx])
    print(f"int8 slices={s}  products={s * (s + 1) // 2:2d}  max|err|/max|y| = {np.max(np.abs(got - ref)) / scale:.2e}")
