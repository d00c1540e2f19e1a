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


# ---------------------------------------------------------------------------------------------------------------
# The gradient pass Gr = psi(Y) Y^T contracts over SAMPLES (K = T): both operands change every iteration and must be
# sliced in the kernel, with one scale per (row, block of B samples); the level sums of a block are exact in s32 as long
# as (d + 1) B 127^2 < 2^31 and are flushed to f64 per block.  How do heavy-tailed rows (Laplace sources: a block's
# maximum is several times its typical entry) affect the error for S = 7 / 8 slices?  Reference: 80-bit long double.
# ---------------------------------------------------------------------------------------------------------------
def slices_blocked(a, block, s):
    """a (n x t): every row scaled per block of `block` columns; returns s int slices and the (n x t/block) exponents."""
    n, t = a.shape
    ab = a.reshape(n, t // block, block)
    m = np.max(np.abs(ab), axis=2, keepdims=True)
    e = np.where(m > 0, np.floor(np.log2(np.where(m > 0, m, 1.0))) + 1, 0.0)
    r = ab / np.exp2(e)
    out = []
    for _ in range(s):
        r = r * 128.0
        q = np.trunc(r)
        out.append(q.astype(np.int64))
        r = r - q
    return out, e[:, :, 0]


def ozaki_gram(psi, y, block, s):
    n, t = y.shape
    ps, ep = slices_blocked(psi, block, s)
    ys, ey = slices_blocked(y, block, s)
    g = np.zeros((n, n))
    for b in range(t // block):
        gb = np.zeros((n, n))
        for level in range(s - 1, -1, -1):
            acc = np.zeros((n, n), dtype=np.int64)
            for p in range(level + 1):
                acc += ps[p][:, b, :] @ ys[level - p][:, b, :].T
            assert np.max(np.abs(acc)) < 2 ** 31
            gb += acc.astype(np.float64) * 2.0 ** (-7 * (level + 2))
        g += gb * np.exp2(ep[:, b])[:, None] * np.exp2(ey[:, b])[None, :]
    return g



def main():
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


    t2 = 1 << 15
    s_lap = rng.laplace(size=(n // 2, t2)) / np.sqrt(2.0)
    s_uni = rng.uniform(-np.sqrt(3.0), np.sqrt(3.0), size=(n - n // 2, t2))
    yy = np.vstack([s_lap, s_uni]) + 0.05 * rng.standard_normal((n, t2))   # nearly separated sources: the hard case (independent rows)
    pp = np.tanh(yy)
    ref_g = (pp.astype(np.longdouble) @ yy.astype(np.longdouble).T).astype(np.float64)
    scale_g = np.max(np.abs(ref_g))
    print(f"gram f64 dgemm            max|err|/max|G| = {np.max(np.abs(pp @ yy.T - ref_g)) / scale_g:.2e}")
    for block in (64, 1024):
        for s in (7, 8):
            gg = ozaki_gram(pp, yy, block, s)
            print(f"gram int8 slices={s} block={block:5d} max|err|/max|G| = {np.max(np.abs(gg - ref_g)) / scale_g:.2e}")


if __name__ == "__main__":
    main()
