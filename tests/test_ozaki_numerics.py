"""CPU check of the error-free INT8 splitting behind i8_loss.cu (tools/ozaki_numerics.py): with 7 signed 7-bit slices per operand
and the slice products p + q <= 6 accumulated exactly per level, Y = W X is reproduced to ~1e-13 of max|y| (parity bar 1e-10),
every level sum fits the s32 accumulator, and the same scheme on the sample contraction of the gradient pass stays below 1e-12."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import ozaki_numerics as oz


def test_seven_slices_reproduce_the_f64_product():
    rng = np.random.default_rng(1)
    n, t = 128, 512
    x = rng.standard_normal((n, n)) @ np.vstack([rng.laplace(size=(n // 2, t)), rng.uniform(-1.7, 1.7, size=(n - n // 2, t))]) / np.sqrt(n)
    w = np.linalg.qr(rng.standard_normal((n, n)))[0]
    idx = [(int(rng.integers(n)), int(rng.integers(t))) for _ in range(60)]
    ref = oz.exact_entries(w, x, idx)
    scale = np.max(np.abs(w @ x))
    err = {}
    for s in (5, 7, 8):
        y = oz.ozaki_matmul(w, x, s)   # asserts |level sum| < 2^31 inside
        err[s] = np.max(np.abs(np.array([y[i, tt] for i, tt in idx]) - ref)) / scale
    assert err[7] < 1e-12 and err[8] < 2e-15 and err[5] > err[7]


def test_slices_are_int8_and_residuals_exact():
    rng = np.random.default_rng(2)
    a = rng.standard_normal((16, 128)) * np.exp2(rng.integers(-20, 20, size=(16, 1)))
    sl, e = oz.slices(a, 1, 7)
    assert all(np.max(np.abs(q)) <= 127 for q in sl)
    back = sum(q.astype(np.float64) * 2.0 ** (-7 * (p + 1)) for p, q in enumerate(sl)) * np.exp2(e)
    assert np.max(np.abs(back - a) / np.max(np.abs(a), axis=1, keepdims=True)) < 2.0 ** -46  # residual < 2^-49 of the scaled value, scale <= 4 max|row|


def test_sample_contraction_of_the_gradient_pass():
    rng = np.random.default_rng(3)
    n, t = 32, 4096
    y = np.vstack([rng.laplace(size=(n // 2, t)) / np.sqrt(2.0), rng.uniform(-1.7, 1.7, size=(n - n // 2, t))])
    psi = np.tanh(y)
    ref = (psi.astype(np.longdouble) @ y.astype(np.longdouble).T).astype(np.float64)
    g = oz.ozaki_gram(psi, y, 64, 7)
    assert np.max(np.abs(g - ref)) / np.max(np.abs(ref)) < 1e-12
