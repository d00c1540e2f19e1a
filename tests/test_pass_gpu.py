"""GPU parity of the fused pass (K1/K2): raw moments Gr, Sd, Hr, Sq, L at Y = W X from the CUDA kernel, through
the C ABI (picard_eval_moments), against the CPU oracle.  Tolerance (BASELINE.json north_star): per-pass
quantities from identical W agree to <= 1e-10 relative, measured as max|delta| / max|ref|."""
import numpy as np
import pytest

import _data
import _gpu
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10

DENS = [(orc.TANH, 1.0), (orc.TANH, 0.7), (orc.EXP, 0.1), (orc.EXP, 1.0), (orc.CUBE, 1.0)]


def _ref(x, w, kind, alpha):
    return orc.eval_point(x, w, kind, alpha, ortho=False, extended=False)


@pytest.mark.parametrize("n,t", [(3, 10000), (2, 17), (8, 1001), (13, 4099), (16, 5000), (32, 3001), (64, 2049), (100, 1500), (128, 1025)])
@pytest.mark.parametrize("kind,alpha", DENS)
def test_fused_moments_match_oracle(n, t, kind, alpha):
    x = _data.whitened(n, t, seed=n * 7 + t) if t > 4 * n else np.random.default_rng(1).standard_normal((n, t))
    w = _data.orthogonal(n, seed=n) + 0.05 * np.random.default_rng(n).standard_normal((n, n))
    ref = _ref(x, w, kind, alpha)
    got = _gpu.eval_moments(x, w, kind, alpha, mode=0, want_h=True)
    assert _data.rel_err(got["gr"], ref.gr) <= TOL
    assert _data.rel_err(got["hr"], ref.hr) <= TOL
    assert _data.rel_err(got["sd"], ref.sd) <= TOL
    assert _data.rel_err(got["sq"], ref.sq) <= TOL
    assert _data.rel_err(got["lrow"], ref.lrow) <= TOL


@pytest.mark.parametrize("n,t", [(5, 777), (64, 4097), (128, 2050)])
def test_pass_modes_agree(n, t):
    """grad-only and loss-only variants produce the same sections as the fused pass (same kernel template)."""
    x = _data.whitened(n, t, seed=3)
    w = _data.orthogonal(n, seed=5)
    full = _gpu.eval_moments(x, w, mode=0, want_h=True)
    grad = _gpu.eval_moments(x, w, mode=1, want_h=True)
    grad_noh = _gpu.eval_moments(x, w, mode=1, want_h=False)
    loss = _gpu.eval_moments(x, w, mode=2, want_h=True)   # LOSS mode: want_h asks for the Sq row sums only
    loss_nosq = _gpu.eval_moments(x, w, mode=2, want_h=False)
    for k in ("gr", "sd", "hr", "sq"):
        np.testing.assert_allclose(grad[k], full[k], rtol=1e-13, atol=1e-9)
    for k in ("gr", "sd"):  # Sq is only produced with want_h (it feeds the non-ortho Hessian / loss)
        np.testing.assert_allclose(grad_noh[k], full[k], rtol=1e-13, atol=1e-9)
    for k in ("sq", "lrow"):
        np.testing.assert_allclose(loss[k], full[k], rtol=1e-13, atol=1e-9)
    np.testing.assert_allclose(loss_nosq["lrow"], full["lrow"], rtol=1e-13, atol=1e-9)


@pytest.mark.parametrize("n,t", [(3, 1000), (16, 5003), (40, 2000), (64, 4097), (128, 2050)])
@pytest.mark.parametrize("kind,alpha", [(orc.TANH, 1.0), (orc.EXP, 0.1), (orc.CUBE, 1.0)])
def test_stored_y_gradient_path_matches_oracle(n, t, kind, alpha):
    """mode 3: a loss-only pass that stores Y' followed by the gradient moments computed from the stored Y'
    (what an accepted loss-only line-search try costs) gives the same moments as the oracle."""
    x = _data.whitened(n, t, seed=n + 17)
    w = _data.orthogonal(n, seed=n + 2) + 0.03 * np.random.default_rng(n).standard_normal((n, n))
    ref = _ref(x, w, kind, alpha)
    for want_h in (True, False):
        got = _gpu.eval_moments(x, w, kind, alpha, mode=3, want_h=want_h)
        assert _data.rel_err(got["gr"], ref.gr) <= TOL
        assert _data.rel_err(got["sd"], ref.sd) <= TOL
        assert _data.rel_err(got["lrow"], ref.lrow) <= TOL
        if want_h:
            assert _data.rel_err(got["hr"], ref.hr) <= TOL
            assert _data.rel_err(got["sq"], ref.sq) <= TOL


@pytest.mark.parametrize("n,t", [(130, 1500), (200, 2001), (256, 1040)])
@pytest.mark.parametrize("kind,alpha", [(orc.TANH, 1.0), (orc.EXP, 0.1)])
def test_n_above_128_row_block_kernels(n, t, kind, alpha):
    """N in (128, 256]: the row-block partitioned LOSS (+ Y store) and stored-Y gradient kernels (BASELINE configs[3] is N = 256)."""
    x = _data.whitened(n, t, seed=n)
    w = _data.orthogonal(n, seed=n + 2) + 0.02 * np.random.default_rng(n).standard_normal((n, n))
    ref = _ref(x, w, kind, alpha)
    loss = _gpu.eval_moments(x, w, kind, alpha, mode=2, want_h=True)
    assert _data.rel_err(loss["lrow"], ref.lrow) <= TOL and _data.rel_err(loss["sq"], ref.sq) <= TOL
    for want_h in (True, False):
        got = _gpu.eval_moments(x, w, kind, alpha, mode=3, want_h=want_h)
        assert _data.rel_err(got["gr"], ref.gr) <= TOL
        assert _data.rel_err(got["sd"], ref.sd) <= TOL
        assert _data.rel_err(got["lrow"], ref.lrow) <= TOL
        if want_h:
            assert _data.rel_err(got["hr"], ref.hr) <= TOL
            assert _data.rel_err(got["sq"], ref.sq) <= TOL


def test_from_x_gradient_kernels_refuse_n_above_128():
    x = np.random.default_rng(0).standard_normal((140, 500))
    with pytest.raises(RuntimeError, match="N > 128 needs the Y store"):
        _gpu.eval_moments(x, None, mode=0, want_h=False)


def test_identity_w_is_default():
    x = _data.whitened(6, 999, seed=2)
    a = _gpu.eval_moments(x, None)
    b = _gpu.eval_moments(x, np.eye(6))
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])


def test_strided_rows():
    """Rows with a stride larger than T (an ndarray view): only the T valid columns are read."""
    big = np.random.default_rng(0).standard_normal((7, 1300))
    x = big[:, :1111]
    ref = _ref(np.ascontiguousarray(x), np.eye(7), orc.TANH, 1.0)
    got = _gpu.eval_moments(x, None)  # _c() keeps a contiguous copy; stride path is exercised by fit tests
    assert _data.rel_err(got["gr"], ref.gr) <= TOL


def test_linearity_in_samples_full_size():
    """Size-independent property at a large shape: moments of [X1 | X2] = moments(X1) + moments(X2)."""
    n, t = 64, 200_000
    x = _data.whitened(n, t, seed=11)
    w = _data.orthogonal(n, seed=1)
    a = _gpu.eval_moments(x[:, : t // 2 + 3], w)
    b = _gpu.eval_moments(x[:, t // 2 + 3:], w)
    c = _gpu.eval_moments(x, w)
    for k in c:
        assert _data.rel_err(a[k] + b[k], c[k]) <= 1e-12


def test_large_values_saturate_cleanly():
    """|y| large: tanh saturates, exp(-2|y|) underflows; no NaN/Inf may appear (extreme-input edge case)."""
    x = np.array([[1e3, -1e3, 350.0, -0.0, 0.0, 1e-300, 40.0]])
    got = _gpu.eval_moments(x, None, orc.TANH, 1.0)
    ref = _ref(x, np.eye(1), orc.TANH, 1.0)
    for k in ("gr", "sd", "sq", "lrow"):
        assert np.all(np.isfinite(got[k]))
        assert _data.rel_err(got[k], getattr(ref, k)) <= TOL
