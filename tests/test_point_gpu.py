"""GPU parity of one iteration front (core.rs:215-293) + loss (core.rs:39-85): projected G, h, h_off, signs,
gradient norm and loss from identical W, CUDA (pass + N x N epilogue kernel) vs the CPU oracle, <= 1e-10."""
import numpy as np
import pytest

import _data
import _gpu
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.mark.parametrize("n,t", [(3, 10000), (16, 5001), (64, 4000), (128, 3000)])
@pytest.mark.parametrize("ortho,extended", [(True, True), (True, False), (False, False), (False, True)])
@pytest.mark.parametrize("kind,alpha", [(orc.TANH, 1.0), (orc.EXP, 0.1), (orc.CUBE, 1.0)])
def test_point_matches_oracle(n, t, ortho, extended, kind, alpha):
    x = _data.whitened(n, t, seed=n + t)
    w = _data.orthogonal(n, seed=n + 1)
    if not ortho:
        w = w + 0.1 * np.random.default_rng(n).standard_normal((n, n))
    c = w @ w.T
    old = np.where(np.arange(n) % 3 == 0, -1.0, 1.0)
    ref = orc.eval_point(x, w, kind, alpha, ortho, extended, 0.01, c=c, old_signs=old)
    got = _gpu.eval_point(x, w, kind, alpha, ortho, extended, 0.01, c=c, old_signs=old)
    np.testing.assert_array_equal(got["signs"], ref.signs)
    assert got["sign_change"] == ref.sign_change
    assert _data.rel_err(got["g"], ref.g) <= TOL
    assert _data.rel_err(got["h"], ref.h) <= TOL
    assert _data.rel_err(got["hoff"], ref.hoff) <= TOL
    assert abs(got["gradient_norm"] - ref.gradient_norm) <= TOL * max(1.0, abs(ref.gradient_norm))
    assert abs(got["loss"] - ref.loss) <= TOL * max(1.0, abs(ref.loss))


def test_first_iteration_never_reports_sign_change():
    x = _data.whitened(8, 2000, seed=4, kind="uniform")  # sub-Gaussian: signs flip to -1
    ref = orc.eval_point(x, None, ortho=True, extended=True)
    got = _gpu.eval_point(x, None, ortho=True, extended=True)
    assert not ref.sign_change and not got["sign_change"]
    np.testing.assert_array_equal(got["signs"], ref.signs)
    assert np.any(ref.signs < 0)


def test_loss_with_explicit_signs():
    """Quirk Q1: the initial loss is evaluated with signs = 1 whatever the data looks like."""
    x = _data.whitened(6, 3000, seed=9, kind="uniform")
    ones = np.ones(6)
    ref = orc.eval_point(x, None, ortho=True, extended=True, loss_signs=ones)
    got = _gpu.eval_point(x, None, ortho=True, extended=True, loss_signs=ones)
    assert abs(got["loss"] - ref.loss) <= TOL * abs(ref.loss)


def test_singular_w_gives_penalty_loss():
    """core.rs:90-96: a singular W' in the line search is a 1e15 loss, not an error."""
    x = _data.whitened(4, 500, seed=1)
    w = np.ones((4, 4))
    ref = orc.eval_point(x, w, ortho=False, extended=False)
    got = _gpu.eval_point(x, w, ortho=False, extended=False)
    assert ref.loss == 1e15 and got["loss"] == 1e15
