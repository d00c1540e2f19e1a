"""GPU parity of the N x N device kernels (K4, K5, K7): matrix_exp (math.rs:38-74), sln_det (math.rs:84-88),
sym_decorrelation (math.rs:12-33), compute_direction (lbfgs.rs:84-150), including the reference's own
known-answer tests (math.rs:100-152) run against the CUDA kernels."""
import numpy as np
import pytest

import _data
import _gpu
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _skew(n, seed, scale):
    a = np.random.default_rng(seed).standard_normal((n, n)) * scale
    return (a - a.T) / 2.0


@pytest.mark.parametrize("n", [2, 3, 16, 33, 64, 128])
@pytest.mark.parametrize("scale", [1e-17, 1e-3, 0.3, 1.0, 7.5])
def test_matrix_exp_matches_oracle(n, scale):
    a = _skew(n, n, scale)
    ref = orc.matrix_exp(a)
    got = _gpu.matrix_exp(a)
    assert _data.rel_err(got, ref) <= 1e-12


def test_matrix_exp_identity():  # math.rs:114-124
    got = _gpu.matrix_exp(np.zeros((3, 3)))
    np.testing.assert_allclose(got, np.eye(3), atol=1e-10)


def test_matrix_exp_general_matrix():
    a = np.random.default_rng(0).standard_normal((10, 10)) * 0.4
    assert _data.rel_err(_gpu.matrix_exp(a), orc.matrix_exp(a)) <= 1e-12


def test_sln_det_known_answers():  # math.rs:127-141
    s, l = _gpu.sln_det(np.array([[1.0, 2.0], [3.0, 4.0]]))
    assert s == -1.0 and abs(l - np.log(2.0)) < 1e-10
    s, l = _gpu.sln_det(np.diag([1e150, 1e150]))
    assert s == 1.0 and abs(l - 2 * np.log(1e150)) < 1e-6


@pytest.mark.parametrize("n", [1, 3, 17, 64, 128])
def test_sln_det_matches_oracle(n):
    m = np.random.default_rng(n).standard_normal((n, n))
    st, rs, rl = orc.sln_det(m)
    s, l = _gpu.sln_det(m)
    assert st == 0 and s == rs and abs(l - rl) <= 1e-11 * max(1.0, abs(rl))


def test_sln_det_singular():
    s, l = _gpu.sln_det(np.ones((4, 4)))
    assert s == 0.0
    st, rs, rl = orc.sln_det(np.ones((4, 4)))
    assert rs == 0.0


@pytest.mark.parametrize("n", [2, 3, 16, 50, 128])
def test_sym_decorrelation(n):  # math.rs:101-111
    w = np.random.default_rng(n).standard_normal((n, n))
    st, got = _gpu.sym_decorrelation(w)
    assert st == 0
    np.testing.assert_allclose(got @ got.T, np.eye(n), atol=1e-10)
    rst, ref = orc.sym_decorrelation(w)
    assert rst == 0 and _data.rel_err(got, ref) <= 1e-10


def test_sym_decorrelation_singular():  # math.rs:21-24
    w = np.ones((3, 3))
    st, _ = _gpu.sym_decorrelation(w)
    assert st == 2
    assert orc.sym_decorrelation(w)[0] == 2


@pytest.mark.parametrize("n", [3, 32, 128])
@pytest.mark.parametrize("ortho", [True, False])
@pytest.mark.parametrize("L", [0, 1, 4, 7])
def test_compute_direction_matches_oracle(n, ortho, L):
    rng = np.random.default_rng(n * 10 + L)
    g = rng.standard_normal((n, n)) * 0.1
    if ortho:
        g = (g - g.T) / 2
        hoff = rng.uniform(0.2, 1.0, n)
        h = np.maximum(rng.uniform(-0.5, 2.0, (n, n)), 0.01)
    else:
        hoff = np.ones(n)
        h = rng.uniform(0.5, 3.0, (n, n))
        h[0, 1] = h[1, 0] = 1.0  # det = 0 for this pair: quirk Q15 (entry zeroed)
    s = [rng.standard_normal((n, n)) * 0.05 for _ in range(L)]
    y = [2.0 * a + rng.standard_normal((n, n)) * 0.01 for a in s]  # positive curvature pairs: well conditioned
    r = [1.0 / float(np.sum(a * b)) for a, b in zip(s, y)]
    ref = orc.compute_direction(g, h, hoff, s, y, r, ortho)
    got = _gpu.compute_direction(g, h, hoff, s, y, r, ortho)
    assert _data.rel_err(got, ref) <= 1e-10
