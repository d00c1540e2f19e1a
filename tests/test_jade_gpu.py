"""GPU parity of the JADE warm start (jade.rs:22-197): cumulant matrices (K8), Jacobi sweeps (K9) and the fit path with
jade_it (BASELINE configs[4]), CUDA through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

import _data
import _gpu
import picard_ica_b200 as P
from oracle import oracle as orc
from picard_ica_b200 import Picard, PicardConfig, PicardError
from picard_ica_b200.utils import amari_distance

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,t", [(2, 500), (3, 1000), (8, 3001), (13, 2000), (16, 4096), (24, 1500), (32, 2500)])
def test_cumulant_matrices_match_oracle(n, t):
    x = _data.whitened(n, t, seed=n)
    ref = orc.cumulants(x)
    got = _gpu.jade_cumulants(x)
    assert got.shape == ref.shape == (n * (n + 1) // 2, n, n)
    assert _data.rel_err(got, ref) <= 1e-10
    np.testing.assert_allclose(got, np.swapaxes(got, 1, 2), atol=1e-13)  # symmetrised (jade.rs:126)


@pytest.mark.parametrize("n,t,max_sweeps,tol", [(3, 1000, 5, 1e-9), (8, 5000, 2, 1e-9), (16, 3000, 1, 1e-8), (32, 2000, 1, 1e-5)])
def test_jade_sweeps_match_oracle(n, t, max_sweeps, tol):
    """The reference's Givens angle rule (jade.rs:170-179: theta = atan2(2 g01, g11 - g00) / 4) does not converge: rotations
    stay O(0.5 rad) sweep after sweep, so the sweep sequence is a chaotic map and ANY two implementations (different
    summation order is enough) drift apart by roughly a factor 1.02 per rotation.  Parity of the
    sweep kernel is therefore checked over a bounded number of rotations, with the tolerance that drift allows."""
    x = _data.whitened(n, t, seed=n + 3)
    rst, rw, rsweeps = orc.jade(x, max_sweeps, 1e-6)
    st, err, w, sweeps = _gpu.jade(x, max_sweeps, 1e-6)
    assert st == 0 and rst == 0, err
    np.testing.assert_allclose(w @ w.T, np.eye(n), atol=1e-6)  # jade.rs:237-255 (test_jade_basic)
    assert sweeps == rsweeps == max_sweeps
    assert np.max(np.abs(w - rw)) <= tol


def test_jade_is_deterministic_and_orthogonal_at_full_length():
    x = _data.whitened(32, 4000, seed=1)
    st, err, w1, s1 = _gpu.jade(x, 50, 1e-6)
    st2, err2, w2, s2 = _gpu.jade(x, 50, 1e-6)
    assert st == 0 and st2 == 0 and s1 == s2 == 50
    np.testing.assert_array_equal(w1, w2)
    np.testing.assert_allclose(w1 @ w1.T, np.eye(32), atol=1e-10)


def test_jade_trivial_sizes():  # jade.rs:25-27: n < 2 -> identity
    st, err, w, sweeps = _gpu.jade(np.random.default_rng(0).standard_normal((1, 100)), 10)
    assert st == 0 and w.shape == (1, 1) and w[0, 0] == 1.0


def test_jade_refuses_more_than_32_components():
    st, err, w, sweeps = _gpu.jade(np.random.default_rng(0).standard_normal((40, 500)), 5)
    assert st == 1 and "32" in err


def test_fit_with_jade_warmstart():  # solver.rs:321-337 (test_fit_with_jade_warmstart) + parity with the oracle
    x, a, _ = _data.mixture(6, 10_000, seed=5, kind="mixed")
    res = Picard.fit_with_config(x, PicardConfig(jade_it=50, max_iter=100, random_state=42))
    ref = orc.fit(x, orc.Config(jade_it=50, max_iter=100, random_state=42))
    assert res.unmixing.shape == (6, 6) and res.sources.shape == (6, 10_000)
    assert res.converged and ref.converged
    # 50 chaotic JADE sweeps: the two warm starts differ, so the Picard iteration counts may too; both runs must end in
    # the same separating solution
    assert amari_distance(res.full_unmixing(), np.linalg.pinv(ref.full_unmixing())) <= 1e-5
    assert amari_distance(res.full_unmixing(), a) < 0.05


def test_jade_warmstart_needs_fewer_iterations():  # solver.rs:340-356 (test_jade_vs_no_warmstart), same weak assertion
    x, a, _ = _data.mixture(8, 20_000, seed=7, kind="mixed")
    w0 = _data.orthogonal(8, 43)
    plain = Picard.fit_with_config(x, PicardConfig(w_init=w0, max_iter=200))
    warm = Picard.fit_with_config(x, PicardConfig(jade_it=50, max_iter=200))
    assert plain.n_iterations <= 200 and warm.n_iterations <= 200
    assert warm.n_iterations <= plain.n_iterations + 2


def test_config5_shape_in_miniature():
    """BASELINE configs[4]: N = 32 with jade_it = 50 then Picard-O (T and the sweep count reduced so the oracle's N^4 T cumulants and from-scratch sweeps finish)."""
    x, a, _ = _data.mixture(32, 2000, seed=9, kind="mixed")
    res = Picard.fit_with_config(x, PicardConfig(jade_it=1, max_iter=100))
    ref = orc.fit(x, orc.Config(jade_it=1, max_iter=100))
    assert res.converged and ref.converged
    assert abs(res.n_iterations - ref.n_iterations) <= 1  # one sweep: the warm starts still agree to ~1e-8
    assert amari_distance(res.full_unmixing(), np.linalg.pinv(ref.full_unmixing())) <= 1e-6
