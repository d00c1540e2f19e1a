"""Whitening (whitening.rs:24-116) on ill-conditioned data: the device path decomposes the Gram matrix, which squares the
condition number; below sigma_min / sigma_max ~ 1e-3 a refinement stage (fit.cu: center_whiten_device) recovers the accuracy of
an SVD of X itself, down to the reference's absolute singularity threshold of 1e-10 (VERDICT r01 missing #5)."""
import numpy as np
import pytest

import _data
import _gpu
import picard_ica_b200 as P
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _graded(n, t, decades, seed):
    """A mixture whose singular values span `decades` orders of magnitude."""
    s = _data.sources(n, t, seed, "mixed")
    q1, q2 = _data.orthogonal(n, seed + 1), _data.orthogonal(n, seed + 2)
    return q1 @ np.diag(np.logspace(0.0, -decades, n)) @ q2 @ s


@pytest.mark.parametrize("n,t,decades", [(8, 20_000, 4.0), (12, 30_000, 6.0), (16, 20_000, 8.0), (40, 10_000, 7.0), (6, 5_000, 9.0)])
def test_graded_spectrum_matches_the_svd_path(n, t, decades):
    x = _graded(n, t, decades, seed=n)
    st, msg, mean, k, data = _gpu.center_whiten(x, n)
    assert st == 0, msg
    xc, _ = orc.center(x)
    st_o, data_o, k_o = orc.whiten(xc, n)
    assert st_o == orc.OK
    # rows of K scale like 1 / sigma_i: compare row by row (relative to the row's own size)
    rel = np.max(np.abs(k - k_o), axis=1) / np.max(np.abs(k_o), axis=1)
    assert np.max(rel) <= 1e-7, rel
    cov = data @ data.T / t
    assert np.max(np.abs(cov - np.eye(n))) <= 1e-7


def test_ill_conditioned_fit_matches_oracle():
    """cond(X) = 1e6: the reference fits this (sigma_min ~ 1e-4 is far above its 1e-10 threshold); so does the device path."""
    n, t = 10, 30_000
    x = _graded(n, t, 6.0, seed=3)
    w0 = _data.orthogonal(n, 43)
    res = P.Picard.fit_with_config(x, P.PicardConfig(w_init=w0))
    ref = orc.fit(x, orc.Config(w_init=w0))
    assert abs(res.n_iterations - ref.n_iterations) <= 1 and res.converged == ref.converged
    assert P.utils.amari_distance(res.full_unmixing(), np.linalg.pinv(ref.full_unmixing())) <= 1e-6


def test_singularity_threshold_is_the_reference_absolute_one():
    """whitening.rs:72-79: SingularMatrix iff min singular value < 1e-10 -- an ABSOLUTE threshold on the centred data."""
    n, t = 6, 4_000
    s = _data.sources(n, t, 5, "laplace")
    s -= s.mean(axis=1, keepdims=True)
    u, sv, vt = np.linalg.svd(s, full_matrices=False)

    def with_min_sv(v):
        sv2 = sv.copy(); sv2[-1] = v
        return (u * sv2) @ vt

    for v, expect_singular in [(1e-6, False), (1e-9, False), (1e-12, True)]:
        x = with_min_sv(v)
        st, msg, mean, k, data = _gpu.center_whiten(x, n)
        xc, _ = orc.center(x)
        st_o = orc.whiten(xc, n)[0]
        assert (st_o != orc.OK) == expect_singular
        assert (st != 0) == expect_singular, (v, st, msg)
        if expect_singular:
            assert st == 2 and "Singular" in msg
