"""End-to-end GPU parity of the fit path (Picard::fit_with_config / transform, solver.rs:45-214) through the
reference-facing API of picard_ica_b200 (ctypes over the C ABI), against the CPU oracle on identical X and
w_init.  Bar (BASELINE.json north_star): Amari distance between the two unmixings <= 1e-6, iteration counts
within +-1.  The shape-level assertions of the reference's own tests (solver.rs:288-408) are repeated too."""
import numpy as np
import pytest

import _data
import picard_ica_b200 as P
from oracle import oracle as orc
from picard_ica_b200 import ConfigBuilder, DensityType, Picard, PicardConfig, PicardError
from picard_ica_b200.utils import amari_distance

pytestmark = pytest.mark.gpu


def _cmp(res, ref, amari_tol=1e-6, iters_tol=1):
    assert abs(res.n_iterations - ref.n_iterations) <= iters_tol, (res.n_iterations, ref.n_iterations)
    assert res.converged == ref.converged
    d = amari_distance(res.full_unmixing(), np.linalg.pinv(ref.full_unmixing()))
    assert d <= amari_tol, d
    return d


def test_config1_reference_bench_case():
    """BASELINE configs[0]: N=3 Laplace, T=10,000, tanh, whiten, ortho=false, data from the reference's own bench
    generator (benches/benchmarks.rs:8-35)."""
    x = _data.lcg_bench_data(3, 10_000, 42)
    w0 = _data.orthogonal(3, 43)
    res = Picard.fit_with_config(x, PicardConfig(ortho=False, w_init=w0))
    ref = orc.fit(x, orc.Config(ortho=False, w_init=w0))
    assert res.signs is None and ref.signs is None
    _cmp(res, ref)
    np.testing.assert_allclose(res.mean, ref.mean, rtol=0, atol=1e-12)
    np.testing.assert_allclose(res.whitening, ref.whitening, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res.sources, ref.sources, rtol=0, atol=1e-6)
    assert abs(res.gradient_norm - ref.gradient_norm) <= 1e-9


@pytest.mark.parametrize("n,t,kind", [(8, 20_000, "mixed"), (16, 30_000, "laplace"), (64, 50_000, "mixed")])
def test_picard_o_extended(n, t, kind):
    """BASELINE configs[1]/[2] at sizes the oracle finishes in seconds: Picard-O, extended, tanh."""
    x, a, _ = _data.mixture(n, t, seed=n, kind=kind)
    w0 = _data.orthogonal(n, 43)
    res = Picard.fit_with_config(x, PicardConfig(w_init=w0))
    ref = orc.fit(x, orc.Config(w_init=w0))
    assert res.converged
    _cmp(res, ref)
    np.testing.assert_array_equal(res.signs, ref.signs)
    # and it actually separates the sources
    assert amari_distance(res.full_unmixing(), a) < 0.05


def test_nonortho_exp_density():
    """BASELINE configs[3] shape in miniature: non-ortho, exp(alpha=0.1), Laplace sources."""
    x, a, _ = _data.mixture(12, 40_000, seed=5, kind="laplace")
    w0 = _data.orthogonal(12, 43)
    cfg = dict(ortho=False, extended=False, w_init=w0)
    res = Picard.fit_with_config(x, PicardConfig(density=DensityType.exp_with_alpha(0.1), **cfg))
    ref = orc.fit(x, orc.Config(density=orc.EXP, alpha=0.1, **cfg))
    _cmp(res, ref)


def test_cube_density_subgaussian():
    x, a, _ = _data.mixture(5, 20_000, seed=8, kind="uniform")
    w0 = _data.orthogonal(5, 43)
    res = Picard.fit_with_config(x, PicardConfig(density=DensityType.cube(), ortho=True, extended=False, w_init=w0, max_iter=200))
    ref = orc.fit(x, orc.Config(density=orc.CUBE, ortho=True, extended=False, w_init=w0, max_iter=200))
    _cmp(res, ref)


@pytest.mark.parametrize("n", [10, 100])
def test_execution_strategies_do_not_change_the_iterates(n):
    """Y store (default), speculative fused first try, and plain loss + from-X gradient passes are execution
    strategies only: same iterate sequence, different pass mix."""
    x, _, _ = _data.mixture(n, 20_000, seed=2)
    w0 = _data.orthogonal(n, 43)
    kw = dict(w_init=w0, max_iter=40)
    a = Picard.fit_with_config(x, PicardConfig(**kw))                                                    # LOSS + Y store, stored-Y gradient
    b = Picard.fit_with_config(x, PicardConfig(flags=P.FLAG_FORCE_SPECULATION, **kw))                   # speculative fused first try
    c = Picard.fit_with_config(x, PicardConfig(flags=P.FLAG_NO_Y_STORE, **kw))                          # no store: speculation + from-X gradient
    d = Picard.fit_with_config(x, PicardConfig(flags=P.FLAG_NO_Y_STORE | P.FLAG_NO_SPECULATION, **kw))  # loss-only tries + from-X gradient
    for r in (b, c, d):
        assert r.n_iterations == a.n_iterations
        np.testing.assert_allclose(r.unmixing, a.unmixing, rtol=0, atol=1e-11)
    assert a.stats["fused_passes"] == 0 and a.stats["grad_passes"] == 0 and a.stats["grady_passes"] > 0
    assert b.stats["fused_passes"] > 0
    assert c.stats["fused_passes"] > 0 and c.stats["grady_passes"] == 0
    assert d.stats["fused_passes"] == 0 and d.stats["grady_passes"] == 0 and d.stats["grad_passes"] > 0


@pytest.mark.parametrize("n,t,kind,dens", [(6, 10_000, "mixed", "tanh"), (20, 20_000, "laplace", "tanh"), (8, 10_000, "uniform", "cube"),
                                          (140, 20_000, "laplace", "tanh")])
def test_fastica_warmstart(n, t, kind, dens):
    """fastica_it (ica_par, solver.rs:218-249): W <- symdecor(E[g(WX)X^T] - diag(E[g'(WX)]) W) from w_init, then Picard."""
    x, a, _ = _data.mixture(n, t, seed=n, kind=kind)
    w0 = _data.orthogonal(n, 43)
    d = DensityType.cube() if dens == "cube" else DensityType.tanh()
    ext = dens != "cube"
    res = Picard.fit_with_config(x, PicardConfig(density=d, fastica_it=5, w_init=w0, extended=ext, max_iter=60))
    ref = orc.fit(x, orc.Config(density=orc.CUBE if dens == "cube" else orc.TANH, fastica_it=5, w_init=w0, extended=ext, max_iter=60))
    _cmp(res, ref)
    # and the warm start really ran: the result differs from a plain fit's iteration count or matches it with the same optimum
    assert amari_distance(res.full_unmixing(), a) < 0.1


def test_fastica_zero_iterations_only_decorrelates():
    """fastica_it = 0: ica_par still applies sym_decorrelation to w_init (solver.rs:226)."""
    x, a, _ = _data.mixture(5, 8000, seed=3)
    w0 = _data.orthogonal(5, 43) @ np.diag([1.0, 2.0, 0.5, 1.5, 1.0])  # not orthogonal
    res = Picard.fit_with_config(x, PicardConfig(fastica_it=0, w_init=w0, max_iter=50))
    ref = orc.fit(x, orc.Config(fastica_it=0, w_init=w0, max_iter=50))
    _cmp(res, ref)


def test_n_above_128_fit():
    """N = 160 (row-block kernels, eigh-based whitening of 160 features): Picard-O extended against the oracle."""
    x, a, _ = _data.mixture(160, 40_000, seed=9, kind="mixed")
    w0 = _data.orthogonal(160, 43)
    res = Picard.fit_with_config(x, PicardConfig(w_init=w0, max_iter=60))
    ref = orc.fit(x, orc.Config(w_init=w0, max_iter=60))
    _cmp(res, ref)


def test_n_above_128_nonortho_exp():
    """BASELINE configs[3] in miniature at N > 128: non-ortho, exp(alpha = 0.1), Laplace sources, H from the stored Y'."""
    x, a, _ = _data.mixture(136, 30_000, seed=11, kind="laplace")
    w0 = _data.orthogonal(136, 43)
    cfg = dict(ortho=False, extended=False, w_init=w0, max_iter=25)
    res = Picard.fit_with_config(x, PicardConfig(density=DensityType.exp_with_alpha(0.1), **cfg))
    ref = orc.fit(x, orc.Config(density=orc.EXP, alpha=0.1, **cfg))
    _cmp(res, ref)


# ---- the reference's own solver tests (solver.rs:288-408), same assertions ---------------------------
def _laplace_mix(n, t, seed):
    return _data.mixture(n, t, seed, "laplace")[0]


def test_fit_default():  # solver.rs:289-301
    x = _laplace_mix(3, 1000, 42)
    r = Picard.fit(x)
    assert r.unmixing.shape == (3, 3) and r.sources.shape == (3, 1000)
    assert r.whitening is not None and r.mean is not None


def test_fit_with_config():  # solver.rs:304-318
    x = _laplace_mix(4, 2000, 42)
    cfg = ConfigBuilder().n_components(3).max_iter(100).tol(1e-6).random_state(42).build()
    r = Picard.fit_with_config(x, cfg)
    assert r.unmixing.shape == (3, 3) and r.sources.shape == (3, 2000) and r.whitening.shape == (3, 4)
    assert r.n_iterations <= 100


def test_transform_matches_sources():  # solver.rs:359-375
    x = _laplace_mix(3, 1000, 7)
    r = Picard.fit_with_config(x, PicardConfig(random_state=1))
    y = Picard.transform(x, r)
    assert y.shape == (3, 1000)
    np.testing.assert_allclose(y, r.sources, rtol=0, atol=1e-9)
    x2 = _laplace_mix(3, 333, 8)
    np.testing.assert_allclose(Picard.transform(x2, r), r.full_unmixing() @ (x2 - r.mean[:, None]), rtol=0, atol=1e-10)


def test_no_whiten():  # solver.rs:378-394 ; quirk Q13: n_components ignored
    x = _laplace_mix(3, 1000, 9)
    r = Picard.fit_with_config(x, PicardConfig(whiten=False, n_components=2, random_state=3, max_iter=50))
    assert r.whitening is None and r.unmixing.shape == (3, 3)


def test_no_centering():
    x = _laplace_mix(3, 1000, 9) + 5.0
    r = Picard.fit_with_config(x, PicardConfig(centering=False, random_state=3, max_iter=20))
    assert r.mean is None


def test_same_seed_same_result():
    x = _laplace_mix(4, 3000, 10)
    a = Picard.fit_with_config(x, PicardConfig(random_state=5))
    b = Picard.fit_with_config(x, PicardConfig(random_state=5))
    np.testing.assert_array_equal(a.unmixing, b.unmixing)


def test_errors():
    x = _laplace_mix(3, 100, 1)
    with pytest.raises(PicardError.InvalidConfig) as e:  # solver.rs:397-407
        Picard.fit_with_config(x, PicardConfig(fastica_it=5, jade_it=5))
    assert e.value.parameter == "jade_it"
    with pytest.raises(PicardError.InvalidDimensions):   # solver.rs:50-54
        Picard.fit(np.zeros((0, 0)))
    with pytest.raises(PicardError.InvalidDimensions):   # solver.rs:100-108
        Picard.fit_with_config(x, PicardConfig(w_init=np.eye(2)))
    with pytest.raises(PicardError.SingularMatrix):      # whitening.rs:72-79 (rank-deficient data)
        Picard.fit(np.vstack([x[0], x[0], x[1]]))
    with pytest.raises(PicardError.InvalidDimensions):   # whitening.rs:51-58 cannot trigger via fit (min()), transform mismatch
        Picard.transform(np.zeros((5, 10)), Picard.fit(x))


def test_not_converged_is_ok_not_error():
    x = _laplace_mix(6, 5000, 3)
    r = Picard.fit_with_config(x, PicardConfig(max_iter=2, random_state=0))
    assert not r.converged and r.n_iterations == 2 and r.gradient_norm > 1e-7


def test_odd_t_and_strided_input():
    big = np.random.default_rng(0).standard_normal((5, 4000))
    x = (_data.orthogonal(5, 1) @ np.sign(big) * np.abs(big) ** 1.5)[:, :3333]  # a view with row stride 4000
    assert x.strides[0] == 4000 * 8
    w0 = _data.orthogonal(5, 43)
    res = Picard.fit_with_config(x, PicardConfig(w_init=w0))
    ref = orc.fit(np.ascontiguousarray(x), orc.Config(w_init=w0))
    _cmp(res, ref)


def test_tiny_shapes():
    """Edge cases: a single partial tile (T < 16), one component, T = N (barely determined)."""
    rng = np.random.default_rng(0)
    x = rng.laplace(size=(2, 13))
    w0 = _data.orthogonal(2, 43)
    res = Picard.fit_with_config(x, PicardConfig(w_init=w0, max_iter=20))
    ref = orc.fit(x, orc.Config(w_init=w0, max_iter=20))
    assert res.sources.shape == (2, 13) and abs(res.n_iterations - ref.n_iterations) <= 1
    np.testing.assert_allclose(res.sources @ res.sources.T / 13, np.eye(2), atol=1e-8)   # Picard-O keeps the whitened scale
    one = Picard.fit_with_config(rng.laplace(size=(1, 100)), PicardConfig(max_iter=5))
    assert one.unmixing.shape == (1, 1) and one.sources.shape == (1, 100)
    np.testing.assert_allclose(np.mean(one.sources ** 2), 1.0, atol=1e-10)


def test_more_features_than_samples_is_singular():
    """n_components = min(n, p) (solver.rs:63); with T < N the centred data has rank < T, so whitening reports SingularMatrix
    (whitening.rs:72-79) exactly like the reference."""
    x = np.random.default_rng(1).standard_normal((6, 4))
    with pytest.raises(PicardError.SingularMatrix):
        Picard.fit(x)
    with pytest.raises(orc.OracleError) as e:
        orc.fit(x)
    assert e.value.code == orc.SINGULAR


def test_rectangular_whitening_and_transform_of_new_data():
    x, a, _ = _data.mixture(10, 5000, seed=4, kind="laplace")
    w0 = _data.orthogonal(4, 43)
    res = Picard.fit_with_config(x, PicardConfig(n_components=4, w_init=w0, max_iter=100))
    ref = orc.fit(x, orc.Config(n_components=4, w_init=w0, max_iter=100))
    assert res.whitening.shape == (4, 10) and res.sources.shape == (4, 5000)
    _cmp(res, ref)
    y = Picard.transform(x[:, :777], res)
    np.testing.assert_allclose(y, res.sources[:, :777], atol=1e-9)


def test_large_sources_come_from_the_pinned_arena_and_are_correct():
    """A `sources` result >= 256 MB is handed out from the library's pinned arena (one DMA into the caller's buffer); a second fit
    while the first result is still alive falls back to malloc + the staged copy; both give the same numbers, and the arena is
    reused once the first result is released."""
    import gc
    n, t = 40, 900_000  # 288 MB of sources
    x, _, _ = _data.mixture(n, t, seed=12, kind="laplace")
    w0 = _data.orthogonal(n, 43)
    cfg = PicardConfig(w_init=w0, max_iter=5)
    a = Picard.fit_with_config(x, cfg)
    b = Picard.fit_with_config(x, cfg)          # `a` still owns the arena
    np.testing.assert_array_equal(a.sources, b.sources)
    np.testing.assert_allclose(a.sources, a.full_unmixing() @ (x - a.mean[:, None]), rtol=0, atol=1e-9)
    keep = a.sources[:, :1000].copy()
    del a
    gc.collect()
    c = Picard.fit_with_config(x, cfg)          # the arena is free again
    np.testing.assert_array_equal(c.sources[:, :1000], keep)
    np.testing.assert_array_equal(c.sources, b.sources)
