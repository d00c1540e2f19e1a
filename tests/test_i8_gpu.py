"""GPU parity of the INT8 tensor-core engines (i8_loss.cu: LOSS pass of a line-search try; i8_grad.cu: stored-Y gradient
pass; tcgen05.mma kind::i8, error-free balanced radix-256 digit splitting) against the CPU oracle, through the C ABI
(picard_eval_moments_ex, Picard.fit_with_config).  Bars (BASELINE.json north_star): per-pass G, h, loss from identical W
<= 1e-10 as max|delta| / max|ref|; whole fits Amari <= 1e-6 and iteration counts +-1.
Covers VERDICT r01 "what's weak" #1: default-mode fits at 64 < N <= 128 against the ORACLE, a T > 1e6 point check, the range
guard with adversarial inputs, and the iterate sequence of the INT8 engines against the FP64 kernels."""
import numpy as np
import pytest

import _data
import _gpu
import picard_ica_b200 as P
from oracle import oracle as orc
from picard_ica_b200.utils import amari_distance

pytestmark = pytest.mark.gpu
TOL = 1e-10
FORCE, NO_I8 = P.FLAG_FORCE_INT8, P.FLAG_NO_INT8


def _ref(x, w, kind, alpha):
    return orc.eval_point(x, w, kind, alpha, ortho=False, extended=False)


def _w(n, eps=0.02):
    return _data.orthogonal(n, seed=n + 2) + eps * np.random.default_rng(n).standard_normal((n, n))


@pytest.mark.parametrize("n,t,kind,alpha", [(128, 2050, orc.TANH, 1.0), (100, 1500, orc.TANH, 0.7), (65, 33, orc.TANH, 1.0),
                                            (128, 4097, orc.EXP, 0.1), (70, 4099, orc.CUBE, 1.0), (96, 31, orc.TANH, 1.0),
                                            (64, 3000, orc.TANH, 1.0), (20, 700, orc.TANH, 1.0)])
def test_int8_loss_pass_matches_oracle(n, t, kind, alpha):
    """Raw LOSS moments (log-likelihood and y^2 row sums) of the INT8 pass: ragged last tile, N < 128 padding, T < one tile; N <= 64
    only runs it when forced (the automatic gate is 64 < N <= 128: no faster than the FP64 kernels below)."""
    x = _data.whitened(n, max(t, 2 * n), seed=n)[:, :t]
    w = _w(n)
    ref = _ref(x, w, kind, alpha)
    got, st = _gpu.eval_moments_ex(x, w, kind, alpha, mode=2, want_h=True, flags=FORCE)
    assert st["i8_loss_passes"] == 1 and st["loss_passes"] == 1
    for k in ("lrow", "sq"):
        assert _data.rel_err(got[k], getattr(ref, k)) <= TOL, k


@pytest.mark.parametrize("n,t,kind,alpha", [(128, 2050, orc.TANH, 1.0), (100, 1500, orc.TANH, 0.7), (65, 33, orc.TANH, 1.0),
                                            (128, 4097, orc.EXP, 0.1), (128, 4097, orc.EXP, 1.0), (96, 20001, orc.TANH, 1.3),
                                            (128, 70_000, orc.TANH, 1.0)])
def test_int8_gradient_pass_matches_oracle(n, t, kind, alpha):
    """Gr = psi(Y) Y^T and Sd from the Y' the INT8 LOSS pass stored, both on the tensor cores (mode 3 = the two-kernel path of an
    accepted line-search try)."""
    x = _data.whitened(n, max(t, 2 * n), seed=n + 1)[:, :t]
    w = _w(n)
    ref = _ref(x, w, kind, alpha)
    got, st = _gpu.eval_moments_ex(x, w, kind, alpha, mode=3, want_h=False, flags=FORCE)
    assert st["i8_loss_passes"] == 1 and st["i8_grad_passes"] == 1
    for k in ("gr", "sd", "lrow"):
        assert _data.rel_err(got[k], getattr(ref, k)) <= TOL, k
    # the same point through the FP64 kernels: the engines agree far inside the bar
    fp, st2 = _gpu.eval_moments_ex(x, w, kind, alpha, mode=3, want_h=False, flags=NO_I8)
    assert st2["i8_loss_passes"] == 0 and st2["i8_grad_passes"] == 0
    assert _data.rel_err(got["gr"], fp["gr"]) <= 1e-11


def test_int8_engines_want_h_falls_back_to_the_fp64_gradient():
    """Non-ortho problems need Hr = psi'(Y) (Y^2)^T: the gradient pass stays on the FP64 kernel, the LOSS pass is INT8."""
    x = _data.whitened(128, 3000, seed=4)
    w = _w(128)
    ref = _ref(x, w, orc.TANH, 1.0)
    got, st = _gpu.eval_moments_ex(x, w, orc.TANH, 1.0, mode=3, want_h=True, flags=FORCE)
    assert st["i8_loss_passes"] == 1 and st["i8_grad_passes"] == 0
    for k in ("gr", "sd", "hr", "sq", "lrow"):
        assert _data.rel_err(got[k], getattr(ref, k)) <= TOL, k


def test_point_check_at_full_depth():
    """N = 128, T = 1.25e6 (ragged): every CTA of the gradient pass crosses an accumulator flush (16384 samples), the LOSS pass
    streams ~39000 tiles.  Raw moments vs the oracle, default engine choice for whitened data (no force flag)."""
    n, t = 128, 1_250_003
    x = _data.whitened(n, t, seed=77)
    w = _w(n, 0.01)
    ref = _ref(x, w, orc.TANH, 1.0)
    got, st = _gpu.eval_moments_ex(x, w, orc.TANH, 1.0, mode=3, want_h=False, whitened=True)
    assert st["i8_loss_passes"] == 1 and st["i8_grad_passes"] == 1 and st["i8_fallbacks"] == 0
    assert 1.0 < st["i8_range"] < 16.0
    for k in ("gr", "sd", "lrow"):
        assert _data.rel_err(got[k], getattr(ref, k)) <= TOL, k


# ---- the range guard (VERDICT r01 weak #1 iii / ADVICE low #3) -------------------------------------------------
def test_whitened_promise_is_checked_not_trusted():
    """`covariance = I` promised for data that is not whitened (one row 1e-6 of the others): the range check refuses the INT8
    engines, the FP64 kernels run (counted in stats.i8_fallbacks), the result meets the bar."""
    n, t = 100, 6000
    x = _data.whitened(n, t, seed=5)
    x[7] *= 1e-6
    w = _w(n)
    ref = _ref(x, w, orc.TANH, 1.0)
    got, st = _gpu.eval_moments_ex(x, w, orc.TANH, 1.0, mode=3, want_h=False, whitened=True)
    assert st["i8_fallbacks"] == 1 and st["i8_loss_passes"] == 0 and st["i8_grad_passes"] == 0 and st["i8_range"] > 64.0
    for k in ("gr", "sd", "lrow"):
        assert _data.rel_err(got[k], getattr(ref, k)) <= TOL, k
    # without the promise nothing is even tried
    _, st0 = _gpu.eval_moments_ex(x, w, orc.TANH, 1.0, mode=3, want_h=False, whitened=False)
    assert st0["i8_fallbacks"] == 0 and st0["i8_loss_passes"] == 0


def test_outlier_sample_and_heavy_tails():
    """(a) one sample with a 1e6 x component (its other components lose 20 bits -- one term of T in every sum);
    (b) heavy-tailed sources (Student t, 3 degrees of freedom, whitened): per-sample bounds vary over 3 orders of magnitude.
    Default engine choice; whichever engine the guard picks, the result meets the bar."""
    n, t = 128, 20_000
    x = _data.whitened(n, t, seed=6)
    x[17, 123] = 1e6
    w = _w(n)
    ref = _ref(x, w, orc.TANH, 1.0)
    got, st = _gpu.eval_moments_ex(x, w, orc.TANH, 1.0, mode=3, want_h=False, whitened=True)
    for k in ("gr", "sd", "lrow"):
        assert _data.rel_err(got[k], getattr(ref, k)) <= TOL, (k, st)
    rng = np.random.default_rng(8)
    s = rng.standard_t(3, size=(n, t))
    xm = _data.orthogonal(n, 3) @ s
    xm -= xm.mean(axis=1, keepdims=True)
    d, e = np.linalg.eigh(xm @ xm.T / t)
    xw = (e / np.sqrt(d)).T @ xm
    ref = _ref(xw, w, orc.TANH, 1.0)
    got, st = _gpu.eval_moments_ex(xw, w, orc.TANH, 1.0, mode=3, want_h=False, whitened=True)
    for k in ("gr", "sd", "lrow"):
        assert _data.rel_err(got[k], getattr(ref, k)) <= TOL, (k, st)


def test_unmixing_rows_with_a_wide_dynamic_range():
    """W' whose rows differ by 1e8 in scale and whose entries spread over 1e8 inside a row (each row has its own exponent)."""
    n, t = 128, 5000
    x = _data.whitened(n, t, seed=9)
    rng = np.random.default_rng(10)
    w = _data.orthogonal(n, 11) * np.exp(rng.uniform(-18.0, 0.0, size=(n, n)))   # entries over ~1e8 inside a row
    w *= np.exp(rng.uniform(-9.0, 9.0, size=(n, 1)))                              # rows over ~1e8
    w[3] = _data.orthogonal(n, 12)[3]
    ref = _ref(x, w, orc.TANH, 1.0)
    got, st = _gpu.eval_moments_ex(x, w, orc.TANH, 1.0, mode=2, want_h=True, flags=FORCE)
    assert st["i8_loss_passes"] == 1
    for k in ("lrow", "sq"):
        e = np.max(np.abs(got[k] - getattr(ref, k)) / np.abs(getattr(ref, k)))    # per ROW here: rows of very different scale
        assert e <= TOL, (k, e)


# ---- whole fits in the default mode (INT8 engines auto-on for whitened data) against the oracle -----------------
@pytest.mark.parametrize("n,t", [(96, 50_000), (128, 60_000)])
def test_default_mode_fit_matches_oracle(n, t):
    """The c3 code path (LOSS + gradient passes on the INT8 tensor cores) end to end: same iteration count (+-1), same signs,
    unmixing within Amari 1e-6 of the oracle's."""
    x, a, _ = _data.mixture(n, t, seed=n, kind="mixed")
    w0 = _data.orthogonal(n, 43)
    res = P.Picard.fit_with_config(x, P.PicardConfig(w_init=w0))
    ref = orc.fit(x, orc.Config(w_init=w0))
    assert res.stats["i8_loss_passes"] == res.stats["loss_passes"] > 0
    assert res.stats["i8_grad_passes"] == res.stats["grady_passes"] > 0
    assert res.stats["i8_fallbacks"] == 0
    assert abs(res.n_iterations - ref.n_iterations) <= 1, (res.n_iterations, ref.n_iterations)
    assert res.converged == ref.converged
    assert amari_distance(res.full_unmixing(), np.linalg.pinv(ref.full_unmixing())) <= 1e-6
    np.testing.assert_array_equal(res.signs, ref.signs)


@pytest.mark.parametrize("n", [96, 128])
def test_int8_and_fp64_engines_give_the_same_iterates(n):
    """PICARD_FLAG_NO_INT8 against the default on the same whitened problem: iteration count and final W."""
    x, _, _ = _data.mixture(n, 30_000, seed=5, kind="mixed")
    w0 = _data.orthogonal(n, 43)
    a = P.Picard.fit_with_config(x, P.PicardConfig(w_init=w0, max_iter=80))
    b = P.Picard.fit_with_config(x, P.PicardConfig(w_init=w0, max_iter=80, flags=NO_I8))
    assert a.stats["i8_loss_passes"] > 0 and b.stats["i8_loss_passes"] == 0 and b.stats["i8_grad_passes"] == 0
    assert abs(a.n_iterations - b.n_iterations) <= 1
    assert amari_distance(a.full_unmixing(), np.linalg.pinv(b.full_unmixing())) <= 1e-6


def test_second_device_in_one_process():
    """Kernel attributes / occupancy are cached per device: a fit on device 1 after one on device 0."""
    if P._ffi.lib().picard_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    x, _, _ = _data.mixture(100, 20_000, seed=3)
    w0 = _data.orthogonal(100, 43)
    a = P.Picard.fit_with_config(x, P.PicardConfig(w_init=w0, device=0))
    b = P.Picard.fit_with_config(x, P.PicardConfig(w_init=w0, device=1))
    assert a.n_iterations == b.n_iterations
    np.testing.assert_allclose(a.unmixing, b.unmixing, rtol=0, atol=1e-12)


def test_int8_engines_are_bit_repeatable_at_the_headline_size():
    """N = 128, T = 1e7 (BASELINE configs[1]) on device-resident data: the LOSS + stored-Y gradient pair of the INT8 engines run
    several times must give bit-identical moments every time, and agree with the FP64 kernels to 1e-12.  The first round-2 version of
    the gradient kernel handed a shared-memory stage back to its producer while loads from it were still in flight: about every second
    launch at this size had one tile's worth of error (1e-6 relative) in a handful of rows -- invisible at the sizes of the other
    tests, and found only by this check."""
    import ctypes as C
    import torch
    from picard_ica_b200 import _ffi
    n, t = 128, 10_000_000
    _ffi.lib().picard_release_cache()  # buffers cached by the fits of the earlier tests
    if torch.cuda.mem_get_info(0)[0] < 60 * 2**30:
        pytest.skip("needs 60 GB of device memory")
    ld = t
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(5)
    x1 = torch.empty((n, ld), dtype=torch.float64, device=dev)
    for r0 in range(0, n, 16):
        x1[r0:r0 + 16].normal_(generator=g)
    x1[::2] = x1[::2].sign() * x1[::2].abs() ** 1.5 * 0.75  # heavier tails on the even rows
    torch.cuda.synchronize()
    w = np.ascontiguousarray(_data.orthogonal(n, 7) + 0.01 * np.random.default_rng(11).standard_normal((n, n)))
    lib = _ffi.lib()

    def hp(a):
        return a.ctypes.data_as(_ffi.dp)

    def moments(flags):
        gr = np.zeros((n, n)); sd = np.zeros(n); hr = np.zeros((n, n)); sq = np.zeros(n); lrow = np.zeros(n)
        stt = _ffi.Stats(); err = C.create_string_buffer(1024)
        rc = lib.picard_eval_moments_device_ex(C.c_void_p(x1.data_ptr()), C.c_int64(n), C.c_int64(t), C.c_int64(ld), hp(w), C.c_int32(0),
                                               C.c_double(1.0), C.c_int32(3), C.c_int32(0), C.c_int32(0), C.c_uint32(flags), C.c_int32(1),
                                               C.c_int32(0), None, hp(gr), hp(sd), hp(hr), hp(sq), hp(lrow), C.byref(stt), err, C.c_size_t(1024))
        assert rc == 0, err.value
        return dict(gr=gr, sd=sd, lrow=lrow), stt.as_dict()

    first, st = moments(0)
    assert st["i8_loss_passes"] == 1 and st["i8_grad_passes"] == 1
    for _ in range(7):
        again, _st = moments(0)
        for k in ("gr", "sd", "lrow"):
            np.testing.assert_array_equal(again[k], first[k])
    fp64, st64 = moments(P.FLAG_NO_INT8)
    assert st64["i8_loss_passes"] == 0 and st64["i8_grad_passes"] == 0
    for k in ("gr", "sd", "lrow"):
        assert _data.rel_err(first[k], fp64[k]) <= 1e-12
    del x1
    torch.cuda.empty_cache()
