"""The device-resident entry points of the C ABI: picard_fit_device (data already in HBM, sources left on the device) and the
resumable core loop picard_core_* (what bench.py times).  Both must agree with the host-buffer fit."""
import ctypes as C

import numpy as np
import pytest

import _data
import picard_ica_b200 as P
from oracle import oracle as orc
from picard_ica_b200 import Picard, PicardConfig
from picard_ica_b200.utils import amari_distance

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _to_dev(x):
    t = x.shape[1]
    ld = (t + 15) // 16 * 16
    buf = torch.zeros((x.shape[0], ld), dtype=torch.float64, device="cuda")
    buf[:, :t] = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return buf[:, :t]


def test_fit_device_matches_host_fit():
    x, a, _ = _data.mixture(9, 20_003, seed=6, kind="mixed")
    w0 = _data.orthogonal(9, 43)
    host = Picard.fit_with_config(x, PicardConfig(w_init=w0))
    res, src = Picard.fit_device(_to_dev(x), PicardConfig(w_init=w0), want_sources=True)
    assert res.n_iterations == host.n_iterations and res.converged == host.converged
    np.testing.assert_allclose(res.unmixing, host.unmixing, atol=1e-12)
    np.testing.assert_allclose(res.whitening, host.whitening, atol=1e-12)
    assert res.sources is None  # PICARD_FLAG_KEEP_SOURCES_ON_DEVICE
    np.testing.assert_allclose(src.cpu().numpy(), host.sources, atol=1e-10)
    res2, src2 = Picard.fit_device(_to_dev(x), PicardConfig(w_init=w0), want_sources=False)
    assert src2 is None and res2.n_iterations == host.n_iterations


def test_core_loop_is_resumable_and_matches_the_oracle():
    n, t = 10, 30_000
    xw = _data.whitened(n, t, seed=12)
    ref = orc.core_run(xw, ortho=True, extended=True, covariance=np.eye(n), want_y=False)
    cfg = PicardConfig()
    core = P.CoreLoop(_to_dev(xw), cfg, covariance_identity=True)
    done = 0
    while True:  # three iterations at a time: the state survives between calls
        d, conv = core.run(3)
        done += d
        if conv or d == 0:
            break
    st = core.state()
    assert st["converged"] and st["n_iterations"] == ref.n_iterations
    assert amari_distance(st["w"], np.linalg.inv(ref.w)) <= 1e-9
    np.testing.assert_array_equal(st["signs"], ref.signs)
    assert abs(st["gradient_norm"] - ref.gradient_norm) <= 1e-9
    s = core.stats()
    assert s["loss_passes"] == s["ls_tries"] + 1 and s["grady_passes"] == st["n_iterations"] and s["fused_passes"] == 0
    # reset -> the same run again
    core.reset()
    d2, conv2 = core.run(1000)
    st2 = core.state()
    assert conv2 and st2["n_iterations"] == st["n_iterations"]
    np.testing.assert_allclose(st2["w"], st["w"], atol=1e-13)
    core.close()


def test_release_cache_is_harmless():
    P._ffi.lib().picard_release_cache()
    x, _, _ = _data.mixture(4, 3000, seed=2)
    assert Picard.fit_with_config(x, PicardConfig(random_state=1)).unmixing.shape == (4, 4)
