"""CPU-side checks of the drop-in boundary: libpicard_b200.so loads and exports every symbol include/picard_b200.h
declares; config defaults / validation mirror config.rs; the host mirror has the reference's names; and compute
entry points fail LOUDLY without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import picard_ica_b200 as P
from picard_ica_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "picard_b200.h")


def _header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(picard_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _ffi.lib()
    names = _header_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_ffi.EXPORTED) == names  # the ctypes binding lists exactly the header's entry points


def test_abi_version_and_status_strings():
    lib = _ffi.lib()
    assert lib.picard_abi_version() == 2
    # Display texts of error.rs:44-74
    assert lib.picard_status_string(2).decode() == "Singular matrix encountered during computation"
    assert lib.picard_status_string(1).decode().startswith("Invalid dimensions")


def test_struct_layouts_match_header_sizes():
    # natural alignment on x86-64: computed by hand from include/picard_b200.h
    assert C.sizeof(_ffi.Config) == 160
    assert C.sizeof(_ffi.Stats) == 176
    assert C.sizeof(_ffi.Result) == 88 + 176


def test_config_defaults():  # config.rs:64-85
    c = _ffi.Config()
    _ffi.lib().picard_config_default(C.byref(c))
    assert (c.density_kind, c.alpha, c.n_components, c.ortho, c.extended, c.whiten, c.centering) == (0, 1.0, -1, 1, -1, 1, 1)
    assert (c.max_iter, c.tol, c.m, c.ls_tries, c.lambda_min) == (500, 1e-7, 7, 10, 0.01)
    assert (c.fastica_it, c.jade_it, c.has_seed, c.verbose) == (-1, -1, 0, 0)
    py = P.PicardConfig()
    assert (py.max_iter, py.tol, py.m, py.ls_tries, py.lambda_min, py.ortho, py.extended, py.whiten, py.centering) == \
        (500, 1e-7, 7, 10, 0.01, True, None, True, True)
    assert py.effective_extended() is True and P.PicardConfig(ortho=False).effective_extended() is False  # config.rs:99-101


@pytest.mark.parametrize("kw,param", [(dict(max_iter=0), "max_iter"), (dict(tol=0.0), "tol"), (dict(tol=-1.0), "tol"),
                                      (dict(lambda_min=0.0), "lambda_min"), (dict(m=0), "m"),
                                      (dict(fastica_it=3, jade_it=3), "jade_it")])
def test_validate_rejects(kw, param):  # config.rs:104-142
    with pytest.raises(P.PicardError.InvalidConfig) as e:
        P.PicardConfig(**kw).validate()
    assert e.value.parameter == param and f"Invalid configuration for '{param}'" in str(e.value)


def test_validate_order_matches_reference():
    """config.rs:104-142 checks max_iter, tol, lambda_min, m, then the warm-start pair: first failure wins."""
    with pytest.raises(P.PicardError.InvalidConfig) as e:
        P.PicardConfig(max_iter=0, tol=0.0, m=0).validate()
    assert e.value.parameter == "max_iter"


def test_builder_mirrors_reference_names():  # config.rs:145-273
    cfg = (P.ConfigBuilder().density(P.DensityType.exp_with_alpha(0.1)).n_components(3).ortho(False).extended(True).whiten(True)
           .centering(False).max_iter(100).tol(1e-6).m(5).ls_tries(4).lambda_min(0.1).w_init(np.eye(3)).jade_it(10)
           .random_state(42).verbose(False).build_validated())
    assert cfg.density == P.Exp(0.1) and cfg.n_components == 3 and not cfg.ortho and cfg.extended and cfg.max_iter == 100
    assert P.DensityType.default() == P.Tanh(1.0) and P.DensityType.tanh_with_alpha(2.0).alpha == 2.0 and P.DensityType.cube().kind == 2
    for name in ("fit", "fit_with_config", "transform"):  # solver.rs:33,45,199
        assert callable(getattr(P.Picard, name))


def test_result_methods():  # result.rs:39-64
    k = np.array([[2.0, 0.0, 0.0], [0.0, 0.5, 0.0]])
    u = np.array([[0.0, 1.0], [1.0, 0.0]])
    r = P.PicardResult(k, u, None, np.zeros(3), 1, True, 0.0, None)
    np.testing.assert_allclose(r.full_unmixing(), u @ k)
    np.testing.assert_allclose(r.mixing(), (u @ k).T)  # W^T W (3 x 3, rank 2) is singular: transpose fallback, result.rs:62-63
    r2 = P.PicardResult(None, u, None, None, 1, True, 0.0, None)
    np.testing.assert_allclose(r2.full_unmixing(), u)
    np.testing.assert_allclose(r2.mixing() @ u, np.eye(2), atol=1e-12)


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point returns ComputationError; with one this test is moot."""
    if _ffi.lib().picard_device_count() > 0:
        pytest.skip("a CUDA device is present")
    x = np.random.default_rng(0).standard_normal((3, 100))
    with pytest.raises(P.PicardError.ComputationError) as e:
        P.Picard.fit(x)
    assert "no usable CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)
    with pytest.raises(P.PicardError.ComputationError):
        P.Picard.transform(x, P.PicardResult(None, np.eye(3), None, None, 1, True, 0.0, None))


def test_config_errors_come_before_device_errors():
    """solver.rs:46: validate() runs first, so an invalid config is reported even on a box without a GPU."""
    with pytest.raises(P.PicardError.InvalidConfig):
        P.Picard.fit_with_config(np.zeros((2, 10)), P.PicardConfig(max_iter=0))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "picard-ica_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("SURVEY", ""), f"{f} mentions the oracle"
