"""picard_ica_b200 -- host-side (Python) mirror of the reference's public interface for the fit path.

Same names, argument meaning and error behaviour as lmmx/picard-ica v0.1.6 (`src/lib.rs:50-60` re-exports):
`Picard.fit / fit_with_config / transform` (solver.rs:33,45,199), `PicardConfig` + `ConfigBuilder`
(config.rs:11-273), `DensityType` (density.rs:137-176), `PicardResult` (result.rs:7-64), `PicardError`
(error.rs:9-42) and `utils.amari_distance / permute` (utils.rs:16-103).  Every computation goes through the
C ABI of libpicard_b200.so (CUDA, sm_100a); there is no CPU fallback -- importing works without a GPU (so
configs can be built and validated), computing without one raises `PicardError.ComputationError`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field, replace
from typing import Optional

import numpy as np

from . import _ffi
from ._ffi import (FLAG_FORCE_INT8, FLAG_FORCE_SPECULATION, FLAG_KEEP_SOURCES_ON_DEVICE, FLAG_NO_INT8, FLAG_NO_SPECULATION,
                   FLAG_NO_Y_STORE, LibraryMissing)

__all__ = [
    "Picard", "PicardConfig", "ConfigBuilder", "DensityType", "Tanh", "Exp", "Cube", "PicardResult", "PicardError", "utils",
    "CoreLoop", "FLAG_NO_SPECULATION", "FLAG_KEEP_SOURCES_ON_DEVICE", "FLAG_NO_Y_STORE", "FLAG_FORCE_SPECULATION", "FLAG_NO_INT8", "FLAG_FORCE_INT8", "LibraryMissing",
]

_dp = _ffi.dp


# ------------------------------------------------------------------------------------------------------
# errors (error.rs:9-42)
# ------------------------------------------------------------------------------------------------------
class PicardError(Exception):
    """Base of the reference's `PicardError` variants; `str(e)` is the reference's Display text."""
    status = -1


class InvalidDimensions(PicardError):
    status = 1


class SingularMatrix(PicardError):
    status = 2


class ComputationError(PicardError):
    status = 3


class InvalidConfig(PicardError):
    status = 4

    def __init__(self, msg, parameter=None):
        super().__init__(msg)
        self.parameter = parameter
        if parameter is None and "'" in msg:
            self.parameter = msg.split("'")[1]


PicardError.InvalidDimensions = InvalidDimensions
PicardError.SingularMatrix = SingularMatrix
PicardError.ComputationError = ComputationError
PicardError.InvalidConfig = InvalidConfig
_BY_STATUS = {1: InvalidDimensions, 2: SingularMatrix, 3: ComputationError, 4: InvalidConfig}


def _raise(status: int, msg: str):
    if not msg:
        msg = _ffi.lib().picard_status_string(status).decode()
    raise _BY_STATUS.get(status, ComputationError)(msg)


# ------------------------------------------------------------------------------------------------------
# densities (density.rs:24-176): a closed set of three kinds with one alpha parameter
# ------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Tanh:
    alpha: float = 1.0
    kind = 0


@dataclass(frozen=True)
class Exp:
    alpha: float = 1.0
    kind = 1


@dataclass(frozen=True)
class Cube:
    kind = 2
    alpha = 1.0


class DensityType:
    """density.rs:137-176 constructors."""

    @staticmethod
    def tanh():
        return Tanh()

    @staticmethod
    def tanh_with_alpha(alpha: float):
        return Tanh(float(alpha))

    @staticmethod
    def exp():
        return Exp()

    @staticmethod
    def exp_with_alpha(alpha: float):
        return Exp(float(alpha))

    @staticmethod
    def cube():
        return Cube()

    @staticmethod
    def default():
        return Tanh()


# ------------------------------------------------------------------------------------------------------
# config (config.rs)
# ------------------------------------------------------------------------------------------------------
@dataclass
class PicardConfig:
    """config.rs:11-85, same fields and defaults.  `device`, `comm`, `flags` are execution placement (not in
    the reference): CUDA ordinal, sample-axis communicator (picard_ica_b200.dist) and PICARD_FLAG_* bits."""
    density: object = field(default_factory=Tanh)
    n_components: Optional[int] = None
    ortho: bool = True
    extended: Optional[bool] = None
    whiten: bool = True
    centering: bool = True
    max_iter: int = 500
    tol: float = 1e-7
    m: int = 7
    ls_tries: int = 10
    lambda_min: float = 0.01
    w_init: Optional[np.ndarray] = None
    fastica_it: Optional[int] = None
    jade_it: Optional[int] = None
    random_state: Optional[int] = None
    verbose: bool = False
    device: int = -1
    comm: object = None
    flags: int = 0

    @staticmethod
    def new():
        return PicardConfig()

    @staticmethod
    def builder():
        return ConfigBuilder()

    def effective_extended(self) -> bool:  # config.rs:99-101
        return self.ortho if self.extended is None else bool(self.extended)

    def _to_c(self):
        c = _ffi.Config()
        _ffi.lib().picard_config_default(C.byref(c))
        keep = None
        c.density_kind = int(self.density.kind)
        c.alpha = float(getattr(self.density, "alpha", 1.0))
        c.n_components = -1 if self.n_components is None else int(self.n_components)
        c.ortho = int(bool(self.ortho))
        c.extended = -1 if self.extended is None else int(bool(self.extended))
        c.whiten = int(bool(self.whiten))
        c.centering = int(bool(self.centering))
        # usize fields: a negative Python int has no Rust counterpart; clamp into the "invalid" range validate() rejects
        c.max_iter = max(int(self.max_iter), 0)
        c.tol = float(self.tol)
        c.m = max(int(self.m), 0)
        c.ls_tries = max(int(self.ls_tries), 0)
        c.lambda_min = float(self.lambda_min)
        if self.w_init is not None:
            keep = np.ascontiguousarray(self.w_init, dtype=np.float64)
            if keep.ndim != 2:
                raise InvalidDimensions(f"Invalid dimensions: w_init must be 2-dimensional, got shape {list(keep.shape)}")
            c.w_init = keep.ctypes.data_as(_dp)
            c.w_init_rows, c.w_init_cols = keep.shape
        c.fastica_it = -1 if self.fastica_it is None else int(self.fastica_it)
        c.jade_it = -1 if self.jade_it is None else int(self.jade_it)
        c.has_seed = int(self.random_state is not None)
        c.seed = int(self.random_state or 0)
        c.verbose = int(bool(self.verbose))
        c.device = int(self.device)
        c.comm = getattr(self.comm, "handle", self.comm)
        c.flags = int(self.flags)
        return c, keep

    def validate(self) -> None:
        """config.rs:104-142; raises PicardError.InvalidConfig with the offending parameter name."""
        c, _keep = self._to_c()
        err = C.create_string_buffer(512)
        st = _ffi.lib().picard_config_validate(C.byref(c), err, C.c_size_t(512))
        if st != 0:
            _raise(st, err.value.decode())


class ConfigBuilder:
    """Fluent builder, config.rs:145-273."""

    def __init__(self):
        self._c = PicardConfig()

    def _set(self, **kw):
        self._c = replace(self._c, **kw)
        return self

    def density(self, d): return self._set(density=d)
    def n_components(self, n): return self._set(n_components=int(n))
    def ortho(self, v): return self._set(ortho=bool(v))
    def extended(self, v): return self._set(extended=bool(v))
    def whiten(self, v): return self._set(whiten=bool(v))
    def centering(self, v): return self._set(centering=bool(v))
    def max_iter(self, v): return self._set(max_iter=int(v))
    def tol(self, v): return self._set(tol=float(v))
    def m(self, v): return self._set(m=int(v))
    def ls_tries(self, v): return self._set(ls_tries=int(v))
    def lambda_min(self, v): return self._set(lambda_min=float(v))
    def w_init(self, w): return self._set(w_init=np.asarray(w, dtype=np.float64))
    def fastica_it(self, v): return self._set(fastica_it=int(v))
    def jade_it(self, v): return self._set(jade_it=int(v))
    def random_state(self, v): return self._set(random_state=int(v))
    def verbose(self, v): return self._set(verbose=bool(v))
    def device(self, v): return self._set(device=int(v))
    def comm(self, v): return self._set(comm=v)
    def flags(self, v): return self._set(flags=int(v))

    def build(self) -> PicardConfig:
        return self._c

    def build_validated(self) -> PicardConfig:
        self._c.validate()
        return self._c


# ------------------------------------------------------------------------------------------------------
# result (result.rs)
# ------------------------------------------------------------------------------------------------------
@dataclass
class PicardResult:
    """result.rs:7-33.  `sources` holds THIS rank's sample columns in a multi-GPU fit."""
    whitening: Optional[np.ndarray]
    unmixing: np.ndarray
    sources: Optional[np.ndarray]
    mean: Optional[np.ndarray]
    n_iterations: int
    converged: bool
    gradient_norm: float
    signs: Optional[np.ndarray]
    stats: dict = field(default_factory=dict)

    def full_unmixing(self) -> np.ndarray:  # result.rs:39-44
        return self.unmixing @ self.whitening if self.whitening is not None else self.unmixing.copy()

    def mixing(self) -> np.ndarray:  # result.rs:49-64 (host-side N x N, as in the reference)
        w = self.full_unmixing()
        inv = _invert_matrix(w.T @ w)
        return inv @ w.T if inv is not None else w.T.copy()  # fallback: transpose (result.rs:62-63)

    def _to_c(self):
        r = _ffi.Result()
        keep = []

        def put(a):
            if a is None:
                return None
            b = np.ascontiguousarray(a, dtype=np.float64)
            keep.append(b)
            return b.ctypes.data_as(_dp)

        r.n_components = self.unmixing.shape[0]
        r.n_features = self.whitening.shape[1] if self.whitening is not None else self.unmixing.shape[0]
        r.whitening = put(self.whitening)
        r.unmixing = put(self.unmixing)
        r.mean = put(self.mean)
        return r, keep


def _invert_matrix(m: np.ndarray):
    """result.rs:67-129: Gauss-Jordan with partial pivoting; None when a pivot is below 1e-15."""
    n = m.shape[0]
    aug = np.hstack([np.array(m, dtype=np.float64), np.eye(n)])
    for i in range(n):
        max_row = i + int(np.argmax(np.abs(aug[i:, i])))
        if max_row != i:
            aug[[i, max_row]] = aug[[max_row, i]]
        if abs(aug[i, i]) < 1e-15:
            return None
        aug[i] /= aug[i, i]
        for k in range(n):
            if k != i:
                aug[k] -= aug[k, i] * aug[i]
    return aug[:, n:].copy()


def _adopt(ptr, shape):
    """Wrap a library-owned buffer as an ndarray WITHOUT copying (sources can be tens of GB); the buffer is released
    through picard_result_free when the array and every view of it are gone."""
    import weakref
    a = np.ctypeslib.as_array(ptr, shape=shape)
    holder = _ffi.Result()
    holder.sources = ptr  # picard_result_free frees every non-null member: a result holding only this buffer
    weakref.finalize(a, lambda h=holder: _ffi.lib().picard_result_free(C.byref(h)))
    return a


def _from_c_result(r: _ffi.Result) -> PicardResult:
    nc, nf, t = r.n_components, r.n_features, r.n_samples

    def arr(ptr, shape):
        return np.ctypeslib.as_array(ptr, shape=shape).copy() if ptr else None

    sources = None
    if r.sources:
        sources = _adopt(r.sources, (nc, t))
        r.sources = None  # ownership moved to the ndarray
    out = PicardResult(arr(r.whitening, (nc, nf)), arr(r.unmixing, (nc, nc)), sources, arr(r.mean, (nf,)),
                       int(r.n_iterations), bool(r.converged), float(r.gradient_norm), arr(r.signs, (nc,)), r.stats.as_dict())
    _ffi.lib().picard_result_free(C.byref(r))
    return out


def _as_matrix(x):
    a = np.asarray(x, dtype=np.float64)
    if a.ndim != 2:
        raise InvalidDimensions(f"Invalid dimensions: expected a 2-D (n_features, n_samples) array, got shape {list(a.shape)}")
    if a.size and a.strides[1] != a.itemsize:  # the library wants a unit inner stride (as_standard_layout in the Rust shim)
        a = np.ascontiguousarray(a)
    return a


# ------------------------------------------------------------------------------------------------------
# solver (solver.rs)
# ------------------------------------------------------------------------------------------------------
class Picard:
    """solver.rs:23-215: static methods only."""

    @staticmethod
    def fit(x) -> PicardResult:  # solver.rs:33
        return Picard.fit_with_config(x, PicardConfig())

    @staticmethod
    def fit_with_config(x, config: PicardConfig) -> PicardResult:  # solver.rs:45
        config.validate()  # solver.rs:46 -- before looking at the data, like the reference
        a = _as_matrix(x)
        n, p = a.shape
        c, _keep = config._to_c()
        r = _ffi.Result()
        err = C.create_string_buffer(1024)
        ptr = a.ctypes.data_as(_dp) if a.size else None
        stride = a.strides[0] // a.itemsize if a.size else max(p, 1)
        st = _ffi.lib().picard_fit(ptr, C.c_int64(n), C.c_int64(p), C.c_int64(stride), C.byref(c), C.byref(r), err, C.c_size_t(1024))
        if st != 0:
            _raise(st, err.value.decode())
        return _from_c_result(r)

    @staticmethod
    def fit_device(x_dev, config: PicardConfig, want_sources: bool = False):
        """`x_dev`: a CUDA torch tensor (n_features, n_samples_local), f64, unit inner stride, even row stride.
        Returns (PicardResult, sources tensor or None).  The product path for data that already lives in HBM."""
        import torch
        config.validate()
        assert x_dev.is_cuda and x_dev.dtype == torch.float64 and x_dev.dim() == 2 and x_dev.stride(1) == 1
        n, p = x_dev.shape
        cfg = replace(config, device=x_dev.device.index, flags=config.flags | FLAG_KEEP_SOURCES_ON_DEVICE)
        c, _keep = cfg._to_c()
        nc = n if not config.whiten else min(n if config.n_components is None else config.n_components, n)
        src = None
        lds = 0
        if want_sources:
            lds = (p + 15) // 16 * 16
            src = torch.empty((nc, lds), dtype=torch.float64, device=x_dev.device)
        torch.cuda.current_stream(x_dev.device).synchronize()
        r = _ffi.Result()
        err = C.create_string_buffer(1024)
        st = _ffi.lib().picard_fit_device(C.c_void_p(x_dev.data_ptr()), C.c_int64(n), C.c_int64(p), C.c_int64(x_dev.stride(0)),
                                          C.byref(c), C.c_void_p(src.data_ptr() if src is not None else 0), C.c_int64(lds),
                                          C.byref(r), err, C.c_size_t(1024))
        if st != 0:
            _raise(st, err.value.decode())
        res = _from_c_result(r)
        return res, (src[:, :p] if src is not None else None)

    @staticmethod
    def transform(x, result: PicardResult, device: int = -1) -> np.ndarray:  # solver.rs:199
        a = _as_matrix(x)
        n, p = a.shape
        r, _keep = result._to_c()
        out = np.empty((int(r.n_components), p), dtype=np.float64)
        err = C.create_string_buffer(1024)
        ptr = a.ctypes.data_as(_dp) if a.size else None
        stride = a.strides[0] // a.itemsize if a.size else max(p, 1)
        st = _ffi.lib().picard_transform(ptr, C.c_int64(n), C.c_int64(p), C.c_int64(stride), C.byref(r), out.ctypes.data_as(_dp),
                                         C.c_int32(device), err, C.c_size_t(1024))
        if st != 0:
            _raise(st, err.value.decode())
        return out


class CoreLoop:
    """The core loop alone (core::run, core.rs:162-401), resumable, on preprocessed device data.
    What bench.py's kernel-level `value` times."""

    def __init__(self, x_dev, config: PicardConfig, covariance_identity: bool = True):
        import torch
        assert x_dev.is_cuda and x_dev.dtype == torch.float64 and x_dev.dim() == 2 and x_dev.stride(1) == 1
        self._x = x_dev  # keep alive
        self.n, self.t = x_dev.shape
        cfg = replace(config, device=x_dev.device.index)
        c, self._keep = cfg._to_c()
        self._h = C.c_void_p()
        err = C.create_string_buffer(1024)
        torch.cuda.current_stream(x_dev.device).synchronize()
        st = _ffi.lib().picard_core_create(C.byref(self._h), C.c_void_p(x_dev.data_ptr()), C.c_int64(self.n), C.c_int64(self.t),
                                           C.c_int64(x_dev.stride(0)), C.byref(c), C.c_int32(int(covariance_identity)), err,
                                           C.c_size_t(1024))
        if st != 0:
            _raise(st, err.value.decode())

    def run(self, max_new_iters: int):
        done = C.c_int64(); conv = C.c_int32()
        err = C.create_string_buffer(1024)
        st = _ffi.lib().picard_core_run(self._h, C.c_int64(max_new_iters), C.byref(done), C.byref(conv), err, C.c_size_t(1024))
        if st != 0:
            _raise(st, err.value.decode())
        return int(done.value), bool(conv.value)

    def reset(self):
        st = _ffi.lib().picard_core_reset(self._h)
        if st != 0:
            _raise(st, "")

    def state(self):
        w = np.empty((self.n, self.n)); signs = np.ones(self.n)
        nit = C.c_int64(); conv = C.c_int32(); gn = C.c_double(); loss = C.c_double()
        st = _ffi.lib().picard_core_state(self._h, w.ctypes.data_as(_dp), signs.ctypes.data_as(_dp), C.byref(nit), C.byref(conv),
                                          C.byref(gn), C.byref(loss))
        if st != 0:
            _raise(st, "")
        return dict(w=w, signs=signs, n_iterations=int(nit.value), converged=bool(conv.value), gradient_norm=gn.value, loss=loss.value)

    def stats(self) -> dict:
        s = _ffi.Stats()
        _ffi.lib().picard_core_stats(self._h, C.byref(s))
        return s.as_dict()

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _ffi.lib().picard_core_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


from . import utils  # noqa: E402
