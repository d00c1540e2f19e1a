"""ctypes binding of the CPU oracle (oracle/picard_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package (picard-ica_b200/) never does.  See the header of picard_oracle.cpp for the
parity-pin status ("parity unpinned" for G/h/loss/iterates: the Rust reference cannot be built here and its
tests hold no golden vectors for them).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpicard_oracle.so")

TANH, EXP, CUBE = 0, 1, 2
OK, INVALID_DIMENSIONS, SINGULAR, COMPUTATION, INVALID_CONFIG = 0, 1, 2, 3, 4
LOSS_VALUE, LOSS_SINGULAR, LOSS_ERROR = 0, 1, 2

_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile the oracle with g++ against the image's OpenBLAS (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "picard_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


class _Config(C.Structure):
    _fields_ = [
        ("density_kind", C.c_int32), ("alpha", C.c_double), ("n_components", C.c_int64),
        ("ortho", C.c_int32), ("extended", C.c_int32), ("whiten", C.c_int32), ("centering", C.c_int32),
        ("max_iter", C.c_int64), ("tol", C.c_double), ("m", C.c_int64), ("ls_tries", C.c_int64),
        ("lambda_min", C.c_double), ("w_init", _dp), ("fastica_it", C.c_int64), ("jade_it", C.c_int64),
        ("has_seed", C.c_int32), ("seed", C.c_uint64), ("verbose", C.c_int32),
    ]


class _Result(C.Structure):
    _fields_ = [
        ("n_components", C.c_int64), ("n_features", C.c_int64), ("n_samples", C.c_int64),
        ("whitening", _dp), ("unmixing", _dp), ("sources", _dp), ("mean", _dp),
        ("n_iterations", C.c_int64), ("converged", C.c_int32), ("gradient_norm", C.c_double), ("signs", _dp),
        ("w_init_used", _dp), ("loss_evals", C.c_int64), ("grad_evals", C.c_int64), ("ls_tries_total", C.c_int64),
        ("fallbacks", C.c_int64), ("trace", _dp), ("trace_rows", C.c_int64), ("core_seconds", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_amari.restype = C.c_double
        _lib.orc_get_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _c(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


def set_threads(n: int):
    lib().orc_set_threads(int(n))


def get_threads() -> int:
    return int(lib().orc_get_threads())


def log_lik(kind, alpha, y):
    y = _c(y); out = np.empty_like(y)
    lib().orc_log_lik(C.c_int(kind), C.c_double(alpha), _p(y), C.c_int64(y.size), _p(out))
    return out


def score_and_der(kind, alpha, y):
    y = _c(np.atleast_2d(y)); a = np.empty_like(y); b = np.empty_like(y)
    lib().orc_score_and_der(C.c_int(kind), C.c_double(alpha), _p(y), C.c_int64(y.shape[0]), C.c_int64(y.shape[1]), _p(a), _p(b))
    return a, b


def sln_det(m):
    m = _c(m); s = C.c_double(); l = C.c_double()
    st = lib().orc_sln_det(_p(m), C.c_int64(m.shape[0]), C.byref(s), C.byref(l))
    return st, s.value, l.value


def sym_decorrelation(w):
    w = _c(w); out = np.empty_like(w)
    st = lib().orc_sym_decorrelation(_p(w), C.c_int64(w.shape[0]), _p(out))
    return st, out


def matrix_exp(a):
    a = _c(a); out = np.empty_like(a)
    lib().orc_matrix_exp(_p(a), C.c_int64(a.shape[0]), _p(out))
    return out


def skew(a):
    a = _c(a); out = np.empty_like(a)
    lib().orc_skew(_p(a), C.c_int64(a.shape[0]), _p(out))
    return out


def regularize_hessian(h, hoff, lambda_min):
    h = _c(h).copy(); hoff = _c(hoff)
    lib().orc_regularize_hessian(_p(h), _p(hoff), C.c_int64(h.shape[0]), C.c_double(lambda_min))
    return h


def compute_direction(g, h, hoff, s_list, y_list, r_list, ortho):
    g = _c(g); h = _c(h); hoff = _c(hoff); n = g.shape[0]
    L = len(r_list)
    s = _c(np.asarray(s_list).reshape(L, n, n)) if L else np.zeros((0, n, n))
    y = _c(np.asarray(y_list).reshape(L, n, n)) if L else np.zeros((0, n, n))
    r = _c(np.asarray(r_list, dtype=np.float64)) if L else np.zeros(0)
    out = np.empty_like(g)
    lib().orc_compute_direction(_p(g), _p(h), _p(hoff), C.c_int64(n), _p(s), _p(y), _p(r), C.c_int64(L), C.c_int(int(ortho)), _p(out))
    return out


def compute_loss(y, w, kind, alpha, signs, ortho, extended):
    y = _c(y); w = _c(w); signs = _c(signs); out = C.c_double()
    st = lib().orc_compute_loss(_p(y), _p(w), C.c_int64(y.shape[0]), C.c_int64(y.shape[1]), C.c_int(kind), C.c_double(alpha),
                                _p(signs), C.c_int(int(ortho)), C.c_int(int(extended)), C.byref(out))
    return st, out.value


@dataclass
class EvalPoint:
    gr: np.ndarray; sd: np.ndarray; hr: np.ndarray; sq: np.ndarray; lrow: np.ndarray
    g: np.ndarray; h: np.ndarray; hoff: np.ndarray; signs: np.ndarray
    sign_change: bool; gradient_norm: float; loss: float; loss_status: int


def eval_point(x, w=None, kind=TANH, alpha=1.0, ortho=True, extended=True, lambda_min=0.01, c=None, old_signs=None,
               loss_signs=None) -> EvalPoint:
    """Raw moments + literal core.rs:215-293 outputs + loss at Y = W X (SURVEY.md §8a contract)."""
    x = _c(x); n, t = x.shape
    w_ = None if w is None else _c(w, (n, n))
    c_ = None if c is None else _c(c, (n, n))
    os_ = None if old_signs is None else _c(old_signs, (n,))
    ls_ = None if loss_signs is None else _c(loss_signs, (n,))
    gr = np.empty((n, n)); hr = np.empty((n, n)); sd = np.empty(n); sq = np.empty(n); lrow = np.empty(n)
    g = np.empty((n, n)); h = np.empty((n, n)); hoff = np.empty(n); signs = np.empty(n)
    sc = C.c_int32(); gn = C.c_double(); loss = C.c_double(); lst = C.c_int32()
    lib().orc_eval_point(_p(x), C.c_int64(n), C.c_int64(t), _p(w_), C.c_int(kind), C.c_double(alpha), C.c_int(int(ortho)),
                         C.c_int(int(extended)), C.c_double(lambda_min), _p(c_), _p(os_), _p(ls_),
                         _p(gr), _p(sd), _p(hr), _p(sq), _p(lrow), _p(g), _p(h), _p(hoff), _p(signs),
                         C.byref(sc), C.byref(gn), C.byref(loss), C.byref(lst))
    return EvalPoint(gr, sd, hr, sq, lrow, g, h, hoff, signs, bool(sc.value), gn.value, loss.value, lst.value)


def center(x):
    x = _c(x); out = np.empty_like(x); mean = np.empty(x.shape[0])
    lib().orc_center(_p(x), C.c_int64(x.shape[0]), C.c_int64(x.shape[1]), _p(out), _p(mean))
    return out, mean


def whiten(x, n_components):
    x = _c(x); nf, t = x.shape
    data = np.empty((n_components, t)); k = np.empty((n_components, nf))
    st = lib().orc_whiten(_p(x), C.c_int64(nf), C.c_int64(t), C.c_int64(n_components), _p(data), _p(k))
    return st, data, k


def cumulants(x):
    x = _c(x); n, t = x.shape
    out = np.empty((n * (n + 1) // 2, n, n))
    lib().orc_cumulants(_p(x), C.c_int64(n), C.c_int64(t), _p(out))
    return out


def jade(x, max_iter, tol=1e-6, verbose=False):
    x = _c(x); n, t = x.shape
    w = np.empty((n, n)); sw = C.c_int64()
    st = lib().orc_jade(_p(x), C.c_int64(n), C.c_int64(t), C.c_int64(max_iter), C.c_double(tol), C.c_int(int(verbose)), _p(w), C.byref(sw))
    return st, w, sw.value


def amari(w, a) -> float:
    w = _c(w); a = _c(a)
    return float(lib().orc_amari(_p(w), _p(a), C.c_int64(w.shape[0])))


def randn(seed, n):
    out = np.empty(n)
    lib().orc_randn(C.c_uint64(seed), C.c_int64(n), _p(out))
    return out


@dataclass
class Config:
    """Mirror of PicardConfig (config.rs:11-85), defaults identical."""
    density: int = TANH
    alpha: float = 1.0
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


@dataclass
class Result:
    whitening: Optional[np.ndarray]; unmixing: np.ndarray; sources: np.ndarray; mean: Optional[np.ndarray]
    n_iterations: int; converged: bool; gradient_norm: float; signs: Optional[np.ndarray]
    w_init_used: np.ndarray = None
    loss_evals: int = 0; grad_evals: int = 0; ls_tries_total: int = 0; fallbacks: int = 0
    trace: np.ndarray = field(default=None); core_seconds: float = 0.0

    def full_unmixing(self):  # result.rs:39-44
        return self.unmixing @ self.whitening if self.whitening is not None else self.unmixing.copy()


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}"); self.code = code; self.msg = msg


def _mk_cfg(cfg: Config):
    c = _Config()
    keep = None
    c.density_kind = cfg.density; c.alpha = cfg.alpha
    c.n_components = -1 if cfg.n_components is None else cfg.n_components
    c.ortho = int(cfg.ortho); c.extended = -1 if cfg.extended is None else int(cfg.extended)
    c.whiten = int(cfg.whiten); c.centering = int(cfg.centering)
    c.max_iter = cfg.max_iter; c.tol = cfg.tol; c.m = cfg.m; c.ls_tries = cfg.ls_tries; c.lambda_min = cfg.lambda_min
    if cfg.w_init is not None:
        keep = _c(cfg.w_init); c.w_init = _p(keep)
    c.fastica_it = -1 if cfg.fastica_it is None else cfg.fastica_it
    c.jade_it = -1 if cfg.jade_it is None else cfg.jade_it
    c.has_seed = int(cfg.random_state is not None); c.seed = cfg.random_state or 0
    c.verbose = int(cfg.verbose)
    return c, keep


def validate(cfg: Config):
    c, _keep = _mk_cfg(cfg)
    err = C.create_string_buffer(256)
    st = lib().orc_validate(C.byref(c), err, C.c_size_t(256))
    return st, err.value.decode()


def fit(x, cfg: Config = None) -> Result:
    """Picard::fit_with_config (solver.rs:45-189)."""
    cfg = cfg or Config()
    x = np.asarray(x, dtype=np.float64)
    n, p = (x.shape + (0, 0))[:2] if x.ndim == 2 else (0, 0)
    x = _c(x) if x.size else np.zeros((max(n, 0), max(p, 0)))
    c, _keep = _mk_cfg(cfg)
    r = _Result(); err = C.create_string_buffer(512)
    st = lib().orc_fit(_p(x), C.c_int64(n), C.c_int64(p), C.byref(c), C.byref(r), err, C.c_size_t(512))
    if st != OK:
        raise OracleError(st, err.value.decode())
    nc, nf, t = r.n_components, r.n_features, r.n_samples

    def arr(ptr, shape):
        if not ptr:
            return None
        return np.ctypeslib.as_array(ptr, shape=shape).copy()

    out = Result(arr(r.whitening, (nc, nf)), arr(r.unmixing, (nc, nc)), arr(r.sources, (nc, t)), arr(r.mean, (nf,)),
                 int(r.n_iterations), bool(r.converged), float(r.gradient_norm), arr(r.signs, (nc,)),
                 arr(r.w_init_used, (nc, nc)), int(r.loss_evals), int(r.grad_evals), int(r.ls_tries_total), int(r.fallbacks),
                 arr(r.trace, (r.trace_rows, 7)) if r.trace_rows else np.zeros((0, 7)), float(r.core_seconds))
    lib().orc_result_free(C.byref(r))
    return out


def transform(x, res: Result):
    x = _c(x); nf, t = x.shape; nc = res.unmixing.shape[0]
    out = np.empty((nc, t))
    mean = None if res.mean is None else _c(res.mean)
    k = None if res.whitening is None else _c(res.whitening)
    u = _c(res.unmixing)
    lib().orc_transform(_p(x), C.c_int64(nf), C.c_int64(t), _p(mean), _p(k), _p(u), C.c_int64(nc), _p(out))
    return out


@dataclass
class CoreResult:
    y: np.ndarray; w: np.ndarray; converged: bool; gradient_norm: float; n_iterations: int; signs: np.ndarray
    trace: np.ndarray; loss_evals: int; grad_evals: int; ls_tries_total: int; fallbacks: int; seconds: float


def core_run(x, kind=TANH, alpha=1.0, ortho=True, extended=True, m=7, max_iter=500, tol=1e-7, lambda_min=0.01, ls_tries=10,
             verbose=False, covariance=None, want_y=True) -> CoreResult:
    """core::run (core.rs:162-401) on already-preprocessed data."""
    x = _c(x); n, t = x.shape
    cov = None if covariance is None else _c(covariance, (n, n))
    y = np.empty((n, t)) if want_y else None
    w = np.empty((n, n)); signs = np.ones(n)
    trace = np.zeros((max_iter, 7)); rows = C.c_int64(); cnt = (C.c_int64 * 4)()
    conv = C.c_int32(); gn = C.c_double(); nit = C.c_int64(); sec = C.c_double()
    st = lib().orc_core_run(_p(x), C.c_int64(n), C.c_int64(t), C.c_int(kind), C.c_double(alpha), C.c_int(int(ortho)),
                            C.c_int(int(extended)), C.c_int64(m), C.c_int64(max_iter), C.c_double(tol), C.c_double(lambda_min),
                            C.c_int64(ls_tries), C.c_int(int(verbose)), _p(cov), _p(y), _p(w), C.byref(conv), C.byref(gn),
                            C.byref(nit), _p(signs), _p(trace), C.byref(rows), cnt, C.byref(sec))
    if st != OK:
        raise OracleError(st, "core_run failed")
    return CoreResult(y, w, bool(conv.value), gn.value, int(nit.value), signs, trace[: rows.value].copy(), cnt[0], cnt[1], cnt[2],
                      cnt[3], sec.value)
