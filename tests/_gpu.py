"""ctypes helpers for the test hooks of the C ABI (include/picard_b200.h): every GPU parity test calls the
product through these entry points -- never through Python-side arithmetic."""
from __future__ import annotations

import ctypes as C

import numpy as np

import picard_ica_b200 as P
from picard_ica_b200 import _ffi

dp = _ffi.dp


def _p(a):
    return None if a is None else a.ctypes.data_as(dp)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _check(st, err=None):
    if st != 0:
        raise RuntimeError(f"status {st}: {err.value.decode() if err is not None else ''}")


def eval_moments(x, w=None, kind=0, alpha=1.0, mode=0, want_h=True, device=0):
    """picard_eval_moments -> dict(gr, sd, hr, sq, lrow)."""
    x = _c(x); w = _c(w); n, t = x.shape
    gr = np.full((n, n), np.nan); hr = np.full((n, n), np.nan)
    sd = np.full(n, np.nan); sq = np.full(n, np.nan); lrow = np.full(n, np.nan)
    err = C.create_string_buffer(1024)
    st = _ffi.lib().picard_eval_moments(_p(x), C.c_int64(n), C.c_int64(t), C.c_int64(x.strides[0] // 8), _p(w), C.c_int32(kind),
                                        C.c_double(alpha), C.c_int32(mode), C.c_int32(int(want_h)), C.c_int32(device),
                                        _p(gr), _p(sd), _p(hr), _p(sq), _p(lrow), err, C.c_size_t(1024))
    _check(st, err)
    return dict(gr=gr, sd=sd, hr=hr, sq=sq, lrow=lrow)


def eval_moments_ex(x, w=None, kind=0, alpha=1.0, mode=0, want_h=True, flags=0, whitened=False, device=0):
    """picard_eval_moments_ex -> (dict(gr, sd, hr, sq, lrow), stats dict): explicit PICARD_FLAG_* bits and the `whitened` promise."""
    x = _c(x); w = _c(w); n, t = x.shape
    gr = np.full((n, n), np.nan); hr = np.full((n, n), np.nan)
    sd = np.full(n, np.nan); sq = np.full(n, np.nan); lrow = np.full(n, np.nan)
    err = C.create_string_buffer(1024)
    stats = _ffi.Stats()
    st = _ffi.lib().picard_eval_moments_ex(_p(x), C.c_int64(n), C.c_int64(t), C.c_int64(x.strides[0] // 8), _p(w), C.c_int32(kind),
                                           C.c_double(alpha), C.c_int32(mode), C.c_int32(int(want_h)), C.c_int32(device),
                                           C.c_uint32(flags), C.c_int32(int(whitened)), _p(gr), _p(sd), _p(hr), _p(sq), _p(lrow),
                                           C.byref(stats), err, C.c_size_t(1024))
    _check(st, err)
    return dict(gr=gr, sd=sd, hr=hr, sq=sq, lrow=lrow), stats.as_dict()


def eval_point(x, w=None, kind=0, alpha=1.0, ortho=True, extended=True, lambda_min=0.01, c=None, old_signs=None, loss_signs=None,
               device=0):
    """picard_eval_point -> dict(g, h, hoff, signs, sign_change, gradient_norm, loss)."""
    x = _c(x); w = _c(w); c = _c(c); old_signs = _c(old_signs); loss_signs = _c(loss_signs)
    n, t = x.shape
    g = np.empty((n, n)); h = np.empty((n, n)); hoff = np.empty(n); signs = np.empty(n)
    sc = C.c_int32(); gn = C.c_double(); loss = C.c_double()
    err = C.create_string_buffer(1024)
    st = _ffi.lib().picard_eval_point(_p(x), C.c_int64(n), C.c_int64(t), C.c_int64(x.strides[0] // 8), _p(w), C.c_int32(kind),
                                      C.c_double(alpha), C.c_int32(int(ortho)), C.c_int32(int(extended)), C.c_double(lambda_min),
                                      _p(c), _p(old_signs), _p(loss_signs), C.c_int32(device), _p(g), _p(h), _p(hoff), _p(signs),
                                      C.byref(sc), C.byref(gn), C.byref(loss), err, C.c_size_t(1024))
    _check(st, err)
    return dict(g=g, h=h, hoff=hoff, signs=signs, sign_change=bool(sc.value), gradient_norm=gn.value, loss=loss.value)


def matrix_exp(a, device=0):
    a = _c(a); out = np.empty_like(a)
    _check(_ffi.lib().picard_matrix_exp(_p(a), C.c_int64(a.shape[0]), _p(out), C.c_int32(device)))
    return out


def sln_det(m, device=0):
    m = _c(m); s = C.c_double(); l = C.c_double()
    _check(_ffi.lib().picard_sln_det(_p(m), C.c_int64(m.shape[0]), C.byref(s), C.byref(l), C.c_int32(device)))
    return s.value, l.value


def sym_decorrelation(w, device=0):
    w = _c(w); out = np.empty_like(w)
    st = _ffi.lib().picard_sym_decorrelation(_p(w), C.c_int64(w.shape[0]), _p(out), C.c_int32(device))
    return st, out


def compute_direction(g, h, hoff, s_list, y_list, r_list, ortho, device=0):
    g = _c(g); h = _c(h); hoff = _c(hoff); n = g.shape[0]; L = len(r_list)
    s = _c(np.asarray(s_list).reshape(L, n, n)) if L else None
    y = _c(np.asarray(y_list).reshape(L, n, n)) if L else None
    r = _c(np.asarray(r_list)) if L else None
    out = np.empty_like(g)
    _check(_ffi.lib().picard_compute_direction(_p(g), _p(h), _p(hoff), C.c_int64(n), _p(s), _p(y), _p(r), C.c_int64(L),
                                               C.c_int32(int(ortho)), _p(out), C.c_int32(device)))
    return out


def center_whiten(x, n_components, centering=True, want_data=True, device=0):
    x = _c(x); nf, t = x.shape
    mean = np.zeros(nf); k = np.empty((n_components, nf)); data = np.empty((n_components, t)) if want_data else None
    err = C.create_string_buffer(1024)
    st = _ffi.lib().picard_center_whiten(_p(x), C.c_int64(nf), C.c_int64(t), C.c_int64(x.strides[0] // 8), C.c_int64(n_components),
                                         C.c_int32(int(centering)), C.c_int32(device), _p(mean), _p(k), _p(data), err, C.c_size_t(1024))
    return st, err.value.decode(), mean, k, data


def jade(x, max_iter, tol=1e-6, device=0):
    x = _c(x); n, t = x.shape
    w = np.empty((n, n)); sw = C.c_int64()
    err = C.create_string_buffer(1024)
    st = _ffi.lib().picard_jade(_p(x), C.c_int64(n), C.c_int64(t), C.c_int64(x.strides[0] // 8), C.c_int64(max_iter), C.c_double(tol),
                                C.c_int32(0), C.c_int32(device), _p(w), C.byref(sw), err, C.c_size_t(1024))
    return st, err.value.decode(), w, sw.value


def jade_cumulants(x, device=0):
    x = _c(x); n, t = x.shape
    out = np.empty((n * (n + 1) // 2, n, n))
    err = C.create_string_buffer(1024)
    st = _ffi.lib().picard_jade_cumulants(_p(x), C.c_int64(n), C.c_int64(t), C.c_int64(x.strides[0] // 8), C.c_int32(device), _p(out), err,
                                          C.c_size_t(1024))
    _check(st, err)
    return out
