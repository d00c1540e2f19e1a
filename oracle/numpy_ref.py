"""Independent numpy restatement of the reference's core loop.  TEST INFRASTRUCTURE ONLY.

Second, separately written restatement of /root/reference/src/{core,density,lbfgs,math}.rs used to
cross-check the C++ oracle (oracle/picard_oracle.cpp) on small problems: two restatements written
independently from the Rust source agreeing to rounding is the strongest pin available while the
reference itself cannot be built here (no cargo/rustc).  Pure numpy/scipy, small sizes only.
"""
from __future__ import annotations

import math

import numpy as np

TANH, EXP, CUBE = 0, 1, 2


def signum(v):
    """f64::signum: +0 -> +1, -0 -> -1 (quirk Q8)."""
    return np.where(np.signbit(v), -1.0, 1.0)


def log_lik(kind, alpha, y):
    if kind == TANH:  # density.rs:50-56
        a = np.abs(y)
        return a + np.log(1.0 + np.exp(-2.0 * alpha * a)) / alpha
    if kind == EXP:  # density.rs:91-94
        return -np.exp(-alpha * y * y / 2.0) / alpha
    return (y * y) * (y * y) / 4.0  # density.rs:122-124


def score_and_der(kind, alpha, y):
    if kind == TANH:  # density.rs:58-63
        s = np.tanh(alpha * y)
        return s, alpha * (1.0 - s * s)
    if kind == EXP:  # density.rs:96-103
        ysq = y * y
        k = np.exp(-alpha / 2.0 * ysq)
        return y * k, (1.0 - alpha * ysq) * k
    return y ** 3, 3.0 * y * y  # density.rs:126-130


def matrix_exp(a):  # math.rs:38-74
    n = a.shape[0]
    norm = np.max(np.abs(a)) if a.size else 0.0
    if norm < 1e-15:
        return np.eye(n)
    s = int(max(math.ceil(math.log2(norm)), 0.0))
    a_s = a / (2.0 ** s)
    result = np.eye(n)
    term = np.eye(n)
    for k in range(1, 31):
        term = term @ a_s / float(k)
        result = result + term
        if np.max(np.abs(term)) < 1e-16:
            break
    for _ in range(s):
        result = result @ result
    return result


def sln_det(w):  # math.rs:84-88 (LAPACK LU); singular -> sign 0
    sign, logabs = np.linalg.slogdet(w)
    return float(sign), float(logabs)


def compute_loss(y, w, kind, alpha, signs, ortho, extended):  # core.rs:39-85
    n, t = y.shape
    loss = 0.0
    if not ortho:
        sg, la = sln_det(w)
        if sg == 0.0:
            return None
        loss = -la
    for i in range(n):
        loss += signs[i] * np.sum(log_lik(kind, alpha, y[i])) / t
        if extended and not ortho:
            loss += 0.5 * np.sum(y[i] * y[i]) / t
    return loss


def regularize_hessian(h, hoff, lam):  # lbfgs.rs:155-171, sequential in place (Q4)
    n = h.shape[0]
    for i in range(n):
        for j in range(n):
            if i != j:
                diff = h[i, j] - h[j, i]
                discr = math.sqrt(diff * diff + 4.0 * hoff[i] * hoff[j])
                ev = 0.5 * (h[i, j] + h[j, i] - discr)
                if ev < lam:
                    h[i, j] += lam - ev


def solve_hessian_system(h, hoff, g):  # lbfgs.rs:136-150
    det = h * h.T - np.outer(hoff, hoff)
    num = h.T * g - hoff[:, None] * g.T
    out = np.zeros_like(g)
    ok = np.abs(det) > 1e-15
    out[ok] = num[ok] / det[ok]
    return out


def compute_direction(g, h, hoff, mem, ortho):  # lbfgs.rs:84-133 ; mem = list of (s, y, r), oldest first
    q = g.copy()
    alphas = []
    for s, y, r in reversed(mem):
        a = r * np.sum(s * q)
        alphas.append(a)
        q = q - a * y
    alphas.reverse()
    if ortho:
        z = q / h
        z = (z - z.T) / 2.0
    else:
        z = solve_hessian_system(h, hoff, q)
    for (s, y, r), a in zip(mem, alphas):
        b = r * np.sum(y * z)
        z = z + (a - b) * s
    return -z


def front(y, kind, alpha, ortho, extended, lam, c, old_signs, first_iter):
    """core.rs:215-293.  Returns g (projected), h, hoff, signs, sign_change, gradient_norm."""
    n, t = y.shape
    psi, psid = score_and_der(kind, alpha, y)
    g = psi @ y.T / t
    ysq = y * y
    signs = np.ones(n)
    sign_change = False
    if extended:
        pm = psid.mean(axis=1)
        k = pm * np.diag(c) - np.diag(g)
        signs = signum(k)
        if not first_iter:
            sign_change = bool(np.any(signs != old_signs))
        g = g * signs[:, None]
        psid = psid * signs[:, None]
        if not ortho:
            g = g + c
            psid = psid + 1.0
    hoff = np.diag(g).copy() if ortho else np.ones(n)
    if ortho:
        pm = psid.mean(axis=1)
        h = 0.5 * (pm[:, None] + pm[None, :] - hoff[:, None] - hoff[None, :])
        h = np.maximum(h, lam)
    else:
        h = psid @ ysq.T / t
        regularize_hessian(h, hoff, lam)
    g = (g - g.T) / 2.0 if ortho else g - np.eye(n)
    gn = float(np.max(np.abs(g))) if g.size else 0.0
    return g, h, hoff, signs, sign_change, gn


def line_search(y, w, kind, alpha, d, signs, cur, tries, ortho, extended):  # core.rs:99-150
    n = w.shape[0]
    a = 1.0
    y_new, w_new, loss = y, w, cur
    n_tries = 0
    for _ in range(tries):
        n_tries += 1
        m = matrix_exp(d * a) if ortho else np.eye(n) + a * d
        y_new = m @ y
        w_new = m @ w
        lv = compute_loss(y_new, w_new, kind, alpha, signs, ortho, extended)
        loss = 1e15 if lv is None else lv
        if loss < cur:
            return True, y_new, w_new, loss, d * a, a, n_tries
        a /= 2.0
    return False, y_new, w_new, loss, d * a, a, n_tries


def core_run(x, kind=TANH, alpha=1.0, ortho=True, extended=True, m=7, max_iter=500, tol=1e-7, lambda_min=0.01, ls_tries=10,
             covariance=None):  # core.rs:162-401
    n, t = x.shape
    w = np.eye(n)
    y = x.copy()
    mem = []
    signs = np.ones(n)
    old_signs = np.ones(n)
    cur = compute_loss(y, w, kind, alpha, signs, ortho, extended)
    if cur is None:
        raise ZeroDivisionError("singular")
    gn = 1.0
    converged = False
    if extended:
        c = covariance.copy() if covariance is not None else y @ y.T / t
    else:
        c = np.eye(n)
    g_old = None
    prev_step = None
    n_iter = 0
    trace = []
    for it in range(max_iter):
        n_iter = it
        g, h, hoff, sg, sign_change, gn = front(y, kind, alpha, ortho, extended, lambda_min, c, old_signs, it == 0)
        if extended:
            signs = sg
            old_signs = sg.copy()
        if gn < tol:
            converged = True
            break
        if it > 0 and prev_step is not None and g_old is not None:
            yd = g - g_old
            with np.errstate(divide="ignore", invalid="ignore"):
                r = np.float64(1.0) / np.sum(prev_step * yd)
            if np.isfinite(r):
                mem.append((prev_step, yd, float(r)))
                if len(mem) > m:
                    mem.pop(0)
            prev_step = None
        g_old = g.copy()
        if extended and sign_change:
            lv = compute_loss(y, w, kind, alpha, signs, ortho, extended)
            cur = 1e15 if lv is None else lv
            mem = []
        d = compute_direction(g, h, hoff, mem, ortho)
        ok, y_new, w_new, loss, step, a_used, tries = line_search(y, w, kind, alpha, d, signs, cur, ls_tries, ortho, extended)
        fb = 0
        if not ok:
            fb = 1
            mem = []
            ok2, y_new, w_new, loss, step, a_used, t2 = line_search(y, w, kind, alpha, -g, signs, cur, 10, ortho, extended)
            tries += t2
        prev_step = step
        y, w = y_new, w_new
        if extended and covariance is not None:
            c = w @ covariance @ w.T
        cur = loss
        trace.append((gn, cur, a_used, tries, fb, int(sign_change), len(mem)))
    return dict(y=y, w=w, converged=converged, gradient_norm=gn, n_iterations=n_iter + 1, signs=signs if extended else None,
                trace=np.array(trace).reshape(-1, 7))


def amari(w, a):  # utils.rs:82-103
    p = np.abs(w @ a)
    n = p.shape[0]

    def s(r):
        r2 = r * r
        mx = r2.max(axis=1)
        ok = mx > 1e-15
        return float(np.sum(r2.sum(axis=1)[ok] / mx[ok] - 1.0))

    return (s(p) + s(p.T)) / (2.0 * n)
