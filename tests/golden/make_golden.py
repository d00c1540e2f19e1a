"""Generates tests/golden/golden_points.json in the build container.

The Rust reference cannot be executed here (no cargo/rustc), so these fixtures are NOT outputs of the reference
itself: each point is computed by the C++ oracle (oracle/picard_oracle.cpp, OpenBLAS) and independently by the
numpy restatement (oracle/numpy_ref.py); the script refuses to write a value unless the two agree to 1e-11.
They pin the oracle against drift and give the GPU tests fixed vectors that travel to the GPU box."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _data  # noqa: E402
from oracle import numpy_ref as npr  # noqa: E402
from oracle import oracle as orc  # noqa: E402

points = []
for (n, t, seed, kind, alpha, ortho, extended) in [(3, 2000, 1, 0, 1.0, True, True), (5, 1501, 2, 0, 1.0, False, False),
                                                   (4, 999, 3, 1, 0.1, False, False), (6, 1200, 4, 2, 1.0, True, False),
                                                   (8, 3000, 5, 0, 0.5, False, True), (16, 2500, 6, 0, 1.0, True, True)]:
    x = _data.whitened(n, t, seed=seed)
    w = _data.orthogonal(n, seed + 1)
    c = w @ w.T
    ep = orc.eval_point(x, w, kind, alpha, ortho, extended, 0.01, c=c)
    g, h, hoff, sg, sc, gn = npr.front(w @ x, kind, alpha, ortho, extended, 0.01, c, np.ones(n), True)
    loss = npr.compute_loss(w @ x, w, kind, alpha, sg, ortho, extended)
    assert _data.rel_err(ep.g, g) < 1e-11 and _data.rel_err(ep.h, h) < 1e-11 and abs(ep.loss - loss) < 1e-11 * max(1, abs(loss)), (n, t)
    points.append(dict(n=n, t=t, seed=seed, kind=kind, alpha=alpha, ortho=ortho, extended=extended, loss=ep.loss,
                       gradient_norm=ep.gradient_norm, g=ep.g.tolist(), h=ep.h.tolist(), signs=ep.signs.tolist()))

fits = []
for (data, n, t, seed, kind, alpha, ortho, extended) in [("lcg", 3, 10000, 42, 0, 1.0, False, False), ("mixed", 6, 8000, 7, 0, 1.0, True, True),
                                                         ("laplace", 5, 6000, 8, 1, 0.1, False, False)]:
    x = _data.lcg_bench_data(n, t, 42) if data == "lcg" else _data.mixture(n, t, seed=seed, kind=data)[0]
    r = orc.fit(x, orc.Config(density=kind, alpha=alpha, ortho=ortho, extended=extended, w_init=_data.orthogonal(n, 43)))
    # independent restatement of the same run (numpy core loop on the oracle's preprocessed data)
    xc = x - x.mean(axis=1, keepdims=True)
    x1 = _data.orthogonal(n, 43) @ (r.whitening @ xc)
    b = npr.core_run(x1, kind, alpha, ortho, extended, covariance=np.eye(n) if extended else None)
    assert b["n_iterations"] == r.n_iterations, (data, b["n_iterations"], r.n_iterations)
    assert orc.amari(b["w"] @ _data.orthogonal(n, 43) @ r.whitening, np.linalg.inv(r.full_unmixing())) < 1e-7
    fits.append(dict(data=data, n=n, t=t, seed=seed, kind=kind, alpha=alpha, ortho=ortho, extended=extended,
                     n_iterations=r.n_iterations, converged=r.converged, gradient_norm=r.gradient_norm,
                     full_unmixing=r.full_unmixing().tolist()))

with open(os.path.join(HERE, "golden_points.json"), "w") as f:
    json.dump(dict(generator="tests/golden/make_golden.py", note="oracle + numpy_ref agreeing; NOT outputs of the Rust reference",
                   points=points, fits=fits), f)
print("wrote", len(points), "points and", len(fits), "fits")
