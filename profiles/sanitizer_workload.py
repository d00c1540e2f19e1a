import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import _data, _gpu
import picard_ica_b200 as P
from oracle import oracle as orc
# small shapes through every kernel family: from-X pass (all modes), row-block loss/grady (64, 128, 256 paths), N x N kernels, fit, jade, fastica
for n, t in [(5, 333), (64, 200), (128, 150), (136, 120)]:
    x = _data.whitened(n, t, seed=n) if t > 2 * n else np.random.default_rng(n).standard_normal((n, t))
    w = _data.orthogonal(n, n + 1)
    for mode in ([0, 1, 2, 3] if n <= 128 else [2, 3]):
        for wh in (True, False):
            _gpu.eval_moments(x, w, 0, 1.0, mode=mode, want_h=wh)
    _gpu.eval_point(x, w, 1, 0.1, ortho=False, extended=True, c=w @ w.T)
    _gpu.eval_point(x, w, 0, 1.0, ortho=True, extended=True)
a = np.random.default_rng(0).standard_normal((20, 20)); a = (a - a.T) / 2
_gpu.matrix_exp(a * 3.0); _gpu.matrix_exp(a * 0.01); _gpu.sln_det(a + np.eye(20)); _gpu.sym_decorrelation(a + 3 * np.eye(20))
_gpu.sym_decorrelation(np.random.default_rng(1).standard_normal((40, 40)))
x, am, _ = _data.mixture(6, 2000, seed=1)
for kw in [dict(), dict(ortho=False, extended=False), dict(jade_it=3), dict(fastica_it=2), dict(whiten=False), dict(flags=P.FLAG_NO_Y_STORE)]:
    r = P.Picard.fit_with_config(x, P.PicardConfig(random_state=1, max_iter=15, **kw))
    P.Picard.transform(x[:, :100], r)
print("sanitizer workload done")
