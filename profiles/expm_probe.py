import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import _gpu
for n, scale in [(128, 0.05), (128, 0.5), (64, 0.05), (64, 0.5), (256, 0.05)]:
    a = np.random.default_rng(n).standard_normal((n, n)) * scale / np.sqrt(n); a = (a - a.T) / 2
    _gpu.matrix_exp(a)
    t0 = time.perf_counter()
    for _ in range(20): _gpu.matrix_exp(a)
    print(n, scale, "max|a|", round(float(np.abs(a).max()), 4), "wall per call (incl. H2D/D2H, alloc)", round((time.perf_counter() - t0) / 20 * 1e3, 3), "ms")
