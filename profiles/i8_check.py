"""Parity of the INT8 LOSS pass (PICARD_I8=1) against the oracle, then its time at c3.  Usage: PICARD_I8=1 python profiles/i8_check.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _data, _gpu
from oracle import oracle as orc
out = {"i8": os.environ.get("PICARD_I8", "0")}
for n, t, kind, alpha in [(128, 2050, orc.TANH, 1.0), (100, 1500, orc.TANH, 1.0), (128, 31, orc.EXP, 0.1), (70, 4099, orc.CUBE, 1.0)]:
    x = _data.whitened(n, t, seed=n)
    w = _data.orthogonal(n, seed=n + 2) + 0.02 * np.random.default_rng(n).standard_normal((n, n))
    ref = orc.eval_point(x, w, kind, alpha, ortho=False, extended=False)
    got = _gpu.eval_moments(x, w, kind, alpha, mode=2, want_h=True)
    e = {k: float(_data.rel_err(got[k], getattr(ref, k))) for k in ("lrow", "sq")}
    got3 = _gpu.eval_moments(x, w, kind, alpha, mode=3, want_h=False)   # gradient from the stored Y'
    e.update({"gr_from_stored_y": float(_data.rel_err(got3["gr"], ref.gr)), "sd": float(_data.rel_err(got3["sd"], ref.sd))})
    out[f"n{n}_t{t}_k{kind}"] = e
print(json.dumps(out))
