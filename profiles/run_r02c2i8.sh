#!/bin/bash
# round 2, GPU call c2i8 (1 GPU): the INT8 LOSS engine forced at N = 64 (BASELINE configs[1] shape) against the default FP64 kernels
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r02c2i8_parity.json
import sys, json
sys.path.insert(0, "tests")
import numpy as np, _data, _gpu
import picard_ica_b200 as P
from oracle import oracle as orc
n, t = 64, 100_000
x, a, _ = _data.mixture(n, t, seed=5, kind="mixed")
xw = np.linalg.cholesky(np.linalg.inv(np.cov(x))).T @ (x - x.mean(1, keepdims=True))
w = _data.orthogonal(n, 7) + 0.01 * np.random.default_rng(11).standard_normal((n, n))
ref = orc.eval_point(xw, w, 0, 1.0, ortho=False, extended=False)
got, st = _gpu.eval_moments_ex(xw, w, 0, 1.0, mode=3, want_h=False, flags=P.FLAG_FORCE_INT8, whitened=True)
print(json.dumps({"n": n, "t": t, "i8_loss_passes": st["i8_loss_passes"], "i8_grad_passes": st["i8_grad_passes"],
                  **{k: _data.rel_err(got[k], getattr(ref, k)) for k in ("gr", "sd", "lrow")}}))
PY
cat gpurun_out/r02c2i8_parity.json
for f in 0 32; do
  timeout -s KILL 600 python bench.py --workload c2 --no-cpu --no-e2e --flags $f > gpurun_out/r02c2i8_flags$f.json 2> gpurun_out/r02c2i8_flags$f.err; echo "flags $f exit $?"
  python -c "
import json; d=json.loads(open('gpurun_out/r02c2i8_flags$f.json').read().strip().splitlines()[-1])
print('flags $f', round(d['value'],1), round(d['ms_per_step'],3), {k:round(v['avg_ms'],3) for k,v in d['roofline']['all_passes'].items()}, d['passes']['i8_loss'], d['passes']['i8_grad'], (d.get('parity') or {}).get('ok'))"
done
exit 0
