"""Full-size repeatability check of the pass engines (round 2): the LOSS + stored-Y gradient pair of the default (INT8) engines and of
the FP64 kernels, and the FP64 from-X gradient pass that does not use the Y store, each run several times on the same device-resident
N=128, T=1e7 data; prints run-to-run differences (expected: exactly 0) and the differences between the engines (expected: ~1e-14)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import picard_ica_b200 as P  # noqa: E402
from picard_ica_b200 import _ffi  # noqa: E402
import _data  # noqa: E402

n = 128
t = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ld = (t + 15) // 16 * 16
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(5)
x1 = torch.empty((n, ld), dtype=torch.float64, device=dev)
for r0 in range(0, n, 16):
    x1[r0:r0 + 16].normal_(generator=g)
x1[::2] = x1[::2].sign() * x1[::2].abs() ** 1.5 * 0.75  # heavier tails on even rows
torch.cuda.synchronize()
lib = _ffi.lib()
wp = np.ascontiguousarray(_data.orthogonal(n, 7) + 0.01 * np.random.default_rng(11).standard_normal((n, n)))


def hp(a):
    return a.ctypes.data_as(_ffi.dp)


def moments(mode, flags):
    gr = np.zeros((n, n)); sd = np.zeros(n); hr = np.zeros((n, n)); sq = np.zeros(n); lrow = np.zeros(n)
    stt = _ffi.Stats(); e2 = C.create_string_buffer(1024)
    rc = lib.picard_eval_moments_device_ex(C.c_void_p(x1.data_ptr()), C.c_int64(n), C.c_int64(t), C.c_int64(ld), hp(wp), C.c_int32(0),
                                           C.c_double(1.0), C.c_int32(mode), C.c_int32(0), C.c_int32(0), C.c_uint32(flags), C.c_int32(1),
                                           C.c_int32(0), None, hp(gr), hp(sd), hp(hr), hp(sq), hp(lrow), C.byref(stt), e2, C.c_size_t(1024))
    assert rc == 0, e2.value
    return dict(gr=gr, sd=sd, lrow=lrow), stt.as_dict()


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


runs = {}
for name, mode, flags in (("int8_pair", 3, 0), ("fp64_pair", 3, P.FLAG_NO_INT8), ("fp64_from_x", 1, P.FLAG_NO_INT8)):
    got = [moments(mode, flags) for _ in range(reps)]
    runs[name] = got[0][0]
    runs[name + "_all"] = [r[0] for r in got]
    for k in ("gr", "sd") + (("lrow",) if mode == 3 else ()):
        d = [rel(r[0][k], got[0][0][k]) for r in got[1:]]
        print(json.dumps({"engine": name, "quantity": k, "run_to_run_rel": d, "i8_loss": got[0][1]["i8_loss_passes"], "i8_grad": got[0][1]["i8_grad_passes"]}), flush=True)
for a, b in (("int8_pair", "fp64_from_x"), ("fp64_pair", "fp64_from_x"), ("int8_pair", "fp64_pair")):
    print(json.dumps({"compare": [a, b], "gr": rel(runs[a]["gr"], runs[b]["gr"]), "sd": rel(runs[a]["sd"], runs[b]["sd"])}), flush=True)
for i, r in enumerate(runs["int8_pair_all"]):
    print(json.dumps({"int8_pair_run": i, "vs_fp64_pair": {k: rel(r[k], runs["fp64_pair"][k]) for k in ("gr", "sd", "lrow")}}), flush=True)
