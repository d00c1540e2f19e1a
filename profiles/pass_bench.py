"""Per-pass timing of the fused kernel on device-resident synthetic data (picard_eval_moments_device).
Usage: python profiles/pass_bench.py [N] [T] [repeats] [density kind] [pass,pass,...]   -> one JSON line per pass mode."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import _data
import picard_ica_b200 as P
from picard_ica_b200 import _ffi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
t = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 5
kind = int(sys.argv[4]) if len(sys.argv) > 4 else 0
only = sys.argv[5].split(",") if len(sys.argv) > 5 else None  # e.g. loss,gradY
lib = _ffi.lib()
ld = (t + 15) // 16 * 16
x = torch.empty((n, ld), dtype=torch.float64, device="cuda")
assert lib.picard_synth_sources(C.c_void_p(x.data_ptr()), C.c_int64(n), C.c_int64(t), C.c_int64(ld), C.c_int64(0), C.c_int64(n // 2),
                                C.c_uint64(42), C.c_int32(0), None) == 0
w = np.ascontiguousarray(_data.orthogonal(n, 43))
peak = 37.19
for mode, name, fl in [(0, "fused", 4.0), (1, "grad", 4.0), (2, "loss", 2.0), (3, "loss+store,gradY", 4.0), (4, "gradY", 2.0),
                       (0, "fused+H", 6.0), (1, "grad+H", 6.0), (4, "gradY+H", 4.0)]:
    if only and name not in only:
        continue
    want_h = name.endswith("+H")
    ms = C.c_double()
    err = C.create_string_buffer(512)
    st = lib.picard_eval_moments_device(C.c_void_p(x.data_ptr()), C.c_int64(n), C.c_int64(t), C.c_int64(ld), w.ctypes.data_as(_ffi.dp),
                                        C.c_int32(kind), C.c_double(1.0 if kind != 1 else 0.1), C.c_int32(mode), C.c_int32(int(want_h)),
                                        C.c_int32(0), C.c_int32(rep), C.byref(ms), None, None, None, None, None, err, C.c_size_t(512))
    if st != 0:
        print(json.dumps({"pass": name, "error": err.value.decode()})); continue
    tf = fl * n * n * t / (ms.value * 1e-3) / 1e12
    print(json.dumps({"pass": name, "n": n, "t": t, "kind": kind, "ms": round(ms.value, 4), "tflops": round(tf, 3), "frac_of_37.19": round(tf / peak, 4),
                      "hbm_gbs": round(8.0 * n * t / (ms.value * 1e-3) / 1e9, 1)}))
