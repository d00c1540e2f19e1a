"""Reads the phase timestamps of the -DPICARD_RB_TRACE build (profiles/rb_trace.sh) after one LOSS pass at c3 and prints,
per warp, the mean duration of the DMMA phase and of the density/store phase, and for each scheduler pair (w, w+4) the
fraction of the first warp's density phase that overlaps the second warp's density phase."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import _data
from picard_ica_b200 import _ffi

n, t = 128, 2_000_000
lib = _ffi.lib()
ld = (t + 15) // 16 * 16
x = torch.empty((n, ld), dtype=torch.float64, device="cuda")
assert lib.picard_synth_sources(C.c_void_p(x.data_ptr()), C.c_int64(n), C.c_int64(t), C.c_int64(ld), C.c_int64(0), C.c_int64(n // 2),
                                C.c_uint64(42), C.c_int32(0), None) == 0
w = np.ascontiguousarray(_data.orthogonal(n, 43))
ms = C.c_double(); err = C.create_string_buffer(512)
st = lib.picard_eval_moments_device(C.c_void_p(x.data_ptr()), C.c_int64(n), C.c_int64(t), C.c_int64(ld), w.ctypes.data_as(_ffi.dp),
                                    C.c_int32(0), C.c_double(1.0), C.c_int32(2), C.c_int32(0), C.c_int32(0), C.c_int32(1), C.byref(ms),
                                    None, None, None, None, None, err, C.c_size_t(512))
assert st == 0, err.value
TILES = 96
buf = np.zeros(16 * TILES * 8, dtype=np.int64)
lib.picard_debug_rb_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.picard_debug_rb_trace(buf.ctypes.data, buf.size) == 0
tr = buf.reshape(16, TILES, 8)[:8, :, :4]  # slots: 0 before mbar wait, 1 after, 2 DMMAs issued + stage released, 3 density/store done
t0 = tr[:, 0, 0].min()
tr = tr - t0
out = {"skew": int(os.environ.get("PICARD_RB_SKEW", "0")), "ms": ms.value, "warps": []}
sl = slice(16, TILES)  # steady state
for wi in range(8):
    wait = (tr[wi, sl, 1] - tr[wi, sl, 0]).mean()
    dm = (tr[wi, sl, 2] - tr[wi, sl, 1]).mean()
    ep = (tr[wi, sl, 3] - tr[wi, sl, 2]).mean()
    per = np.diff(tr[wi, sl, 0]).mean()
    out["warps"].append({"warp": wi, "wait": float(wait), "dmma": float(dm), "epilogue": float(ep), "period": float(per)})
for wi in range(4):
    a, b = tr[wi], tr[wi + 4]
    ov = 0.0; tot = 0.0
    for it in range(16, TILES):
        s, e = a[it, 2], a[it, 3]
        tot += e - s
        for jt in range(TILES):
            s2, e2 = b[jt, 2], b[jt, 3]
            ov += max(0, min(e, e2) - max(s, s2))
    out.setdefault("pair_epilogue_overlap", []).append(float(ov / tot))
out["raw_first_tiles"] = tr[:, 16:24, :].tolist()
json.dump(out, open(sys.argv[1], "w"))
print(json.dumps({k: v for k, v in out.items() if k != "raw_first_tiles"}))
