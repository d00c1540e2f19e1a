"""Times the preprocessing entry points on device-resident synthetic data. Usage: python profiles/whiten_bench.py [N] [T]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import picard_ica_b200 as P
from picard_ica_b200 import _ffi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
t = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
lib = _ffi.lib(); ld = (t + 15) // 16 * 16
x = torch.empty((n, ld), dtype=torch.float64, device="cuda")
lib.picard_synth_sources(C.c_void_p(x.data_ptr()), C.c_int64(n), C.c_int64(t), C.c_int64(ld), C.c_int64(0), C.c_int64(n // 2), C.c_uint64(1), C.c_int32(0), None)
mean = np.zeros(n); k = np.zeros((n, n)); err = C.create_string_buffer(512)
for r in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = lib.picard_center_whiten_device(C.c_void_p(x.data_ptr()), C.c_int64(n), C.c_int64(t), C.c_int64(ld), C.c_int64(n), C.c_int32(1), None, C.c_int32(0),
                                         mean.ctypes.data_as(_ffi.dp), k.ctypes.data_as(_ffi.dp), err, C.c_size_t(512))
    torch.cuda.synchronize(); print("center_whiten_device", st, err.value, round(1e3 * (time.perf_counter() - t0), 2), "ms")
y = torch.empty((n, ld), dtype=torch.float64, device="cuda")
for r in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lib.picard_apply_device(k.ctypes.data_as(_ffi.dp), mean.ctypes.data_as(_ffi.dp), C.c_int64(n), C.c_int64(n), C.c_void_p(x.data_ptr()), C.c_int64(ld),
                            C.c_void_p(y.data_ptr()), C.c_int64(ld), C.c_int64(t), C.c_int32(0), None)
    torch.cuda.synchronize(); print("apply_device", round(1e3 * (time.perf_counter() - t0), 2), "ms")
c = (y[:, :t] @ y[:, :t].T / t).cpu().numpy()
print("whiteness error", np.abs(c - np.eye(n)).max())
