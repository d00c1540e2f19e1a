import ctypes as C, os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import picard_ica_b200 as P
from picard_ica_b200 import _ffi
n, t = 128, 2_000_000
lib = _ffi.lib(); ld = (t + 15) // 16 * 16
x = torch.empty((n, ld), dtype=torch.float64, device="cuda")
lib.picard_synth_sources(C.c_void_p(x.data_ptr()), C.c_int64(n), C.c_int64(t), C.c_int64(ld), C.c_int64(0), C.c_int64(n // 2), C.c_uint64(1), C.c_int32(0), None)
mean = np.zeros(n); k = np.zeros((n, n)); err = C.create_string_buffer(512)
big = torch.empty(1 << 28, dtype=torch.float64, pin_memory=True)
dbig = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
def run(tag):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lib.picard_center_whiten_device(C.c_void_p(x.data_ptr()), C.c_int64(n), C.c_int64(t), C.c_int64(ld), C.c_int64(n), C.c_int32(1), None, C.c_int32(0),
                                    mean.ctypes.data_as(_ffi.dp), k.ctypes.data_as(_ffi.dp), err, C.c_size_t(512))
    torch.cuda.synchronize(); print(tag, round(1e3 * (time.perf_counter() - t0), 2), "ms", file=sys.stderr)
for i in range(4): run("back-to-back")
for i in range(3):
    time.sleep(0.5); run("after 0.5 s idle")
for i in range(3):
    dbig.copy_(big, non_blocking=True); torch.cuda.synchronize(); run("after 2 GB H2D")
for i in range(3):
    time.sleep(2.0); run("after 2 s idle")
