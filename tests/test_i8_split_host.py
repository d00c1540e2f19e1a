"""The error-free digit splitting behind the INT8 tensor-core passes (picard-ica_b200/csrc/i8_split.h) is plain host-callable
C++: compile it with g++ and run the kernels' integer arithmetic on the CPU (tests/host/i8_split_check.cpp) -- balanced radix-256
digits, the 21 slice products accumulated per level in int32, the exact level combination -- against long-double references.
This is a CPU test of PRODUCT code, no oracle involved.  Parity bar of the passes: 1e-10 (BASELINE.json)."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_digit_splitting_on_host(tmp_path):
    exe = str(tmp_path / "i8chk")
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "host", "i8_split_check.cpp"), "-o", exe], check=True)
    out = json.loads(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)
    assert out["overflow"] == 0                  # every level sum fits the s32 accumulator (K = 128; 16384 samples between flushes)
    assert out["round_trip_quanta"] <= 0.5       # digits reproduce rint(v 2^(47 - e)) exactly: half a quantum at most
    assert out["loss_err"] < 2e-12               # y = sum_k w_k x_k, K = 128, relative to max|w| max|x|
    assert out["loss_err_adversarial"] < 2e-12   # W' row spread over 1e8, one component of the sample 1e6 x larger
    assert out["grad_err"] < 1e-13               # sum_t psi(y_it) y_jt over 2e5 samples with fixed exponents, relative to max|G|
    assert out["grad_diag_err"] < 1e-13
