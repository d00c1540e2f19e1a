"""The table-driven exp / log / tanh / log-likelihood routines of picard-ica_b200/csrc/density.cuh are plain
host-callable C++: compile them with g++ and check them against long-double libm on millions of points
(density.rs:50-63, 91-103 give the formulas).  This is a CPU test of PRODUCT code, no oracle involved."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_density_polynomials_on_host(tmp_path):
    exe = str(tmp_path / "dhc")
    cuda_inc = "/usr/local/cuda/include"
    subprocess.run(["g++", "-O2", "-std=c++17", f"-I{cuda_inc}", os.path.join(ROOT, "tests", "host", "density_host_check.cpp"), "-o", exe],
                   check=True)
    lines = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    outs = [json.loads(l) for l in lines]
    assert sorted(o["set"] for o in outs) == ["big", "small"]  # both table sets (4 KB and 80 KB)
    for out in outs:
        assert out["exp_err_units"] < 1.0    # exp(s z), s z in [-700, 0]: relative error in units of (3 + |s z|) 2^-53
        assert out["log_abs"] < 1e-15        # log on [1, 2], absolute
        assert out["tanh_psi_abs"] < 1e-15   # tanh(alpha y), absolute
        assert out["tanh_psid_abs"] < 3e-15  # alpha (1 - tanh^2), absolute (alpha up to 2.5)
        assert out["tanh_ll_rel"] < 1e-15    # |y| + log(1 + exp(-2 alpha |y|)) / alpha, relative
        assert out["expdens_abs"] < 1e-15    # exp density: psi, psi', log-lik, absolute up to their polynomial prefactors


def test_tables_are_reproducible():
    """density_tables.inc is generated (tools/gen_density_tables.py); regenerate and compare when mpmath is present."""
    mp = pytest.importorskip("mpmath")
    path = os.path.join(ROOT, "picard-ica_b200", "csrc", "density_tables.inc")
    before = open(path).read()
    subprocess.run(["python", os.path.join(ROOT, "tools", "gen_density_tables.py")], check=True, capture_output=True)
    assert open(path).read() == before
