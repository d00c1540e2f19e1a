"""The transform kernels divide every Taylor term by its index with a reciprocal and two FMAs instead of the division sequence
(picard-ica_b200/csrc/exact_div.h, host-callable).  Compile the header with gcc and compare it with the hardware division, bit for
bit, on random operands over 600 binades, exact quotients and ties, signed zeros, denormals, infinities and NaN.  A CPU test of
PRODUCT code, no oracle involved."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_division_by_the_term_index_is_exact(tmp_path):
    exe = str(tmp_path / "dbc")
    subprocess.run(["gcc", "-O2", "-march=native", os.path.join(ROOT, "tests", "host", "div_by_count_check.c"), "-o", exe, "-lm"], check=True)
    out = json.loads(subprocess.run([exe, "400000"], check=True, capture_output=True, text=True).stdout)
    assert out["cases"] > 1.5e7
    assert out["mismatches"] == 0
