"""The C++ host-side mirror of the reference's public interface (picard-ica_b200/host/picard.hpp, header-only over the C
ABI): compiled with g++ and run.  CPU mode: config / validation / error mapping / utils (and the loud no-GPU error);
GPU mode: the reference's own solver tests (solver.rs:288-408) through the mirror."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "picard_hpp_check")
    libdir = os.path.join(ROOT, "picard-ica_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I/usr/local/cuda/include", os.path.join(ROOT, "tests", "host", "picard_hpp_check.cpp"),
                    "-o", exe, f"-L{libdir}", "-lpicard_b200", f"-Wl,-rpath,{libdir}"], check=True)
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_cpp_mirror_host_logic(tmp_path):
    out = subprocess.run([_build(tmp_path), "cpu"], check=True, capture_output=True, text=True).stdout
    assert out.startswith("ok ")


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_cpp_mirror_reference_solver_tests(tmp_path):
    r = subprocess.run([_build(tmp_path), "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("ok ")
