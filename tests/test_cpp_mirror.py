"""The C++ mirror of zoe's interface (include/zoe_cuda.hpp): compiles and links on CPU; runs on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_hpp")


def _build():
    from zoe_b200.build import build
    lib = build()
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_hpp.cpp"), "-o", EXE, lib, f"-Wl,-rpath,{os.path.dirname(lib)}"]
    subprocess.check_call(cmd)


def test_cpp_mirror_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_known_answers():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "cpp mirror ok" in out.stdout
