"""CPU-only: the C-ABI library builds, loads and exports every symbol include/zoe_cuda.h declares."""
import ctypes
import os
import re

from zoe_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "zoe_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zoe_cuda_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_listed_in_binding():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from zoe_b200.build import build
    path = build()
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.zoe_cuda_create(ctypes.byref(h), None, 1)
    assert rc == _lib.E_CUDA and not h.value  # no CPU fallback


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "zoe_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(root, f), errors="replace").read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"


def test_rust_sys_crate_binds_every_declared_symbol():
    # the -sys crate cannot be compiled here (no rustc): at least keep its extern block in step with the header
    text = open(os.path.join(ROOT, "rust", "zoe-cuda-sys", "src", "lib.rs")).read()
    bound = sorted(set(re.findall(r"pub fn (zoe_cuda_[a-z0-9_]+)", text)))
    assert bound == declared_symbols()
