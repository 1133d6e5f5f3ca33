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


# ---- signature-level agreement: include/zoe_cuda.h  vs  the Rust extern block  vs  the ctypes argtypes ----------
# Every parameter is reduced to a canonical (pointer depth, const, base type) form; an argument-order or type drift in
# either binding then fails here, not at a maintainer's first call (no rustc in this image to compile the -sys crate).
_C_BASE = {"int": "i32", "int8_t": "i8", "uint8_t": "u8", "uint32_t": "u32", "uint64_t": "u64", "float": "f32",
           "double": "f64", "char": "char", "void": "void", "zoe_cuda_ctx": "ctx", "zoe_cuda_stats": "stats"}
_RUST_BASE = {"c_int": "i32", "i8": "i8", "u8": "u8", "u32": "u32", "u64": "u64", "f32": "f32", "f64": "f64",
              "c_char": "char", "c_void": "void", "zoe_cuda_ctx": "ctx", "zoe_cuda_stats": "stats"}


def _canon_c(param: str):
    param = re.sub(r"\[\d*\]", "*", param.strip())            # `const uint8_t byte_to_index[256]` decays to a pointer
    depth = param.count("*")
    toks = [t for t in re.sub(r"\*", " ", param).split()]
    const = "const" in toks
    toks = [t for t in toks if t != "const"]
    base = toks[0] if toks[0] in _C_BASE else None
    assert base, param
    return (depth, const and depth > 0, _C_BASE[base])


def _canon_rust(ty: str):
    ty = ty.strip()
    depth, const = 0, False
    while ty.startswith("*"):
        m = re.match(r"\*(const|mut)\s+", ty)
        if depth == 0:
            pass
        const = const or (m.group(1) == "const" and not ty[m.end():].startswith("*"))
        depth += 1
        ty = ty[m.end():]
    return (depth, const, _RUST_BASE[ty])


def header_signatures():
    text = open(os.path.join(ROOT, "include", "zoe_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for ret, name, params in re.findall(r"([A-Za-z_][\w \*]*?)\s*\b(zoe_cuda_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text):
        plist = [p for p in (x.strip() for x in params.replace("\n", " ").split(",")) if p and p != "void"]
        out[name] = ([_canon_c(p) for p in plist], _canon_c(ret + " x")[:1] + _canon_c(ret + " x")[1:])
    return out


def rust_signatures():
    text = open(os.path.join(ROOT, "rust", "zoe-cuda-sys", "src", "lib.rs")).read()
    out = {}
    for name, params, ret in re.findall(r"pub fn (zoe_cuda_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", text, flags=re.S):
        plist = [p.strip() for p in params.replace("\n", " ").split(",") if p.strip()]
        args = [_canon_rust(p.split(":", 1)[1]) for p in plist]
        out[name] = (args, _canon_rust(ret) if ret.strip() else (0, False, "void"))
    return out


def test_rust_extern_block_matches_header_signatures():
    h, r = header_signatures(), rust_signatures()
    assert sorted(h) == sorted(r) == declared_symbols()
    for name in h:
        assert h[name][0] == r[name][0], (name, h[name][0], r[name][0])
        assert h[name][1] == r[name][1], (name, "return", h[name][1], r[name][1])


def test_rust_and_ctypes_stats_struct_match_header():
    text = open(os.path.join(ROOT, "include", "zoe_cuda.h")).read()
    body = re.search(r"typedef struct\s*\{(.*?)\}\s*zoe_cuda_stats;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [f.strip() for decl in re.findall(r"uint64_t([^;]*);", body) for f in decl.split(",")]
    assert fields == [n for n, _ in _lib.Stats._fields_]
    rust = open(os.path.join(ROOT, "rust", "zoe-cuda-sys", "src", "lib.rs")).read()
    rbody = re.search(r"pub struct zoe_cuda_stats\s*\{(.*?)\}", rust, flags=re.S).group(1)
    assert re.findall(r"pub (\w+): u64", rbody) == fields


def test_ctypes_argtypes_match_header_signatures():
    import ctypes as C
    lib = _lib.load()
    simple = {C.c_int: (0, "i32"), C.c_int8: (0, "i8"), C.c_uint32: (0, "u32"), C.c_uint64: (0, "u64"),
              C.c_float: (0, "f32"), C.c_void_p: (1, "ctx")}
    ptr_base = {C.c_uint8: "u8", C.c_int8: "i8", C.c_uint32: "u32", C.c_uint64: "u64", C.c_int: "i32", C.c_float: "f32",
                C.c_double: "f64", C.c_void_p: "ctx", _lib.Stats: "stats"}
    for name, (args, _ret) in header_signatures().items():
        at = getattr(lib, name).argtypes
        assert at is not None and len(at) == len(args), (name, at, args)
        for k, (a, (depth, _const, base)) in enumerate(zip(at, args)):
            if a in simple:
                d, b = simple[a]
            else:
                d, b = 1, ptr_base[a._type_]
                if a._type_ is C.c_void_p:
                    d = 2
            assert (d, b) == (depth, base), (name, k, a, (depth, base))
