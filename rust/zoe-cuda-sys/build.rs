// build.rs: compile the CUDA library with nvcc for sm_100a and link it.  No Triton, no multi-backend
// dispatch, no CPU fallback.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let src = root.join("zoe_b200/csrc/zoe_cuda.cu");
    let lib = out.join("libzoe_cuda.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let status = Command::new(nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math"])
        .args(["-Xcompiler", "-fPIC", "-shared", "-cudart", "shared", "-o"])
        .arg(&lib)
        .arg(&src)
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=zoe_cuda");
    println!("cargo:rerun-if-changed={}", root.join("zoe_b200/csrc").display());
    println!("cargo:rerun-if-changed={}", root.join("include/zoe_cuda.h").display());
}
