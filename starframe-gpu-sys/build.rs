// UNCOMPILED HERE (no Rust toolchain in this image). Builds libstarframe_gpu.so from the CUDA sources with the same nvcc
// command line as __graft_entry__.build() and links it.
//
// The CUDA sources live in this repository under moleengine_b200/csrc (sfg_api.cu includes the .cuh files beside it) and
// include/ (starframe_gpu.h, sfg_hash.h); STARFRAME_GPU_SRC overrides the repository root.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let root = env::var("STARFRAME_GPU_SRC").map(PathBuf::from).unwrap_or_else(|_| manifest.join(".."));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let so = out.join("libstarframe_gpu.so");
    let status = Command::new(env::var("NVCC").unwrap_or_else(|_| "nvcc".into()))
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"])
        // `/` and sqrtf as the 2-ulp MUFU sequences: the solver kernel is code-size bound; everything that must be
        // bit-exact (tick-time boxes, the `near` hint) uses explicit _rn intrinsics
        .args(["-prec-div=false", "-prec-sqrt=false"])
        .args(["-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&so)
        .arg(root.join("moleengine_b200/csrc/sfg_api.cu"))
        .status()
        .expect("nvcc not found (set NVCC)");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=starframe_gpu");
    println!("cargo:rerun-if-changed={}", root.join("moleengine_b200/csrc").display());
    println!("cargo:rerun-if-changed={}", root.join("include").display());
    println!("cargo:rerun-if-env-changed=STARFRAME_GPU_SRC");
}
