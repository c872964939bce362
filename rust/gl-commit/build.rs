// Compiles the CUDA library for sm_100a with nvcc and links it statically; with the `prebuilt` feature links the
// in-tree libgl_commit.so instead.  The repository root is two levels up (rust/gl-commit/).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("plonky2.5_b200/csrc");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/gl_commit.h").display());
    if env::var("CARGO_FEATURE_PREBUILT").is_ok() {
        let dir = root.join("plonky2.5_b200");
        println!("cargo:rustc-link-search=native={}", dir.display());
        println!("cargo:rustc-link-lib=dylib=gl_commit");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
        return;
    }
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let obj = out.join("gl_commit.o");
    let st = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-c"])
        .arg(csrc.join("gl_commit.cu"))
        .arg("-o")
        .arg(&obj)
        .status()
        .expect("nvcc not found (set NVCC)");
    assert!(st.success(), "nvcc failed");
    let st = Command::new("ar").arg("crs").arg(out.join("libgl_commit.a")).arg(&obj).status().expect("ar");
    assert!(st.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=gl_commit");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
