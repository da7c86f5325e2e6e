// build.rs of the (new) gtars-overlaprs-sys crate: compiles the CUDA sources of this repository for sm_100a with
// nvcc and links the result.  WRITTEN, NOT BUILT: there is no cargo/rustc in the environment this repository was
// developed in; the nvcc command line is the one gtars_b200/csrc/Makefile uses.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("GTARS_GPU_ROOT").expect("GTARS_GPU_ROOT = checkout of the gtars-b200 repository"));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = vec![];
    // every source of gtars_b200/csrc/Makefile's SRCS (tests/test_abi.py keeps the two lists identical)
    for f in ["api", "index", "kernels", "igd", "fragments", "scoring", "ingest", "comm", "sort", "build", "marshal", "inflate"] {
        let src = root.join(format!("gtars_b200/csrc/cuda/{f}.cu"));
        let obj = out.join(format!("{f}.o"));
        let ok = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--use_fast_math"])
            .args(["-Xcompiler", "-fPIC", "-DGT_FUSED_MINBLOCKS=4", "-c", "-o"])
            .arg(&obj)
            .arg(&src)
            .status()
            .expect("nvcc")
            .success();
        assert!(ok, "nvcc failed on {src:?}");
        println!("cargo:rerun-if-changed={}", src.display());
        objs.push(obj);
    }
    let lib = out.join("libgtars_gpu.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=gtars_gpu");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=cudart_static");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rustc-link-lib=dylib=rt");
    println!("cargo:rustc-link-lib=dylib=pthread");
}
