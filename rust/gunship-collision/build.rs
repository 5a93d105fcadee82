// build.rs -- compiles the CUDA sources for sm_100a with nvcc and links them into the crate.
// Same flags as gunship_rs_b200/csrc/Makefile: IEEE f32 exactly as the Rust reference computes
// (no FMA contraction, precise divide, no flush-to-zero).
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("gunship_rs_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let obj = out.join("gsc_api.o");
    let status = Command::new(&nvcc)
        .args(&["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
                "-Xcompiler", "-fPIC", "-c"])
        .arg(csrc.join("gsc_api.cu"))
        .arg("-o").arg(&obj)
        .status().expect("nvcc not found: the collision path has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    let lib = out.join("libgunship_collision.a");
    let status = Command::new("ar").arg("crs").arg(&lib).arg(&obj).status().expect("ar");
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=gunship_collision");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    for f in &["gsc_api.cu", "gsc_kernels.cuh", "gsc_pairs.cuh", "gsc_sort.cuh", "gsc_math.cuh", "gsc_hierarchy.cuh", "gsc_group.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/gunship_collision.h").display());
}
