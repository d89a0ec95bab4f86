// Materialised from INTEGRATION.md section 1 (round 2).  Not compiled in the authoring image: no rustc / cargo.
// build.rs — compile the CUDA side with nvcc for sm_100a and link it statically.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    let nvcc = format!("{cuda}/bin/nvcc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let obj = out.join("fm_gpu.o");
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
               "-fmad=false",                       // keep the reference's FP64 operation order
               "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Iinclude",
               "-c", "cuda/fm_gpu.cu", "-o"])
        .arg(&obj)
        .status()
        .expect("nvcc not found: ferromic's GPU path has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    // the host packer of the 2-bit ingest format is plain C++ (AVX-512 / AVX2 chosen at run time)
    let pack = out.join("fm_host_pack.o");
    let cxx = env::var("CXX").unwrap_or_else(|_| "g++".into());
    assert!(Command::new(&cxx)
        .args(["-O3", "-std=c++17", "-fPIC", "-pthread", "-Iinclude", "-c", "cuda/fm_host_pack.cpp", "-o"])
        .arg(&pack).status().expect("g++ not found").success());
    let lib = out.join("libferromic_gpu.a");
    assert!(Command::new("ar").args(["crs"]).arg(&lib).arg(&obj).arg(&pack).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=static=ferromic_gpu");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=stdc++");
    for f in ["cuda/fm_gpu.cu", "cuda/fm_kernels.cuh", "cuda/fm_wc.cuh", "cuda/fm_device.cuh", "cuda/fm_comm.cuh", "cuda/fm_multi.cuh",
              "cuda/fm_falsta.cuh", "cuda/fm_vcf.cuh", "cuda/fm_host_pack.cpp", "include/ferromic_gpu.h"] {
        println!("cargo:rerun-if-changed={f}");
    }
}
