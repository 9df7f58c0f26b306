// build.rs for the `zlib-cuda` feature. UNCOMPILED in this image (no rustc/cargo) — see rust/README.md.
// Compiles the CUDA sources of compu-b200 with nvcc for sm_100a into libcompu_b200.so and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    if env::var_os("CARGO_FEATURE_ZLIB_CUDA").is_none() {
        return;
    }
    let root = PathBuf::from(env::var("COMPU_B200_DIR").expect("COMPU_B200_DIR = checkout of compu-b200"));
    let csrc = root.join("compu_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = Vec::new();
    for src in ["inflate.cu", "host.cu", "deflate.cu", "synth.cu"] {
        let obj = out.join(src).with_extension("o");
        let ok = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
                   "-Xcompiler", "-fPIC,-fopenmp", "-c"])
            .arg(csrc.join(src)).arg("-o").arg(&obj)
            .status().expect("nvcc not found").success();
        assert!(ok, "nvcc failed on {src}");
        objs.push(obj);
    }
    let so = out.join("libcompu_b200.so");
    let ok = Command::new(&nvcc).args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"]).arg(&so).args(&objs)
        .arg("-lgomp").status().unwrap().success();
    assert!(ok, "link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=compu_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
