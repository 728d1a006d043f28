// build.rs -- compiles the CUDA sources of thermite_b200/csrc with nvcc for sm_100a and links them into the crate.
// UNCOMPILED in this repository's environment (no cargo/rustc in the image); kept as the integration recipe.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("THERMITE_GPU_CSRC").unwrap_or_else(|_| "thermite_b200/csrc".into()));
    let lib = out.join("libthermite_gpu.a");
    let objs = ["thermite_gpu.cu", "tg_sa.cu", "tg_paf.cu", "tg_multi.cpp", "host_index.cpp", "host_io.cpp", "host_stream.cpp", "host_bam.cpp", "host_batcher.cpp"];
    let mut obj_paths = Vec::new();
    for src in objs.iter() {
        let obj = out.join(format!("{}.o", src));
        let status = Command::new("nvcc")
            .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(src))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found: the GPU aligner has no CPU fallback");
        assert!(status.success(), "nvcc failed on {}", src);
        obj_paths.push(obj);
    }
    let status = Command::new("ar").arg("crs").arg(&lib).args(&obj_paths).status().unwrap();
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=thermite_gpu");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=z");  // BGZF blocks of the BAM writer (host_bam.cpp)
    println!("cargo:rustc-link-lib=dylib=dl");  // tg_multi.cpp loads libnccl.so.2 at run time (index broadcast)
    println!("cargo:rustc-link-lib=dylib=stdc++");
    for src in objs.iter() {
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
    }
    // headers every object depends on (the C ABI header lives in include/)
    for h in ["tg_core.h", "tg_rounds.h", "tg_dpt.h", "tg_internal.h", "host_text.h", "tg_textfmt.h"].iter() {
        println!("cargo:rerun-if-changed={}", csrc.join(h).display());
    }
    println!("cargo:rerun-if-changed={}", csrc.join("../../include").join("thermite_gpu.h").display());
}
