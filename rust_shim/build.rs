//! build.rs -- what `nafcodec/build.rs` would do under feature `cuda` (north star: "a thin extern "C" layer over .cu
//! files that build.rs compiles for sm_100a").  UNBUILT here; nafcodec_b200/csrc/Makefile performs the same steps.
fn main() {
    if std::env::var("CARGO_FEATURE_CUDA").is_err() {
        return;
    }
    let cuda = std::env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    let src = "csrc";
    cc::Build::new()
        .cuda(true)
        .cudart("static")
        .flag("-gencode")
        .flag("arch=compute_100a,code=sm_100a")
        .flag("-lineinfo")
        .flag("-O3")
        .flag("-std=c++17")
        .include("include")
        .files(["zstd_kernels.cu", "naf_kernels.cu", "naf_text.cu", "naf_pack.cu", "nafgpu_api.cu"].iter().map(|f| format!("{src}/{f}")))
        .files(["frame_walk.cpp", "naf_parse.cpp", "nafgpu_pipeline.cpp"].iter().map(|f| format!("{src}/{f}")))
        .compile("nafgpu");
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rerun-if-changed={src}");
}
