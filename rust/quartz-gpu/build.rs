fn main() {
    // libquartz_gpu.so is built by quartz_b200/build.sh (nvcc, sm_100a)
    let dir = std::env::var("QUARTZ_GPU_LIB_DIR").unwrap_or_else(|_| "../../quartz_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=quartz_gpu");
}
