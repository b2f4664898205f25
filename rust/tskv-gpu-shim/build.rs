// Links libtskv_gpu.so (built by `python -m cnosdb_b200.build`, sm_100a). TSKV_GPU_LIB_DIR points at the directory
// holding it; libnccl is loaded by the library itself with dlopen on first use, so nothing else is linked here.
fn main() {
    if let Ok(dir) = std::env::var("TSKV_GPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=tskv_gpu");
    println!("cargo:rerun-if-env-changed=TSKV_GPU_LIB_DIR");
}
