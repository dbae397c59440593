// Links libquadrs_gpu.so; QUADRS_GPU_LIB_DIR points at the directory holding it (quadrs_b200/ in this repo).
fn main() {
    if let Ok(dir) = std::env::var("QUADRS_GPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=quadrs_gpu");
}
