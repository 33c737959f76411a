// Links libtsgpu.so, built by `make -C multilinear-map-cryptography_b200`
// (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo).  No CUDA toolkit is needed on the Rust side.
fn main() {
    let dir = std::env::var("TSGPU_LIB_DIR")
        .expect("set TSGPU_LIB_DIR to the directory that holds libtsgpu.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=tsgpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=TSGPU_LIB_DIR");
    println!("cargo:rerun-if-changed=build.rs");
}
