// build.rs -- links the B200 pivoting engine when the `gpu` feature is enabled.
//   ELLP_B200_LIB_DIR=/path/to/repo/ellp_b200 cargo test --features gpu
fn main() {
    println!("cargo:rerun-if-env-changed=ELLP_B200_LIB_DIR");
    if std::env::var("CARGO_FEATURE_GPU").is_ok() {
        let dir = std::env::var("ELLP_B200_LIB_DIR")
            .expect("feature `gpu`: set ELLP_B200_LIB_DIR to the directory that holds libellp_b200.so");
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-lib=dylib=ellp_b200");
        // so that `cargo test` finds the library without LD_LIBRARY_PATH
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
}
