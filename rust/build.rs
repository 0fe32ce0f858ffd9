// build.rs for the reference crate when the `b200` feature is on: link libanemoi_b200.so.
// ANEMOI_B200_LIB_DIR points at the directory holding the library (anemoi_rust_b200/ in this repo).
fn main() {
    if std::env::var("CARGO_FEATURE_B200").is_ok() {
        let dir = std::env::var("ANEMOI_B200_LIB_DIR").unwrap_or_else(|_| "/usr/local/lib".to_string());
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-lib=dylib=anemoi_b200");
        println!("cargo:rerun-if-env-changed=ANEMOI_B200_LIB_DIR");
    }
}
