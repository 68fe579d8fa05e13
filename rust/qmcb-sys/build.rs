// The qmc crate is `#![forbid(unsafe_code)]` (src/lib.rs:1), so the extern block lives in this
// separate -sys crate.  No bindgen: the extern block in src/lib.rs is hand-written against
// include/qmcb.h.  libqmcb.so is built by `make -C isingmontecarlo_b200/csrc` (nvcc, sm_100a).
fn main() {
    let dir = std::env::var("QMCB_LIB_DIR").unwrap_or_else(|_| "../../isingmontecarlo_b200/_build".into());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=qmcb");
    println!("cargo:rerun-if-env-changed=QMCB_LIB_DIR");
}
