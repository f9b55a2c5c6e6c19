// Link against cv_b200/libcvb200.so: CVB200_LIB_DIR overrides the in-tree location (repo root/cv_b200, built by `make -C cv_b200/csrc`).
fn main() {
    let dir = std::env::var("CVB200_LIB_DIR")
        .unwrap_or_else(|_| format!("{}/../../../cv_b200", std::env::var("CARGO_MANIFEST_DIR").unwrap()));
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=cvb200");
    println!("cargo:rerun-if-env-changed=CVB200_LIB_DIR");
}
