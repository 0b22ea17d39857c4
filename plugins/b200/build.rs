// Link against openzl_b200/libozl_b200.so (built by `make -C openzl_b200/csrc`).
fn main() {
    let dir = std::env::var("OZL_B200_LIB_DIR").unwrap_or_else(|_| "../../openzl_b200".into());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=ozl_b200");
    println!("cargo:rerun-if-env-changed=OZL_B200_LIB_DIR");
}
