// Point the linker at the directory that holds libmsm_b200.so (built by `make -C msm_b200/csrc`).
fn main() {
    let dir = std::env::var("MSM_B200_LIB_DIR").unwrap_or_else(|_| "../../msm_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=msm_b200");
    println!("cargo:rerun-if-env-changed=MSM_B200_LIB_DIR");
}
