// Links libfdeflate_b200.so (built by `make -C fdeflate_b200/csrc` at the repository root).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("FDEFLATE_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../fdeflate_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=fdeflate_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=FDEFLATE_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/fdeflate_b200.h");
}
