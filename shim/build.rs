// Link against libzerocaf_b200.so.  ZEROCAF_B200_LIB_DIR points at the directory that holds it (the repository builds it
// in-tree as dusk_zerocaf_b200/libzerocaf_b200.so); the CUDA runtime is a dependency of that library, not of this crate.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("ZEROCAF_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..").join("dusk_zerocaf_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=zerocaf_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=ZEROCAF_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../include/zerocaf_b200.h");
}
