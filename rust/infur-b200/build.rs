// Links libinfur_b200.so.  INFUR_B200_LIB_DIR points at the directory holding it (default: the in-tree build output
// infur_b200/lib of the infur_b200 repository next to this crate).
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("INFUR_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../infur_b200/lib")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=infur_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=INFUR_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=build.rs");
}
