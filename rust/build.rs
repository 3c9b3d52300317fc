// build.rs for the kzg crate with the `b200` feature: link the prebuilt C-ABI library.
// (cc is not needed: the CUDA code is built by the Makefile of kzg-b200 with nvcc -gencode arch=compute_100a,code=sm_100a.)
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("KZG_B200_LIB_DIR").expect("set KZG_B200_LIB_DIR to the directory holding libkzg_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=kzg_b200");
    println!("cargo:rerun-if-env-changed=KZG_B200_LIB_DIR");

    // Optional: regenerate src/ffi_bindgen.rs from the header instead of using the hand-written src/ffi.rs.
    if let Ok(header) = env::var("KZG_B200_HEADER") {
        let bindings = bindgen::Builder::default()
            .header(header)
            .allowlist_function("kzgb_.*")
            .allowlist_type("kzgb_.*")
            .allowlist_var("KZGB_.*")
            .generate()
            .expect("bindgen failed on kzg_b200.h");
        let out = PathBuf::from(env::var("OUT_DIR").unwrap());
        bindings.write_to_file(out.join("ffi_bindgen.rs")).unwrap();
    }
}
