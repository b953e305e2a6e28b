// build.rs — link the prebuilt C-ABI library and generate the FFI declarations from include/flowgger_cuda.h.
use std::env;
use std::path::PathBuf;

fn main() {
    let root = PathBuf::from(env::var("FLOWGGER_B200_ROOT").unwrap_or_else(|_| "../..".into()));
    let lib_dir = root.join("flowgger_b200/lib");
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=flowgger_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());
    println!("cargo:rerun-if-changed={}", root.join("include/flowgger_cuda.h").display());
    let bindings = bindgen::Builder::default()
        .header(root.join("include/flowgger_cuda.h").to_str().unwrap())
        .allowlist_function("fg_.*")
        .allowlist_type("fg_.*")
        .allowlist_var("FG_.*")
        .generate()
        .expect("bindgen failed on flowgger_cuda.h");
    bindings
        .write_to_file(PathBuf::from(env::var("OUT_DIR").unwrap()).join("ffi.rs"))
        .unwrap();
}
