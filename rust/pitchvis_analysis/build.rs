// Links libpvqt.so.  PVQT_LIB_DIR points at the directory holding it
// (<repo>/pitchvis_b200/lib after `python -m pitchvis_b200.build`).
fn main() {
    if let Ok(dir) = std::env::var("PVQT_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=pvqt");
    println!("cargo:rerun-if-env-changed=PVQT_LIB_DIR");
}
