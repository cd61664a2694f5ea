// Links libcm31.so (built by `python cairo-m_b200/build.py`).  CM31_LIB_DIR points at the directory holding it.
fn main() {
    let dir = std::env::var("CM31_LIB_DIR").unwrap_or_else(|_| "../../cairo-m_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=cm31");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=CM31_LIB_DIR");
}
