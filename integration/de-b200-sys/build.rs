// Links libde_b200.so (built by `python -c "import __graft_entry__ as g; g.build()"` in the backend repository).
fn main() {
    let dir = std::env::var("DE_B200_LIB_DIR").expect("set DE_B200_LIB_DIR to the directory holding libde_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=de_b200");
    println!("cargo:rerun-if-env-changed=DE_B200_LIB_DIR");
}
