// Links libpb254.so. Set PB254_LIB_DIR to the directory that holds it
// (…/plonky2_bn254_b200 after `python -m plonky2_bn254_b200.build`).
fn main() {
    let dir = std::env::var("PB254_LIB_DIR").expect("set PB254_LIB_DIR to the directory of libpb254.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=pb254");
    println!("cargo:rerun-if-env-changed=PB254_LIB_DIR");
}
