// UNCOMPILED (no Rust toolchain in the build image).  Links libredux_b200.so from REDUX_B200_LIB_DIR.
fn main() {
    let dir = std::env::var("REDUX_B200_LIB_DIR").expect("set REDUX_B200_LIB_DIR to the directory of libredux_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=redux_b200");
    println!("cargo:rerun-if-env-changed=REDUX_B200_LIB_DIR");
}
