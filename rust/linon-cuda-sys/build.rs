// UNVERIFIED (no Rust toolchain in the build image).  LINON_B200_LIB_DIR = directory of liblinon_b200.so
// (linon_b200/lib after `python -m linon_b200.build`).
fn main() {
    if let Ok(dir) = std::env::var("LINON_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=linon_b200");
    println!("cargo:rerun-if-env-changed=LINON_B200_LIB_DIR");
}
