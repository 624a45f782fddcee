// Links the prebuilt CUDA library.  B200ZK_LIB_DIR points at the directory holding libb200zk.so
// (built by `make -C ethrex_b200/csrc`); the CUDA runtime is linked statically into that library.
fn main() {
    println!("cargo:rerun-if-env-changed=B200ZK_LIB_DIR");
    if let Ok(dir) = std::env::var("B200ZK_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=b200zk");
}
