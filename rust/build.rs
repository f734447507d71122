// UNCOMPILED (no rustc in this image).  Links the in-tree shared library built by `python -m mixlab_b200.build`.
fn main() {
    let dir = std::env::var("MIXLAB_B200_LIB_DIR").unwrap_or_else(|_| format!("{}/../mixlab_b200", env!("CARGO_MANIFEST_DIR")));
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=mixlab_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-changed=../include/mixlab_b200.h");
}
