// Links libsdf2mesh_b200.so (built by `make -C sdf2mesh_b200/csrc`).  NOT COMPILED in the build image (no rustc).
fn main() {
    let dir = std::env::var("SDF2MESH_B200_LIB_DIR").unwrap_or_else(|_| "../../../sdf2mesh_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=sdf2mesh_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=SDF2MESH_B200_LIB_DIR");
}
