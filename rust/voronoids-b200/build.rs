// Links libvoronoids_b200.so.  VORONOIDS_B200_DIR = directory that holds it (default: ../../voronoids_b200, the in-tree build of
// `python -m voronoids_b200.build`: nvcc -gencode arch=compute_100a,code=sm_100a ... vor_lib.cu).
fn main() {
    let dir = std::env::var("VORONOIDS_B200_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{here}/../../voronoids_b200")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=voronoids_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=VORONOIDS_B200_DIR");
}
