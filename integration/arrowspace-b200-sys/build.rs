// Links libarrowspace_b200.so (built by `python -c "import __graft_entry__ as g; g.build()"` in the B200 repository:
// nvcc -gencode arch=compute_100a,code=sm_100a).  ARROWSPACE_B200_LIB_DIR = the directory that holds the .so.
fn main() {
    println!("cargo:rerun-if-env-changed=ARROWSPACE_B200_LIB_DIR");
    if let Ok(dir) = std::env::var("ARROWSPACE_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=arrowspace_b200");
}
