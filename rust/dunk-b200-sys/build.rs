// Links the prebuilt libdunk_b200.so (make -C cubesat-apds_b200/csrc).  DUNK_B200_LIB_DIR = directory holding it.
fn main() {
    let dir = std::env::var("DUNK_B200_LIB_DIR").expect("set DUNK_B200_LIB_DIR to the directory of libdunk_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=dunk_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=DUNK_B200_LIB_DIR");
}
