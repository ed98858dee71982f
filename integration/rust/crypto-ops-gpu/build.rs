// SOURCE ONLY -- see Cargo.toml.  MPTV_LIB_DIR = directory that holds libmptv.so.
fn main() {
    let dir = std::env::var("MPTV_LIB_DIR").expect("set MPTV_LIB_DIR to the directory of libmptv.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=mptv");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
