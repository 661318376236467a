// build.rs of the crate once it links the B200 library (include/rtp.h).
// librtp_b200.so is built by `make -C raytracing-potato_b200/csrc` (nvcc, sm_100a).
fn main() {
    let dir = std::env::var("RTP_B200_LIB_DIR").expect("set RTP_B200_LIB_DIR to the directory holding librtp_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rtp_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=RTP_B200_LIB_DIR");
}
