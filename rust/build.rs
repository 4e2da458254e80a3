// cc-built: compiles the CUDA sources of ../plonky2_bn254_pairing_b200/csrc into a static library with nvcc.
// No bindgen - the extern "C" block in src/ffi.rs is written by hand against ../include/bnp.h.
// Prerequisite: `python -m plonky2_bn254_pairing_b200.microcode.gen` has generated csrc/microcode_gen.cpp.
fn main() {
    let csrc = "../plonky2_bn254_pairing_b200/csrc";
    cc::Build::new()
        .cuda(true)
        .flag("-gencode").flag("arch=compute_100a,code=sm_100a")
        .flag("-O3").flag("-std=c++17").flag("-lineinfo")
        .file(format!("{csrc}/bnp.cu"))
        .file(format!("{csrc}/microcode_gen.cpp"))
        .compile("bnp");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rerun-if-changed={csrc}");
    println!("cargo:rerun-if-changed=../include/bnp.h");
}
