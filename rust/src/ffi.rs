//! Hand-written `extern "C"` declarations for ../include/bnp.h (no bindgen-generated dispatch layer).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const BNP_VARIANT_REFERENCE: c_int = 0;
pub const BNP_VARIANT_ARK: c_int = 1;
pub const BNP_WIRE_ARK_UNCOMPRESSED: c_int = 0;
pub const BNP_WIRE_ARK_COMPRESSED: c_int = 1;
pub const BNP_WIRE_EIP197: c_int = 2;
pub const BNP_POINT_OK: u8 = 0;
pub const BNP_POINT_INFINITY: u8 = 1;
pub const BNP_EMALFORMED: c_int = -6;
pub const BNP_PREP_FQ: usize = 546;

extern "C" {
    pub fn bnp_init(devices: *const c_int, n_devices: c_int) -> c_int;
    pub fn bnp_shutdown();
    /// "nccl" or "peer-copy": how bnp_pairing_product gathers the per-device partials of a multi-device context
    pub fn bnp_gather_transport() -> *const c_char;
    pub fn bnp_strerror(code: c_int) -> *const c_char;
    pub fn bnp_last_error() -> *const c_char;
    pub fn bnp_miller_loop_batch(g1: *const u64, g2: *const u64, out: *mut u64, n: usize) -> c_int;
    pub fn bnp_multi_miller_loop_batch(g1: *const u64, g2: *const u64, out: *mut u64, n: usize, k: c_int) -> c_int;
    pub fn bnp_final_exp_batch(input: *const u64, out: *mut u64, n: usize, variant: c_int) -> c_int;
    /// out: [60][4][n] = easy part, m^x, m^(x^2), m^(x^3), final_exp_native (BNP_WITNESS_* offsets of include/bnp.h)
    pub fn bnp_final_exp_witness_batch(input: *const u64, out: *mut u64, n: usize) -> c_int;
    pub fn bnp_pairing_batch(g1: *const u64, g2: *const u64, out: *mut u64, n: usize, variant: c_int) -> c_int;
    pub fn bnp_multi_pairing_batch(g1: *const u64, g2: *const u64, out: *mut u64, n: usize, k: c_int, variant: c_int) -> c_int;
    pub fn bnp_pairing_product(g1: *const u64, g2: *const u64, out: *mut u64, n: usize, variant: c_int) -> c_int;
    pub fn bnp_frobenius_batch(input: *const u64, out: *mut u64, n: usize, power: usize) -> c_int;
    pub fn bnp_fq12_mul_batch(a: *const u64, b: *const u64, out: *mut u64, n: usize) -> c_int;
    /// pow_native (final_exp_native.rs:56): `exp` = n_limbs little-endian u64 limbs shared by the batch
    pub fn bnp_pow_u64_batch(input: *const u64, out: *mut u64, n: usize, exp: *const u64, n_limbs: usize) -> c_int;
    /// prepared G2 points (include/bnp.h): BNP_PREP_FQ Fq of line coefficients per point
    pub fn bnp_g2_prepare_batch(g2: *const u64, coeffs: *mut u64, n: usize) -> c_int;
    pub fn bnp_pairing_prepared_batch(g1: *const u64, g2: *const u64, prepared: *const u64, out: *mut u64, n: usize,
                                      kv: c_int, kp: c_int, variant: c_int) -> c_int;
    /// wire formats (include/bnp.h): bytes -> SoA + one status byte per element, decoded on the device
    pub fn bnp_decode_g1_batch(fmt: c_int, input: *const u8, n: usize, g1: *mut u64, status: *mut u8) -> c_int;
    pub fn bnp_decode_g2_batch(fmt: c_int, input: *const u8, n: usize, g2: *mut u64, status: *mut u8, check_subgroup: c_int) -> c_int;
    pub fn bnp_encode_fq12_batch(f12: *const u64, n: usize, out: *mut u8) -> c_int;
    pub fn bnp_decode_fq12_batch(input: *const u8, n: usize, f12: *mut u64, status: *mut u8) -> c_int;
    /// scalar multiplication on the device (include/bnp.h): group 1 = G1, 2 = G2; scalars [4][n] plain little-endian limbs
    pub fn bnp_scalar_mul_batch(group: c_int, pts: *const u64, scalars: *const u64, out: *mut u64, inf: *mut u8, n: usize) -> c_int;
    pub fn bnp_eip197_pairing_check(input: *const u8, k: usize, result: *mut c_int) -> c_int;
    pub fn bnp_pairing_dev(device: c_int, stream: *mut c_void, g1: *const u64, g2: *const u64, out: *mut u64,
                           n: usize, k: c_int, variant: c_int) -> c_int;
}
