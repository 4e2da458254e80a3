//! Drop-in for the native path of qope/plonky2-bn254-pairing: same names, same ark-bn254 signatures,
//! plus the batched slice variants; all arithmetic happens in libbnp (CUDA, sm_100a).
//!
//!   miller_loop_native        <- src/miller_loop_native.rs:320
//!   multi_miller_loop_native  <- src/miller_loop_native.rs:324
//!   final_exp_native          <- src/final_exp_native.rs:209
//!   frobenius_map_native      <- src/final_exp_native.rs:17
//!   pow_native                <- src/final_exp_native.rs:56
//!   get_naf, frob_coeffs      <- src/final_exp_native.rs:86, :183   (host-side table generators; final_exp_target.rs:18 imports frob_coeffs)
//!   conjugate_fp2, neg_conjugate_fp2 <- src/miller_loop_native.rs:284, :291
//!   pairing                   <- src/pairing.rs:20
//!
//! Marshalling: ark's `Fq` is `Fp(BigInt<4>([u64; 4]), _)` in Montgomery form with R = 2^256 - exactly the
//! limbs libbnp consumes - so it is a limb copy into structure-of-arrays buffers `[K][4][n]`, no conversion.
//! Rust structs are `repr(Rust)`: they are never passed by pointer, limbs are copied field by field.
//! Errors: the reference panics (assert!/unwrap); these wrappers panic with libbnp's message.
mod ffi;

use ark_bn254::{Fq, Fq12, Fq2, Fr, G1Affine, G2Affine};
use ark_ff::{BigInt, Fp, PrimeField};
use plonky2_bn254::fields::native::MyFq12;
use std::ffi::CStr;

pub const BN_X: u64 = 4965661367192848881; // final_exp_native.rs:15
pub const SIX_U_PLUS_2_NAF: [i8; 65] = [
    0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 1, -1, 0, 0, 1, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, -1, 0, 0, 0, 0, 1, 1, 1, 0, 0, -1, 0, 0,
    1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 1, 0, 0, -1, 0, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, 1, 1,
]; // miller_loop_native.rs:314

fn check(rc: i32) {
    if rc != 0 {
        let (a, b) = unsafe {
            (CStr::from_ptr(ffi::bnp_strerror(rc)).to_string_lossy().into_owned(),
             CStr::from_ptr(ffi::bnp_last_error()).to_string_lossy().into_owned())
        };
        panic!("libbnp: {a} ({rc}) {b}");
    }
}

fn init() {
    use std::sync::Once;
    static ONCE: Once = Once::new();
    ONCE.call_once(|| check(unsafe { ffi::bnp_init(core::ptr::null(), 0) }));
}

#[inline]
fn put(buf: &mut [u64], k: usize, n: usize, e: usize, v: &Fq) {
    for j in 0..4 {
        buf[(k * 4 + j) * n + e] = (v.0).0[j];
    }
}
#[inline]
fn get(buf: &[u64], k: usize, n: usize, e: usize) -> Fq {
    let mut l = [0u64; 4];
    for j in 0..4 {
        l[j] = buf[(k * 4 + j) * n + e];
    }
    Fp::new_unchecked(BigInt(l)) // already Montgomery, canonical
}

fn pack_g1(ps: &[G1Affine]) -> Vec<u64> {
    let n = ps.len();
    let mut b = vec![0u64; 8 * n];
    for (e, p) in ps.iter().enumerate() {
        put(&mut b, 0, n, e, &p.x);
        put(&mut b, 1, n, e, &p.y);
    }
    b
}
fn pack_g2(qs: &[G2Affine]) -> Vec<u64> {
    let n = qs.len();
    let mut b = vec![0u64; 16 * n];
    for (e, q) in qs.iter().enumerate() {
        put(&mut b, 0, n, e, &q.x.c0);
        put(&mut b, 1, n, e, &q.x.c1);
        put(&mut b, 2, n, e, &q.y.c0);
        put(&mut b, 3, n, e, &q.y.c1);
    }
    b
}
fn pack_fq12(fs: &[MyFq12]) -> Vec<u64> {
    let n = fs.len();
    let mut b = vec![0u64; 48 * n];
    for (e, f) in fs.iter().enumerate() {
        for k in 0..12 {
            put(&mut b, k, n, e, &f.coeffs[k]);
        }
    }
    b
}
fn unpack_fq12(b: &[u64], n: usize) -> Vec<MyFq12> {
    (0..n)
        .map(|e| MyFq12 { coeffs: core::array::from_fn(|k| get(b, k, n, e)) })
        .collect()
}

// ---- scalar multiplication ---------------------------------------------------
fn pack_scalars(ks: &[Fr]) -> Vec<u64> {
    let n = ks.len();
    let mut b = vec![0u64; 4 * n];
    for (e, k) in ks.iter().enumerate() {
        let l = k.into_bigint().0; // the plain integer, not the Montgomery residue
        for j in 0..4 {
            b[j * n + e] = l[j];
        }
    }
    b
}

/// `ps[i] * ks[i]` for the whole slice on the device - what `(G1Affine * Fr).into()` gives one at a time
/// (`G1.mul(s).into()`, final_exp_native.rs:247); the identity comes back as `G1Affine::identity()`.
pub fn g1_scalar_mul_batch(ps: &[G1Affine], ks: &[Fr]) -> Vec<G1Affine> {
    assert_eq!(ps.len(), ks.len());
    init();
    let n = ps.len();
    let (g1, sc) = (pack_g1(ps), pack_scalars(ks));
    let mut out = vec![0u64; 8 * n];
    let mut inf = vec![0u8; n];
    check(unsafe { ffi::bnp_scalar_mul_batch(1, g1.as_ptr(), sc.as_ptr(), out.as_mut_ptr(), inf.as_mut_ptr(), n) });
    (0..n)
        .map(|e| if inf[e] != 0 { G1Affine::identity() } else { G1Affine::new_unchecked(get(&out, 0, n, e), get(&out, 1, n, e)) })
        .collect()
}

/// `qs[i] * ks[i]` on G2 (`G2.mul(t).into()`, final_exp_native.rs:248)
pub fn g2_scalar_mul_batch(qs: &[G2Affine], ks: &[Fr]) -> Vec<G2Affine> {
    assert_eq!(qs.len(), ks.len());
    init();
    let n = qs.len();
    let (g2, sc) = (pack_g2(qs), pack_scalars(ks));
    let mut out = vec![0u64; 16 * n];
    let mut inf = vec![0u8; n];
    check(unsafe { ffi::bnp_scalar_mul_batch(2, g2.as_ptr(), sc.as_ptr(), out.as_mut_ptr(), inf.as_mut_ptr(), n) });
    (0..n)
        .map(|e| {
            if inf[e] != 0 {
                G2Affine::identity()
            } else {
                G2Affine::new_unchecked(Fq2::new(get(&out, 0, n, e), get(&out, 1, n, e)), Fq2::new(get(&out, 2, n, e), get(&out, 3, n, e)))
            }
        })
        .collect()
}

// ---- batched slice variants ----------------------------------------------------------------
pub fn miller_loop_native_batch(qs: &[G2Affine], ps: &[G1Affine]) -> Vec<MyFq12> {
    assert_eq!(qs.len(), ps.len());
    init();
    let n = ps.len();
    let (g1, g2) = (pack_g1(ps), pack_g2(qs));
    let mut out = vec![0u64; 48 * n];
    check(unsafe { ffi::bnp_miller_loop_batch(g1.as_ptr(), g2.as_ptr(), out.as_mut_ptr(), n) });
    unpack_fq12(&out, n)
}

pub fn final_exp_native_batch(fs: &[MyFq12]) -> Vec<MyFq12> {
    init();
    let n = fs.len();
    let inp = pack_fq12(fs);
    let mut out = vec![0u64; 48 * n];
    check(unsafe { ffi::bnp_final_exp_batch(inp.as_ptr(), out.as_mut_ptr(), n, ffi::BNP_VARIANT_REFERENCE) });
    unpack_fq12(&out, n)
}

/// Native values the final-exponentiation circuit takes as witnesses (final_exp_target.rs:65-185), one GPU pass:
/// the easy part `m`, the three `Fq12ExpU64` stark outputs `m^x`, `m^(x^2)`, `m^(x^3)` and `final_exp_native(a)`.
pub struct FinalExpWitness {
    pub m: MyFq12,
    pub mx: MyFq12,
    pub mx2: MyFq12,
    pub mx3: MyFq12,
    pub out: MyFq12,
}

pub fn final_exp_witness_batch(fs: &[MyFq12]) -> Vec<FinalExpWitness> {
    init();
    let n = fs.len();
    let inp = pack_fq12(fs);
    let mut out = vec![0u64; 4 * 60 * n];
    check(unsafe { ffi::bnp_final_exp_witness_batch(inp.as_ptr(), out.as_mut_ptr(), n) });
    let part = |e: usize, base: usize| MyFq12 { coeffs: core::array::from_fn(|k| get(&out, base + k, n, e)) };
    (0..n)
        .map(|e| FinalExpWitness { m: part(e, 0), mx: part(e, 12), mx2: part(e, 24), mx3: part(e, 36), out: part(e, 48) })
        .collect()
}

pub fn pairing_batch(ps: &[G1Affine], qs: &[G2Affine]) -> Vec<Fq12> {
    pairing_batch_variant(ps, qs, ffi::BNP_VARIANT_REFERENCE)
}

/// variant 0: the reference's exponent (p^12-1)/r; variant 1: ark-ec 0.4.2's `Bn254::pairing` exponent.
pub fn pairing_batch_variant(ps: &[G1Affine], qs: &[G2Affine], variant: i32) -> Vec<Fq12> {
    assert_eq!(qs.len(), ps.len());
    init();
    let n = ps.len();
    let (g1, g2) = (pack_g1(ps), pack_g2(qs));
    let mut out = vec![0u64; 48 * n];
    check(unsafe { ffi::bnp_pairing_batch(g1.as_ptr(), g2.as_ptr(), out.as_mut_ptr(), n, variant) });
    unpack_fq12(&out, n).into_iter().map(Into::into).collect() // MyFq12 -> Fq12, as pairing.rs:21
}

/// `K`-way products (Groth16-verify shape): one Fq12 per group of K pairs, K <= 4.
pub fn multi_pairing_batch<const K: usize>(groups: &[[(G1Affine, G2Affine); K]]) -> Vec<Fq12> {
    init();
    let n = groups.len();
    let mut g1 = vec![0u64; 8 * K * n];
    let mut g2 = vec![0u64; 16 * K * n];
    for (e, g) in groups.iter().enumerate() {
        for (j, (p, q)) in g.iter().enumerate() {
            put(&mut g1, 2 * j, n, e, &p.x);
            put(&mut g1, 2 * j + 1, n, e, &p.y);
            put(&mut g2, 4 * j, n, e, &q.x.c0);
            put(&mut g2, 4 * j + 1, n, e, &q.x.c1);
            put(&mut g2, 4 * j + 2, n, e, &q.y.c0);
            put(&mut g2, 4 * j + 3, n, e, &q.y.c1);
        }
    }
    let mut out = vec![0u64; 48 * n];
    check(unsafe {
        ffi::bnp_multi_pairing_batch(g1.as_ptr(), g2.as_ptr(), out.as_mut_ptr(), n, K as i32, ffi::BNP_VARIANT_REFERENCE)
    });
    unpack_fq12(&out, n).into_iter().map(Into::into).collect()
}

// ---- the reference's scalar signatures -------------------------------------------------------
#[allow(non_snake_case)]
pub fn miller_loop_native(Q: &G2Affine, P: &G1Affine) -> MyFq12 {
    miller_loop_native_batch(core::slice::from_ref(Q), core::slice::from_ref(P)).remove(0)
}

pub fn multi_miller_loop_native(pairs: Vec<(&G1Affine, &G2Affine)>) -> MyFq12 {
    // k <= 4: one shared-squaring launch; larger k: product of single loops (equal by miller_loop_native.rs:336-348)
    let ps: Vec<G1Affine> = pairs.iter().map(|p| *p.0).collect();
    let qs: Vec<G2Affine> = pairs.iter().map(|p| *p.1).collect();
    if pairs.len() <= 4 {
        init();
        let k = pairs.len();
        let (mut g1, mut g2) = (vec![0u64; 8 * k], vec![0u64; 16 * k]);
        for j in 0..k {
            put(&mut g1, 2 * j, 1, 0, &ps[j].x);
            put(&mut g1, 2 * j + 1, 1, 0, &ps[j].y);
            put(&mut g2, 4 * j, 1, 0, &qs[j].x.c0);
            put(&mut g2, 4 * j + 1, 1, 0, &qs[j].x.c1);
            put(&mut g2, 4 * j + 2, 1, 0, &qs[j].y.c0);
            put(&mut g2, 4 * j + 3, 1, 0, &qs[j].y.c1);
        }
        let mut out = vec![0u64; 48];
        check(unsafe { ffi::bnp_multi_miller_loop_batch(g1.as_ptr(), g2.as_ptr(), out.as_mut_ptr(), 1, k as i32) });
        unpack_fq12(&out, 1).remove(0)
    } else {
        miller_loop_native_batch(&qs, &ps).into_iter().reduce(|a, b| a * b).unwrap()
    }
}

pub fn final_exp_native(a: MyFq12) -> MyFq12 {
    final_exp_native_batch(core::slice::from_ref(&a)).remove(0)
}

pub fn frobenius_map_native(a: MyFq12, power: usize) -> MyFq12 {
    init();
    let inp = pack_fq12(core::slice::from_ref(&a));
    let mut out = vec![0u64; 48];
    check(unsafe { ffi::bnp_frobenius_batch(inp.as_ptr(), out.as_mut_ptr(), 1, power) });
    unpack_fq12(&out, 1).remove(0)
}

pub fn pairing(p: G1Affine, q: G2Affine) -> Fq12 {
    pairing_batch(&[p], &[q]).remove(0)
}

/// `a^exp` for any `a` (final_exp_native.rs:56-84), `exp` little-endian u64 limbs.  Same quirk as the reference for a
/// zero exponent: its accumulator starts at `a` and the loop never starts, so `a` itself comes back.
pub fn pow_native_batch(fs: &[MyFq12], exp: &[u64]) -> Vec<MyFq12> {
    init();
    let n = fs.len();
    let inp = pack_fq12(fs);
    let mut out = vec![0u64; 48 * n];
    check(unsafe { ffi::bnp_pow_u64_batch(inp.as_ptr(), out.as_mut_ptr(), n, exp.as_ptr(), exp.len()) });
    unpack_fq12(&out, n)
}

pub fn pow_native(a: MyFq12, exp: Vec<u64>) -> MyFq12 {
    pow_native_batch(core::slice::from_ref(&a), &exp).remove(0)
}

/// Signed-digit (non-adjacent form) expansion, least significant digit first (final_exp_native.rs:86-128).
/// Pure host-side table generation - the GPU programs have the digits of BN_X and 6x+2 baked in at build time, and
/// `bnp_pow_u64_batch` derives them itself - kept so that `use ...::get_naf` keeps compiling.
pub fn get_naf(mut exp: Vec<u64>) -> Vec<i8> {
    let mut naf: Vec<i8> = Vec::with_capacity(64 * exp.len());
    let len = exp.len();
    for idx in 0..len {
        let mut e: u64 = exp[idx];
        for _ in 0..64 {
            if e & 1 == 1 {
                let z = 2i8 - (e % 4) as i8;
                e /= 2;
                if z == -1 {
                    e += 1;
                }
                naf.push(z);
            } else {
                naf.push(0);
                e /= 2;
            }
        }
        if e != 0 {
            let mut j = idx + 1;
            while j < exp.len() && exp[j] == u64::MAX {
                exp[j] = 0;
                j += 1;
            }
            if j < exp.len() {
                exp[j] += 1;
            } else {
                exp.push(1);
            }
        }
    }
    // the reference asserts `len == exp.len() + 1` here (:123), which cannot hold once `exp` has grown: it panics on
    // a carry out of the top limb, and so does this
    assert!(exp.len() == len, "get_naf: carry out of the top limb");
    naf
}

/// xi^((p^index - 1) / 6), xi = 9 + u (final_exp_native.rs:183-192; imported by final_exp_target.rs:18).
pub fn frob_coeffs(index: usize) -> Fq2 {
    use ark_ff::{Field, PrimeField};
    use num_bigint::BigUint;
    let modulus: BigUint = Fq::MODULUS.into();
    let num = modulus.pow(index as u32) - 1u32;
    let k = num / 6u32;
    Fq2::new(Fq::from(9u64), Fq::from(1u64)).pow(k.to_u64_digits())
}

// conjugate_fp2 / neg_conjugate_fp2 (miller_loop_native.rs:284-296) are two field negations: kept on the CPU.
pub fn conjugate_fp2(x: Fq2) -> Fq2 {
    Fq2::new(x.c0, -x.c1)
}
pub fn neg_conjugate_fp2(x: Fq2) -> Fq2 {
    Fq2::new(-x.c0, x.c1)
}

// ---- wire formats decoded on the device -------------------------------------
/// `CanonicalSerialize::serialize_compressed` bytes of `n` G1Affine / G2Affine points (32 / 64 bytes each) straight
/// into the batched pairing: decompression (the square roots), curve and subgroup checks run on the GPU.
/// Panics - like `G1Affine::deserialize_compressed(..).unwrap()` would - if a point is malformed; points at infinity are
/// not accepted (the reference's Miller loop has no meaning for them).
pub fn pairing_batch_from_compressed(g1_bytes: &[u8], g2_bytes: &[u8]) -> Vec<Fq12> {
    assert!(g1_bytes.len() % 32 == 0 && g2_bytes.len() == 2 * g1_bytes.len());
    let n = g1_bytes.len() / 32;
    init();
    let (mut g1, mut g2) = (vec![0u64; 8 * n], vec![0u64; 16 * n]);
    let (mut s1, mut s2) = (vec![0u8; n], vec![0u8; n]);
    check(unsafe { ffi::bnp_decode_g1_batch(ffi::BNP_WIRE_ARK_COMPRESSED, g1_bytes.as_ptr(), n, g1.as_mut_ptr(), s1.as_mut_ptr()) });
    check(unsafe { ffi::bnp_decode_g2_batch(ffi::BNP_WIRE_ARK_COMPRESSED, g2_bytes.as_ptr(), n, g2.as_mut_ptr(), s2.as_mut_ptr(), 1) });
    for e in 0..n {
        assert!(s1[e] == ffi::BNP_POINT_OK && s2[e] == ffi::BNP_POINT_OK, "pair {e}: G1 status {}, G2 status {}", s1[e], s2[e]);
    }
    let mut out = vec![0u64; 48 * n];
    check(unsafe { ffi::bnp_pairing_batch(g1.as_ptr(), g2.as_ptr(), out.as_mut_ptr(), n, ffi::BNP_VARIANT_REFERENCE) });
    unpack_fq12(&out, n).into_iter().map(|f| f.into()).collect()
}

/// `serialize_compressed` bytes (384 per element) of the ark `Fq12` values a batch of `MyFq12` converts into.
pub fn fq12_to_bytes_batch(fs: &[MyFq12]) -> Vec<u8> {
    init();
    let mut out = vec![0u8; 384 * fs.len()];
    check(unsafe { ffi::bnp_encode_fq12_batch(pack_fq12(fs).as_ptr(), fs.len(), out.as_mut_ptr()) });
    out
}

/// The Ethereum pairing precompile (address 0x08, EIP-197): `Some(true / false)`, or `None` where the precompile fails
/// (a coordinate >= p, a point off its curve, a G2 point outside the subgroup, a length that is not a multiple of 192).
pub fn eip197_pairing_check(input: &[u8]) -> Option<bool> {
    if input.len() % 192 != 0 {
        return None;
    }
    init();
    let mut res: core::ffi::c_int = 0;
    let rc = unsafe { ffi::bnp_eip197_pairing_check(input.as_ptr(), input.len() / 192, &mut res) };
    if rc == ffi::BNP_EMALFORMED {
        return None;
    }
    check(rc);
    Some(res == 1)
}

// ---- prepared G2 points ------------------------------------------------------
/// The engine's `G2Prepared`: the line coefficients of fixed G2 points (a Groth16 verifying key's beta, gamma, delta),
/// computed once on the GPU and shared by every proof of a batch.  Layout `[kp * BNP_PREP_FQ][4][1]` u64.
pub struct PreparedG2 {
    coeffs: Vec<u64>,
    kp: usize,
}

pub fn prepare_g2(qs: &[G2Affine]) -> PreparedG2 {
    init();
    let kp = qs.len();
    let mut per_point = vec![0u64; ffi::BNP_PREP_FQ * 4 * kp];           // [BNP_PREP_FQ][4][kp]
    check(unsafe { ffi::bnp_g2_prepare_batch(pack_g2(qs).as_ptr(), per_point.as_mut_ptr(), kp) });
    let mut coeffs = vec![0u64; ffi::BNP_PREP_FQ * 4 * kp];              // [kp * BNP_PREP_FQ][4][1]
    for j in 0..kp {
        for row in 0..ffi::BNP_PREP_FQ * 4 {
            coeffs[j * ffi::BNP_PREP_FQ * 4 + row] = per_point[row * kp + j];
        }
    }
    PreparedG2 { coeffs, kp }
}

/// out[i] = final_exp( e-Miller(ps_live[i], qs_live[i]) * prod_j e-Miller(ps_fixed[i][j], key.point[j]) ): one live pair and
/// `key.kp` (2 or 3) prepared pairs per proof - the Groth16 verification equation with the key prepared once.
pub fn groth16_shaped_batch(ps_live: &[G1Affine], qs_live: &[G2Affine], ps_fixed: &[Vec<G1Affine>], key: &PreparedG2) -> Vec<Fq12> {
    let n = ps_live.len();
    assert!(qs_live.len() == n && ps_fixed.len() == n && ps_fixed.iter().all(|v| v.len() == key.kp));
    init();
    let k = 1 + key.kp;
    let mut g1 = vec![0u64; 8 * k * n];
    for e in 0..n {
        put(&mut g1, 0, n, e, &ps_live[e].x);
        put(&mut g1, 1, n, e, &ps_live[e].y);
        for j in 0..key.kp {
            put(&mut g1, 2 * (j + 1), n, e, &ps_fixed[e][j].x);
            put(&mut g1, 2 * (j + 1) + 1, n, e, &ps_fixed[e][j].y);
        }
    }
    let mut out = vec![0u64; 48 * n];
    check(unsafe {
        ffi::bnp_pairing_prepared_batch(g1.as_ptr(), pack_g2(qs_live).as_ptr(), key.coeffs.as_ptr(), out.as_mut_ptr(), n, 1,
                                        key.kp as i32, ffi::BNP_VARIANT_REFERENCE)
    });
    unpack_fq12(&out, n).into_iter().map(|f| f.into()).collect()
}
