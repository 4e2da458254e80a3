//! The pin this repository cannot run itself (no cargo / rustc in its build image): the GPU engine against the REAL
//! reference crate and against arkworks, on the committed golden inputs and on random points.
//!
//!     cd rust && cargo test --release -- --nocapture        (needs a CUDA device, nvcc, and network for the git deps)
//!
//! What each test closes:
//!   * `golden_vectors_match_the_reference`  - tests/golden/survey_kat.json was produced by a transcription of the
//!     reference; here the reference itself recomputes it.
//!   * `engine_equals_reference_on_random_inputs` - pairing / miller_loop_native / final_exp_native / pow_native /
//!     frobenius_map_native bit for bit, incl. the external `MyFq12 -> Fq12` slot permutation (pairing() returns Fq12).
//!   * `ark_variant_equals_bn254_pairing` - variant 1 of the engine against `Bn254::pairing`.
use ark_bn254::{Bn254, Fq12, G1Affine, G2Affine};
use ark_ec::{pairing::Pairing, AffineRepr};
use ark_ff::PrimeField;
use ark_std::UniformRand;
use bn254_pairing_b200 as gpu;
use plonky2_bn254::fields::native::MyFq12;
use plonky2_bn254_pairing as reference;

fn hex_of(f: &MyFq12) -> Vec<String> {
    f.coeffs.iter().map(|c| format!("{:064x}", num_bigint::BigUint::from(c.into_bigint()))).collect()
}

#[test]
fn golden_vectors_match_the_reference() {
    let golden: serde_json::Value =
        serde_json::from_str(&std::fs::read_to_string("../tests/golden/survey_kat.json").unwrap()).unwrap();
    let (p, q) = (G1Affine::generator(), G2Affine::generator());
    let m = reference::miller_loop_native::miller_loop_native(&q, &p);
    let want: Vec<String> = golden["kat1_miller"].as_array().unwrap().iter().map(|v| v.as_str().unwrap().to_string()).collect();
    assert_eq!(hex_of(&m), want, "kat1_miller");
    let e = reference::final_exp_native::final_exp_native(m);
    let want: Vec<String> = golden["kat1_pairing"].as_array().unwrap().iter().map(|v| v.as_str().unwrap().to_string()).collect();
    assert_eq!(hex_of(&e), want, "kat1_pairing");
}

#[test]
fn engine_equals_reference_on_random_inputs() {
    let rng = &mut ark_std::test_rng();
    let ps: Vec<G1Affine> = (0..64).map(|_| G1Affine::rand(rng)).collect();
    let qs: Vec<G2Affine> = (0..64).map(|_| G2Affine::rand(rng)).collect();
    let got = gpu::pairing_batch(&ps, &qs);
    let ml = gpu::miller_loop_native_batch(&qs, &ps);
    for i in 0..64 {
        let want: Fq12 = reference::pairing::pairing(ps[i], qs[i]);
        assert_eq!(got[i], want, "pairing {i} (includes the MyFq12 -> Fq12 permutation)");
        assert_eq!(ml[i], reference::miller_loop_native::miller_loop_native(&qs[i], &ps[i]), "miller {i}");
    }
    let x: MyFq12 = Fq12::rand(rng).into();
    assert_eq!(gpu::final_exp_native(x), reference::final_exp_native::final_exp_native(x));
    assert_eq!(gpu::pow_native(x, vec![gpu::BN_X]), reference::final_exp_native::pow_native(x, vec![gpu::BN_X]));
    for k in [0usize, 1, 2, 3, 6, 11, 12, 25] {
        assert_eq!(gpu::frobenius_map_native(x, k), reference::final_exp_native::frobenius_map_native(x, k), "frobenius {k}");
    }
    assert_eq!(gpu::get_naf(vec![gpu::BN_X]), reference::final_exp_native::get_naf(vec![gpu::BN_X]));
    for i in 0..4 {
        assert_eq!(gpu::frob_coeffs(i), reference::final_exp_native::frob_coeffs(i));
    }
    let pairs: Vec<(&G1Affine, &G2Affine)> = ps.iter().zip(qs.iter()).take(3).collect();
    assert_eq!(gpu::multi_miller_loop_native(pairs.clone()), reference::miller_loop_native::multi_miller_loop_native(pairs));
}

#[test]
fn ark_variant_equals_bn254_pairing() {
    // variant 1 = ark-ec 0.4.2's final exponentiation; libbnp returns MyFq12 order, `.into()` is the permutation under test
    let rng = &mut ark_std::test_rng();
    let (p, q) = (G1Affine::rand(rng), G2Affine::rand(rng));
    let want = Bn254::pairing(p, q).0;
    let got = gpu::pairing_batch_variant(&[p], &[q], 1).remove(0);
    assert_eq!(got, want);
}


/// the byte formats csrc/wire.cuh decodes are ark-serialize's - pinned here against the crate itself
/// (oracle/wire_formats.py restates them without being able to run it).
#[test]
fn wire_formats_match_ark_serialize() {
    use ark_serialize::CanonicalSerialize;
    let mut rng = ark_std::test_rng();
    let n = 32;
    let ps: Vec<G1Affine> = (0..n).map(|_| G1Affine::rand(&mut rng)).collect();
    let qs: Vec<G2Affine> = (0..n).map(|_| G2Affine::rand(&mut rng)).collect();
    let (mut b1, mut b2) = (Vec::new(), Vec::new());
    for (p, q) in ps.iter().zip(qs.iter()) {
        p.serialize_compressed(&mut b1).unwrap();
        q.serialize_compressed(&mut b2).unwrap();
    }
    let got = gpu::pairing_batch_from_compressed(&b1, &b2);
    for i in 0..n {
        assert_eq!(got[i], reference::pairing::pairing(ps[i], qs[i]));
    }
    // Fq12 bytes
    let fs: Vec<_> = ps.iter().zip(qs.iter()).map(|(p, q)| reference::miller_loop_native::miller_loop_native(q, p)).collect();
    let bytes = gpu::fq12_to_bytes_batch(&fs);
    for (i, f) in fs.iter().enumerate() {
        let mut want = Vec::new();
        let a: Fq12 = f.clone().into();
        a.serialize_compressed(&mut want).unwrap();
        assert_eq!(&bytes[384 * i..384 * (i + 1)], &want[..]);
    }
}

/// scalar multiplication on the device against ark-ec's (`(P * k).into()`, the expression the reference's
/// test_to_one builds its points with, final_exp_native.rs:245-250), the identity included.
#[test]
fn scalar_mul_matches_ark_ec() {
    use ark_bn254::Fr;
    use ark_ec::CurveGroup;
    use ark_ff::{One, Zero};
    let mut rng = ark_std::test_rng();
    let n = 64;
    let ps: Vec<G1Affine> = (0..n).map(|_| G1Affine::rand(&mut rng)).collect();
    let qs: Vec<G2Affine> = (0..n).map(|_| G2Affine::rand(&mut rng)).collect();
    let mut ks: Vec<Fr> = (0..n).map(|_| Fr::rand(&mut rng)).collect();
    ks[0] = Fr::zero();
    ks[1] = Fr::one();
    ks[2] = -Fr::one();
    let g1 = gpu::g1_scalar_mul_batch(&ps, &ks);
    let g2 = gpu::g2_scalar_mul_batch(&qs, &ks);
    for i in 0..n {
        assert_eq!(g1[i], (ps[i] * ks[i]).into_affine());
        assert_eq!(g2[i], (qs[i] * ks[i]).into_affine());
    }
    assert!(g1[0].is_zero() && g2[0].is_zero());
}
