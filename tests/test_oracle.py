"""The Python oracle against (a) the survey's independent known-answer vectors and (b) every relation the
reference's own tests assert.  CPU only."""
import random

import bn254_oracle as O


def H(xs):
    return ["%064x" % x for x in xs]


def test_survey_kat1(golden):
    m = O.miller_loop_native(O.G2_GEN, O.G1_GEN)
    assert H(m) == golden["survey"]["kat1_miller"]
    e = O.final_exp_native(m)
    assert H(e) == golden["survey"]["kat1_pairing"]
    assert H(O.final_exp_ark(m)) == golden["survey"]["kat1_pairing_ark"]
    assert O.pairing(O.G1_GEN, O.G2_GEN) == e


def test_survey_kat2(golden):
    p5, q6 = O.g1_mul(O.G1_GEN, 5), O.g2_mul(O.G2_GEN, 6)
    assert H(O.miller_loop_native(q6, p5)) == golden["survey"]["kat2_miller"]
    e2 = O.pairing(p5, q6)
    assert H(e2) == golden["survey"]["kat2_pairing"]
    e1 = [int(x, 16) for x in golden["survey"]["kat1_pairing"]]
    assert e2 == O.fq12_pow(e1, 30)  # bilinearity


def test_oracle_vectors_are_current(golden):
    """tests/golden/oracle_vectors.json is what make_golden.py would write today."""
    pts = O.seeded_points(golden["oracle"]["seed"], 2)
    for (p, q), case in zip(pts, golden["oracle"]["cases"]):
        assert H([p[0], p[1]]) == case["g1"]
        assert H(O.miller_loop_native(q, p)) == case["miller"]
        assert H(O.pairing(p, q)) == case["pairing"]


def test_multi_miller_loop_native():
    """miller_loop_native.rs:336-348"""
    (p0, q0), (p1, q1) = O.seeded_points(11, 2)
    r_expected = O.fq12_mul(O.miller_loop_native(q0, p0), O.miller_loop_native(q1, p1))
    assert O.multi_miller_loop_native([(p0, q0), (p1, q1)]) == r_expected


def test_to_one():
    """final_exp_native.rs:240-264 (plus the == 1 the reference never asserts)"""
    s, t = 5, 6
    p0, q0 = O.g1_mul(O.G1_GEN, s), O.g2_mul(O.G2_GEN, t)
    p1, q1 = O.g1_mul(O.G1_GEN, s * t), O.g2_neg(O.G2_GEN)
    m = O.multi_miller_loop_native([(p0, q0), (p1, q1)])
    m0, m1 = O.miller_loop_native(q0, p0), O.miller_loop_native(q1, p1)
    assert m == O.fq12_mul(m0, m1)
    r_sep = O.fq12_mul(O.final_exp_native(m0), O.final_exp_native(m1))
    r_mul = O.final_exp_native(m)
    assert r_sep == r_mul
    assert r_mul == O.FQ12_ONE


def test_pow():
    """final_exp_native.rs:266-286: pow_native == generic pow on a NON-cyclotomic element, and
    final_exp_native(x) == x^((p^12-1)/r) exactly."""
    rnd = random.Random(5)
    x = [rnd.randrange(O.P) for _ in range(12)]
    assert O.pow_native(x, [O.BN_X]) == O.fq12_pow(x, O.BN_X)
    exp = (O.P ** 12 - 1) // O.R_ORDER
    assert O.final_exp_native(x) == O.fq12_pow(x, exp)


def test_ark_variant_is_lambda_power():
    """ark's final exponentiation is the reference's raised to 2x(6x^2+3x+1)."""
    p, q = O.seeded_points(12, 1)[0]
    m = O.miller_loop_native(q, p)
    assert O.final_exp_ark(m) == O.fq12_pow(O.final_exp_native(m), O.ARK_LAMBDA)


def test_pairing_order_and_bilinearity():
    p, q = O.seeded_points(13, 1)[0]
    e = O.pairing(p, q)
    assert e != O.FQ12_ONE
    assert O.fq12_pow(e, O.R_ORDER) == O.FQ12_ONE
    assert O.pairing(O.g1_mul(p, 7), q) == O.fq12_pow(e, 7)
    assert O.pairing(p, O.g2_mul(q, 3)) == O.fq12_pow(e, 3)


def test_get_naf():
    """final_exp_native.rs:86-128: digits in {-1,0,1}, non-adjacent, sum d_i 2^i == value (incl. limb carries)."""
    for exp in ([O.BN_X], [2 ** 64 - 1], [2 ** 64 - 1, 2 ** 64 - 1], [3, 0, 7], [0xFFFFFFFFFFFFFFFF, 5]):
        naf = O.get_naf(exp)
        val = sum(v << (64 * i) for i, v in enumerate(exp))
        assert sum(d << i for i, d in enumerate(naf)) == val
        assert all(d in (-1, 0, 1) for d in naf)
        assert all(not (naf[i] and naf[i + 1]) for i in range(len(naf) - 1))
    assert sum(d << i for i, d in enumerate(O.SIX_U_PLUS_2_NAF)) == 6 * O.BN_X + 2


def test_published_constants():
    """c^2, c^3 are ark's TWIST_MUL_BY_Q_X / _Y; gamma_2^1 is in Fq."""
    c = O._expected_c()
    c2 = O.fq2_mul(c, c)
    c3 = O.fq2_mul(c2, c)
    assert c2 == (0x2fb347984f7911f74c0bec3cf559b143b78cc310c2c3330c99e39557176f553d,
                  0x16c9e55061ebae204ba4cc8bd75a079432ae2a1d0b7c9dce1665d51c640fcba2)
    assert c3 == (0x063cf305489af5dcdc5ec698b6e2f9b9dbaae0eda9c95998dc54014671a0135a,
                  0x07c03cbcac41049a0704b5a7ec796f2b21807dc98fa25bd282d37f632623b0e3)
    assert O.frob_coeffs(2) == (0x30644e72e131a0295e6dd9e7e0acccb0c28f069fbb966e3de4bd44e5607cfd49, 0)
    assert O.g2_on_curve(O.G2_GEN) and O.g1_on_curve(O.G1_GEN)
    assert O.g2_mul(O.G2_GEN, O.R_ORDER) is None


def test_frobenius_is_p_power():
    rnd = random.Random(9)
    x = [rnd.randrange(O.P) for _ in range(12)]
    assert O.frobenius_map_native(x, 1) == O.fq12_pow(x, O.P)
    assert O.frobenius_map_native(x, 14) == O.frobenius_map_native(x, 2)
    assert O.frobenius_map_native(O.frobenius_map_native(x, 5), 7) == x


def test_fq12_inverse_and_ark_layout():
    rnd = random.Random(10)
    x = [rnd.randrange(O.P) for _ in range(12)]
    assert O.fq12_mul(x, O.fq12_inv(x)) == O.FQ12_ONE
    assert O.ark_to_myfq12(O.myfq12_to_ark(x)) == x


def test_montgomery_limbs_roundtrip():
    for v in (0, 1, O.P - 1, 0x1234567890abcdef << 100):
        assert O.from_mont_limbs(O.to_mont_limbs(v)) == v
    assert O.to_mont_limbs(1) == [0xd35d438dc58f0d9d, 0x0a78eb28f5c70b3d, 0x666ea36f7879462c, 0x0e0a77c19a07df2f]
