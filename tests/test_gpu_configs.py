"""BASELINE.json configurations 2-5 at their FULL per-GPU sizes, as SURVEY 8(d)'s table specifies them:

    2  2^16 fused pairings            every element against the C oracle
    3  2^20 Miller loops              2^12 sampled indices against the C oracle
       2^20 final exponentiations     inputs = the GPU Miller outputs of the same indices (2^12 sampled), plus 2^10
                                      uniformly random Fq12 (final_exp_native.rs:266-286 feeds a random element too)
    4  2^18 Groth16-shaped 4-way products   2^10 sampled against the oracle, and == product of single pairings
    5  2^22 pairings                  2^12 sampled against the oracle; one checksum over the WHOLE batch

plus pow_native (final_exp_native.rs:56-84) with the reference's own test (:266-273).  Inputs come from the pool of
K = 4096 G1 x 4096 G2 subgroup points of plonky2_bn254_pairing_b200/workload.py (all index pairs distinct up to 2^24).
Everything goes through the C ABI with host buffers.  Bit-exact is the bar.  Run on the B200 box: pytest -m gpu.
"""
import random

import numpy as np
import pytest

import bn254_oracle as O
from plonky2_bn254_pairing_b200 import api, native
from plonky2_bn254_pairing_b200 import workload as wl

pytestmark = pytest.mark.gpu

K = 4096
TOP_LIMB_MAX = np.uint64(0x30644E72E131A029)


@pytest.fixture(scope="module")
def lib(built):
    return native.init([0])


def _sample(n, m, seed):
    return np.sort(np.random.RandomState(seed).choice(n, m, replace=False))


def _cols(a, idx):
    return np.ascontiguousarray(a[:, :, idx])


def _canonical(a):
    """every Fq of the array is a canonical residue (cheap necessary condition over the whole batch)"""
    return bool((a[:, 3, :] <= TOP_LIMB_MAX).all())


def test_config2_every_one_of_2e16_pairings_against_the_oracle(lib, cref):
    n = 1 << 16
    g1, g2, _ = wl.pairing_inputs(n, K=K)
    got = api.pairing_soa(g1, g2)
    want = cref.pairing(g1, g2)  # all 65 536, on every host core
    assert np.array_equal(got, want)


def test_config3_2e20_miller_loops_and_final_exponentiations(lib, cref):
    n = 1 << 20
    g1, g2, _ = wl.pairing_inputs(n, K=K, offset=1 << 16)
    ml = api.miller_loop_soa(g1, g2)
    assert _canonical(ml)
    s = _sample(n, 1 << 12, 3)
    want_ml = cref.miller(_cols(g1, s), _cols(g2, s))
    assert np.array_equal(_cols(ml, s), want_ml)
    # final exponentiation of the same 2^20 Miller outputs, with 2^10 uniformly random Fq12 spliced in
    rnd = random.Random(31)
    rows = [[rnd.randrange(O.P) for _ in range(12)] for _ in range(1 << 10)]
    r_idx = _sample(n, 1 << 10, 4)
    ml[:, :, r_idx] = api.pack_soa(rows)
    fe = api.final_exp_soa(ml)
    assert _canonical(fe)
    s2 = np.unique(np.concatenate([s, r_idx]))
    assert np.array_equal(_cols(fe, s2), cref.final_exp(_cols(ml, s2)))
    # the random inputs against the plain exponent x^((p^12-1)/r) (final_exp_native.rs:274-285), big integers
    e = (O.P ** 12 - 1) // O.R_ORDER
    out = api.unpack_soa(_cols(fe, r_idx[:3]))
    for k in range(3):
        assert out[k] == O.fq12_pow(rows[k], e)


def test_config4_2e18_groth16_shaped_products(lib, cref):
    n, k = 1 << 18, 4
    g1, g2, _ = wl.pairing_inputs(n, K=K, k=k, offset=5 << 20)
    got = api.pairing_soa(g1, g2, k=k)
    assert _canonical(got)
    s = _sample(n, 1 << 10, 5)
    assert np.array_equal(_cols(got, s), cref.pairing(_cols(g1, s), _cols(g2, s), k=k))
    # == product of the four single pairings (final_exp_native.rs:258-263), on the GPU for 2^12 elements
    s = _sample(n, 1 << 12, 6)
    acc = None
    for j in range(k):
        e = api.pairing_soa(_cols(g1[2 * j:2 * j + 2], s), _cols(g2[4 * j:4 * j + 4], s))
        acc = e if acc is None else api.fq12_mul_soa(acc, e)
    assert np.array_equal(_cols(got, s), acc)


def test_config5_2e22_pairings(lib, cref):
    import torch

    n = 1 << 22
    g1, g2, _ = wl.pairing_inputs(n, K=K, offset=9 << 20)
    got = api.pairing_soa(g1, g2)
    assert _canonical(got)
    s = _sample(n, 1 << 12, 7)
    assert np.array_equal(_cols(got, s), cref.pairing(_cols(g1, s), _cols(g2, s)))
    # checksum of checksums over the WHOLE batch: the product of all 2^22 results (tree product on the GPU) equals
    # ONE final exponentiation of the product of all Miller values (bilinearity / final_exp is a homomorphism)
    d = torch.from_numpy(got.view(np.int64)).cuda()
    d_out = torch.zeros((12, 4, 1), dtype=torch.int64, device="cuda")
    native.check(lib.bnp_fq12_product_dev(0, None, d.data_ptr(), d_out.data_ptr(), n))
    torch.cuda.synchronize()
    prod = d_out.cpu().numpy().view(np.uint64)
    assert np.array_equal(prod, api.pairing_product_soa(g1, g2))


def test_pow_native_like_the_reference_test(lib, cref):
    """final_exp_native.rs:266-273: a uniformly random (non-cyclotomic) x, pow_native(x, [BN_X]) == x^BN_X; then other
    exponents through the run-time NAF walk: one limb, several limbs, zero, one, a -1 digit in the top position."""
    rnd = random.Random(12)
    rows = [[rnd.randrange(O.P) for _ in range(12)] for _ in range(40)]
    f = api.pack_soa(rows)
    got = api.pow_soa(f, [O.BN_X])
    assert np.array_equal(got, cref.pow_u64(f, O.BN_X))
    assert api.unpack_soa(_cols(got, [0]))[0] == O.pow_native(rows[0], [O.BN_X]) == O.fq12_pow(rows[0], O.BN_X)
    for e in (0, 1, 2, 3, 0xFFFFFFFFFFFFFFFF, 0xB2540000DEADBEEF):
        assert np.array_equal(api.pow_soa(f, [e]), cref.pow_u64(f, e)), hex(e)
    # multi-limb exponents against big integers
    for limbs in ([5, 1], [0xFFFFFFFFFFFFFFFF, 0x7FFFFFFFFFFFFFFF], [0, 0, 3]):
        e = sum(v << (64 * i) for i, v in enumerate(limbs))
        out = api.unpack_soa(api.pow_soa(_cols(f, [1, 2]), limbs))
        assert out[0] == O.fq12_pow(rows[1], e) and out[1] == O.fq12_pow(rows[2], e)
    assert api.pow_native(rows[3], [O.BN_X]) == O.pow_native(rows[3], [O.BN_X])
    # the host-side table generators of the reference's public API
    assert api.get_naf([O.BN_X]) == O.get_naf([O.BN_X])
    assert [api.frob_coeffs(i) for i in range(4)] == [O.frob_coeffs(i) for i in range(4)]
    assert api.SIX_U_PLUS_2_NAF == O.SIX_U_PLUS_2_NAF and api.BN_X == O.BN_X


# ----------------------------------------------------------------------------- on-device input validation
def _fq2_sqrt(a):
    """Square root in Fq2 for p = 3 mod 4 (Adj-Rodriguez-Henriquez, eprint 2012/685, algorithm 9), None if a non-residue."""
    mul, P = O.fq2_mul, O.P

    def pw(x, e):
        r = (1, 0)
        while e:
            if e & 1:
                r = mul(r, x)
            x = mul(x, x)
            e >>= 1
        return r

    a1 = pw(a, (P - 3) // 4)
    alpha = mul(a1, mul(a1, a))
    if mul(pw(alpha, P), alpha) == (P - 1, 0):
        return None
    x0 = mul(a1, a)
    if alpha == (P - 1, 0):
        return mul((0, 1), x0)
    return mul(pw(((alpha[0] + 1) % P, alpha[1]), (P - 1) // 2), x0)


def test_validate_batch_flags_what_g2affine_new_rejects(lib):
    """`G2Affine::new` (behind miller_loop_native.rs:303,311) asserts on-curve AND in the r-torsion subgroup."""
    rnd = random.Random(77)
    Ps, Qs = wl.point_pool(24)
    want = [True] * 24
    Ps, Qs = list(Ps), list(Qs)
    # off-curve G1 / G2, the (0, 0) encoding of the identity, and points ON the twist but OUTSIDE the subgroup
    Ps[1] = (Ps[1][0], (Ps[1][1] + 1) % O.P); want[1] = False
    Qs[2] = (Qs[2][0], ((Qs[2][1][0] + 1) % O.P, Qs[2][1][1])); want[2] = False
    Ps[3] = (0, 0); want[3] = False
    Qs[4] = ((0, 0), (0, 0)); want[4] = False
    b_twist = O.fq2_mul((3, 0), O.fq2_inv((9, 1)))
    k = 5
    while k < 12:
        x = (rnd.randrange(O.P), rnd.randrange(O.P))
        rhs = O.fq2_mul(O.fq2_mul(x, x), x)
        rhs = ((rhs[0] + b_twist[0]) % O.P, (rhs[1] + b_twist[1]) % O.P)
        y = _fq2_sqrt(rhs)
        if y is None or O.fq2_mul(y, y) != rhs:
            continue
        Qs[k] = (x, y)          # on the twist; in the subgroup with probability 1 / (2p - r): never
        want[k] = False
        k += 1
    assert api.validate_batch(Ps, Qs) == want
    assert api.validate_batch(Ps, None) == [i not in (1, 3) for i in range(24)]
    assert api.validate_batch(None, Qs) == [i not in (2, 4, 5, 6, 7, 8, 9, 10, 11) for i in range(24)]
    # a full-size batch of valid points: all ones
    g1, g2, _ = wl.pairing_inputs(1 << 16, K=K)
    assert api.validate_soa(g1, g2).all()
