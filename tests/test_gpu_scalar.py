"""Scalar multiplication on the device (csrc/scalar.cuh): `G1Affine * Fr` / `G2Affine * Fr` of the reference's
test_to_one (`G1.mul(s).into()`, `G2.mul(t).into()`: final_exp_native.rs:245-250), against the oracle's affine double-and-add and, at size, through the pairing:
e(k P, Q) == e(P, k Q)."""
import numpy as np
import pytest

import bn254_oracle as O
from conftest import point_pool
from plonky2_bn254_pairing_b200 import api, native

pytestmark = pytest.mark.gpu

R = O.R_ORDER
EDGE = [0, 1, 2, 3, R - 1, R, R + 1, 2 * R + 5, (1 << 256) - 1, 1 << 255, 0x8000000000000000, 0xFFFFFFFFFFFFFFFF + 1]


@pytest.fixture(scope="module")
def lib(built):
    return native.init([0])


def _scalars(seed, n):
    return O.seeded_scalars(seed, n)


def test_g1_scalar_mul_against_the_oracle(lib):
    Ps, _ = point_pool(40)
    ks = EDGE + _scalars(0xB2540F01, 40 - len(EDGE))
    got = api.g1_scalar_mul_batch(Ps, ks)
    for p, k, g in zip(Ps, ks, got):
        assert g == O.g1_mul(p, k), hex(k)
    assert got[0] is None and got[EDGE.index(R)] is None and got[1] == Ps[1]


def test_g2_scalar_mul_against_the_oracle(lib):
    _, Qs = point_pool(24)
    ks = EDGE + _scalars(0xB2540F02, 24 - len(EDGE))
    got = api.g2_scalar_mul_batch(Qs, ks)
    for q, k, g in zip(Qs, ks, got):
        assert g == O.g2_mul(q, k), hex(k)
    assert got[0] is None and got[EDGE.index(R)] is None


def test_identity_in_identity_out_and_ragged_sizes(lib):
    Ps, Qs = point_pool(3)
    assert api.g1_scalar_mul_batch([None, Ps[0]], [5, 7]) == [None, O.g1_mul(Ps[0], 7)]
    assert api.g2_scalar_mul_batch([None], [5]) == [None]
    assert api.g1_scalar_mul_batch([], []) == []
    for n in (1, 129):
        ks = _scalars(0xB2540F03 + n, n)
        got = api.g1_scalar_mul_batch([Ps[1]] * n, ks)
        assert got[0] == O.g1_mul(Ps[1], ks[0]) and got[-1] == O.g1_mul(Ps[1], ks[-1])


def test_scalar_mul_commutes_with_the_pairing_at_size(lib):
    """e(k P, Q) == e(P, k Q) for 4 096 (P, Q, k): both kernels against the pairing engine, bit for bit"""
    n = 4096
    Ps, Qs = point_pool(64)
    idx = np.arange(n)
    g1 = np.ascontiguousarray(api.pack_soa(api.g1_rows(Ps))[:, :, idx % 64])
    g2 = np.ascontiguousarray(api.pack_soa(api.g2_rows(Qs))[:, :, (idx // 64 + 7 * idx) % 64])
    ks = api._scalar_rows(_scalars(0xB2540F04, n))
    kp, inf1 = api.scalar_mul_soa(1, g1, ks)
    kq, inf2 = api.scalar_mul_soa(2, g2, ks)
    assert not inf1.any() and not inf2.any()
    assert api.validate_soa(kp, kq).all()          # still on the curves and in the subgroups
    a = api.pairing_soa(kp, g2)
    b = api.pairing_soa(g1, kq)
    assert np.array_equal(a, b)
    # and one of them against the oracle end to end
    k0 = sum(int(ks[j, 17]) << (64 * j) for j in range(4))
    assert api.unpack_soa(a[:, :, 17:18])[0] == O.pairing(O.g1_mul(Ps[17], k0), Qs[(7 * 17) % 64])


def test_to_one_with_points_made_on_the_device(lib):
    """the reference's test_to_one (final_exp_native.rs:241-263) with `G1.mul(s).into()` / `G2.mul(t).into()` done by the
    device: P0 = s G1, Q0 = t G2, P1 = s t G1, Q1 = -G2  =>  e(P0, Q0) e(P1, Q1) = 1, and the multi Miller loop equals
    the product of the single ones"""
    s, t = 5, 6
    P0, P1 = api.g1_scalar_mul_batch([O.G1_GEN, O.G1_GEN], [s, s * t])
    (Q0,) = api.g2_scalar_mul_batch([O.G2_GEN], [t])
    Q1 = O.g2_neg(O.G2_GEN)
    assert (P0, P1, Q0) == (O.g1_mul(O.G1_GEN, s), O.g1_mul(O.G1_GEN, s * t), O.g2_mul(O.G2_GEN, t))
    m = api.multi_miller_loop_native([(P0, Q0), (P1, Q1)])
    m0, m1 = api.miller_loop_native(Q0, P0), api.miller_loop_native(Q1, P1)
    assert m == api.fq12_mul_batch([m0], [m1])[0]
    r_sep = api.fq12_mul_batch([api.final_exp_native(m0)], [api.final_exp_native(m1)])[0]
    assert r_sep == api.final_exp_native(m)
    one = [1] + [0] * 11
    assert r_sep == one
