"""The C oracle (oracle/bn254_ref.c, the CPU baseline) against the Python oracle and the golden vectors."""
import random

import numpy as np

import bn254_oracle as O
from plonky2_bn254_pairing_b200 import api


def ints(hexes):
    return [int(h, 16) for h in hexes]


def test_c_oracle_matches_golden(cref, golden):
    cases = golden["oracle"]["cases"]
    g1 = api.pack_soa([ints(c["g1"]) for c in cases])
    g2 = api.pack_soa([ints(c["g2"]) for c in cases])
    for faithful in (0, 1):
        assert api.unpack_soa(cref.miller(g1, g2, faithful=faithful)) == [ints(c["miller"]) for c in cases]
        assert api.unpack_soa(cref.pairing(g1, g2, faithful=faithful, threads=2)) == [ints(c["pairing"]) for c in cases]
    fe = cref.final_exp(api.pack_soa([ints(c["miller"]) for c in cases]), faithful=1)
    assert api.unpack_soa(fe) == [ints(c["pairing"]) for c in cases]


def test_c_oracle_survey_kat(cref, golden):
    g1 = api.pack_soa([[O.G1_GEN[0], O.G1_GEN[1]]])
    g2 = api.pack_soa(api.g2_rows([O.G2_GEN]))
    assert api.unpack_soa(cref.miller(g1, g2))[0] == ints(golden["survey"]["kat1_miller"])
    assert api.unpack_soa(cref.pairing(g1, g2, faithful=1))[0] == ints(golden["survey"]["kat1_pairing"])


def test_c_oracle_multi_and_random(cref, golden):
    cases = golden["oracle"]["cases"]
    g1 = api.pack_soa([[v for c in cases[0:3] for v in ints(c["g1"])]])
    g2 = api.pack_soa([[v for c in cases[0:3] for v in ints(c["g2"])]])
    assert api.unpack_soa(cref.miller(g1, g2, k=3))[0] == ints(golden["oracle"]["multi3"]["miller"])
    assert api.unpack_soa(cref.pairing(g1, g2, k=3))[0] == ints(golden["oracle"]["multi3"]["pairing"])
    r = golden["oracle"]["random_fq12"]
    x = api.pack_soa([ints(r["x"])])
    assert api.unpack_soa(cref.final_exp(x, faithful=1))[0] == ints(r["final_exp"])
    assert api.unpack_soa(cref.pow_u64(x, O.BN_X))[0] == ints(r["pow_x"])
    for k, v in r["frobenius"].items():
        for faithful in (0, 1):
            assert api.unpack_soa(cref.frobenius(x, int(k), faithful=faithful))[0] == ints(v)


def test_c_oracle_field_ops_random(cref):
    rnd = random.Random(3)
    rows_a = [[rnd.randrange(O.P) for _ in range(12)] for _ in range(9)]
    rows_b = [[rnd.randrange(O.P) for _ in range(12)] for _ in range(9)]
    rows_a[0] = [0] * 12
    rows_a[1] = [O.P - 1] * 12
    got = api.unpack_soa(cref.fq12_mul(api.pack_soa(rows_a), api.pack_soa(rows_b), threads=3))
    assert got == [O.fq12_mul(a, b) for a, b in zip(rows_a, rows_b)]


def test_c_oracle_threads_agree(cref):
    """The pthread work splitting must not change results (ragged n, more threads than work)."""
    pts = O.seeded_points(21, 3)
    idx = np.arange(11) % 3
    g1 = np.ascontiguousarray(api.pack_soa(api.g1_rows([p for p, _ in pts]))[:, :, idx])
    g2 = np.ascontiguousarray(api.pack_soa(api.g2_rows([q for _, q in pts]))[:, :, idx])
    a = cref.pairing(g1, g2, threads=1)
    b = cref.pairing(g1, g2, threads=16)
    assert np.array_equal(a, b)
    assert api.unpack_soa(a[:, :, :3].copy()) == [O.pairing(p, q) for p, q in pts]
