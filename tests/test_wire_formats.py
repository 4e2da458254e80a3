"""Wire formats, CPU side: the restatement in oracle/wire_formats.py against the published constants
and against itself (round trips, every status), and the EIP-197 precompile semantics on the oracle's pairing."""
import random

import bn254_oracle as O
import wire_formats as W

PTS = O.seeded_points(0xB2540F30, 6)


def off_subgroup_twist_point():
    """A point on the twist that is not in the r-torsion (the cofactor is ~2^254: any point found by solving the curve
    equation is outside with overwhelming probability)."""
    x0 = 1
    while True:
        x = (x0, 0)
        y = W.f2_sqrt(W.f2_add(W.f2_mul(W.f2_mul(x, x), x), W.B2))
        if y is not None and not W.in_subgroup_g2((x, y)):
            return (x, y)
        x0 += 1


def test_published_generators_and_eip_layout():
    # EIP-197's generators; the G2 constants are the ones every BN254 library carries
    assert O.G1_GEN == (1, 2)
    g2 = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))
    assert O.G2_GEN == g2
    assert W.on_curve_g2(g2) and W.in_subgroup_g2(g2)
    b = W.encode_g1(O.G1_GEN, W.EIP197)
    assert b == (1).to_bytes(32, "big") + (2).to_bytes(32, "big")
    b = W.encode_g2(g2, W.EIP197)
    # imaginary part first (EIP-197: "a * i + b is encoded as a, b")
    assert int.from_bytes(b[:32], "big") == g2[0][1] and int.from_bytes(b[32:64], "big") == g2[0][0]
    assert int.from_bytes(b[64:96], "big") == g2[1][1] and int.from_bytes(b[96:], "big") == g2[1][0]
    # the twist constant 3 / (9 + u)
    assert W.f2_mul(W.B2, (9, 1)) == (3, 0)


def test_round_trips_every_format():
    for fmt in (W.UNCOMPRESSED, W.COMPRESSED, W.EIP197):
        for p, q in PTS:
            for pt in (p, (p[0], (-p[1]) % O.P)):       # both signs of y
                assert W.decode_g1(W.encode_g1(pt, fmt), fmt) == (W.OK, pt)
            for qt in (q, (q[0], W.f2_neg(q[1]))):
                assert W.decode_g2(W.encode_g2(qt, fmt), fmt) == (W.OK, qt)
        assert W.decode_g1(W.encode_g1(None, fmt), fmt) == (W.INFINITY, None)
        assert W.decode_g2(W.encode_g2(None, fmt), fmt) == (W.INFINITY, None)
    sizes = {W.UNCOMPRESSED: (64, 128), W.COMPRESSED: (32, 64), W.EIP197: (64, 128)}
    for fmt, (s1, s2) in sizes.items():
        assert len(W.encode_g1(PTS[0][0], fmt)) == s1 and len(W.encode_g2(PTS[0][1], fmt)) == s2


def test_sign_flag_is_the_larger_root():
    p, q = PTS[0]
    b = W.encode_g1(p, W.COMPRESSED)
    assert bool(b[-1] & 0x80) == (p[1] > O.P - p[1])
    b = W.encode_g2(q, W.COMPRESSED)
    y = q[1]
    larger = (y[1] > O.P - y[1]) if y[1] else (y[0] > O.P - y[0])
    assert bool(b[-1] & 0x80) == larger


def test_malformed_inputs():
    p, q = PTS[1]
    # coordinate >= p
    bad = (O.P).to_bytes(32, "big") + (2).to_bytes(32, "big")
    assert W.decode_g1(bad, W.EIP197)[0] == W.NOT_CANONICAL
    bad = (O.P + 1).to_bytes(32, "little") + bytes(32)
    assert W.decode_g1(bad, W.UNCOMPRESSED)[0] == W.NOT_CANONICAL
    # both flags set
    b = bytearray(W.encode_g1(p, W.COMPRESSED))
    b[-1] |= 0xC0
    assert W.decode_g1(bytes(b), W.COMPRESSED)[0] == W.NOT_CANONICAL
    # off the curve
    assert W.decode_g1(W.encode_g1((p[0], (p[1] + 1) % O.P), W.UNCOMPRESSED), W.UNCOMPRESSED)[0] == W.NOT_ON_CURVE
    assert W.decode_g2(W.encode_g2((q[0], W.f2_add(q[1], (1, 0))), W.EIP197), W.EIP197)[0] == W.NOT_ON_CURVE
    # x with no point above it (compressed)
    x = 1
    while W.fp_sqrt((x * x * x + 3) % O.P) is not None:
        x += 1
    assert W.decode_g1(x.to_bytes(32, "little"), W.COMPRESSED)[0] == W.NOT_ON_CURVE
    # on the twist, outside the subgroup
    t = off_subgroup_twist_point()
    for fmt in (W.UNCOMPRESSED, W.COMPRESSED, W.EIP197):
        assert W.decode_g2(W.encode_g2(t, fmt), fmt)[0] == W.NOT_IN_SUBGROUP
        assert W.decode_g2(W.encode_g2(t, fmt), fmt, check_subgroup=False) == (W.OK, t)


def test_fq2_sqrt_on_random_squares():
    rnd = random.Random(7)
    for _ in range(20):
        a = (rnd.randrange(O.P), rnd.randrange(O.P))
        s = W.f2_mul(a, a)
        r = W.f2_sqrt(s)
        assert r is not None and W.f2_mul(r, r) == s
    assert W.f2_sqrt((4, 0)) in ((2, 0), (O.P - 2, 0))
    assert W.f2_mul(W.f2_sqrt((O.P - 4, 0)), W.f2_sqrt((O.P - 4, 0))) == (O.P - 4, 0)   # a non-residue of Fq: root is imaginary


def test_fq12_bytes_follow_the_ark_tower_order():
    from plonky2_bn254_pairing_b200 import api
    rnd = random.Random(3)
    c = [rnd.randrange(O.P) for _ in range(12)]
    b = W.encode_fq12(c)
    assert len(b) == 384
    ark = api.myfq12_to_ark(c)      # [c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2], each (re, im)
    flat = [v for pair in ark for v in pair]
    assert [int.from_bytes(b[32 * k:32 * k + 32], "little") for k in range(12)] == flat
    assert W.decode_fq12(b) == (W.OK, c)
    assert W.decode_fq12((O.P).to_bytes(32, "little") + b[32:])[0] == W.NOT_CANONICAL


def test_eip197_semantics_on_the_oracle():
    p, q = PTS[2]
    neg_p = (p[0], (-p[1]) % O.P)
    pair = lambda a, b: W.encode_g1(a, W.EIP197) + W.encode_g2(b, W.EIP197)
    assert W.eip197_pairing_check(b"") is True
    assert W.eip197_pairing_check(pair(p, q) + pair(neg_p, q)) is True          # e(P,Q) e(-P,Q) = 1
    assert W.eip197_pairing_check(pair(p, q) + pair(p, q)) is False
    assert W.eip197_pairing_check(pair(None, q) + pair(p, None)) is True        # infinity pairs drop out
    assert W.eip197_pairing_check(pair(p, q) + pair(None, q) + pair(neg_p, q)) is True
    assert W.eip197_pairing_check(pair(p, off_subgroup_twist_point())) is None  # precompile failure
