"""Wire formats decoded / encoded on the device against oracle/wire_formats.py: every format, both signs
of the compressed forms, infinity, every failure status, Fq12 bytes, and the EIP-197 precompile end to end."""
import random

import numpy as np
import pytest

import bn254_oracle as O
import wire_formats as W
from conftest import point_pool
from plonky2_bn254_pairing_b200 import api, native
from test_wire_formats import off_subgroup_twist_point

pytestmark = pytest.mark.gpu
FORMATS = [W.UNCOMPRESSED, W.COMPRESSED, W.EIP197]


@pytest.fixture(scope="module")
def lib(built):
    return native.init([0])


def _neg1(p):
    return (p[0], (-p[1]) % O.P)


def _neg2(q):
    return (q[0], W.f2_neg(q[1]))


def _check_points(group, fmt, blobs, subgroup=True):
    """decode on the device, compare status and coordinates with the oracle's decoder element by element"""
    data = b"".join(blobs)
    if group == 1:
        soa, st = api.decode_g1_soa(fmt, data)
        want = [W.decode_g1(b, fmt) for b in blobs]
        rows = [[pt[0], pt[1]] if pt else [0, 0] for _, pt in want]
    else:
        soa, st = api.decode_g2_soa(fmt, data, check_subgroup=subgroup)
        want = [W.decode_g2(b, fmt, check_subgroup=subgroup) for b in blobs]
        rows = [[pt[0][0], pt[0][1], pt[1][0], pt[1][1]] if pt else [0, 0, 0, 0] for _, pt in want]
    assert list(st) == [s for s, _ in want]
    assert np.array_equal(soa, api.pack_soa(rows))
    return st


@pytest.mark.parametrize("fmt", FORMATS)
def test_decode_valid_points(lib, fmt):
    Ps, Qs = point_pool(40)
    g1 = [W.encode_g1(p, fmt) for p in Ps] + [W.encode_g1(_neg1(p), fmt) for p in Ps[:8]] + [W.encode_g1(None, fmt)]
    g2 = [W.encode_g2(q, fmt) for q in Qs] + [W.encode_g2(_neg2(q), fmt) for q in Qs[:8]] + [W.encode_g2(None, fmt)]
    st1 = _check_points(1, fmt, g1)
    st2 = _check_points(2, fmt, g2)
    assert list(st1[:-1]) == [W.OK] * 48 and st1[-1] == W.INFINITY
    assert list(st2[:-1]) == [W.OK] * 48 and st2[-1] == W.INFINITY


@pytest.mark.parametrize("fmt", FORMATS)
def test_decode_malformed_points(lib, fmt):
    Ps, Qs = point_pool(4)
    be = fmt == W.EIP197
    word = (lambda v: v.to_bytes(32, "big")) if be else (lambda v: v.to_bytes(32, "little"))
    off = off_subgroup_twist_point()
    g1 = [W.encode_g1(Ps[0], fmt)]
    g2 = [W.encode_g2(Qs[0], fmt), W.encode_g2(off, fmt)]
    if fmt == W.COMPRESSED:
        x = 1
        while W.fp_sqrt((x * x * x + 3) % O.P) is not None:
            x += 1
        g1 += [word(x), word(O.P), bytes(31) + b"\xc0"]
        g2 += [word(O.P) + word(1), word(1) + word(O.P)[:-1] + bytes([word(O.P)[-1] | 0x00])]
    else:
        g1 += [word(O.P) + word(2), word(1) + word(O.P + 5), W.encode_g1((Ps[1][0], (Ps[1][1] + 1) % O.P), fmt)]
        bad_y = (Qs[1][0], W.f2_add(Qs[1][1], (0, 1)))
        g2 += [W.encode_g2(bad_y, fmt), word(O.P) + W.encode_g2(Qs[2], fmt)[32:]]
        if not be:
            both = bytearray(W.encode_g1(Ps[2], fmt))
            both[-1] |= 0xC0
            g1.append(bytes(both))
    st1 = _check_points(1, fmt, g1)
    st2 = _check_points(2, fmt, g2)
    assert st1[0] == W.OK and all(s >= W.NOT_CANONICAL for s in st1[1:])
    assert st2[0] == W.OK and st2[1] == W.NOT_IN_SUBGROUP and all(s >= W.NOT_CANONICAL for s in st2[2:])
    # without the subgroup test the off-subgroup point decodes
    st = _check_points(2, fmt, g2[:2], subgroup=False)
    assert list(st) == [W.OK, W.OK]


def test_decoded_points_feed_the_pairing(lib):
    """bytes in, bytes out: compressed ark points -> device decode -> fused pairing -> ark Fq12 bytes"""
    Ps, Qs = point_pool(16)
    g1, s1 = api.decode_g1_soa(W.COMPRESSED, b"".join(W.encode_g1(p, W.COMPRESSED) for p in Ps))
    g2, s2 = api.decode_g2_soa(W.COMPRESSED, b"".join(W.encode_g2(q, W.COMPRESSED) for q in Qs))
    assert not s1.any() and not s2.any()
    out = api.pairing_soa(g1, g2)
    blob = api.encode_fq12_soa(out)
    vals = api.unpack_soa(out)
    for i in range(16):
        assert vals[i] == O.pairing(Ps[i], Qs[i]) if i < 2 else True      # two against the Python oracle (slow)
        assert blob[384 * i:384 * (i + 1)] == W.encode_fq12(vals[i])
    back, st = api.decode_fq12_soa(blob)
    assert not st.any() and np.array_equal(back, out)
    bad = bytearray(blob[:768])
    bad[384:416] = (O.P + 3).to_bytes(32, "little")
    back, st = api.decode_fq12_soa(bytes(bad))
    assert list(st) == [W.OK, W.NOT_CANONICAL]
    assert np.array_equal(back[:, :, 0], out[:, :, 0]) and not back[:, :, 1].any()


def test_eip197_precompile(lib):
    Ps, Qs = point_pool(6)
    pair = lambda a, b: W.encode_g1(a, W.EIP197) + W.encode_g2(b, W.EIP197)
    cases = [
        b"",
        pair(Ps[0], Qs[0]) + pair(_neg1(Ps[0]), Qs[0]),
        pair(Ps[0], Qs[0]) + pair(Ps[0], Qs[0]),
        pair(None, Qs[1]) + pair(Ps[1], None),
        pair(Ps[2], Qs[2]) + pair(None, Qs[3]) + pair(_neg1(Ps[2]), Qs[2]),
        pair(Ps[3], Qs[3]),
    ]
    # bilinearity with scalars: e(aP, bQ) e(-abP, Q) = 1
    a, b = 0x1234567, 0x89abcdef
    cases.append(pair(O.g1_mul(O.G1_GEN, a), O.g2_mul(O.G2_GEN, b)) +
                 pair(_neg1(O.g1_mul(O.G1_GEN, a * b % W.R_ORDER)), O.G2_GEN))
    for data in cases:
        assert api.eip197_pairing_check(data) == W.eip197_pairing_check(data), len(data)
    assert api.eip197_pairing_check(cases[1]) is True and api.eip197_pairing_check(cases[2]) is False
    for bad in (pair(Ps[0], off_subgroup_twist_point()),
                (O.P).to_bytes(32, "big") + bytes(32) + W.encode_g2(Qs[0], W.EIP197),
                pair((Ps[0][0], (Ps[0][1] + 1) % O.P), Qs[0])):
        assert W.eip197_pairing_check(bad) is None
        with pytest.raises(native.BnpError):
            api.eip197_pairing_check(bad)
    with pytest.raises(native.BnpError):
        api.eip197_pairing_check(bytes(100))


@pytest.mark.parametrize("n", [1, 33, 1000])
def test_compressed_decode_at_size(lib, n):
    """ragged and larger batches of compressed points (every point needs its square root): coordinates against the
    oracle's decoder, then the decoded batch through the fused pairing against the C oracle"""
    Ps, Qs = point_pool(64)
    idx = np.arange(n)
    ps = [Ps[i % 64] if i % 3 else _neg1(Ps[i % 64]) for i in idx]
    qs = [Qs[(5 * i + 1) % 64] if i % 2 else _neg2(Qs[(5 * i + 1) % 64]) for i in idx]
    g1, s1 = api.decode_g1_soa(W.COMPRESSED, b"".join(W.encode_g1(p, W.COMPRESSED) for p in ps))
    g2, s2 = api.decode_g2_soa(W.COMPRESSED, b"".join(W.encode_g2(q, W.COMPRESSED) for q in qs), check_subgroup=(n <= 33))
    assert not s1.any() and not s2.any()
    assert np.array_equal(g1, api.pack_soa(api.g1_rows(ps))) and np.array_equal(g2, api.pack_soa(api.g2_rows(qs)))
