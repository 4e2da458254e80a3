"""Synthetic inputs for bench.py and the large parity tests.

Generating 2^16..2^22 random G2 points costs more CPU than the pairings themselves, so inputs are
drawn from a pool: K G1 and K G2 points (non-identity, in the r-torsion subgroups like `rand()` in the
reference tests) made once by an additive walk  P_{j+1} = P_j + D,  Q_{j+1} = Q_j + E  from fixed
multiples of the generators; pairing i uses (P[i mod K], Q[(i div K + 7 i) mod K]) - all index pairs
are distinct up to K^2.  Kernel time is data-independent (the schedule is a compile-time constant).

This module has its own few lines of affine curve arithmetic so that the product/bench path never
touches oracle/.
"""
import numpy as np

from .api import P, pack_soa

G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


def _f2mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def _f2sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def _f2inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % P, P - 2, P)
    return (a[0] * n % P, (-a[1]) * n % P)


def _g1_add(a, b):
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        lam = 3 * x1 * x1 * pow(2 * y1 % P, P - 2, P) % P
    else:
        lam = (y2 - y1) * pow((x2 - x1) % P, P - 2, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def _g2_add(a, b):
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        lam = _f2mul(_f2mul((3, 0), _f2mul(x1, x1)), _f2inv(((2 * y1[0]) % P, (2 * y1[1]) % P)))
    else:
        lam = _f2mul(_f2sub(y2, y1), _f2inv(_f2sub(x2, x1)))
    x3 = _f2sub(_f2sub(_f2mul(lam, lam), x1), x2)
    return (x3, _f2sub(_f2mul(lam, _f2sub(x1, x3)), y1))


def _mul(add, pt, k):
    acc = None
    for bit in bin(k)[2:]:
        if acc is not None:
            acc = add(acc, acc)
        if bit == "1":
            acc = pt if acc is None else add(acc, pt)
    return acc


def point_pool(K=256, seed=0xB2540000):
    """K G1 points and K G2 points (plain affine integer coordinates)."""
    a0, a1 = 0x1234567 + seed, 0x89ABCDE + 3 * seed
    p = _mul(_g1_add, G1_GEN, a0)
    d = _mul(_g1_add, G1_GEN, a1)
    q = _mul(_g2_add, G2_GEN, a1 + 11)
    e = _mul(_g2_add, G2_GEN, a0 + 5)
    Ps, Qs = [p], [q]
    for _ in range(K - 1):
        Ps.append(_g1_add(Ps[-1], d))
        Qs.append(_g2_add(Qs[-1], e))
    return Ps, Qs


def pool_indices(n, K, offset=0):
    i = np.arange(offset, offset + n, dtype=np.int64)
    return i % K, (i // K + 7 * i) % K


def pairing_inputs(n, K=256, offset=0, k=1, seed=0xB2540000):
    """SoA Montgomery inputs for n elements of k pairs each: g1 [2k][4][n], g2 [4k][4][n] (uint64),
    plus the pool and index arrays so that callers can look up the plain points of any element."""
    Ps, Qs = point_pool(K, seed)
    g1p = pack_soa([[x, y] for (x, y) in Ps])
    g2p = pack_soa([[q[0][0], q[0][1], q[1][0], q[1][1]] for q in Qs])
    g1s, g2s, idx = [], [], []
    for j in range(k):
        i1, i2 = pool_indices(n, K, offset + j * 1000003)
        g1s.append(g1p[:, :, i1])
        g2s.append(g2p[:, :, i2])
        idx.append((i1, i2))
    g1 = np.ascontiguousarray(np.concatenate(g1s, axis=0))
    g2 = np.ascontiguousarray(np.concatenate(g2s, axis=0))
    return g1, g2, (Ps, Qs, idx)
