"""CPU restatement of the wire formats csrc/wire.cuh decodes (test infrastructure only, like the rest of oracle/).

ark-serialize 0.4 `CanonicalSerialize` for short-Weierstrass affine points (ark-ec 0.4.2
models/short_weierstrass/mod.rs::serialize_with_mode / deserialize_with_mode, serialization_flags.rs::SWFlags) and for
Fq12, and the EIP-196 / EIP-197 encodings of the Ethereum precompiles.  **Parity unpinned** for the ark formats: the
crate is a dependency that is not vendored under /root/reference, so this file restates its published behaviour; the
EIP encodings follow the EIPs' text.  rust/tests/parity.rs is where a maintainer with cargo pins them.
"""
import bn254_oracle as O

P = O.P
R_ORDER = 21888242871839275222246405745257275088548364400416034343698204186575808495617
B1 = 3
B2 = O.fq2_mul((3, 0), O.fq2_inv((9, 1))) if hasattr(O, "fq2_inv") else None

UNCOMPRESSED, COMPRESSED, EIP197 = 0, 1, 2
OK, INFINITY, NOT_CANONICAL, NOT_ON_CURVE, NOT_IN_SUBGROUP = 0, 1, 2, 3, 4
FLAG_NEG, FLAG_INF = 0x80, 0x40


# ------------------------------------------------------------------ field helpers (plain integers / pairs)
def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_neg(a):
    return ((-a[0]) % P, (-a[1]) % P)


def f2_inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % P, P - 2, P)
    return (a[0] * n % P, (-a[1]) * n % P)


if B2 is None:
    B2 = f2_mul((3, 0), f2_inv((9, 1)))


def fp_sqrt(a):
    """A square root of a in Fq or None (p = 3 mod 4)."""
    c = pow(a, (P + 1) // 4, P)
    return c if c * c % P == a % P else None


def f2_sqrt(a):
    """A square root of a in Fq2 = Fq[u]/(u^2+1) or None (complex method, as csrc/wire.cuh::w_fp2_sqrt)."""
    a0, a1 = a[0] % P, a[1] % P
    if a1 == 0:
        s = fp_sqrt(a0)
        if s is not None:
            return (s, 0)
        s = fp_sqrt((-a0) % P)
        return None if s is None else (0, s)
    s = fp_sqrt((a0 * a0 + a1 * a1) % P)
    if s is None:
        return None
    inv2 = pow(2, P - 2, P)
    x0 = fp_sqrt((a0 + s) * inv2 % P)
    if x0 is None:
        x0 = fp_sqrt((a0 - s) * inv2 % P)
        if x0 is None:
            return None
    x1 = a1 * pow(2 * x0, P - 2, P) % P
    r = (x0, x1)
    return r if f2_mul(r, r) == (a0, a1) else None


def fp_is_larger(y):
    """ark's `y > -y` on Fq: the integer order of the canonical representatives."""
    return y % P > (P - 1) // 2


def f2_is_larger(y):
    """ark-ff's Ord on a quadratic extension compares c1 first, then c0."""
    return fp_is_larger(y[1]) if y[1] % P else fp_is_larger(y[0])


def on_curve_g1(pt):
    x, y = pt
    return (y * y - x * x * x - B1) % P == 0


def on_curve_g2(pt):
    x, y = pt
    return f2_mul(y, y) == f2_add(f2_mul(f2_mul(x, x), x), B2)


# ------------------------------------------------------------------ G2 arithmetic on the twist (subgroup test by [r]Q = O)
def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0):
            return None
        lam = f2_mul(f2_mul((3, 0), f2_mul(x1, x1)), f2_inv(f2_add(y1, y1)))
    else:
        lam = f2_mul(f2_add(y2, f2_neg(y1)), f2_inv(f2_add(x2, f2_neg(x1))))
    x3 = f2_add(f2_mul(lam, lam), f2_neg(f2_add(x1, x2)))
    y3 = f2_add(f2_mul(lam, f2_add(x1, f2_neg(x3))), f2_neg(y1))
    return (x3, y3)


def g2_mul(k, pt):
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, pt)
        pt = g2_add(pt, pt)
        k >>= 1
    return acc


def in_subgroup_g2(pt):
    return g2_mul(R_ORDER, pt) is None


# ------------------------------------------------------------------ encoders (None = point at infinity)
def _le(v):
    return (v % P).to_bytes(32, "little")


def _be(v):
    return (v % P).to_bytes(32, "big")


def _with_flags(b, flags):
    return b[:-1] + bytes([b[-1] | flags])


def encode_g1(pt, fmt):
    if fmt == EIP197:
        return bytes(64) if pt is None else _be(pt[0]) + _be(pt[1])
    if pt is None:
        return _with_flags(bytes(32 if fmt == COMPRESSED else 64), FLAG_INF)
    flags = FLAG_NEG if fp_is_larger(pt[1]) else 0
    if fmt == COMPRESSED:
        return _with_flags(_le(pt[0]), flags)
    return _le(pt[0]) + _with_flags(_le(pt[1]), flags)


def encode_g2(pt, fmt):
    if fmt == EIP197:
        if pt is None:
            return bytes(128)
        (x0, x1), (y0, y1) = pt
        return _be(x1) + _be(x0) + _be(y1) + _be(y0)
    if pt is None:
        return _with_flags(bytes(64 if fmt == COMPRESSED else 128), FLAG_INF)
    (x0, x1), (y0, y1) = pt
    flags = FLAG_NEG if f2_is_larger((y0, y1)) else 0
    if fmt == COMPRESSED:
        return _le(x0) + _with_flags(_le(x1), flags)
    return _le(x0) + _le(x1) + _le(y0) + _with_flags(_le(y1), flags)


# ------------------------------------------------------------------ decoders: (status, point or None)
def _split_flags(b):
    last = b[-1]
    return b[:-1] + bytes([last & 0x3F]), last & 0xC0


def decode_g1(b, fmt):
    if fmt == EIP197:
        x, y = int.from_bytes(b[:32], "big"), int.from_bytes(b[32:64], "big")
        if x == 0 and y == 0:
            return INFINITY, None
        if x >= P or y >= P:
            return NOT_CANONICAL, None
        return (OK, (x, y)) if on_curve_g1((x, y)) else (NOT_ON_CURVE, None)
    b, flags = _split_flags(b)
    if flags == FLAG_NEG | FLAG_INF:
        return NOT_CANONICAL, None
    if flags & FLAG_INF:
        return INFINITY, None
    x = int.from_bytes(b[:32], "little")
    if fmt == COMPRESSED:
        if x >= P:
            return NOT_CANONICAL, None
        y = fp_sqrt((x * x * x + B1) % P)
        if y is None:
            return NOT_ON_CURVE, None
        if fp_is_larger(y) != bool(flags & FLAG_NEG):
            y = (-y) % P
        return OK, (x, y)
    y = int.from_bytes(b[32:64], "little")
    if x >= P or y >= P:
        return NOT_CANONICAL, None
    return (OK, (x, y)) if on_curve_g1((x, y)) else (NOT_ON_CURVE, None)


def decode_g2(b, fmt, check_subgroup=True):
    def finish(pt):
        if not on_curve_g2(pt):
            return NOT_ON_CURVE, None
        if check_subgroup and not in_subgroup_g2(pt):
            return NOT_IN_SUBGROUP, None
        return OK, pt

    if fmt == EIP197:
        x1, x0, y1, y0 = (int.from_bytes(b[32 * i:32 * i + 32], "big") for i in range(4))
        if x0 == x1 == y0 == y1 == 0:
            return INFINITY, None
        if max(x0, x1, y0, y1) >= P:
            return NOT_CANONICAL, None
        return finish(((x0, x1), (y0, y1)))
    b, flags = _split_flags(b)
    if flags == FLAG_NEG | FLAG_INF:
        return NOT_CANONICAL, None
    if flags & FLAG_INF:
        return INFINITY, None
    x0, x1 = int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little")
    if fmt == COMPRESSED:
        if x0 >= P or x1 >= P:
            return NOT_CANONICAL, None
        x = (x0, x1)
        y = f2_sqrt(f2_add(f2_mul(f2_mul(x, x), x), B2))
        if y is None:
            return NOT_ON_CURVE, None
        if f2_is_larger(y) != bool(flags & FLAG_NEG):
            y = f2_neg(y)
        return finish((x, y))
    y0, y1 = int.from_bytes(b[64:96], "little"), int.from_bytes(b[96:128], "little")
    if max(x0, x1, y0, y1) >= P:
        return NOT_CANONICAL, None
    return finish(((x0, x1), (y0, y1)))


# ------------------------------------------------------------------ Fq12 (MyFq12 coefficient list of 12 ints <-> ark bytes)
ARK_ORDER = [0, 2, 4, 1, 3, 5]   # ark's c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2 = w^0, w^2, w^4, w^1, w^3, w^5


def encode_fq12(coeffs):
    out = b""
    for i in ARK_ORDER:
        out += _le(coeffs[i]) + _le(coeffs[i + 6])
    return out


def decode_fq12(b):
    vals = [int.from_bytes(b[32 * k:32 * k + 32], "little") for k in range(12)]
    if max(vals) >= P:
        return NOT_CANONICAL, None
    coeffs = [0] * 12
    for k, i in enumerate(ARK_ORDER):
        coeffs[i], coeffs[i + 6] = vals[2 * k], vals[2 * k + 1]
    return OK, coeffs


# ------------------------------------------------------------------ EIP-197
def eip197_pairing_check(data):
    """True / False, or None where the precompile fails (malformed input)."""
    assert len(data) % 192 == 0
    acc = None
    for i in range(len(data) // 192):
        s1, p = decode_g1(data[192 * i:192 * i + 64], EIP197)
        s2, q = decode_g2(data[192 * i + 64:192 * i + 192], EIP197, check_subgroup=True)
        if s1 >= NOT_CANONICAL or s2 >= NOT_CANONICAL:
            return None
        if p is None or q is None:
            continue
        m = O.miller_loop_native(q, p)
        acc = m if acc is None else O.fq12_mul(acc, m)
    if acc is None:
        return True
    return O.final_exp_native(acc) == [1] + [0] * 11
