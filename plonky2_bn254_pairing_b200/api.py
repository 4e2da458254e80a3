"""Host-side mirror of the reference's native-path interface, on top of the C ABI (include/bnp.h).

Same names, argument order and meaning as the reference's public functions:

    miller_loop_native(Q, P)              /root/reference/src/miller_loop_native.rs:320   (G2 first!)
    pow_native(a, exp)                    /root/reference/src/final_exp_native.rs:56      exp = list of u64 limbs
    get_naf(exp) / frob_coeffs(index)     /root/reference/src/final_exp_native.rs:86,183  (host-side constants)
    conjugate_fp2 / neg_conjugate_fp2     /root/reference/src/miller_loop_native.rs:284,291
    SIX_U_PLUS_2_NAF, BN_X                /root/reference/src/miller_loop_native.rs:314; final_exp_native.rs:15
    multi_miller_loop_native(pairs)       /root/reference/src/miller_loop_native.rs:324   pairs = [(P, Q), ...]
    final_exp_native(a)                   /root/reference/src/final_exp_native.rs:209
    frobenius_map_native(a, power)        /root/reference/src/final_exp_native.rs:17
    pairing(p, q)                         /root/reference/src/pairing.rs:20               (G1 first)

plus the batched slice variants the north star adds (`*_batch`).  Values use the oracle's plain-integer
conventions (Fq = int, Fq2 = (c0, c1), G1 = (x, y), G2 = (x, y) over Fq2, MyFq12 = 12 ints); this module
converts to the ABI's Montgomery structure-of-arrays layout and back, nothing else - all arithmetic
happens in libbnp.so on the GPU.  Errors surface as BnpError (the reference panics).
"""
import ctypes

import numpy as np

from . import native

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
MONT_R = (1 << 256) % P
MONT_RINV = pow(MONT_R, P - 2, P)

VARIANT_REFERENCE = 0
VARIANT_ARK = 1


# ----------------------------------------------------------------------------- marshalling
def pack_soa(rows):
    """rows: n sequences of K Fq ints -> uint64 array [K][4][n] of Montgomery limbs (ark's Fp.0.0)."""
    n = len(rows)
    K = len(rows[0]) if n else 0
    buf = bytearray()
    for r in rows:
        assert len(r) == K
        for v in r:
            buf += (v * MONT_R % P).to_bytes(32, "little")
    a = np.frombuffer(bytes(buf), dtype=np.uint64).reshape(n, K, 4)
    return np.ascontiguousarray(a.transpose(1, 2, 0))


def unpack_soa(arr):
    """uint64 [K][4][n] Montgomery limbs -> n lists of K plain Fq ints (asserts canonical residues)."""
    K, four, n = arr.shape
    assert four == 4
    a = np.ascontiguousarray(arr.transpose(2, 0, 1)).tobytes()
    out = []
    for e in range(n):
        row = []
        for k in range(K):
            off = (e * K + k) * 32
            m = int.from_bytes(a[off:off + 32], "little")
            if m >= P:
                raise native.BnpError("non-canonical residue returned by the device")
            row.append(m * MONT_RINV % P)
        out.append(row)
    return out


def g1_rows(points):
    return [[p[0], p[1]] for p in points]


def g2_rows(points):
    return [[q[0][0], q[0][1], q[1][0], q[1][1]] for q in points]


def _ptr(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


# ----------------------------------------------------------------------------- SoA-level calls (numpy in, numpy out)
def miller_loop_soa(g1, g2, k=1):
    lib = native.lib()
    n = g1.shape[2]
    out = np.empty((12, 4, n), dtype=np.uint64)
    if k == 1:
        native.check(lib.bnp_miller_loop_batch(_ptr(g1), _ptr(g2), _ptr(out), n))
    else:
        native.check(lib.bnp_multi_miller_loop_batch(_ptr(g1), _ptr(g2), _ptr(out), n, k))
    return out


def final_exp_soa(f, variant=VARIANT_REFERENCE):
    lib = native.lib()
    n = f.shape[2]
    out = np.empty((12, 4, n), dtype=np.uint64)
    native.check(lib.bnp_final_exp_batch(_ptr(f), _ptr(out), n, variant))
    return out


WITNESS_FIELDS = ("m", "mx", "mx2", "mx3", "out")   # 12 Fq each, in this order (include/bnp.h BNP_WITNESS_*)


def final_exp_witness_soa(f):
    """[12][4][n] -> [60][4][n]: easy part, its three BN_X powers and final_exp_native, one GPU pass."""
    lib = native.lib()
    n = f.shape[2]
    out = np.empty((12 * len(WITNESS_FIELDS), 4, n), dtype=np.uint64)
    native.check(lib.bnp_final_exp_witness_batch(_ptr(f), _ptr(out), n))
    return out


def pairing_soa(g1, g2, variant=VARIANT_REFERENCE, k=1):
    lib = native.lib()
    n = g1.shape[2]
    out = np.empty((12, 4, n), dtype=np.uint64)
    if k == 1:
        native.check(lib.bnp_pairing_batch(_ptr(g1), _ptr(g2), _ptr(out), n, variant))
    else:
        native.check(lib.bnp_multi_pairing_batch(_ptr(g1), _ptr(g2), _ptr(out), n, k, variant))
    return out


def pairing_product_soa(g1, g2, variant=VARIANT_REFERENCE):
    lib = native.lib()
    n = g1.shape[2]
    out = np.empty((12, 4, 1), dtype=np.uint64)
    native.check(lib.bnp_pairing_product(_ptr(g1), _ptr(g2), _ptr(out), n, variant))
    return out


def frobenius_soa(f, power):
    lib = native.lib()
    n = f.shape[2]
    out = np.empty((12, 4, n), dtype=np.uint64)
    native.check(lib.bnp_frobenius_batch(_ptr(f), _ptr(out), n, power))
    return out


def fq12_mul_soa(a, b):
    lib = native.lib()
    n = a.shape[2]
    out = np.empty((12, 4, n), dtype=np.uint64)
    native.check(lib.bnp_fq12_mul_batch(_ptr(a), _ptr(b), _ptr(out), n))
    return out


def pow_soa(f, exp):
    """f ^ exp for every element; exp: list of little-endian u64 limbs (the reference's Vec<u64>), shared by the batch."""
    lib = native.lib()
    n = f.shape[2]
    out = np.empty((12, 4, n), dtype=np.uint64)
    limbs = np.array([int(x) for x in exp], dtype=np.uint64)
    native.check(lib.bnp_pow_u64_batch(_ptr(f), _ptr(out), n, limbs.ctypes.data_as(ctypes.c_void_p) if len(limbs) else None,
                                       len(limbs)))
    return out


def validate_soa(g1=None, g2=None):
    """uint8 flags [n]: 1 iff the element's points pass `G1Affine::new` / `G2Affine::new` (on-curve, G2 in the subgroup)."""
    lib = native.lib()
    n = (g1 if g1 is not None else g2).shape[2]
    ok = np.zeros(n, dtype=np.uint8)
    native.check(lib.bnp_validate_batch(_ptr(g1) if g1 is not None else None, _ptr(g2) if g2 is not None else None,
                                        ok.ctypes.data_as(ctypes.c_void_p), n))
    return ok


def scalar_mul_soa(group, pts, scalars):
    """group 1: pts uint64 [2][4][n] (G1), 2: [4][4][n] (G2); scalars uint64 [4][n], plain little-endian 256-bit
    integers -> (points of the same shape, uint8 [n] infinity flags): out_i = scalars_i * pts_i."""
    lib = native.lib()
    n = pts.shape[2]
    out = np.empty_like(pts)
    inf = np.zeros(n, dtype=np.uint8)
    sc = np.ascontiguousarray(scalars, dtype=np.uint64)
    assert sc.shape == (4, n)
    native.check(lib.bnp_scalar_mul_batch(group, _ptr(pts), _ptr(sc), _ptr(out), inf.ctypes.data_as(ctypes.c_void_p), n))
    return out, inf


# ----------------------------------------------------------------------------- prepared G2 points
PREP_FQ = 546


def g2_prepare_soa(g2):
    """uint64 [4][4][n] -> [PREP_FQ][4][n]: the line coefficients of every point (the engine's `G2Prepared`)."""
    lib = native.lib()
    n = g2.shape[2]
    out = np.empty((PREP_FQ, 4, n), dtype=np.uint64)
    native.check(lib.bnp_g2_prepare_batch(_ptr(g2), _ptr(out), n))
    return out


def pairing_prepared_soa(g1, g2, prepared, kv, kp, variant=VARIANT_REFERENCE):
    """g1 [2 (kv + kp)][4][n], g2 [4 kv][4][n] or None, prepared [kp * PREP_FQ][4][1] (shared by the batch) ->
    [12][4][n]: the product of kv + kp pairings per element, the last kp against the prepared points."""
    lib = native.lib()
    n = g1.shape[2]
    assert g1.shape[0] == 2 * (kv + kp) and prepared.shape == (kp * PREP_FQ, 4, 1)
    out = np.empty((12, 4, n), dtype=np.uint64)
    native.check(lib.bnp_pairing_prepared_batch(_ptr(g1), _ptr(g2) if kv else None, _ptr(prepared), _ptr(out), n, kv, kp,
                                                variant))
    return out


# ----------------------------------------------------------------------------- wire formats
WIRE_ARK_UNCOMPRESSED, WIRE_ARK_COMPRESSED, WIRE_EIP197 = 0, 1, 2
POINT_OK, POINT_INFINITY, POINT_NOT_CANONICAL, POINT_NOT_ON_CURVE, POINT_NOT_IN_SUBGROUP = 0, 1, 2, 3, 4


def _point_bytes(group, fmt):
    full = 64 if group == 1 else 128
    return full // 2 if fmt == WIRE_ARK_COMPRESSED else full


def decode_g1_soa(fmt, data):
    """bytes (n elements of 64 / 32 / 64 bytes for ark uncompressed / ark compressed / EIP-196) -> (uint64 [2][4][n]
    Montgomery SoA, uint8 status [n]); decoding, range and curve checks and the compressed form's square root run on
    the device (csrc/wire.cuh)."""
    lib = native.lib()
    size = _point_bytes(1, fmt)
    assert len(data) % size == 0
    n = len(data) // size
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    out = np.zeros((2, 4, n), dtype=np.uint64)
    st = np.zeros(n, dtype=np.uint8)
    native.check(lib.bnp_decode_g1_batch(fmt, buf.ctypes.data_as(ctypes.c_void_p), n, _ptr(out),
                                         st.ctypes.data_as(ctypes.c_void_p)))
    return out, st


def decode_g2_soa(fmt, data, check_subgroup=True):
    """bytes (128 / 64 / 128 per element) -> (uint64 [4][4][n], uint8 status [n]); `check_subgroup` adds the r-torsion
    test of `G2Affine::new` (miller_loop_native.rs:303,311)."""
    lib = native.lib()
    size = _point_bytes(2, fmt)
    assert len(data) % size == 0
    n = len(data) // size
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    out = np.zeros((4, 4, n), dtype=np.uint64)
    st = np.zeros(n, dtype=np.uint8)
    native.check(lib.bnp_decode_g2_batch(fmt, buf.ctypes.data_as(ctypes.c_void_p), n, _ptr(out),
                                         st.ctypes.data_as(ctypes.c_void_p), 1 if check_subgroup else 0))
    return out, st


def encode_fq12_soa(f):
    """uint64 [12][4][n] MyFq12 SoA -> n * 384 bytes: ark-serialize of the ark Fq12 that `MyFq12::into()` gives."""
    lib = native.lib()
    n = f.shape[2]
    out = np.zeros(384 * n, dtype=np.uint8)
    native.check(lib.bnp_encode_fq12_batch(_ptr(f), n, out.ctypes.data_as(ctypes.c_void_p)))
    return out.tobytes()


def decode_fq12_soa(data):
    lib = native.lib()
    assert len(data) % 384 == 0
    n = len(data) // 384
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    out = np.zeros((12, 4, n), dtype=np.uint64)
    st = np.zeros(n, dtype=np.uint8)
    native.check(lib.bnp_decode_fq12_batch(buf.ctypes.data_as(ctypes.c_void_p), n, _ptr(out),
                                           st.ctypes.data_as(ctypes.c_void_p)))
    return out, st


def eip197_pairing_check(data):
    """The Ethereum pairing precompile (EIP-197) on len(data) / 192 pairs: True iff the product of pairings is one.
    Raises BnpError (BNP_EMALFORMED) where the precompile fails: a non-canonical coordinate, a point off its curve, a G2
    point outside the subgroup."""
    lib = native.lib()
    if len(data) % 192:
        raise native.BnpError("EIP-197 input length must be a multiple of 192 bytes")
    k = len(data) // 192
    buf = np.frombuffer(bytes(data), dtype=np.uint8) if k else np.zeros(1, dtype=np.uint8)
    res = ctypes.c_int(0)
    native.check(lib.bnp_eip197_pairing_check(buf.ctypes.data_as(ctypes.c_void_p), k, ctypes.byref(res)))
    return bool(res.value)


# ----------------------------------------------------------------------------- batched slice variants (north star)
def miller_loop_native_batch(Qs, Ps):
    """[miller_loop_native(Q_i, P_i)]"""
    if len(Qs) != len(Ps):
        raise ValueError("length mismatch")
    if not Qs:
        return []
    return unpack_soa(miller_loop_soa(pack_soa(g1_rows(Ps)), pack_soa(g2_rows(Qs))))


def multi_miller_loop_native_batch(products):
    """products: list of equal-length lists of (P, Q) tuples -> one MyFq12 per product (k <= 4)."""
    if not products:
        return []
    k = len(products[0])
    g1 = pack_soa([[c for (p, _) in prod for c in (p[0], p[1])] for prod in products])
    g2 = pack_soa([[c for (_, q) in prod for c in (q[0][0], q[0][1], q[1][0], q[1][1])] for prod in products])
    return unpack_soa(miller_loop_soa(g1, g2, k=k))


def final_exp_native_batch(fs, variant=VARIANT_REFERENCE):
    if not fs:
        return []
    return unpack_soa(final_exp_soa(pack_soa(fs), variant))


def final_exp_witness_batch(fs):
    """The native values the final-exponentiation circuit consumes (final_exp_target.rs:65-185), per input:
    {"m": easy part, "mx": m^x, "mx2": m^(x^2), "mx3": m^(x^3), "out": final_exp_native(a)}."""
    if not fs:
        return []
    flat = unpack_soa(final_exp_witness_soa(pack_soa(fs)))
    return [{name: row[12 * i:12 * i + 12] for i, name in enumerate(WITNESS_FIELDS)} for row in flat]


def pairing_batch(Ps, Qs, variant=VARIANT_REFERENCE):
    """[pairing(P_i, Q_i)] in MyFq12 coefficient order (see myfq12_to_ark for ark's nesting)."""
    if len(Qs) != len(Ps):
        raise ValueError("length mismatch")
    if not Ps:
        return []
    return unpack_soa(pairing_soa(pack_soa(g1_rows(Ps)), pack_soa(g2_rows(Qs)), variant))


def multi_pairing_batch(products, variant=VARIANT_REFERENCE):
    """products: list of k-lists of (P, Q) -> prod_j pairing(P_j, Q_j) per product (Groth16-verify shape)."""
    if not products:
        return []
    k = len(products[0])
    g1 = pack_soa([[c for (p, _) in prod for c in (p[0], p[1])] for prod in products])
    g2 = pack_soa([[c for (_, q) in prod for c in (q[0][0], q[0][1], q[1][0], q[1][1])] for prod in products])
    return unpack_soa(pairing_soa(g1, g2, variant, k=k))


def pairing_product(pairs, variant=VARIANT_REFERENCE):
    """final_exp(prod_i miller(Q_i, P_i)) over ALL pairs (any count), split across the initialised GPUs."""
    g1 = pack_soa(g1_rows([p for (p, _) in pairs]))
    g2 = pack_soa(g2_rows([q for (_, q) in pairs]))
    return unpack_soa(pairing_product_soa(g1, g2, variant))[0]


def validate_batch(Ps=None, Qs=None):
    """[bool]: would `G1Affine::new(P_i)` / `G2Affine::new(Q_i)` accept the coordinates?  (The check the reference
    gets from ark inside twisted_frobenius, miller_loop_native.rs:303,311; the pairing calls here do not repeat it.)"""
    g1 = pack_soa(g1_rows(Ps)) if Ps else None
    g2 = pack_soa(g2_rows(Qs)) if Qs else None
    if g1 is None and g2 is None:
        return []
    return [bool(v) for v in validate_soa(g1, g2)]


def _scalar_rows(ks):
    a = np.zeros((4, len(ks)), dtype=np.uint64)
    for i, k in enumerate(ks):
        assert 0 <= k < 1 << 256
        for j in range(4):
            a[j, i] = (k >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return a


def g1_scalar_mul_batch(Ps, ks):
    """[k_i * P_i] on G1 - `G1.mul(s).into()` of the reference's test_to_one (final_exp_native.rs:247), batched.  Points are
    (x, y) integer pairs; the point at infinity is None (in and out)."""
    if not Ps:
        return []
    out, inf = scalar_mul_soa(1, pack_soa(g1_rows([p if p is not None else (0, 0) for p in Ps])), _scalar_rows(ks))
    rows = unpack_soa(out)
    return [None if f else (r[0], r[1]) for r, f in zip(rows, inf)]


def g2_scalar_mul_batch(Qs, ks):
    """[k_i * Q_i] on G2 (`G2.mul(t).into()`, final_exp_native.rs:248); points are ((x0, x1), (y0, y1)) or None."""
    if not Qs:
        return []
    zero = ((0, 0), (0, 0))
    out, inf = scalar_mul_soa(2, pack_soa(g2_rows([q if q is not None else zero for q in Qs])), _scalar_rows(ks))
    rows = unpack_soa(out)
    return [None if f else ((r[0], r[1]), (r[2], r[3])) for r, f in zip(rows, inf)]


def frobenius_map_native_batch(fs, power):
    if not fs:
        return []
    return unpack_soa(frobenius_soa(pack_soa(fs), power))


def fq12_mul_batch(As, Bs):
    if not As:
        return []
    return unpack_soa(fq12_mul_soa(pack_soa(As), pack_soa(Bs)))


# ----------------------------------------------------------------------------- the reference's scalar signatures (n = 1 calls)
def miller_loop_native(Q, Pt):
    return miller_loop_native_batch([Q], [Pt])[0]


def multi_miller_loop_native(pairs):
    if len(pairs) > 4:
        # shared-squaring programs exist for k <= 4; larger products: product of single loops, which the
        # reference's own test pins as equal (miller_loop_native.rs:336-348)
        ms = miller_loop_native_batch([q for (_, q) in pairs], [p for (p, _) in pairs])
        acc = ms[0]
        for m in ms[1:]:
            acc = fq12_mul_batch([acc], [m])[0]
        return acc
    return multi_miller_loop_native_batch([list(pairs)])[0]


def final_exp_native(a):
    return final_exp_native_batch([a])[0]


def frobenius_map_native(a, power):
    return frobenius_map_native_batch([a], power)[0]


def pow_native_batch(fs, exp):
    if not fs:
        return []
    return unpack_soa(pow_soa(pack_soa(fs), exp))


def pow_native(a, exp):
    """final_exp_native.rs:56-84: a^exp for any MyFq12 a, exp = list of u64 limbs (little endian)."""
    return pow_native_batch([a], exp)[0]


# ---- host-side constants of the reference's public API (no device work: these are table generators)
BN_X = 4965661367192848881                       # final_exp_native.rs:15
SIX_U_PLUS_2_NAF = [                             # miller_loop_native.rs:314-318
    0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 1, -1, 0, 0, 1, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, -1, 0, 0, 0, 0,
    1, 1, 1, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 1, 0, 0, -1, 0, 0, 0, 1, 1, 0, -1, 0,
    0, 1, 0, 1, 1,
]


def get_naf(exp):
    """final_exp_native.rs:86-128, statement by statement (exp: list of u64 limbs, least significant first):
    every limb contributes exactly 64 digits, the carry out of a limb is added into the next one, and a carry out
    of the top limb appends one more digit - after tripping the reference's own `assert_eq!(len, exp.len() + 1)`
    (:123), which can never hold there: the reference PANICS on such exponents, and so does this (ValueError)."""
    exp = [int(x) for x in exp]
    naf = []
    length = len(exp)
    for idx in range(length):
        e = exp[idx]
        for _ in range(64):
            if e & 1:
                z = 2 - (e % 4)
                e //= 2
                if z == -1:
                    e += 1
                naf.append(z)
            else:
                naf.append(0)
                e //= 2
        if e != 0:
            j = idx + 1
            while j < len(exp) and exp[j] == (1 << 64) - 1:
                exp[j] = 0
                j += 1
            if j < len(exp):
                exp[j] += 1
            else:
                exp.append(1)
    if len(exp) != length:
        raise ValueError("get_naf: carry out of the top limb (the reference panics here, final_exp_native.rs:123)")
    return naf


def frob_coeffs(index):
    """final_exp_native.rs:183-192: xi^((p^index - 1) / 6) as an Fq2 (c0, c1), xi = 9 + u."""
    e = (P ** index - 1) // 6
    r, b = (1, 0), (9, 1)
    while e:
        if e & 1:
            r = ((r[0] * b[0] - r[1] * b[1]) % P, (r[0] * b[1] + r[1] * b[0]) % P)
        b = ((b[0] * b[0] - b[1] * b[1]) % P, (2 * b[0] * b[1]) % P)
        e >>= 1
    return r


def conjugate_fp2(x):
    """miller_loop_native.rs:284-289"""
    return (x[0] % P, (-x[1]) % P)


def neg_conjugate_fp2(x):
    """miller_loop_native.rs:291-296"""
    return ((-x[0]) % P, x[1] % P)


def pairing(p, q):
    return pairing_batch([p], [q])[0]


def myfq12_to_ark(a):
    """`impl From<MyFq12> for Fq12`: MyFq12 coefficient order -> ark's [c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2]."""
    c = [(a[i], a[i + 6]) for i in range(6)]
    return [c[0], c[2], c[4], c[1], c[3], c[5]]
