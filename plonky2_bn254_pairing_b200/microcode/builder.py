"""SSA builder for sequencer programs plus the Fq6/Fq12 tower algorithms expressed in it.

Everything here is build-time Python: it produces the straight-line Fq2 instruction streams that
csrc/vm.cu executes (`gen.py` writes them into csrc/microcode_gen.h).  The same streams can be run on
the CPU by `interp.py` (big integers), which is how the schedules are validated against the oracle
without a GPU.

Tower (identical to the reference's `MyFq12`, /root/reference/src/miller_loop_native.rs:46-96):
    Fq2  = Fq[u]/(u^2+1),  xi = 9+u
    Fq12 = Fq2[w]/(w^6 - xi);  a 12-element value is a list of six Fq2 SSA values [a0..a5] (w^0..w^5).
    Internally A = (a0,a2,a4), B = (a1,a3,a5) are its Fq6 = Fq2[v]/(v^3-xi) halves, v = w^2.
"""
import os

from . import isa

P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
BN_X = 4965661367192848881
XI = (9, 1)


# ----------------------------------------------------------------------------- plain Fq2 helpers (build-time constants)
def c_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def c_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = c_mul(r, a)
        a = c_mul(a, a)
        e >>= 1
    return r


def c_inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % P, P - 2, P)
    return (a[0] * n % P, (-a[1]) * n % P)


class ConstPool:
    """Fq2 constants shared by all programs of one library build (index = LDC immediate)."""

    def __init__(self):
        self.values = []
        self.index = {}

    def get(self, c):
        c = (c[0] % P, c[1] % P)
        if c not in self.index:
            self.index[c] = len(self.values)
            self.values.append(c)
        return self.index[c]


class Op:
    __slots__ = ("op", "dst", "srcs", "imm", "f_lo", "f_hi")

    def __init__(self, op, dst, srcs, imm=0, f_lo=0, f_hi=0):
        self.op, self.dst, self.srcs, self.imm, self.f_lo, self.f_hi = op, dst, list(srcs), imm, f_lo, f_hi


class V:
    """An SSA Fq2 value."""
    __slots__ = ("b", "id")

    def __init__(self, b, vid):
        self.b, self.id = b, vid

    def __mul__(self, o):
        return self.b.mul(self, o)

    def __add__(self, o):
        return self.b.add(self, o)

    def __sub__(self, o):
        return self.b.sub(self, o)

    def __neg__(self):
        return self.b.neg(self)

    def sqr(self):
        return self.b.sqr(self)

    def conj(self):
        return self.b.conj(self)

    def mulxi(self):
        return self.b.mulxi(self)

    def dbl(self):
        return self.b.dbl(self)

    def inv(self):
        return self.b.inv(self)

    def mulfp(self, s, half):
        return self.b.mulfp(self, s, half)


class Builder:
    def __init__(self, pool):
        self.pool = pool
        self.ops = []
        self.nvals = 0
        self._const_cache = {}

    # ---- primitive emitters
    def _emit(self, op, srcs, imm=0, f_lo=0, f_hi=0, has_dst=True):
        dst = None
        if has_dst:
            dst = self.nvals
            self.nvals += 1
        self.ops.append(Op(op, dst, [s.id for s in srcs], imm, f_lo, f_hi))
        return V(self, dst) if has_dst else None

    def mul(self, a, b):
        return self._emit("MUL", [a, b])

    def sqr(self, a):
        return self._emit("SQR", [a])

    def mulfp(self, a, s, half):
        """a * (half 0: s.c0, half 1: s.c1) - multiplication of an Fq2 by an Fq scalar."""
        return self._emit("MULFP", [a, s], imm=half)

    def add(self, a, b):
        return self._emit("ADD", [a, b])

    def sub(self, a, b):
        return self._emit("SUB", [a, b])

    def neg(self, a):
        return self._emit("NEG", [a])

    def conj(self, a):
        return self._emit("CONJ", [a])

    def mulxi(self, a):
        return self._emit("MULXI", [a])

    def dbl(self, a):
        return self._emit("DBL", [a])

    def inv(self, a):
        return self._emit("INV", [a])

    def const(self, c):
        """Fq2 constant.  Each call site gets its own rematerialisable value (an LDC is one cheap
        instruction), which keeps live ranges short."""
        return self._emit("LDC", [], imm=self.pool.get(c))

    def ldg(self, arr, f_lo, f_hi):
        return self._emit("LDG", [], imm=arr, f_lo=f_lo, f_hi=f_hi)

    def stg(self, arr, f_lo, f_hi, v):
        self._emit("STG", [v], imm=arr, f_lo=f_lo, f_hi=f_hi, has_dst=False)

    def cut(self):
        """Candidate phase boundary (microcode/phases.py): the program may be split here into separately
        scheduled tasks that hand their live values over through the per-element state array."""
        self._emit("CUT", [], has_dst=False)

    # ---- small multiples: a linear pseudo-op that the fusion pass folds into whatever consumes it
    def times(self, a, k):
        assert 1 <= k <= 31
        if k == 1:
            return a
        return self._emit("MULK", [a], imm=k)

    # ---- Fq12 load / store in MyFq12 coefficient order: coeffs[i] + coeffs[i+6] u  <->  w^i
    def ld_fq12(self, arr, base=0):
        return [self.ldg(arr, base + i, base + i + 6) for i in range(6)]

    def st_fq12(self, arr, f, base=0):
        for i in range(6):
            self.stg(arr, base + i, base + i + 6, f[i])

    def fq12_one(self):
        return [self.const((1, 0))] + [self.const((0, 0)) for _ in range(5)]

    # ------------------------------------------------------------------------- Fq6 = Fq2[v]/(v^3 - xi)
    def fq6_add(self, a, b):
        return [a[i] + b[i] for i in range(3)]

    def fq6_sub(self, a, b):
        return [a[i] - b[i] for i in range(3)]

    def fq6_neg(self, a):
        return [-a[i] for i in range(3)]

    def fq6_mul_v(self, a):
        return [a[2].mulxi(), a[0], a[1]]

    def fq6_mul(self, a, b):
        """Karatsuba, 6 Fq2 multiplications."""
        v0, v1, v2 = a[0] * b[0], a[1] * b[1], a[2] * b[2]
        t0 = (a[1] + a[2]) * (b[1] + b[2]) - v1 - v2
        t1 = (a[0] + a[1]) * (b[0] + b[1]) - v0 - v1
        t2 = (a[0] + a[2]) * (b[0] + b[2]) - v0 - v2
        return [v0 + t0.mulxi(), t1 + v2.mulxi(), t2 + v1]

    def fq6_sqr(self, a):
        """Chung-Hasan SQR2: 2 multiplications + 3 squarings."""
        s0 = a[0].sqr()
        ab = a[0] * a[1]
        s1 = ab.dbl()
        s2 = (a[0] - a[1] + a[2]).sqr()
        bc = a[1] * a[2]
        s3 = bc.dbl()
        s4 = a[2].sqr()
        return [s0 + s3.mulxi(), s1 + s4.mulxi(), s1 + s2 + s3 - s0 - s4]

    def fq6_mul_fq2(self, a, s):
        return [a[i] * s for i in range(3)]

    def fq6_mul_by_01(self, x, c0, c1):
        """x * (c0 + c1 v): 5 Fq2 multiplications."""
        a_a = x[0] * c0
        b_b = x[1] * c1
        t1 = ((x[1] + x[2]) * c1 - b_b).mulxi() + a_a
        t2 = (x[0] + x[1]) * (c0 + c1) - a_a - b_b
        t3 = (x[0] + x[2]) * c0 - a_a + b_b
        return [t1, t2, t3]

    def fq6_inv(self, a):
        c0, c1, c2 = a
        t0 = c0.sqr() - (c1 * c2).mulxi()
        t1 = c2.sqr().mulxi() - c0 * c1
        t2 = c1.sqr() - c0 * c2
        d = c0 * t0 + (c2 * t1 + c1 * t2).mulxi()
        di = d.inv()
        return [t0 * di, t1 * di, t2 * di]

    # ------------------------------------------------------------------------- Fq12 in the w-power basis
    @staticmethod
    def _split(f):
        return [f[0], f[2], f[4]], [f[1], f[3], f[5]]

    @staticmethod
    def _join(A, B):
        return [A[0], B[0], A[1], B[1], A[2], B[2]]

    def fq12_mul(self, f, g):
        """18 Fq2 multiplications (Karatsuba over Fq6)."""
        A, B = self._split(f)
        C, D = self._split(g)
        AC = self.fq6_mul(A, C)
        BD = self.fq6_mul(B, D)
        M = self.fq6_mul(self.fq6_add(A, B), self.fq6_add(C, D))
        even = self.fq6_add(AC, self.fq6_mul_v(BD))
        odd = self.fq6_sub(self.fq6_sub(M, AC), BD)
        return self._join(even, odd)

    def fq12_mul_conj(self, f, g):
        """f * conj(g) without materialising conj(g): conj negates the odd half D of g."""
        A, B = self._split(f)
        C, D = self._split(g)
        AC = self.fq6_mul(A, C)
        BD = self.fq6_mul(B, D)
        M = self.fq6_mul(self.fq6_add(A, B), self.fq6_sub(C, D))
        even = self.fq6_sub(AC, self.fq6_mul_v(BD))
        odd = self.fq6_add(self.fq6_sub(M, AC), BD)
        return self._join(even, odd)

    def fq12_sqr(self, f):
        """Complex squaring: 12 Fq2 multiplications."""
        A, B = self._split(f)
        AB = self.fq6_mul(A, B)
        T = self.fq6_mul(self.fq6_add(A, B), self.fq6_add(A, self.fq6_mul_v(B)))
        even = self.fq6_sub(self.fq6_sub(T, AB), self.fq6_mul_v(AB))
        odd = [x.dbl() for x in AB]
        return self._join(even, odd)

    def fq12_mul_034(self, f, e0, e1, e3):
        """f * (e0 + e1 w + e3 w^3): 13 Fq2 multiplications.
        The line is E0 + E1 w with E0 = (e0,0,0), E1 = (e1,e3,0) over v = w^2."""
        A, B = self._split(f)
        a = self.fq6_mul_fq2(A, e0)
        b = self.fq6_mul_by_01(B, e1, e3)
        e = self.fq6_mul_by_01(self.fq6_add(A, B), e0 + e1, e3)
        odd = self.fq6_sub(self.fq6_sub(e, a), b)
        even = self.fq6_add(a, self.fq6_mul_v(b))
        return self._join(even, odd)

    def fq12_conj(self, f):
        """a^(p^6): negate the odd powers of w (final_exp_native.rs:171-181)."""
        return [f[i] if i % 2 == 0 else -f[i] for i in range(6)]

    def fq12_inv(self, f):
        A, B = self._split(f)
        n = self.fq6_sub(self.fq6_sqr(A), self.fq6_mul_v(self.fq6_sqr(B)))
        ni = self.fq6_inv(n)
        return self._join(self.fq6_mul(A, ni), self.fq6_neg(self.fq6_mul(B, ni)))

    def fq12_rot(self, f, k):
        """f * w^k for 0 <= k < 6."""
        out = [None] * 6
        for i in range(6):
            j = i + k
            out[j % 6] = f[i] if j < 6 else f[i].mulxi()
        return out

    def fq12_mul_fq2(self, f, s):
        return [x * s for x in f]

    def fq12_frobenius(self, f, power):
        """a^(p^power): coefficient i -> conj^power(a_i) * gamma_power^i, gamma_k = xi^((p^k-1)/6)
        (final_exp_native.rs:17-54; the constants there are recomputed per call, here they are table entries)."""
        pw = power % 12
        gamma = c_pow(XI, (P ** pw - 1) // 6)
        out = []
        for i in range(6):
            g = c_pow(gamma, i)
            a = f[i].conj() if pw % 2 else f[i]
            if g == (1, 0):
                out.append(a)
            elif g[1] == 0:
                out.append(a.mulfp(self.const(g), 0))
            else:
                out.append(a * self.const(g))
        return out

    def fq12_cyclo_sqr(self, f):
        """Granger-Scott squaring, valid in the cyclotomic subgroup (after the easy part).
        With s = w^3 (s^2 = xi): Fq4 = Fq2[s], f = A + B w + C w^2 over Fq4, A=(f0,f3), B=(f1,f4), C=(f2,f5).
        A' = 3A^2 - 2 conj(A),  B' = 3 s C^2 + 2 conj(B),  C' = 3 B^2 - 2 conj(C).   9 Fq2 squarings."""

        def fq4_sqr(a, b):  # (a + b s)^2 = (a^2 + xi b^2) + ((a+b)^2 - a^2 - b^2) s
            a2, b2 = a.sqr(), b.sqr()
            return a2 + b2.mulxi(), (a + b).sqr() - a2 - b2

        a0, a1 = fq4_sqr(f[0], f[3])
        b0, b1 = fq4_sqr(f[1], f[4])
        c0, c1 = fq4_sqr(f[2], f[5])
        # s * C^2 = xi c1 + c0 s
        sc0, sc1 = c1.mulxi(), c0

        def three_minus_two(x, y):  # 3x - 2y   (x used once: the whole expression folds into x's product)
            return self.times(x, 3) - y.dbl()

        def three_plus_two(x, y):  # 3x + 2y
            return self.times(x, 3) + y.dbl()

        n0 = three_minus_two(a0, f[0])
        n3 = three_plus_two(a1, f[3])
        n1 = three_plus_two(sc0, f[1])
        n4 = three_minus_two(sc1, f[4])
        n2 = three_minus_two(b0, f[2])
        n5 = three_plus_two(b1, f[5])
        return [n0, n1, n2, n3, n4, n5]

    def fq12_pow_x_cyclo(self, m, window=None):
        """m^BN_X for cyclotomic m.  The reference walks the NAF of the exponent (pow_native, final_exp_native.rs:56-84:
        62 squarings, 23 multiplications / divisions); m^BN_X is one well-defined group element, so any addition chain
        gives the same bits, and in the cyclotomic subgroup (after the easy part) squarings are Granger-Scott and an
        inverse is a conjugation.  Here: width-4 signed windows - digits +-1, +-3, +-5, +-7 on a table m, m^3, m^5, m^7
        (one squaring and three multiplications to build) - 13 multiplications in the walk instead of 23:
        16 Fq12 multiplications per exponentiation instead of 23 (-6 % of a pairing's MACs).  window=2 is the
        reference's NAF walk."""
        if window is None:
            window = int(os.environ.get("BNP_POWX_WINDOW", "4"))
        digits = wnaf_digits(BN_X, window)
        table = {1: m}
        top = max(abs(z) for z in digits)
        if top > 1:
            m2 = self.fq12_cyclo_sqr(m)
            for k in range(3, top + 1, 2):
                table[k] = self.fq12_mul(table[k - 2], m2)
        res = None
        for z in reversed(digits):
            if res is not None:
                self.cut()
                res = self.fq12_cyclo_sqr(res)
            if z != 0:
                t = table[abs(z)]
                if res is None:
                    assert z > 0
                    res = t
                else:
                    res = self.fq12_mul(res, t) if z > 0 else self.fq12_mul_conj(res, t)
        return res


def wnaf_digits(e, w):
    """LSB-first width-w non-adjacent form of a positive integer: odd digits in (-2^(w-1), 2^(w-1)), at most one non-zero
    digit in any w consecutive positions (w = 2 is the NAF of get_naf, final_exp_native.rs:86-128)."""
    out = []
    while e:
        if e & 1:
            z = e % (1 << w)
            if z >= 1 << (w - 1):
                z -= 1 << w
            e -= z
        else:
            z = 0
        out.append(z)
        e >>= 1
    return out


def naf_digits(e):
    """LSB-first NAF of a non-negative integer (same digits as get_naf, final_exp_native.rs:86-128)."""
    out = []
    while e:
        if e & 1:
            z = 2 - (e % 4)
            e -= z
        else:
            z = 0
        out.append(z)
        e //= 2
    return out
