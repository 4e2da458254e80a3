"""The pairing-level schedules, written against builder.Builder.

Reference being restated (results must be bit-identical):
    miller_loop_BN_native        /root/reference/src/miller_loop_native.rs:112-190
    multi_miller_loop_BN_native  /root/reference/src/miller_loop_native.rs:192-282
    easy_part / hard_part_BN_native / final_exp_native   /root/reference/src/final_exp_native.rs:130-213
    frobenius_map_native         /root/reference/src/final_exp_native.rs:17-54
    pairing                      /root/reference/src/pairing.rs:20-22

B200-first differences (all value-preserving):
  * R is kept in homogeneous projective coordinates; the reference normalises to affine after every
    step (one Fq2 inversion each, miller_loop_native.rs:157,167,186).  Our tangent line is
    xi*Z^2/w^3 times the reference's and our chord line Z/w^2 times the reference's; the accumulated
    factor is S * xi^a * w^-b with S in Fq2 tracked at run time and (a, b) known at build time.
      - standalone Miller loop: one Fq2 inversion at the end restores the reference's exact value;
      - fused pairing: the factor lies in a proper subfield and is annihilated by the easy part
        (p^6-1)(p^2+1), so it is not tracked at all.
  * lines are multiplied in "034" form (13 Fq2 products) instead of the reference's 18-product
    schoolbook (miller_loop_native.rs:46-96).
  * the hard part runs in the cyclotomic subgroup: Granger-Scott squarings, conj() for the inverse
    (the reference divides, final_exp_native.rs:72-75), Frobenius constants from a table
    (the reference recomputes them on every call, final_exp_native.rs:27,183-192).
"""
from . import isa
from .builder import Builder, ConstPool, P, XI, BN_X, c_mul, c_pow, c_inv, naf_digits

# miller_loop_native.rs:314-318
SIX_U_PLUS_2_NAF = [
    0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 1, -1, 0, 0, 1, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, -1, 0, 0, 0, 0,
    1, 1, 1, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 1, 0, 0, -1, 0, 0, 0, 1, 1, 0, -1, 0,
    0, 1, 0, 1, 1,
]

# xi^((p-1)/6): `expected_c` of miller_loop_native.rs:176-178; c2, c3 at :180-181
_C1 = c_pow(XI, (P - 1) // 6)
_C2 = c_mul(_C1, _C1)
_C3 = c_mul(_C2, _C1)


class _Pair:
    """Per-pair state of the projective Miller loop.  Every step yields a P-INDEPENDENT coefficient triple (c0, c1, c2)
    - what ark calls an `EllCoeff` of `G2Prepared`, in this engine's normalisation - and its evaluation at P,
    (e0, e1, e3) = (c0 * s0, c1 * s1, c2) with Fq scalars s0, s1 taken from P.  `with_p=False` builds the coefficients
    only (program g2_prepare)."""

    def __init__(self, b, pair_index, with_p=True, g2_index=None):
        k = pair_index
        self.b = b
        if with_p:
            self.Pv = b.ldg(isa.ARR_G1, 2 * k, 2 * k + 1)  # (xP, yP) packed in one slot
            # (-3 xP, -yP): scalars for the tangent's w^4 coefficient and the chord's w^2 coefficient
            self.Pn = -b.times(self.Pv, 3)  # c0 = -3 xP (c1 = -3 yP unused)
            self.Pm = -self.Pv  # c1 = -yP
        j = k if g2_index is None else g2_index
        self.Qx = b.ldg(isa.ARR_G2, 4 * j, 4 * j + 1)
        self.Qy = b.ldg(isa.ARR_G2, 4 * j + 2, 4 * j + 3)
        self.nQy = None
        self.X, self.Y, self.Z = self.Qx, self.Qy, b.const((1, 0))

    def neg_qy(self):
        if self.nQy is None:
            self.nQy = -self.Qy
        return self.nQy

    # ---- P-independent halves
    def tangent_coeffs(self):
        """Tangent at R and R <- 2R.  Returns (c0, c1, c2) = (xi 2YZ, xi X^2, xi Y^2 - 9 Z^2):
              tangent_ref*Z^2 = l0 + l3 w^3 + l4 w^4,  l0 = xi Y^2 - 9 Z^2, l3 = 2YZ yP, l4 = -3 X^2 xP
              (miller_loop_native.rs:30-44 with the curve equation substituted)
           and the line multiplied into f is xi*Z^2/w^3 * tangent_ref = (xi l3) + (xi l4) w + l0 w^3.
        Doubling: Costello-Lange-Naehrig homogeneous formulas for y^2 = x^3 + 3/xi, with the whole
        triple scaled by 4 xi^2 so that neither a halving nor the constant 3/xi is needed."""
        b = self.b
        X, Y, Z = self.X, self.Y, self.Z
        B = Y.sqr()
        C = Z.sqr()
        J = X.sqr()
        H = (Y * Z).dbl()  # 2YZ
        Bx = B.mulxi()  # xi Y^2
        E9 = b.times(C, 9)  # 9 Z^2
        F27 = b.times(E9, 3)
        xiH = H.mulxi()
        l0 = Bx - E9
        xiJ = J.mulxi()
        XY2 = (X * Y).dbl()
        self.X = (XY2 * (Bx - F27)).mulxi()
        G2 = Bx + F27
        self.Y = G2.sqr() - b.times(E9.sqr(), 12)
        self.Z = b.times(Bx * xiH, 4)
        return xiH, xiJ, l0

    def chord_coeffs(self, x2, y2, update=True):
        """Chord through R and the affine point (x2, y2), then R <- R + (x2, y2).  Returns (c0, c1, c2) =
        (X - x2 Z, Y - y2 Z, X y2 - x2 Y):
              chord_ref*Z = l2 w^2 + l3 w^3 + l5 w^5,
              l2 = (x2 Z - X) yP,  l3 = (Y - y2 Z) xP,  l5 = X y2 - x2 Y     (miller_loop_native.rs:10-28)
        and the line multiplied into f is Z/w^2 * chord_ref = l2 + l3 w + l5 w^3."""
        X, Y, Z = self.X, self.Y, self.Z
        theta = Y - y2 * Z
        lam = X - x2 * Z
        l5 = X * y2 - x2 * Y
        if update:
            c = theta.sqr()
            d = lam.sqr()
            e = lam * d
            f = Z * c
            g = X * d
            h = e + f - g.dbl()
            self.X = lam * h
            self.Y = theta * (g - h) - e * Y
            self.Z = Z * e
        return lam, theta, l5

    # ---- the Miller loop's view: one line per step, evaluated at P
    def eval_tangent(self, c):
        return c[0].mulfp(self.Pv, 1), c[1].mulfp(self.Pn, 0), c[2]     # xi 2YZ * yP,  xi X^2 * (-3 xP),  l0

    def eval_chord(self, c):
        return c[0].mulfp(self.Pm, 1), c[1].mulfp(self.Pv, 0), c[2]     # (X - x2 Z) * (-yP),  (Y - y2 Z) * xP,  l5

    def double_step(self):
        return self.eval_tangent(self.tangent_coeffs())

    def add_step(self, x2, y2, update=True):
        return self.eval_chord(self.chord_coeffs(x2, y2, update))

    def add_q(self, sign):
        """chord through R and +-Q, R <- R +- Q (miller_loop_native.rs:160-168)"""
        return self.add_step(self.Qx, self.Qy if sign == 1 else self.neg_qy())

    def frobenius_coeffs(self):
        """the two Frobenius endpoints (:176-187, :266-280): chords through pi(Q) and -pi^2(Q)"""
        b = self.b
        q1x = b.const(_C2) * self.Qx.conj()
        q1y = b.const(_C3) * self.Qy.conj()
        c1 = self.chord_coeffs(q1x, q1y)
        q2x = b.const(_C2) * q1x.conj()
        nq2y = -(b.const(_C3) * q1y.conj())
        c2 = self.chord_coeffs(q2x, nq2y, update=False)
        return c1, c2

    def frobenius_steps(self):
        """the same two chords, each evaluated as soon as it exists (keeps the live program's instruction order)"""
        b = self.b
        q1x = b.const(_C2) * self.Qx.conj()
        q1y = b.const(_C3) * self.Qy.conj()
        yield self.add_step(q1x, q1y)
        q2x = b.const(_C2) * q1x.conj()
        nq2y = -(b.const(_C3) * q1y.conj())
        yield self.add_step(q2x, nq2y, update=False)


def line_schedule():
    """The lines of one Miller loop in order: 'T' tangent, +1 / -1 chord through +-Q, 'F' the two Frobenius chords -
    the order in which g2_prepare stores coefficient triples and a prepared pair reads them."""
    naf = SIX_U_PLUS_2_NAF
    top = len(naf) - 1
    while naf[top] == 0:
        top -= 1
    out = ["T"]
    for i in range(top - 1, -1, -1):
        if i != top - 1:
            out.append("T")
        if naf[i]:
            out.append(naf[i])
    return out + ["F", "F"]


PREP_LINES = len(line_schedule())       # coefficient triples per prepared G2 point
PREP_FQ = PREP_LINES * 6                # Fq elements per prepared G2 point (include/bnp.h: BNP_PREP_FQ)


class _PreparedPair:
    """A pair whose G2 side arrives as line coefficients (program g2_prepare, the engine's `G2Prepared`): no point
    arithmetic, every step is three loads and two Fq-scalar products."""

    def __init__(self, b, pair_index, prep_index):
        k = pair_index
        self.b = b
        self.Pv = b.ldg(isa.ARR_G1, 2 * k, 2 * k + 1)
        self.Pn = -b.times(self.Pv, 3)
        self.Pm = -self.Pv
        self.base = prep_index * PREP_FQ
        self.line = 0
        self.Z = None

    def _next(self):
        f = self.base + 6 * self.line
        self.line += 1
        return tuple(self.b.ldg(isa.ARR_AUX, f + 2 * j, f + 2 * j + 1) for j in range(3))

    eval_tangent = _Pair.eval_tangent
    eval_chord = _Pair.eval_chord

    def double_step(self):
        return self.eval_tangent(self._next())

    def add_q(self, sign):
        return self.eval_chord(self._next())

    def frobenius_steps(self):
        yield self.eval_chord(self._next())
        yield self.eval_chord(self._next())


def _miller_core(b, n_pairs, track_scale, pairs=None):
    """Shared-squaring Miller loop over n_pairs pairs (n_pairs = 1 is miller_loop_BN_native).
    Returns (g, S, exp_xi, exp_w): g = f_ref * S * xi^exp_xi * w^-exp_w  with S in Fq2 (None when untracked).
    `pairs`: pair objects (live `_Pair`s or `_PreparedPair`s); default: n_pairs live pairs."""
    naf = SIX_U_PLUS_2_NAF
    top = len(naf) - 1
    while naf[top] == 0:
        top -= 1
    assert naf[top] == 1  # multi_miller_loop_BN_native asserts this (:201); R starts at +Q
    if pairs is None:
        pairs = [_Pair(b, k) for k in range(n_pairs)]
    assert not track_scale or all(isinstance(pr, _Pair) for pr in pairs)

    exp_xi, exp_w = 0, 0
    S = None

    def scale_tangent(pr):
        nonlocal S
        if track_scale:
            z2 = pr.Z.sqr()
            S = z2 if S is None else S * z2

    def scale_chord(pr):
        nonlocal S
        if track_scale:
            S = S * pr.Z

    # initial f = product of the tangents at R = Q (miller_loop_native.rs:127-149, :206-233)
    g = None
    for pr in pairs:
        scale_tangent(pr)
        e0, e1, e3 = pr.double_step()
        exp_xi += 1
        exp_w += 3
        if g is None:
            zero = b.const((0, 0))
            g = [e0, e1, zero, e3, zero, zero]
        else:
            g = b.fq12_mul_034(g, e0, e1, e3)

    i = top - 1
    first = True
    while True:
        if not first:
            # f <- f^2, then one tangent per pair (:152-155, :238-243); R <- 2R (:157, :244-246)
            b.cut()
            g = b.fq12_sqr(g)
            exp_xi *= 2
            exp_w *= 2
            if track_scale:
                S = S.sqr()
            for pr in pairs:
                scale_tangent(pr)
                e0, e1, e3 = pr.double_step()
                exp_xi += 1
                exp_w += 3
                g = b.fq12_mul_034(g, e0, e1, e3)
        first = False
        if naf[i] != 0:
            if i < 8:
                b.cut()   # finer candidates near the end: the last phases of a split program taper off (phases.py)
            for pr in pairs:
                scale_chord(pr)
                e0, e1, e3 = pr.add_q(naf[i])
                exp_w += 2
                g = b.fq12_mul_034(g, e0, e1, e3)
        if i == 0:
            break
        i -= 1

    # Frobenius endpoints (:176-187, :266-280)
    for pr in pairs:
        b.cut()
        if track_scale:
            # the scale of the first chord is Z before it, of the second Z after it
            q1x = b.const(_C2) * pr.Qx.conj()
            q1y = b.const(_C3) * pr.Qy.conj()
            scale_chord(pr)
            e0, e1, e3 = pr.add_step(q1x, q1y)
            exp_w += 2
            g = b.fq12_mul_034(g, e0, e1, e3)
            b.cut()
            q2x = b.const(_C2) * q1x.conj()
            nq2y = -(b.const(_C3) * q1y.conj())
            scale_chord(pr)
            e0, e1, e3 = pr.add_step(q2x, nq2y, update=False)
            exp_w += 2
            g = b.fq12_mul_034(g, e0, e1, e3)
        else:
            for e0, e1, e3 in pr.frobenius_steps():
                exp_w += 2
                g = b.fq12_mul_034(g, e0, e1, e3)
    return g, S, exp_xi, exp_w


def _miller_exact(b, n_pairs):
    """Miller loop whose output equals the reference's bit for bit."""
    g, S, exp_xi, exp_w = _miller_core(b, n_pairs, track_scale=True)
    # f_ref = g * w^exp_w / (S * xi^exp_xi);  w^exp_w = xi^(exp_w div 6) * w^(exp_w mod 6)
    order = P * P - 1
    k = c_pow(XI, (exp_w // 6 - exp_xi) % order)
    b.cut()
    corr = S.inv() * b.const(k)
    b.cut()
    g = b.fq12_rot(g, exp_w % 6)
    return b.fq12_mul_fq2(g, corr)


def _easy_part(b, a):
    """final_exp_native.rs:195-206"""
    f2 = b.fq12_mul(b.fq12_conj(a), b.fq12_inv(a))
    f3 = b.fq12_frobenius(f2, 2)
    return b.fq12_mul(f3, f2)


def _hard_part_ref(b, m, tap=None):
    """final_exp_native.rs:130-169, in the cyclotomic subgroup.  `tap(name, value)` is called with the three
    BN_X powers as soon as each exists (witness program: they are stored and their slots freed early)."""
    tap = tap or (lambda name, v: None)
    mp = b.fq12_frobenius(m, 1)
    mp2 = b.fq12_frobenius(m, 2)
    mp3 = b.fq12_frobenius(m, 3)
    y0 = b.fq12_mul(mp, b.fq12_mul(mp2, mp3))
    mx = b.fq12_pow_x_cyclo(m)
    tap("mx", mx)
    mxp = b.fq12_frobenius(mx, 1)
    mx2 = b.fq12_pow_x_cyclo(mx)
    tap("mx2", mx2)
    mx2p = b.fq12_frobenius(mx2, 1)
    y2 = b.fq12_frobenius(mx2, 2)
    mx3 = b.fq12_pow_x_cyclo(mx2)
    tap("mx3", mx3)
    mx3p = b.fq12_frobenius(mx3, 1)
    # y1 = conj(m), y3 = conj(mxp), y4 = conj(mx * mx2p), y5 = conj(mx2), y6 = conj(mx3 * mx3p)
    # (cut candidates between the steps of the vectorial chain: the LAST phase of a split program wants to be short,
    # phases.py)
    y4c = b.fq12_mul(mx, mx2p)       # conj(y4)
    b.cut()
    y6c = b.fq12_mul(mx3, mx3p)      # conj(y6)
    b.cut()
    # T0 = y6^2 * y4 * y5 = conj(y6c^2 * y4c * mx2)
    T0c = b.fq12_cyclo_sqr(y6c)
    T0c = b.fq12_mul(T0c, y4c)
    b.cut()
    T0c = b.fq12_mul(T0c, mx2)
    b.cut()
    # T1 = y3 * y5 * T0 = conj(mxp * mx2 * T0c)
    T1c = b.fq12_mul(b.fq12_mul(mxp, mx2), T0c)
    b.cut()
    # T0 = y2 * T0
    T0 = b.fq12_mul_conj(y2, T0c)
    b.cut()
    T1 = b.fq12_conj(b.fq12_cyclo_sqr(T1c))   # T1^2
    T1 = b.fq12_mul(T1, T0)
    b.cut()
    T1 = b.fq12_cyclo_sqr(T1)
    T0 = b.fq12_mul_conj(T1, m)               # T1 * y1
    b.cut()
    T1 = b.fq12_mul(T1, y0)
    b.cut()
    T0 = b.fq12_cyclo_sqr(T0)
    return b.fq12_mul(T0, T1)


def _hard_part_ark(b, e):
    """ark-ec 0.4.2 models/bn final_exponentiation hard part (Fuentes-Castaneda), restated; equals the
    reference's result raised to 2x(6x^2+3x+1).  f^(-x) = conj(f^x) since x > 0."""
    y0c = b.fq12_pow_x_cyclo(e)                # conj(y0)
    y1c = b.fq12_cyclo_sqr(y0c)                # conj(y1)
    y2c = b.fq12_cyclo_sqr(y1c)
    y3c = b.fq12_mul(y2c, y1c)                 # conj(y3 before its own conjugation) => equals final y3
    y3 = y3c
    # y4 = (y3_old)^(-x) = conj(y3_old^x) = conj(conj(y3c)^x) = y3c^x
    y4 = b.fq12_pow_x_cyclo(y3c)
    y5 = b.fq12_cyclo_sqr(y4)
    y6c_old = b.fq12_pow_x_cyclo(y5)           # y5^x = conj(y6_old)  => final y6 = conj(y6_old) = y5^x
    y6 = y6c_old
    y7 = b.fq12_mul(y6, y4)
    b.cut()
    y8 = b.fq12_mul(y7, y3)
    b.cut()
    y9 = b.fq12_mul_conj(y8, y1c)              # y8 * y1
    y10 = b.fq12_mul(y8, y4)
    b.cut()
    y11 = b.fq12_mul(y10, e)
    y12 = b.fq12_frobenius(y9, 1)
    b.cut()
    y13 = b.fq12_mul(y12, y11)
    y8f = b.fq12_frobenius(y8, 2)
    b.cut()
    y14 = b.fq12_mul(y8f, y13)
    b.cut()
    y15 = b.fq12_frobenius(b.fq12_mul_conj(y9, e), 3)
    return b.fq12_mul(y15, y14)


def _final_exp(b, a, variant):
    b.cut()
    m = _easy_part(b, a)
    b.cut()
    return _hard_part_ref(b, m) if variant == 0 else _hard_part_ark(b, m)


# ----------------------------------------------------------------------------- program table
def prog_miller(b, n_pairs=1):
    """arr G1/G2 -> OUT: reference-exact (multi_)miller_loop_native."""
    b.st_fq12(isa.ARR_OUT, _miller_exact(b, n_pairs))


def prog_miller_fused(b, n_pairs=1):
    """arr G1/G2 -> OUT: Miller value up to a proper-subfield factor (input to a later final exp only)."""
    g, _, _, _ = _miller_core(b, n_pairs, track_scale=False)
    b.st_fq12(isa.ARR_OUT, g)


def prog_final_exp(b, variant):
    """arr F12 -> OUT: final_exp_native (variant 0) or the ark-compatible exponent (variant 1)."""
    b.st_fq12(isa.ARR_OUT, _final_exp(b, b.ld_fq12(isa.ARR_F12), variant))


# Fq offsets of the five MyFq12 values in the witness program's output (include/bnp.h: BNP_WITNESS_*)
WITNESS_LAYOUT = {"m": 0, "mx": 12, "mx2": 24, "mx3": 36, "out": 48}
WITNESS_FQ = 60


def prog_final_exp_witness(b):
    """arr F12 -> OUT[60 Fq]: every native value the final-exponentiation CIRCUIT takes from the CPU today
    (SURVEY 8(f).1): m = easy part (final_exp_target.rs:152-161 recomputes it in-circuit, the value is the input of
    the first exponentiation stark), m^x, m^(x^2), m^(x^3) = the outputs of the three Fq12ExpU64 starks with
    offset 1 (final_exp_target.rs:89-117), and final_exp_native(a) itself (the expected public output,
    final_exp_target.rs:240)."""
    a = b.ld_fq12(isa.ARR_F12)
    b.cut()
    m = _easy_part(b, a)
    b.st_fq12(isa.ARR_OUT, m, base=WITNESS_LAYOUT["m"])
    b.cut()
    out = _hard_part_ref(b, m, tap=lambda name, v: b.st_fq12(isa.ARR_OUT, v, base=WITNESS_LAYOUT[name]))
    b.st_fq12(isa.ARR_OUT, out, base=WITNESS_LAYOUT["out"])


def prog_pairing(b, variant, n_pairs=1):
    """arr G1/G2 -> OUT: pairing.rs:20-22 fused in one launch (n_pairs > 1: product of pairings)."""
    g, _, _, _ = _miller_core(b, n_pairs, track_scale=False)
    b.st_fq12(isa.ARR_OUT, _final_exp(b, g, variant))


def prog_g2_prepare(b):
    """arr G2 -> OUT[PREP_FQ Fq]: the line coefficients of one G2 point in `line_schedule()` order - the engine's
    `G2Prepared` (ark-ec's `G2Prepared::from` builds the same list in its own normalisation)."""
    pr = _Pair(b, 0, with_p=False)
    line = 0

    def put(c):
        nonlocal line
        for j in range(3):
            b.stg(isa.ARR_OUT, 6 * line + 2 * j, 6 * line + 2 * j + 1, c[j])
        line += 1

    for n_done, step in enumerate(line_schedule()):
        if step == "T":
            if n_done:
                b.cut()
            put(pr.tangent_coeffs())
        elif step == "F":
            if line == PREP_LINES - 2:
                c1, c2 = pr.frobenius_coeffs()
                put(c1)
                put(c2)
        else:
            put(pr.chord_coeffs(pr.Qx, pr.Qy if step == 1 else pr.neg_qy()))
    assert line == PREP_LINES


def prog_pairing_prepared(b, variant, kv, kp):
    """arr G1 (kv + kp points), G2 (kv points), AUX (kp prepared points) -> OUT: the product of kv + kp pairings whose
    last kp G2 points arrive as line coefficients - a Groth16 verifier's fixed verifying-key points: their point
    arithmetic is done once, not once per proof."""
    pairs = [_Pair(b, k) for k in range(kv)] + [_PreparedPair(b, kv + j, j) for j in range(kp)]
    g, _, _, _ = _miller_core(b, kv + kp, track_scale=False, pairs=pairs)
    b.st_fq12(isa.ARR_OUT, _final_exp(b, g, variant))


def prog_frobenius(b, power):
    b.st_fq12(isa.ARR_OUT, b.fq12_frobenius(b.ld_fq12(isa.ARR_F12), power))


def prog_fq12_mul(b):
    """OUT = F12 * AUX (MyFq12 `Mul`); used for product reductions."""
    b.st_fq12(isa.ARR_OUT, b.fq12_mul(b.ld_fq12(isa.ARR_F12), b.ld_fq12(isa.ARR_AUX)))


def _pow_naf_general(b, a, e):
    """pow_native (final_exp_native.rs:56-84) for a build-time exponent and ANY Fq12 input (the reference's own test
    feeds a non-cyclotomic element, :266-273): left-to-right walk over the NAF digits of `e`, general squarings, and
    the reference's `res / a` for a -1 digit done as a multiplication by a^-1, computed once (field division is
    exact, so the bits are the same)."""
    naf = naf_digits(e)
    ainv = b.fq12_inv(a) if -1 in naf else None
    res = None
    for z in reversed(naf):
        if res is not None:
            b.cut()
            res = b.fq12_sqr(res)
        if z:
            m = a if z == 1 else ainv
            res = m if res is None else b.fq12_mul(res, m)
    return res


def prog_pow_bnx(b):
    """OUT = F12 ^ BN_X for general F12 (pow_native(a, vec![BN_X]))."""
    b.st_fq12(isa.ARR_OUT, _pow_naf_general(b, b.ld_fq12(isa.ARR_F12), BN_X))


def prog_fq12_sqr(b):
    """OUT = F12^2 (general element): one step of the run-time exponent walk of bnp_pow_u64_batch."""
    b.st_fq12(isa.ARR_OUT, b.fq12_sqr(b.ld_fq12(isa.ARR_F12)))


def prog_fq12_inv(b):
    """OUT = 1 / F12 (zero maps to zero; the reference panics on a division by zero)."""
    b.st_fq12(isa.ARR_OUT, b.fq12_inv(b.ld_fq12(isa.ARR_F12)))


# ----------------------------------------------------------------------------- input validation
_B_TWIST = c_mul((3, 0), c_inv(XI))          # b' = 3 / (9 + u): the twist is y^2 = x^3 + b'
_B3_TWIST = c_mul((3, 0), _B_TWIST)
SIX_X_SQUARED = 6 * BN_X * BN_X


def _g2_double_complete(b, X, Y, Z):
    """Renes-Costello-Batina complete doubling for a = 0 (Algorithm 9 of eprint 2015/1060), homogeneous projective."""
    b3 = b.const(_B3_TWIST)
    t0 = Y.sqr()
    Z3 = b.times(t0, 8)
    t1 = Y * Z
    t2 = Z.sqr() * b3
    X3 = t2 * Z3
    Y3 = t0 + t2
    Z3 = t1 * Z3
    t0 = t0 - b.times(t2, 3)
    Y3 = X3 + t0 * Y3
    X3 = (t0 * (X * Y)).dbl()
    return X3, Y3, Z3


def _g2_add_mixed_complete(b, X1, Y1, Z1, x2, y2):
    """Complete mixed addition for a = 0 (Algorithm 8 of eprint 2015/1060): (X1 : Y1 : Z1) + (x2, y2, 1).  No exceptional
    cases on a curve without points of order two - the twist has odd order r (2p - r) - so the straight-line program
    is right for EVERY input on the twist, including points outside the r-torsion and the identity as accumulator."""
    b3 = b.const(_B3_TWIST)
    t0 = X1 * x2
    t1 = Y1 * y2
    t3 = (x2 + y2) * (X1 + Y1) - (t0 + t1)
    t4 = y2 * Z1 + Y1
    Y3 = x2 * Z1 + X1
    t0 = b.times(t0, 3)
    t2 = Z1 * b3
    Z3 = t1 + t2
    t1 = t1 - t2
    Y3 = Y3 * b3
    X3 = t3 * t1 - t4 * Y3
    Y3 = t1 * Z3 + Y3 * t0
    Z3 = Z3 * t4 + t0 * t3
    return X3, Y3, Z3


def prog_validate_g1(b):
    """arr G1 -> OUT[1 Fq2]: (y^2 - x^3 - 3, same) - zero iff the point is on the curve (G1 has cofactor 1, so that is
    the whole of `G1Affine::new`'s check).  The coordinates are loaded twice over so that the Fq arithmetic runs in both
    halves of the Fq2 slots."""
    X = b.ldg(isa.ARR_G1, 0, 0)
    Y = b.ldg(isa.ARR_G1, 1, 1)
    x3 = X.mulfp(X, 0).mulfp(X, 0)
    b.stg(isa.ARR_OUT, 0, 1, Y.mulfp(Y, 0) - x3 - b.const((3, 3)))


def prog_validate_g2(b):
    """arr G2 -> OUT[3 Fq2]: three residuals that are all zero iff `G2Affine::new(x, y)` would accept the point
    (the hidden assertion behind miller_loop_native.rs:303,311): on the twist, and in the r-torsion subgroup by
    ark-bn254 0.4's own test [6 x^2] Q == psi(Q) (eprint 2022/352, section 4.3), psi = the twisted Frobenius of
    miller_loop_native.rs:298-304.  [6 x^2] Q is a fixed NAF walk with complete projective formulas."""
    x = b.ldg(isa.ARR_G2, 0, 1)
    y = b.ldg(isa.ARR_G2, 2, 3)
    on_curve = y.sqr() - x.sqr() * x - b.const(_B_TWIST)
    b.stg(isa.ARR_OUT, 0, 1, on_curve)
    ny = -y
    naf = naf_digits(SIX_X_SQUARED)
    X, Y, Z = x, y, b.const((1, 0))
    assert naf[-1] == 1
    for z in reversed(naf[:-1]):
        b.cut()
        X, Y, Z = _g2_double_complete(b, X, Y, Z)
        if z:
            X, Y, Z = _g2_add_mixed_complete(b, X, Y, Z, x, y if z == 1 else ny)
    px = b.const(_C2) * x.conj()
    py = b.const(_C3) * y.conj()
    b.stg(isa.ARR_OUT, 2, 3, X - px * Z)
    b.stg(isa.ARR_OUT, 4, 5, Y - py * Z)


OPTEST_OUTPUTS = 36


def prog_optest(b):
    """Every opcode and operand form once, for op-level GPU parity tests.  F12 holds 6 input slots x0..x5;
    OUT gets OPTEST_OUTPUTS slots.  The expected values are spelled out in tests/optest_expect.py."""
    x = b.ld_fq12(isa.ARR_F12)
    outs = [
        x[0] * x[1], x[2].sqr(), x[3].mulfp(x[4], 0), x[3].mulfp(x[4], 1), x[0] + x[5], x[1] - x[2],
        -x[3], x[4].conj(), x[5].mulxi(), x[0].dbl(), x[1].inv(), x[2] * b.const(_C3),
        (x[0] + x[1]) * (x[2] + x[3]), (x[0] - x[1]) * (x[2] - x[3]), (x[0] + x[1]) * x[2], x[0] * (x[2] - x[3]),
        (x[4] + x[5]).sqr(), (x[4] - x[5]).sqr(),
        x[0] - x[1] - x[2] + x[3],
        x[0] + (x[1] - x[2] - x[3]).mulxi(),
        b.times(x[0], 27) - x[1].mulxi(),
        (x[2] - x[3]).dbl() + x[2],
        x[0].conj() + x[1].mulxi() - x[2].dbl(),
        b.times(x[5], 12) - x[4].conj(),
        # product-class epilogues: hi terms (1, 2 and 3 of them, lazy and canonical operands), post LIN with and
        # without xi, r' stored next to the post value, SQR / MULFP producers
        x[0] * x[1] - x[2],
        (x[0] + x[1]) * (x[2] + x[3]) - x[4] - x[5],
        (x[0] + x[1]) * (x[2] + x[3]) - x[4] - x[5] + x[0],
        x[2] * x[3] + x[0] - x[1] + x[5],
        x[4] + (x[0] * x[5] - x[1] - x[2]).mulxi(),
        (x[1] + x[2]) * (x[3] - x[4]) - x[0] - x[5] + x[2].mulxi(),
        b.times(x[3].sqr(), 3) - x[4].dbl(),
        b.times((x[0] + x[1]).sqr() - x[2] - x[3], 3) + x[5].dbl(),
        x[1].mulfp(x[2], 1) - x[3] + x[4],
        (x[0] - x[1]) * x[2] - x[3] - x[4] - x[5],
    ]
    hp = x[1] * x[4] - x[0] - x[2]          # hi-form value with two readers: stored AND fed to a post stage
    outs += [hp, x[3] + hp.mulxi()]
    assert len(outs) == OPTEST_OUTPUTS
    for i, v in enumerate(outs):
        b.stg(isa.ARR_OUT, 2 * i, 2 * i + 1, v)


def prog_opbench(b, op, count=2048):
    """`count` back-to-back instances of one opcode over a rotating set of slots (cost-model tool)."""
    x = b.ld_fq12(isa.ARR_F12)
    vals = list(x)
    for i in range(count):
        a, c = vals[i % 6], vals[(i + 1) % 6]
        if op == "MUL":
            r = a * c
        elif op == "SQR":
            r = a.sqr()
        elif op == "MULFP":
            r = a.mulfp(c, i & 1)
        elif op == "ADD":
            r = a + c
        elif op == "SUB":
            r = a - c
        elif op == "DBL":
            r = a.dbl()
        elif op == "NEG":
            r = -a
        elif op == "MULXI":
            r = a.mulxi()
        elif op == "LIN4":
            r = a - c - vals[(i + 2) % 6] + vals[(i + 3) % 6]
        elif op == "LIN4XI":
            r = a + (c - vals[(i + 2) % 6] - vals[(i + 3) % 6]).mulxi()
        elif op == "MULS":
            r = (a + c) * (vals[(i + 2) % 6] + vals[(i + 3) % 6])
        elif op == "MIX":
            # the pairing's opcode proportions (MUL 4 : SQR 2 : ADD/SUB 12 : MULXI 2 : DBL 2), on six slots
            k = i % 22
            if k < 4:
                r = a * c
            elif k < 6:
                r = a.sqr()
            elif k < 12:
                r = a + c
            elif k < 18:
                r = a - c
            elif k < 20:
                r = a.mulxi()
            else:
                r = a.dbl()
        else:
            raise ValueError(op)
        vals[i % 6] = r
    b.st_fq12(isa.ARR_OUT, vals)


# (pairs with a live G2 point, pairs with a prepared one): a single prepared pairing, and the Groth16 shape
# e(A, B) e(L, gamma) e(C, delta) e(alpha, beta) with B per proof and gamma, delta, beta from the verifying key
PREPARED_SHAPES = [(0, 1), (1, 2), (1, 3)]

PROGRAMS = [
    # (name, build function, kwargs)
    ("miller", prog_miller, {}),
    ("miller_fused", prog_miller_fused, {}),
    ("final_exp_v0", prog_final_exp, {"variant": 0}),
    ("final_exp_v1", prog_final_exp, {"variant": 1}),
    ("final_exp_witness", prog_final_exp_witness, {}),
    ("pairing_v0", prog_pairing, {"variant": 0}),
    ("pairing_v1", prog_pairing, {"variant": 1}),
    ("fq12_mul", prog_fq12_mul, {}),
    ("fq12_sqr", prog_fq12_sqr, {}),
    ("fq12_inv", prog_fq12_inv, {}),
    ("pow_bnx", prog_pow_bnx, {}),
    ("validate_g1", prog_validate_g1, {}),
    ("validate_g2", prog_validate_g2, {}),
    ("optest", prog_optest, {}),
] + [("opbench_" + o.lower(), prog_opbench, {"op": o}) for o in ("MUL", "SQR", "MULFP", "ADD", "SUB", "DBL", "NEG", "MULXI", "LIN4", "LIN4XI", "MULS", "MIX")] \
  + [("frobenius_%d" % k, prog_frobenius, {"power": k}) for k in range(12)] \
  + [("miller_x%d" % k, prog_miller, {"n_pairs": k}) for k in (2, 3, 4)] \
  + [("pairing_x%d_v%d" % (k, v), prog_pairing, {"variant": v, "n_pairs": k}) for k in (2, 3, 4) for v in (0, 1)] \
  + [("g2_prepare", prog_g2_prepare, {})] \
  + [("pairing_p%d_%d_v%d" % (kv, kp, v), prog_pairing_prepared, {"variant": v, "kv": kv, "kp": kp})
     for kv, kp in PREPARED_SHAPES for v in (0, 1)]


def build_program(name, pool):
    for n, fn, kw in PROGRAMS:
        if n == name:
            b = Builder(pool)
            fn(b, **kw)
            return b
    raise KeyError(name)
