"""Fusion pass: SSA program of elementary Fq2 ops  ->  SSA program of product-class instructions
(MUL / SQR / MULFP with pre-additions, hi terms and a post LIN stage), INV, loads/stores and the few linear
instructions that could not be attached to a product.

Why (measured on B200, profiles/): a stand-alone linear opcode costs ~140-240 sub-partition cycles per warp-op,
nearly all of it dispatch, slot traffic and the serial carry chain of the modular step, and the tower formulas contain ~2.6 of them per multiplication.  Almost every linear value in a
pairing is `product +- older values` (Karatsuba recombination, v-multiplication, 3x-2y in the cyclotomic
squaring, the Fq12-level even/odd halves), so it is computed in the epilogue of the product that completes it:

  * hi terms   r' = redc(T + 2^256 * sum(+-s_i)):  canonical slots added into the upper half of the wide
               product before the Montgomery reduction - eight adds per component, no extra reduction;
  * post LIN   v = LIN(r', slots ...): a lazily accumulated linear combination (small integer coefficients,
               optional xi = 9 + u factor) of r' - still in registers - and older slots, reduced once.

Every Fq-linear map of an Fq2 value x = (x0, x1) is a 2x2 integer matrix; the pass tracks one matrix per leaf
and decomposes  M = diag(p, q) + XI * diag(r, s),  XI = [[9,-1],[1,9]]  (always possible and unique), which is
exactly what the kernel's LIN core evaluates.

Range contracts (checked here for ALL inputs, the kernel does not check):
  * wide values stay below 2^512 - p 2^256 = 4.29 p 2^256 so that redc's result fits eight limbs;
  * the result bound selects the number of conditional subtractions (`canon` levels, per component);
  * a LIN accumulator plus K p lies in [0, 1024 p).
"""
from . import isa

LINEAR = ("ADD", "SUB", "NEG", "DBL", "MULXI", "CONJ", "MOV", "MULK")
I2 = (1, 0, 0, 1)
NEG_I2 = (-1, 0, 0, -1)
U = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47 / 2.0 ** 256   # p / 2^256
WIDE_LIMIT = 2.0 ** 256 / 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47 - 1.0 - 1e-6
MAX_SRCS = 10          # distinct source slots one instruction may name (slot budget is 14)


def m_add(a, b):
    return tuple(x + y for x, y in zip(a, b))


def m_scale(a, k):
    return tuple(k * x for x in a)


def m_xi(a):  # XI * a
    return (9 * a[0] - a[2], 9 * a[1] - a[3], a[0] + 9 * a[2], a[1] + 9 * a[3])


def m_conj(a):  # diag(1,-1) * a
    return (a[0], a[1], -a[2], -a[3])


def decompose(M):
    """M -> ((p, q), (r, s)) with M = diag(p,q) + XI*diag(r,s)."""
    s = -M[1]
    r = M[2]
    return (M[0] - 9 * r, M[3] - 9 * s), (r, s)


class FOp:
    __slots__ = ("op", "dst", "srcs", "imm", "f_lo", "f_hi", "terms", "K", "flags",
                 "hi", "post", "dst2", "store_r", "canon")

    def __init__(self, op, dst, srcs, imm=0, f_lo=0, f_hi=0, terms=None, K=0, flags=0):
        self.op, self.dst, self.srcs, self.imm, self.f_lo, self.f_hi = op, dst, list(srcs), imm, f_lo, f_hi
        self.terms, self.K, self.flags = terms, K, flags   # terms: (ent0, ent1) of a LIN
        # product-class epilogue
        self.hi = []          # [(value, negate)]
        self.post = None      # (ent0, ent1); a leaf of None names r'
        self.dst2 = None      # value produced by the post stage
        self.store_r = False  # r' (= dst) is written to a slot (set by the first reader / the final fix-up)
        self.canon = (0, 0)

    def all_srcs(self):
        s = list(self.srcs) + [v for v, _ in self.hi]
        if self.post:
            s += [t[0] for t in self.post[0] + self.post[1] if t[0] is not None]
        return s

    def dsts(self):
        out = []
        if self.dst is not None and (self.op not in isa.PRODUCT_OPS or self.store_r):
            out.append(self.dst)
        if self.dst2 is not None:
            out.append(self.dst2)
        return out


def wide_bounds(op, flags):
    """Upper bounds of (T0, T1) in units of p 2^256 for canonical slot inputs."""
    if op == "MUL":
        la = 2 if (flags & isa.MUL_B and not flags & isa.MUL_BCANON) else 1
        lb = 2 if flags & isa.MUL_E else 1
        return (1.0, 2 * la * lb * U)   # T0 = x0 y0 - x1 y1 (+ p 2^256 if negative) < p 2^256, T1 = x0 y1 + x1 y0
    if op == "SQR":
        return (2 * U, 2 * U)
    if op == "MULFP":
        return (U, U)
    raise ValueError(op)


def canon_levels(op, flags, n_hi):
    lv = []
    for b in wide_bounds(op, flags):
        tot = b + n_hi
        assert tot < WIDE_LIMIT, "wide value out of range"
        r_bound = tot + 1.0          # redc adds < p
        lv.append(0 if r_bound <= 2.0 else 1 if r_bound <= 4.0 else 2)
    return tuple(lv)


def max_hi_terms(op, flags):
    b = max(wide_bounds(op, flags))
    n = 0
    while n < 3 and b + n + 1 < WIDE_LIMIT:
        n += 1
    return n


def entries_of(expr):
    """expr: {leaf: M (2x2 integer matrix, row-major)} -> (ent0, ent1) or None if not encodable.
    ent_c = [(leaf, half, mult, neg)]: output component c = sum of mult * (neg ? p - z : z), z = leaf[half]."""
    ents = ([], [])
    for leaf, M in expr.items():
        for c in (0, 1):
            for half in (0, 1):
                k = M[2 * c + half]
                if k == 0:
                    continue
                if abs(k) > isa.LIN_MAX_MULT:
                    return None
                ents[c].append((leaf, half, abs(k), k < 0))
    for e in ents:
        if len(e) > isa.LIN_MAX_ENT or sum(t[2] for t in e) > isa.LIN_MAX_SUM:
            return None
    return ents


def fuse(ops, enable=True, attach=True, max_srcs=MAX_SRCS):
    """enable=False: no fusion at all (every linear op stays elementary, no pre-additions) - the baseline
    lowering the tests compare against.  attach=False: pre-additions and LIN trees only."""
    uses = {}
    user = {}
    defop = {}
    for o in ops:
        if o.dst is not None:
            defop[o.dst] = o
        for s in o.srcs:
            uses[s] = uses.get(s, 0) + 1
            user[s] = o

    def deferrable(v):
        o = defop[v]
        if not enable or o.op not in LINEAR or uses.get(v, 0) != 1:
            return False
        u = user[v]
        if u.op in LINEAR:
            return True
        return u.op in ("MUL", "SQR") and o.op in ("ADD", "SUB") and u.srcs.count(v) == 1

    out = []
    alias = {}
    pos = {}         # materialised value -> index in `out` of the instruction producing it
    prim_of = {}     # value that is the r' of a product instruction -> that FOp
    pending = {}     # deferred linear value -> its Op

    def res(v):
        while v in alias:
            v = alias[v]
        return v

    def expand(v):
        """Linear expression of v over materialised leaves: {leaf: M}."""
        v = res(v)
        if v not in pending:
            return {v: I2}
        o = pending[v]
        if o.op == "ADD":
            parts = [(o.srcs[0], 1), (o.srcs[1], 1)]
        elif o.op == "SUB":
            parts = [(o.srcs[0], 1), (o.srcs[1], -1)]
        elif o.op == "NEG":
            parts = [(o.srcs[0], -1)]
        elif o.op == "DBL":
            parts = [(o.srcs[0], 2)]
        elif o.op == "MULK":
            parts = [(o.srcs[0], o.imm)]
        else:  # MULXI, CONJ, MOV
            parts = [(o.srcs[0], 1)]
        expr = {}
        for src, k in parts:
            for leaf, M in expand(src).items():
                M = m_scale(M, k)
                expr[leaf] = m_add(expr[leaf], M) if leaf in expr else M
        if o.op == "MULXI":
            expr = {l: m_xi(M) for l, M in expr.items()}
        elif o.op == "CONJ":
            expr = {l: m_conj(M) for l, M in expr.items()}
        return {l: M for l, M in expr.items() if any(M)}

    def consume(v):
        """Drop v and every deferred value inlined beneath it from `pending`."""
        o = pending.pop(v)
        for s in o.srcs:
            s = res(s)
            if s in pending:
                consume(s)

    def touch(v):
        """v is read from a slot by some instruction: its producer must store it."""
        f = prim_of.get(v)
        if f is not None:
            f.store_r = True
        return v

    touched = set()

    def touch_t(v):
        touched.add(v)
        return touch(v)

    def try_attach(v, expr):
        """Compute linear value v in the epilogue of the product instruction that completes it."""
        if not attach or len(expr) < 1:
            return False
        leaves = list(expr)
        if any(l not in pos for l in leaves):
            return False
        last = max(leaves, key=lambda l: pos[l])
        f = prim_of.get(last)
        if f is None or f.post is not None:
            return False
        k = pos[last]
        others = [l for l in leaves if l is not last]
        if any(pos[l] >= k for l in others):
            return False
        n_srcs = len(set(f.all_srcs()) | set(others))
        if n_srcs > max_srcs:
            return False
        # hi terms: leaves whose matrix equals +-M_r ride through the Montgomery reduction with r
        # (v = M_r (r +- s1 +- s2 +- s3) + rest); needs the raw r not to be wanted by anything else
        Mr = expr[last]
        negMr = m_scale(Mr, -1)
        hi = []
        flags = f.flags
        if not f.hi and uses.get(last, 0) == 1 and last not in touched:
            cand = [l for l in others if expr[l] in (Mr, negMr)]
            room = max_hi_terms(f.op, flags)
            if len(cand) > room and f.op == "MUL" and flags & isa.MUL_B:
                # both operands are lazy sums: canonicalising the first one (one conditional subtraction
                # per component) makes room for a third hi term
                if max_hi_terms(f.op, flags | isa.MUL_BCANON) > room:
                    flags |= isa.MUL_BCANON
                    room = max_hi_terms(f.op, flags)
            hi = [(l, expr[l] == negMr) for l in cand[:room]]
        rest = {l: expr[l] for l in others if l not in [h[0] for h in hi]}
        if not rest and Mr == I2:
            post = None
        else:
            pexpr = dict(rest)
            pexpr[None] = Mr
            post = entries_of(pexpr)
            if post is None:
                return False
        if hi:
            f.flags = flags
            f.hi = hi
            f.canon = canon_levels(f.op, f.flags, len(hi))
            del prim_of[last]
            del pos[last]
            if post is None:
                f.dst = v               # v itself is the hi-form value
                prim_of[v] = f
                pos[v] = k
                f.store_r = False       # set by the first reader (touch) or by the end-of-pass fix-up
                for l, _ in hi:
                    touch_t(l)
                return True
            f.dst = ("hi", last)        # r' exists only inside this instruction
            last = f.dst
            pos[last] = k
        f.post = post
        f.dst2 = v
        f.store_r = last in touched
        for l in others:
            touch_t(l)
        pos[v] = k
        return True

    def emit_lin(v):
        """Materialise linear value v (removing it from pending)."""
        v = res(v)
        o = pending[v]
        if not enable:
            pending.pop(v)
            if o.op == "MULK":
                src = need(o.srcs[0])
                out.append(FOp("LIN", v, [src], terms=([(src, 0, o.imm, False)], [(src, 1, o.imm, False)])))
            else:
                out.append(FOp(o.op, v, [need(s) for s in o.srcs]))
            pos[v] = len(out) - 1
            return
        expr = expand(v)
        assert expr, "zero linear expression"
        if len(expr) == 1:
            (leaf, M), = expr.items()
            if M == I2:
                consume(v)
                alias[v] = leaf
                return
        if try_attach(v, expr):
            consume(v)
            return
        enc = entries_of(expr)
        if enc is None or len(expr) > max_srcs:
            # too big: materialise deferred children one at a time and retry
            kids = [res(s) for s in o.srcs if res(s) in pending]
            assert kids, "single linear op not encodable"
            for kid in kids:
                emit_lin(kid)
            return emit_lin(v)
        consume(v)
        # single elementary operation on one or two slots: use the dedicated opcode
        elem = _as_elementary(expr)
        if elem is not None:
            opname, srcs = elem
            out.append(FOp(opname, v, [touch_t(s) for s in srcs]))
        else:
            out.append(FOp("LIN", v, [touch_t(l) for l in expr], terms=enc))
        pos[v] = len(out) - 1

    def need(v):
        v = res(v)
        if v in pending:
            emit_lin(v)
            v = res(v)
        return touch_t(v)

    for o in ops:
        if o.op in LINEAR:
            pending[o.dst] = o
            if not deferrable(o.dst):
                emit_lin(o.dst)
            continue
        if o.op in ("MUL", "SQR"):
            operands = []  # (a, b or None, b_negative)
            for s in (o.srcs if o.op == "MUL" else o.srcs[:1]):
                s = res(s)
                if s in pending and pending[s].op in ("ADD", "SUB") and deferrable(s):
                    p = pending.pop(s)
                    a, b = need(p.srcs[0]), need(p.srcs[1])
                    operands.append((a, b, p.op == "SUB"))
                else:
                    operands.append((need(s), None, False))
            flags = 0
            srcs = []
            a, b, bn = operands[0]
            srcs.append(a)
            if b is not None:
                flags |= isa.MUL_B | (isa.MUL_BNEG if bn else 0)
                srcs.append(b)
            if o.op == "MUL":
                c, e, en = operands[1]
                srcs.append(c)
                if e is not None:
                    flags |= isa.MUL_E | (isa.MUL_ENEG if en else 0)
                    srcs.append(e)
            f = FOp(o.op, o.dst, srcs, flags=flags)
            f.canon = canon_levels(o.op, flags, 0)
            out.append(f)
            pos[o.dst] = len(out) - 1
            prim_of[o.dst] = f
            continue
        srcs = [need(s) for s in o.srcs]
        f = FOp(o.op, o.dst, srcs, imm=o.imm, f_lo=o.f_lo, f_hi=o.f_hi)
        if o.op == "MULFP":
            f.canon = canon_levels("MULFP", 0, 0)
            prim_of[o.dst] = f
        out.append(f)
        if o.dst is not None:
            pos[o.dst] = len(out) - 1
    assert not pending, "dangling deferred values"
    for f in out:
        if f.op in isa.PRODUCT_OPS and f.post is None:
            f.store_r = True   # a product without a post stage always writes its result
        if f.op in isa.PRODUCT_OPS and f.post is not None and not f.store_r:
            assert f.dst2 is not None
    return out


def _as_elementary(expr):
    """{leaf: M} -> (opcode, [srcs]) when the expression is exactly one elementary linear operation."""
    items = list(expr.items())
    if len(items) == 1:
        (a, M), = items
        if M == (2, 0, 0, 2):
            return "DBL", [a]
        if M == NEG_I2:
            return "NEG", [a]
        if M == (1, 0, 0, -1):
            return "CONJ", [a]
        if M == m_xi(I2):
            return "MULXI", [a]
        return None
    if len(items) == 2:
        (a, Ma), (b, Mb) = items
        if Ma == I2 and Mb == I2:
            return "ADD", [a, b]
        if Ma == I2 and Mb == NEG_I2:
            return "SUB", [a, b]
        if Ma == NEG_I2 and Mb == I2:
            return "SUB", [b, a]
    return None
