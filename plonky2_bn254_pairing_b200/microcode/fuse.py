"""Fusion pass: SSA program of elementary Fq2 ops  ->  SSA program of MUL/SQR (with pre-additions), MULFP,
INV, loads/stores and LIN (one variable-length linear-combination instruction).

Why: measured on B200 (profiles/opbench_r1.txt) every elementary linear op (add, sub, double, neg, *xi)
costs ~215 SM-sub-partition cycles per warp, independent of occupancy - it moves three 64-byte slots per
thread through shared memory (128 B/clk/SM shared by four sub-partitions) and performs its own modular
reduction.  Trees of linear ops with single-use intermediates are therefore collapsed into ONE instruction
that reads each leaf once, accumulates lazily in registers (signed, unreduced) and reduces once.

Every Fq-linear map of an Fq2 value x = (x0, x1) is a 2x2 integer matrix; the pass tracks one matrix per
leaf and finally decomposes  M = diag(p, q) + XI * diag(r, s),  XI = [[9,-1],[1,9]]  (always possible and
unique), which is exactly what the kernel's LIN handler evaluates.
"""
from . import isa

LINEAR = ("ADD", "SUB", "NEG", "DBL", "MULXI", "CONJ", "MOV")
I2 = (1, 0, 0, 1)


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
    __slots__ = ("op", "dst", "srcs", "imm", "f_lo", "f_hi", "terms", "K", "flags")

    def __init__(self, op, dst, srcs, imm=0, f_lo=0, f_hi=0, terms=None, K=0, flags=0):
        self.op, self.dst, self.srcs, self.imm, self.f_lo, self.f_hi = op, dst, list(srcs), imm, f_lo, f_hi
        self.terms, self.K, self.flags = terms, K, flags


def _terms_of(expr):
    """expr: {leaf: M} -> (terms [(leaf, xi, m0, m1)], K, Ktot) or None if not encodable."""
    terms = []
    neg = [0, 0]
    pos = [0, 0]
    for leaf, M in expr.items():
        (p, q), (r, s) = decompose(M)
        if max(abs(p), abs(q), abs(r), abs(s)) > isa.LIN_MAX_MULT:
            return None
        if p or q:
            terms.append((leaf, False, p, q))
            for c, m in ((0, p), (1, q)):
                (neg if m < 0 else pos)[c] += abs(m)
        if r or s:
            terms.append((leaf, True, r, s))
            # R0 += 9 r x0 - s x1 ; R1 += r x0 + 9 s x1
            for c, m in ((0, 9 * r), (0, -s), (1, r), (1, 9 * s)):
                (neg if m < 0 else pos)[c] += abs(m)
    K = max(neg)
    ktot = K + max(pos)
    if len(terms) > isa.LIN_MAX_TERMS or K > isa.LIN_MAX_K or ktot > 1000:
        return None
    return terms, K, ktot


def fuse(ops, enable=True, lin_trees=False):
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
            return lin_trees
        return u.op in ("MUL", "SQR") and o.op in ("ADD", "SUB") and u.srcs.count(v) == 1

    out = []
    alias = {}
    done = set()     # values materialised in `out`
    pending = {}     # deferred linear value -> its Op

    def res(v):
        while v in alias:
            v = alias[v]
        return v

    def expand(v, budget):
        """Linear expression of v over materialised leaves: {leaf: M}.  Deferred children are inlined
        while the result stays encodable; otherwise the child is materialised and becomes a leaf."""
        v = res(v)
        if v not in pending:
            return {v: I2}
        o = pending[v]
        parts = []
        if o.op == "ADD":
            parts = [(o.srcs[0], 1), (o.srcs[1], 1)]
        elif o.op == "SUB":
            parts = [(o.srcs[0], 1), (o.srcs[1], -1)]
        elif o.op == "NEG":
            parts = [(o.srcs[0], -1)]
        elif o.op == "DBL":
            parts = [(o.srcs[0], 2)]
        elif o.op in ("MULXI", "CONJ", "MOV"):
            parts = [(o.srcs[0], 1)]
        expr = {}
        for src, k in parts:
            sub = expand(src, budget)
            for leaf, M in sub.items():
                M = m_scale(M, k)
                expr[leaf] = m_add(expr[leaf], M) if leaf in expr else M
        if o.op == "MULXI":
            expr = {l: m_xi(M) for l, M in expr.items()}
        elif o.op == "CONJ":
            expr = {l: m_conj(M) for l, M in expr.items()}
        expr = {l: M for l, M in expr.items() if any(M)}
        return expr

    def consume(v):
        """Drop v and every deferred value inlined beneath it from `pending`."""
        o = pending.pop(v)
        for s in o.srcs:
            s = res(s)
            if s in pending:
                consume(s)

    def emit_lin(v):
        """Materialise linear value v (removing it from pending)."""
        v = res(v)
        o = pending[v]
        if not lin_trees:
            # elementary form: one dedicated opcode per modular step
            pending.pop(v)
            if o.op == "MOV":
                alias[v] = need(o.srcs[0])
                return
            out.append(FOp(o.op, v, [need(s) for s in o.srcs]))
            done.add(v)
            return
        expr = expand(v, None)
        enc = _terms_of(expr) if expr else None
        if expr and enc is None:
            # too big: materialise deferred children one at a time (largest first) and retry
            kids = [res(s) for s in o.srcs if res(s) in pending]
            assert kids, "single linear op not encodable"
            for k in kids:
                emit_lin(k)
                expr = expand(v, None)
                enc = _terms_of(expr)
                if enc is not None:
                    break
            assert enc is not None
        consume(v)
        if not expr:
            # identically zero: load the constant instead (never happens in the shipped programs)
            raise AssertionError("zero linear expression")
        terms, K, _ = enc
        if len(terms) == 1 and terms[0][1:] == (False, 1, 1):
            alias[v] = terms[0][0]
            return
        out.append(FOp("LIN", v, [t[0] for t in terms], terms=terms, K=K))
        done.add(v)

    def need(v):
        v = res(v)
        if v in pending:
            emit_lin(v)
            v = res(v)
        return v

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
            out.append(FOp(o.op, o.dst, srcs, flags=flags))
            done.add(o.dst)
            continue
        srcs = [need(s) for s in o.srcs]
        out.append(FOp(o.op, o.dst, srcs, imm=o.imm, f_lo=o.f_lo, f_hi=o.f_hi))
        if o.dst is not None:
            done.add(o.dst)
    assert not pending, "dangling deferred values"
    return out
