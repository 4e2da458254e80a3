"""Slot allocation: fused SSA program (fuse.py) -> instruction words over `n_slots` shared-memory slots.

The programs are straight-line, so every value's future uses are known exactly; eviction takes the
resident value whose next use is furthest away (Belady).  Evicted values go to per-thread scratch in
global memory (L2-resident: the kernel is persistent, so scratch is sized by the resident threads),
except constants and inputs, which are simply re-loaded (LDC / LDG) when needed again.
"""
from collections import defaultdict

from . import isa

REMAT = ("LDC", "LDG")
INF = 1 << 60


class Allocated:
    def __init__(self, words, n_slots, n_scratch, stats):
        self.words, self.n_slots, self.n_scratch, self.stats = words, n_slots, n_scratch, stats


def allocate(ops, n_slots):
    uses = defaultdict(list)
    defop = {}
    for i, o in enumerate(ops):
        for s in o.srcs:
            uses[s].append(i)
        if o.dst is not None:
            defop[o.dst] = o
    upos = defaultdict(int)

    def next_use(v, i):
        """first use at index >= i"""
        u = uses[v]
        k = upos[v]
        while k < len(u) and u[k] < i:
            k += 1
        upos[v] = k
        return u[k] if k < len(u) else INF

    loc = {}                      # value -> slot
    slot_val = [None] * n_slots
    free_slots = list(range(n_slots - 1, -1, -1))
    scratch_of = {}               # value -> scratch index (copy stays valid for the value's whole life: SSA)
    free_scratch = []
    n_scratch = 0
    words = []
    stats = defaultdict(int)

    def emit(op, **kw):
        words.append(isa.encode(op, **kw))
        stats[op] += 1

    def release(v):
        s = loc.pop(v, None)
        if s is not None:
            slot_val[s] = None
            free_slots.append(s)
        sc = scratch_of.pop(v, None)
        if sc is not None:
            free_scratch.append(sc)

    def take_slot(i, protect):
        nonlocal n_scratch
        if free_slots:
            return free_slots.pop()
        best, best_key = None, -1
        for s in range(n_slots):
            v = slot_val[s]
            if v in protect:
                continue
            nu = next_use(v, i)
            cheap = defop[v].op in REMAT or v in scratch_of  # eviction needs no store
            key = nu * 2 + (1 if cheap else 0)
            if key > best_key:
                best, best_key = s, key
        assert best is not None, "not enough slots for one instruction"
        v = slot_val[best]
        if next_use(v, i) != INF and defop[v].op not in REMAT and v not in scratch_of:
            if free_scratch:
                sc = free_scratch.pop()
            else:
                sc = n_scratch
                n_scratch += 1
            scratch_of[v] = sc
            emit("SPILL", a=best, imm=sc)
        del loc[v]
        slot_val[best] = None
        return best

    def materialise(v, i, protect):
        if v in loc:
            return
        s = take_slot(i, protect)
        o = defop[v]
        if o.op == "LDC":
            emit("LDC", d=s, imm=o.imm)
        elif o.op == "LDG":
            emit("LDG", d=s, a=o.f_lo, b=o.f_hi, imm=o.imm)
        else:
            emit("FILL", d=s, imm=scratch_of[v])
        loc[v] = s
        slot_val[s] = v

    for i, o in enumerate(ops):
        if o.op in REMAT:
            continue  # materialised at first use
        protect = set(o.srcs)
        assert len(protect) + 1 <= n_slots, "instruction needs more slots than available"
        for v in o.srcs:
            materialise(v, i, protect)
        src_slots = [loc[v] for v in o.srcs]
        for v in set(o.srcs):
            if next_use(v, i + 1) == INF:
                release(v)  # last use: the destination may reuse the slot (handlers read before they write)
        d = 0
        if o.dst is not None:
            d = take_slot(i + 1, protect=set(v for v in o.srcs if v in loc))
            loc[o.dst] = d
            slot_val[d] = o.dst
        if o.op == "MUL":
            k = 0
            a = src_slots[k]; k += 1
            b = 0
            if o.flags & isa.MUL_B:
                b = src_slots[k]; k += 1
            c = src_slots[k]; k += 1
            e = 0
            if o.flags & isa.MUL_E:
                e = src_slots[k]; k += 1
            emit("MUL", d=d, a=a, b=b, c=c, e=e, imm=o.flags)
        elif o.op == "SQR":
            b = src_slots[1] if o.flags & isa.MUL_B else 0
            emit("SQR", d=d, a=src_slots[0], b=b, imm=o.flags)
        elif o.op == "MULFP":
            emit("MULFP", d=d, a=src_slots[0], b=src_slots[1], imm=o.imm)
        elif o.op in ("INV", "DBL", "NEG", "CONJ", "MULXI"):
            emit(o.op, d=d, a=src_slots[0])
        elif o.op in ("ADD", "SUB"):
            emit(o.op, d=d, a=src_slots[0], b=src_slots[1])
        elif o.op == "STG":
            emit("STG", d=o.f_lo, a=src_slots[0], b=o.f_hi, imm=o.imm)
        elif o.op == "LIN":
            emit("LIN", d=d, a=len(o.terms), imm=o.K)
            tw = [isa.encode_term(loc_s, xi, m0, m1) for loc_s, (_, xi, m0, m1) in zip(src_slots, o.terms)]
            if len(tw) % 2:
                tw.append(0)
            for j in range(0, len(tw), 2):
                words.append(tw[j] | (tw[j + 1] << 32))
        else:
            raise ValueError(o.op)
        if o.dst is not None and next_use(o.dst, i + 1) == INF:
            release(o.dst)  # result never used
    emit("END")
    return Allocated(words, n_slots, n_scratch, dict(stats))
