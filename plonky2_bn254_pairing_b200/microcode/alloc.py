"""Slot allocation: SSA program -> instruction words over `n_slots` shared-memory slots.

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
    scratch_of = {}               # value -> scratch index (copy is valid for the value's whole life: SSA)
    free_scratch = []
    n_scratch = 0
    words = []
    stats = defaultdict(int)

    def emit(op, d=0, a=0, b=0, imm=0):
        words.append(isa.encode(op, d, a, b, imm))
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
        best, best_nu = None, -1
        for s in range(n_slots):
            v = slot_val[s]
            if v in protect:
                continue
            nu = next_use(v, i)
            # prefer victims that need no store (rematerialisable or already in scratch)
            cheap = defop[v].op in REMAT or v in scratch_of
            key = nu * 2 + (1 if cheap else 0)
            if key > best_nu:
                best, best_nu = s, key
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
        if o.op in REMAT and next_use(o.dst, i + 1) == INF:
            continue  # dead load
        if o.op in REMAT:
            # defer: materialised at first use (keeps slots free until the value is needed)
            continue
        protect = set(o.srcs)
        for v in o.srcs:
            materialise(v, i, protect)
        src_slots = [loc[v] for v in o.srcs]
        # sources whose last use is this instruction free their slot before the destination is chosen
        for v in set(o.srcs):
            if next_use(v, i + 1) == INF:
                release(v)
        if o.dst is not None:
            d = take_slot(i + 1, protect=set(v for v in o.srcs if v in loc))
            loc[o.dst] = d
            slot_val[d] = o.dst
        if o.op in ("MUL", "ADD", "SUB"):
            emit(o.op, d=d, a=src_slots[0], b=src_slots[1])
        elif o.op == "MULFP":
            emit(o.op, d=d, a=src_slots[0], b=src_slots[1], imm=o.imm)
        elif o.op in ("SQR", "NEG", "CONJ", "MULXI", "MOV", "INV", "DBL"):
            emit(o.op, d=d, a=src_slots[0])
        elif o.op == "STG":
            emit("STG", d=o.f_lo, a=src_slots[0], b=o.f_hi, imm=o.imm)
        else:
            raise ValueError(o.op)
        if o.dst is not None and next_use(o.dst, i + 1) == INF:
            release(o.dst)  # result never used
    emit("END")
    return Allocated(words, n_slots, n_scratch, dict(stats))
