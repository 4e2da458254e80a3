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
        for s in o.all_srcs():
            uses[s].append(i)
        for v in ([o.dst] if o.dst is not None else []) + ([o.dst2] if o.dst2 is not None else []):
            defop[v] = o
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

    def remat(v):
        return defop[v].op in REMAT

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
            cheap = remat(v) or v in scratch_of  # eviction needs no store
            key = nu * 2 + (1 if cheap else 0)
            if key > best_key:
                best, best_key = s, key
        assert best is not None, "not enough slots for one instruction"
        v = slot_val[best]
        if next_use(v, i) != INF and not remat(v) and v not in scratch_of:
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
            lo, hi, imm = isa.ldst_fields(o.f_lo, o.f_hi, o.imm)
            emit("LDG", d=s, a=lo, b=hi, imm=imm)
        else:
            emit("FILL", d=s, imm=scratch_of[v])
        loc[v] = s
        slot_val[s] = v

    for i, o in enumerate(ops):
        if o.op in REMAT or o.op == "CUT":
            continue  # REMAT: materialised at first use; CUT: a phase-boundary marker (phases.py), no code
        all_srcs = o.all_srcs()
        protect = set(all_srcs)
        dsts = o.dsts()
        assert len(protect) + len(dsts) <= n_slots, "instruction needs more slots than available"
        for v in all_srcs:
            materialise(v, i, protect)
        sl = {v: loc[v] for v in all_srcs}
        # last use: a destination may reuse the slot (handlers read everything before they write) - except for
        # the sources of a post stage, which is evaluated AFTER r' has been parked in its destination slot
        late = set(t[0] for t in o.post[0] + o.post[1] if t[0] is not None) if o.post else set()
        dying = [v for v in protect if next_use(v, i + 1) == INF]
        for v in dying:
            if v not in late:
                release(v)
        dslot = []
        keep = set(v for v in all_srcs if v in loc)
        for v in dsts:
            d = take_slot(i + 1, protect=keep)
            loc[v] = d
            slot_val[d] = v
            keep = keep | {v}
            dslot.append(d)
        for v in dying:
            if v in late:
                release(v)
        d = dslot[0] if dslot else 0
        src_slots = [sl[v] for v in o.srcs]
        if o.op in isa.PRODUCT_OPS:
            imm = (o.canon[0] << isa.MUL_CANON_SHIFT) | (o.canon[1] << (isa.MUL_CANON_SHIFT + 2))
            ext = bool(o.hi) or o.post is not None
            if ext:
                imm |= isa.MUL_EXT
            d_r = dslot[0] if o.store_r else 0
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
                emit("MUL", d=d_r, a=a, b=b, c=c, e=e, imm=imm | o.flags)
            elif o.op == "SQR":
                b = src_slots[1] if o.flags & isa.MUL_B else 0
                emit("SQR", d=d_r, a=src_slots[0], b=b, imm=imm | o.flags)
            else:
                emit("MULFP", d=d_r, a=src_slots[0], b=src_slots[1], imm=imm | (isa.MULFP_HALF if o.imm else 0))
            if ext:
                d2 = dslot[-1] if o.dst2 is not None else 0
                post = o.post or ([], [])
                r_slot = d_r if o.store_r else d2

                def enc(lst):
                    return [isa.encode_entry(r_slot if leaf is None else sl[leaf], half, mult, neg) for leaf, half, mult, neg in lst]

                pairs = isa.pair_entries(enc(post[0]), enc(post[1]), r_slot)
                words.append(isa.encode_ext(d2, [(sl[v], neg) for v, neg in o.hi], o.store_r, len(pairs) // 2))
                words.extend(isa.pack_entries(pairs))
                stats["hi_terms"] += len(o.hi)
                stats["post_stages"] += 1 if o.post else 0
                stats["post_entries"] += len(post[0]) + len(post[1])
        elif o.op in ("INV", "DBL", "NEG", "CONJ", "MULXI"):
            emit(o.op, d=d, a=src_slots[0])
        elif o.op in ("ADD", "SUB"):
            emit(o.op, d=d, a=src_slots[0], b=src_slots[1])
        elif o.op == "STG":
            lo, hi, imm = isa.ldst_fields(o.f_lo, o.f_hi, o.imm)
            emit("STG", d=lo, a=src_slots[0], b=hi, imm=imm)
        elif o.op == "LIN":

            def enc(lst):
                return [isa.encode_entry(sl[leaf], half, mult, neg) for leaf, half, mult, neg in lst]

            pairs = isa.pair_entries(enc(o.terms[0]), enc(o.terms[1]), src_slots[0])
            emit("LIN", d=d, a=len(pairs) // 2)
            words.extend(isa.pack_entries(pairs))
            stats["lin_entries"] += len(o.terms[0]) + len(o.terms[1])
        else:
            raise ValueError(o.op)
        for v in dsts:
            if next_use(v, i + 1) == INF:
                release(v)  # result never used
    emit("END")
    return Allocated(words, n_slots, n_scratch, dict(stats))
