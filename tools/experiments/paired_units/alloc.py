"""Slot allocation: fused SSA program (fuse.py)  ->  bundles (sched.py)  ->  instruction words over `n_slots`
shared-memory slots.

The programs are straight-line, so every value's future uses are known exactly; eviction takes the
resident value whose next use is furthest away (Belady).  Evicted values go to per-pairing scratch in
global memory (L2-resident: the kernel is persistent, so scratch is sized by the resident pairings),
except constants and inputs, which are simply re-loaded (LDC / LDG) when needed again.

Both units of a pairing share one slot file.  The handlers read every main-stage operand of BOTH units before
either unit stores (a warp barrier sits in front of every store), so a destination may reuse the slot of a
source that dies in the bundle - except the sources of the post stages, which are read after r' has been parked.
The spills / re-loads a bundle needs are collected and emitted as bundles of their own (two per bundle).
"""
from collections import defaultdict

from . import isa, sched

REMAT = ("LDC", "LDG")
INF = 1 << 60


class Allocated:
    def __init__(self, words, n_slots, n_scratch, stats):
        self.words, self.n_slots, self.n_scratch, self.stats = words, n_slots, n_scratch, stats


def _post_leaves(o):
    return set(t[0] for t in o.post[0] + o.post[1] if t[0] is not None) if o.post else set()


def allocate(ops, n_slots, pair=True, trace=None):
    """ops: fused program (list of FOp).  pair=False keeps one instruction per bundle (unit B idles)."""
    remat_ops, bundles = sched.schedule(ops, n_slots) if pair else sched.singles(ops)
    defop = {}
    for o in remat_ops:
        defop[o.dst] = o
    uses = defaultdict(list)
    for i, (a, b) in enumerate(bundles):
        for o in (a, b):
            if o is None:
                continue
            for s in o.all_srcs():
                if not uses[s] or uses[s][-1] != i:
                    uses[s].append(i)
            for v in ([o.dst] if o.dst is not None else []) + ([o.dst2] if o.dst2 is not None else []):
                defop[v] = o
    upos = defaultdict(int)

    def next_use(v, i):
        """first use at bundle index >= i"""
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
    pre_spills = []               # (slot, scratch index) collected for the current bundle
    pre_loads = []                # ("FILL", slot, sc) | ("LDC", slot, imm) | ("LDG", slot, f_lo, f_hi, arr)
    freed_scratch = []            # scratch indices released while the current bundle is being prepared

    def emit_bundle(wa, wb, extra=()):
        """wa, wb: (op, fields dict); wb None: unit B idles (it shadows unit A's word)."""
        a = isa.encode(wa[0], **wa[1])
        b = isa.encode(wb[0], **wb[1]) if wb is not None else (a | isa.UNIT_IDLE)
        words.append(a)
        words.append(b)
        words.extend(extra)
        if len(extra) & 1:
            words.append(0)
        stats[wa[0]] += 1 if wb is None else 2
        stats["bundles"] += 1
        if wb is None:
            stats["idle_units"] += 1

    def remat(v):
        return defop[v].op in REMAT

    def release(v):
        s = loc.pop(v, None)
        if s is not None:
            slot_val[s] = None
            free_slots.append(s)
        sc = scratch_of.pop(v, None)
        if sc is not None:
            freed_scratch.append(sc)   # reusable only after this bundle's re-loads have been emitted (flush_pre)

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
        assert best is not None, "not enough slots for one bundle"
        v = slot_val[best]
        if next_use(v, i) != INF and not remat(v) and v not in scratch_of:
            if free_scratch:
                sc = free_scratch.pop()
            else:
                sc = n_scratch
                n_scratch += 1
            scratch_of[v] = sc
            pre_spills.append((best, sc))
        del loc[v]
        slot_val[best] = None
        return best

    def materialise(v, i, protect):
        if v in loc:
            return
        s = take_slot(i, protect)
        o = defop[v]
        if o.op == "LDC":
            pre_loads.append(("LDC", s, o.imm))
        elif o.op == "LDG":
            pre_loads.append(("LDG", s, o.f_lo, o.f_hi, o.imm))
        else:
            pre_loads.append(("FILL", s, scratch_of[v]))
        loc[v] = s
        slot_val[s] = v

    def flush_pre():
        """Emit the collected spills, then the collected loads, two per bundle where the kinds agree."""
        for k in range(0, len(pre_spills), 2):
            pr = pre_spills[k:k + 2]
            ws = [("SPILL", dict(a=s, imm=sc)) for s, sc in pr]
            emit_bundle(ws[0], ws[1] if len(ws) > 1 else None)
        del pre_spills[:]
        kinds = defaultdict(list)
        for ld in pre_loads:
            kinds[(ld[0], ld[4] if ld[0] == "LDG" else 0)].append(ld)
        for (kind, _), lst in sorted(kinds.items()):
            for k in range(0, len(lst), 2):
                ws = []
                for ld in lst[k:k + 2]:
                    if kind == "LDC":
                        ws.append(("LDC", dict(d=ld[1], imm=ld[2])))
                    elif kind == "LDG":
                        ws.append(("LDG", dict(d=ld[1], a=ld[2], b=ld[3], imm=ld[4])))
                    else:
                        ws.append(("FILL", dict(d=ld[1], imm=ld[2])))
                emit_bundle(ws[0], ws[1] if len(ws) > 1 else None)
        del pre_loads[:]
        free_scratch.extend(freed_scratch)
        del freed_scratch[:]

    def unit_fields(o, sl, dslot_of, canon, bcanon, n_hi_b, np_b, ext):
        """(main word, extension word or None, (ent0, ent1) 16-bit entries, pad slot) of one unit."""
        src_slots = [sl[v] for v in o.srcs]
        dsl = [dslot_of[v] for v in o.dsts()]
        d = dsl[0] if dsl else 0
        if o.op in isa.PRODUCT_OPS:
            imm = (canon[0] << isa.MUL_CANON_SHIFT) | (canon[1] << (isa.MUL_CANON_SHIFT + 2))
            if ext:
                imm |= isa.MUL_EXT
            d_r = dsl[0] if o.store_r else 0
            if o.op == "MUL":
                k = 0
                a = src_slots[k]; k += 1
                b = a
                if o.flags & isa.MUL_B:
                    b = src_slots[k]; k += 1
                c = src_slots[k]; k += 1
                e = c
                if o.flags & isa.MUL_E:
                    e = src_slots[k]; k += 1
                fl = o.flags & (isa.MUL_B | isa.MUL_BNEG | isa.MUL_E | isa.MUL_ENEG)
                main = ("MUL", dict(d=d_r, a=a, b=b, c=c, e=e, imm=imm | fl | (isa.MUL_BCANON if bcanon else 0)))
            elif o.op == "SQR":
                b = src_slots[1] if o.flags & isa.MUL_B else src_slots[0]
                main = ("SQR", dict(d=d_r, a=src_slots[0], b=b, imm=imm | (o.flags & (isa.MUL_B | isa.MUL_BNEG))))
            else:
                main = ("MULFP", dict(d=d_r, a=src_slots[0], b=src_slots[1], imm=imm | (isa.MULFP_HALF if o.imm else 0)))
            xw, ents, pad = None, ([], []), src_slots[0]
            if ext:
                d2 = dsl[-1] if o.dst2 is not None else d_r
                post = o.post or ([], [])
                r_slot = d_r if o.store_r else d2

                def enc(lst):
                    return [isa.encode_entry(r_slot if leaf is None else sl[leaf], half, mult, neg) for leaf, half, mult, neg in lst]

                ents = (enc(post[0]), enc(post[1]))
                np_own = max(len(ents[0]), len(ents[1]))
                store_r = o.store_r or np_own == 0
                xw = isa.encode_ext(d2, [(sl[v], neg) for v, neg in o.hi], store_r, np_b, n_hi_b, np_own)
                stats["hi_terms"] += len(o.hi)
                stats["post_stages"] += 1 if o.post else 0
                stats["post_entries"] += len(post[0]) + len(post[1])
            return main, xw, ents, pad
        if o.op in ("INV", "DBL", "NEG", "CONJ", "MULXI"):
            return (o.op, dict(d=d, a=src_slots[0])), None, ([], []), 0
        if o.op in ("ADD", "SUB"):
            return (o.op, dict(d=d, a=src_slots[0], b=src_slots[1])), None, ([], []), 0
        if o.op == "STG":
            return ("STG", dict(d=o.f_lo, a=src_slots[0], b=o.f_hi, imm=o.imm)), None, ([], []), 0
        if o.op == "LIN":

            def enc(lst):
                return [isa.encode_entry(sl[leaf], half, mult, neg) for leaf, half, mult, neg in lst]

            stats["lin_entries"] += len(o.terms[0]) + len(o.terms[1])
            return ("LIN", dict(d=d, a=np_b)), None, (enc(o.terms[0]), enc(o.terms[1])), src_slots[0]
        raise ValueError(o.op)

    for i, (oa, ob) in enumerate(bundles):
        units = [o for o in (oa, ob) if o is not None]
        all_srcs = []
        for o in units:
            all_srcs += o.all_srcs()
        protect = set(all_srcs)
        dsts = []
        for o in units:
            dsts += o.dsts()
        assert len(protect) + len(dsts) <= n_slots, "bundle needs more slots than available"
        for v in all_srcs:
            materialise(v, i, protect)
        sl = {v: loc[v] for v in all_srcs}
        # last use: a destination may reuse the slot (the handlers read every main-stage operand of both units before
        # either stores) - except for the sources of a post stage, which is evaluated AFTER r' has been parked
        late = set()
        for o in units:
            late |= _post_leaves(o)
        dying = [v for v in protect if next_use(v, i + 1) == INF]
        for v in dying:
            if v not in late:
                release(v)
        dslot_of = {}
        keep = set(v for v in all_srcs if v in loc)
        for v in dsts:
            d = take_slot(i + 1, protect=keep)
            loc[v] = d
            slot_val[d] = v
            keep = keep | {v}
            dslot_of[v] = d
        for v in dying:
            if v in late:
                release(v)
        flush_pre()
        # bundle-level (uniform) fields
        if oa.op in isa.PRODUCT_OPS:
            canon = tuple(max(o.canon[k] for o in units) for k in (0, 1))
            bcanon = any(o.op == "MUL" and o.flags & isa.MUL_BCANON for o in units)
            ext = any(bool(o.hi) or o.post is not None for o in units)
            n_hi_b = max(len(o.hi) for o in units)
            np_b = max(max(len(o.post[0]), len(o.post[1])) if o.post else 0 for o in units)
        elif oa.op == "LIN":
            canon, bcanon, ext, n_hi_b = (0, 0), False, False, 0
            np_b = max(max(len(o.terms[0]), len(o.terms[1])) for o in units)
        else:
            canon, bcanon, ext, n_hi_b, np_b = (0, 0), False, False, 0, 0
        enc = [unit_fields(o, sl, dslot_of, canon, bcanon, n_hi_b, np_b, ext) for o in units]
        extra = []
        if ext:
            extra.append(enc[0][1])
            extra.append(enc[1][1] if len(enc) > 1 else enc[0][1])
        if np_b:
            ea = enc[0][2]
            eb = enc[1][2] if len(enc) > 1 else ([], [])
            extra += isa.entry_words(ea, eb, enc[0][3], enc[1][3] if len(enc) > 1 else enc[0][3])
        if trace is not None:
            trace.append((len(words), [(v, dslot_of[v]) for v in dsts], units))
        emit_bundle(enc[0][0], enc[1][0] if len(enc) > 1 else None, extra)
        for v in dsts:
            if next_use(v, i + 1) == INF:
                release(v)  # result never used
    assert not pre_spills and not pre_loads
    emit_bundle(("END", {}), ("END", {}))
    return Allocated(words, n_slots, n_scratch, dict(stats))
