"""Pairing scheduler: fused program (fuse.py / phases.py)  ->  list of bundles (op_a, op_b | None).

The kernel gives every pairing TWO lanes of a warp ("units").  Both units run the same handler on the same slot
file, each on its own Fq2 instruction, so a bundle is two independent instructions of the same opcode (the
handlers branch only on warp-uniform fields; per-unit differences are operand fields, masks and selects).  The programs are
straight-line SSA, so the only ordering constraints are true data dependences.

Algorithm: greedy list scheduling in program order.  The earliest unscheduled instruction A is always ready
(everything before it is scheduled); its partner is the first later instruction B inside a look-ahead window that
  * has the same shape (`shape_key`),
  * depends on nothing unscheduled (in particular not on A),
  * and fits the slot file together with A (distinct sources + destinations <= n_slots).
Pulling B forward lengthens live ranges a little; the window bounds that.
"""
from . import isa

REMAT = ("LDC", "LDG")
WINDOW = 96
PATIENCE = 12


def shape_key(o):
    """Fields that steer uniform branches in the handlers: two instructions can share a bundle iff equal."""
    # MUL / SQR: presence and sign of the pre-additions, hi terms and post stages may differ between the units
    # (masks / selects in the handler when a bundle is mixed), so every MUL pairs with every MUL
    if o.op == "STG":
        return ("STG", o.imm)
    if o.op in ("INV", "CUT"):
        return None
    return (o.op,)


def n_values(a, b=None):
    vals = set(a.all_srcs()) | set(a.dsts())
    if b is not None:
        vals |= set(b.all_srcs()) | set(b.dsts())
    return len(vals)


def schedule(fops, n_slots, window=WINDOW, patience=PATIENCE):
    ops = [o for o in fops if o.op not in REMAT and o.op != "CUT"]
    remat = [o for o in fops if o.op in REMAT]
    producer = {}
    for i, o in enumerate(ops):
        for v in ([o.dst] if o.dst is not None else []) + ([o.dst2] if o.dst2 is not None else []):
            producer[v] = i
    deps = [sorted(set(producer[v] for v in o.all_srcs() if v in producer)) for o in ops]
    keys = [shape_key(o) for o in ops]
    n = len(ops)
    done = [False] * n
    waited = 0          # bundles emitted while the oldest instruction was passed over
    bundles = []
    base = 0
    while True:
        while base < n and done[base]:
            base += 1
            waited = 0
        if base >= n:
            break
        lim = min(n, base + window)
        ready = [i for i in range(base, lim) if not done[i] and all(done[k] for k in deps[i])]
        # a store to global memory never moves ahead of earlier work: programs may run in place (the output array
        # aliasing an input that is re-loaded late), so every STG stays behind all instructions that precede it
        ready = [i for i in ready
                 if ops[i].op != "STG" or all(done[k] or ops[k].op == "STG" for k in range(base, i))]
        # ready instructions grouped by shape, program order inside a group
        groups = {}
        for i in ready:
            if keys[i] is not None:
                groups.setdefault(keys[i], []).append(i)

        def partner_of(i):
            for j in groups.get(keys[i], ()):
                if j != i and n_values(ops[i], ops[j]) <= n_slots:
                    return j
            return None

        pick = None
        j = partner_of(base) if keys[base] is not None else None
        if j is not None:
            pick = (base, j)
        elif keys[base] is not None and waited < patience:
            # pass over the oldest instruction: run the earliest ready pair instead (it may unlock a partner)
            for i in ready:
                if i == base or keys[i] is None:
                    continue
                jj = partner_of(i)
                if jj is not None and jj != base:
                    pick = (min(i, jj), max(i, jj))
                    break
            if pick is not None:
                waited += 1
        if pick is None:
            pick = (base, None)
        a, b = pick
        done[a] = True
        if b is not None:
            done[b] = True
        bundles.append((ops[a], ops[b] if b is not None else None))
    return remat, bundles


def singles(fops):
    """No pairing: one instruction per bundle (unit B idles) - the baseline the tests compare against."""
    return ([o for o in fops if o.op in REMAT],
            [(o, None) for o in fops if o.op not in REMAT and o.op != "CUT"])


def stats(bundles):
    from collections import Counter
    pairs, singles = Counter(), Counter()
    for a, b in bundles:
        (pairs if b is not None else singles)[a.op] += 1
    return pairs, singles
