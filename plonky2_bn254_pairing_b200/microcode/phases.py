"""Phase splitting: one straight-line program -> K shorter programs ("phases") that hand their live values
over through a per-element state array in global memory.

Why: a fused pairing is ~10 ms of warp time and the kernel keeps 8 warps per SM resident (the slots fill
shared memory), so a 2^16 batch is 2048 warp-tasks on 1184 warp slots - 1.73 rounds that cost 2.  With
every task cut into K phases and the tasks handed out breadth-first (all chunks' phase 0, then phase 1, ...),
only the last phase pays the partial round: 1.73 + 0.27/K.  The per-phase state is a dozen Fq2 values per
element (f, the running point), a few hundred bytes against ~0.5 M MACs of work.

The split is automatic: `Builder.cut()` marks candidate boundaries; `split` picks K-1 of them that balance the
algorithmic work, computes the values live across each boundary, gives each a state index (indices are reused
once a value is dead) and adds the LDG / STG instructions.  Constants and program inputs are simply re-loaded.
"""
import os

from . import isa
from .fuse import FOp

REMAT = ("LDC", "LDG")
COST = {"MUL": 336, "SQR": 272, "MULFP": 272, "INV": 60000}
TAIL = "0.5,0.25,0.12"   # lengths of the last phases relative to the others


def split(fops, n_phases):
    """fops: fused program (with CUT markers).  Returns (list of per-phase FOp lists without markers, n_state)."""
    cuts = [i for i, o in enumerate(fops) if o.op == "CUT"]
    if n_phases <= 1 or not cuts:
        return [[o for o in fops if o.op != "CUT"]], 0
    # cumulative work at every candidate
    cum, acc = [], 0
    for o in fops:
        cum.append(acc)
        acc += COST.get(o.op, 20)
    total = acc
    # Relative lengths of the phases: equal, except that the LAST ones taper off (TAIL, overridden by BNP_PHASE_TAIL).
    # Tasks are handed out phase by phase, so the warps that find the queue empty at the end of a launch wait for the
    # last-phase tasks still running - half a task each on average.  With equal phases that was 3 % of a 2^16 launch
    # (ncu: EXIT + barrier samples); a last phase of 1/80 of the program instead of 1/13 brings back 1.5 %
    # (profiles/experiments_r2.txt, A/B on one box).
    weights = [1.0] * n_phases
    tail = [float(x) for x in os.environ.get("BNP_PHASE_TAIL", TAIL).split(",") if x]
    if tail and 4 * len(tail) <= n_phases:   # a taper needs a body of equal phases in front of it
        weights[-len(tail):] = tail
    # boundary k goes to the candidate nearest its target, among those that leave enough candidates on either side for
    # the other boundaries (the candidates thin out towards the end of a program, the targets of the taper do not)
    chosen, lo = [], 0
    if len(cuts) >= n_phases - 1:
        for k in range(1, n_phases):
            target = total * sum(weights[:k]) / sum(weights)
            hi = len(cuts) - (n_phases - 1 - k)          # exclusive
            j = min(range(lo, hi), key=lambda j: abs(cum[cuts[j]] - target))
            chosen.append(cuts[j])
            lo = j + 1
    chosen.sort()
    bounds = [0] + chosen + [len(fops)]
    segs = [[o for o in fops[bounds[k]:bounds[k + 1]] if o.op != "CUT"] for k in range(len(bounds) - 1)]

    seg_of = {}          # value -> segment that defines it
    defop = {}
    for k, seg in enumerate(segs):
        for o in seg:
            for v in ([o.dst] if o.dst is not None else []) + ([o.dst2] if o.dst2 is not None else []):
                seg_of[v] = k
                defop[v] = o
    last_use = {}        # value -> last segment that reads it
    first_seg_uses = [set() for _ in segs]
    for k, seg in enumerate(segs):
        for o in seg:
            for v in o.all_srcs():
                last_use[v] = k
                first_seg_uses[k].add(v)

    # a value crosses boundaries seg_of[v] .. last_use[v]-1; give it a state index for that interval
    crossing = sorted((v for v in seg_of if last_use.get(v, -1) > seg_of[v] and defop[v].op not in REMAT),
                      key=lambda v: (seg_of[v], str(v)))
    free_at = {}         # state index -> first segment in which it may be written again
    state_of = {}
    n_state = 0
    for v in crossing:
        idx = None
        for i in range(n_state):
            # the index is free if its previous tenant was last read in a segment <= the one that stores v
            # (that segment's loads all precede its stores - the STGs are appended at its end)
            if free_at[i] <= seg_of[v]:
                idx = i
                break
        if idx is None:
            idx = n_state
            n_state += 1
        state_of[v] = idx
        free_at[idx] = last_use[v]

    out = []
    for k, seg in enumerate(segs):
        prog = []
        used = first_seg_uses[k]
        # re-create the loads of constants / inputs defined in earlier phases
        for v in sorted((v for v in used if seg_of.get(v, k) < k and defop[v].op in REMAT), key=str):
            o = defop[v]
            prog.append(FOp(o.op, v, [], imm=o.imm, f_lo=o.f_lo, f_hi=o.f_hi))
        # values handed over from earlier phases: a re-loadable LDG from the state array
        for v in sorted((v for v in used if seg_of.get(v, k) < k and defop[v].op not in REMAT), key=str):
            i = state_of[v]
            prog.append(FOp("LDG", v, [], imm=isa.ARR_STATE, f_lo=2 * i, f_hi=2 * i + 1))
        prog.extend(seg)
        for v in crossing:
            if seg_of[v] == k:
                i = state_of[v]
                prog.append(FOp("STG", None, [v], imm=isa.ARR_STATE, f_lo=2 * i, f_hi=2 * i + 1))
        out.append(prog)
    return out, n_state
