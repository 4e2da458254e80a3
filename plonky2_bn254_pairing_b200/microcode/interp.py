"""Big-integer interpreter for allocated sequencer programs (build/test tool, never on the product path).

Executes exactly the instruction words the GPU runs - including spills, fills and re-loads - on one
"thread", with plain (non-Montgomery) residues.  Montgomery form is a representation detail of the
kernel: every opcode is a ring operation, so the plain-domain run is isomorphic to the device run.
For LIN it additionally checks the lazy-accumulation contract the kernel relies on (K makes the
accumulated value non-negative; the total stays below the reduction's range).
Also counts the algorithmic work of a program (Fp products / Montgomery reductions / MACs).
"""
from . import isa
from .builder import P

# popcount(p-2) and bit length, for the Fermat ladder of csrc/fp2.cuh::fp_inv
_INV_SQR = 254
_INV_MUL = bin(P - 2).count("1")

# (wide products, Montgomery reductions) per opcode: the ALGORITHMIC count (Karatsuba Fq2 product),
# which is what roofline.achieved is quoted on ...
WORK = {
    "MUL": (3, 2), "SQR": (2, 2), "MULFP": (2, 2),
    "INV": (2 + 2 + (_INV_SQR + _INV_MUL), 2 + 2 + (_INV_SQR + _INV_MUL)),
}
# ... and what the kernel actually issues.  One thread runs a whole Fq2 operation (Karatsuba product, one
# inversion), so the two counts coincide; the field is kept because the bench line reports both.
WORK_EXECUTED = dict(WORK)


def _pre(slots, a, b, has_b, neg_b):
    x = slots[a]
    if has_b:
        y = slots[b]
        x = ((x[0] - y[0]) % P, (x[1] - y[1]) % P) if neg_b else ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
    return x


def lin_eval(ents, slots):
    """One output component of a LIN: sum of mult * (neg ? p - z : z).  Checks the kernel's contract
    (canonical inputs, the lazy sum stays below 1024 p - true for ALL inputs because it only depends on
    the multipliers)."""
    assert sum(m for _, _, m, _ in ents) <= isa.LIN_MAX_SUM and len(ents) <= isa.LIN_MAX_ENT
    acc = 0
    for slot, half, mult, neg in ents:
        if mult == 0:
            assert slots[slot] is not None  # padding entry: the kernel still loads the slot
            continue
        z = slots[slot][half]
        assert 0 <= z < P
        acc += mult * ((P - z) if neg else z)
    return acc % P


def _lin(ins, slots):
    return (lin_eval(ins.ent0, slots), lin_eval(ins.ent1, slots))


# wide-value range contract of the product-class epilogue, in exact integers
_WIDE_MAX = (1 << 512) - (P << 256)


def run(words, consts, arrays, n_slots, n_scratch):
    """arrays: {arr_id: list of Fq ints}; STG writes into arrays[arr_id] (a dict or list)."""
    slots = [None] * n_slots
    scratch = [None] * max(n_scratch, 1)
    for ins in isa.parse(words):
        op, d, a, b, c, e, imm = ins.op, ins.d, ins.a, ins.b, ins.c, ins.e, ins.imm
        if op == "END":
            break
        if op in isa.PRODUCT_OPS:
            # mirror the kernel's unreduced arithmetic closely enough to check its range contract
            if op == "MUL":
                x, y = slots[a], slots[c]
                if imm & isa.MUL_B:
                    t = slots[b]
                    x = (x[0] - t[0] + P, x[1] - t[1] + P) if imm & isa.MUL_BNEG else (x[0] + t[0], x[1] + t[1])
                    if imm & isa.MUL_BCANON:
                        x = (x[0] % P, x[1] % P)
                if imm & isa.MUL_E:
                    t = slots[e]
                    y = (y[0] - t[0] + P, y[1] - t[1] + P) if imm & isa.MUL_ENEG else (y[0] + t[0], y[1] + t[1])
                # Karatsuba: x1 y1 < 4 p^2 < p 2^256, so one conditional p 2^256 makes the difference non-negative
                T0 = x[0] * y[0] - x[1] * y[1]
                if T0 < 0:
                    T0 += P << 256
                T1 = x[0] * y[1] + x[1] * y[0]
            elif op == "SQR":
                x = _pre(slots, a, b, imm & isa.MUL_B, imm & isa.MUL_BNEG)
                T0 = (x[0] + x[1]) * ((x[0] - x[1]) % P)
                T1 = 2 * x[0] * x[1]
            else:
                x, s = slots[a], slots[b][1 if imm & isa.MULFP_HALF else 0]
                T0, T1 = x[0] * s, x[1] * s
            # hi terms: on the device (Montgomery domain) h * 2^256 is added to the wide value, which adds h to
            # the reduced result; in this plain-residue model the value gets h and the range check gets h * 2^256
            H0 = H1 = 0
            for slot, neg in ins.hi:
                h = slots[slot]
                assert 0 <= h[0] < P and 0 <= h[1] < P
                H0 += (P - h[0]) if neg else h[0]
                H1 += (P - h[1]) if neg else h[1]
            lv0, lv1 = (imm >> isa.MUL_CANON_SHIFT) & 3, (imm >> (isa.MUL_CANON_SHIFT + 2)) & 3
            for T, lv in ((T0 + (H0 << 256), lv0), (T1 + (H1 << 256), lv1)):
                assert 0 <= T < _WIDE_MAX, "wide value out of range"
                # redc gives T / 2^256 + (< p); the canon ladder handles [0, 2p (lv + 1))
                assert (T >> 256) + P < 2 * P * (lv + 1), "canon level too small"
            r = ((T0 + H0) % P, (T1 + H1) % P)
            # the kernel parks r' in its slot (d, or d2 when r' is not wanted for itself), then runs the post LIN
            slots[ins.r_slot()] = r
            if ins.has_post():
                slots[ins.d2] = _lin(ins, slots)
        elif op == "LIN":
            slots[d] = _lin(ins, slots)
        elif op == "LDC":
            slots[d] = consts[imm]
        elif op == "LDG":
            pg = (imm >> 8) << 8
            slots[d] = (arrays[imm & 0xFF][pg + a], arrays[imm & 0xFF][pg + b])
        elif op == "STG":
            pg = (imm >> 8) << 8
            arrays[imm & 0xFF][pg + d], arrays[imm & 0xFF][pg + b] = slots[a]
        elif op == "SPILL":
            scratch[imm] = slots[a]
        elif op == "FILL":
            slots[d] = scratch[imm]
        elif op == "ADD":
            slots[d] = ((slots[a][0] + slots[b][0]) % P, (slots[a][1] + slots[b][1]) % P)
        elif op == "SUB":
            slots[d] = ((slots[a][0] - slots[b][0]) % P, (slots[a][1] - slots[b][1]) % P)
        elif op == "DBL":
            slots[d] = (2 * slots[a][0] % P, 2 * slots[a][1] % P)
        elif op == "NEG":
            slots[d] = (-slots[a][0] % P, -slots[a][1] % P)
        elif op == "CONJ":
            slots[d] = (slots[a][0], -slots[a][1] % P)
        elif op == "MULXI":
            slots[d] = ((9 * slots[a][0] - slots[a][1]) % P, (slots[a][0] + 9 * slots[a][1]) % P)
        elif op == "INV":
            x = slots[a]
            n = (x[0] * x[0] + x[1] * x[1]) % P
            ni = pow(n, P - 2, P)  # 0 -> 0, like the device ladder
            slots[d] = (x[0] * ni % P, (-x[1]) * ni % P)
        else:
            raise ValueError(op)
    return arrays


def walk(words):
    """Yield (opname, n_slot_moves) per instruction (a LIN entry moves half a slot)."""
    for ins in isa.parse(words):
        if ins.op in ("LDC", "LDG", "FILL"):
            yield ins.op, 1
            continue
        nent = len(ins.ent0) + len(ins.ent1)
        full_reads = len(ins.slots_read()) - nent
        yield ins.op, full_reads + 0.5 * nent + len(ins.slots_written()) + (1 if ins.has_post() and not ins.store_r else 0)


def work(words):
    """Algorithmic work of one program run: dict with Fp products, reductions, and 32x32 MACs
    (one product = 64 MACs, one reduction = 72), plus the opcode histogram and the
    number of 64-byte shared-memory slot moves."""
    prod = red = xprod = xred = 0
    hist = {}
    moves = 0
    for op, mv in walk(words):
        hist[op] = hist.get(op, 0) + 1
        moves += mv
        if op in WORK:
            prod += WORK[op][0]
            red += WORK[op][1]
            xprod += WORK_EXECUTED[op][0]
            xred += WORK_EXECUTED[op][1]
    return {"products": prod, "reductions": red, "macs": 64 * prod + 72 * red,
            "macs_executed": 64 * xprod + 72 * xred, "hist": hist, "slot_moves": moves}
