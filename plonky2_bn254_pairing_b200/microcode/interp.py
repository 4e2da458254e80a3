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

# (wide products, Montgomery reductions) per opcode
WORK = {
    "MUL": (3, 2), "SQR": (2, 2), "MULFP": (2, 2),
    "INV": (2 + 2 + (_INV_SQR + _INV_MUL), 2 + 2 + (_INV_SQR + _INV_MUL)),
}


def _pre(slots, a, b, has_b, neg_b):
    x = slots[a]
    if has_b:
        y = slots[b]
        x = ((x[0] - y[0]) % P, (x[1] - y[1]) % P) if neg_b else ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
    return x


def lin_eval(terms, K, values):
    """terms: [(xi, m0, m1)], values: [(x0, x1)] canonical.  Returns the Fq2 result and checks the
    kernel's contract: with every input in [0, p) the lazy sum + K*p lies in [0, 1024 p)."""
    a0 = a1 = x0 = x1 = 0
    for (xi, m0, m1), v in zip(terms, values):
        assert 0 <= v[0] < P and 0 <= v[1] < P
        if xi:
            x0 += m0 * v[0]
            x1 += m1 * v[1]
        else:
            a0 += m0 * v[0]
            a1 += m1 * v[1]
    r0 = a0 + 9 * x0 - x1 + K * P
    r1 = a1 + x0 + 9 * x1 + K * P
    assert 0 <= r0 < 1024 * P and 0 <= r1 < 1024 * P, "LIN lazy-accumulation contract violated"
    return (r0 % P, r1 % P)


def lin_worst_case_ok(terms, K):
    """Static check of the same contract for ALL inputs in [0,p): evaluate at the extreme points."""
    lo = [0, 0]
    hi = [0, 0]
    for xi, m0, m1 in terms:
        contrib = ((0, 9 * m0), (0, -m1), (1, m0), (1, 9 * m1)) if xi else ((0, m0), (1, m1))
        for c, m in contrib:
            if m < 0:
                lo[c] += m
            else:
                hi[c] += m
    return all(lo[c] + K >= 0 and hi[c] + K < 1024 for c in (0, 1))


def run(words, consts, arrays, n_slots, n_scratch):
    """arrays: {arr_id: list of Fq ints}; STG writes into arrays[arr_id] (a dict or list)."""
    slots = [None] * n_slots
    scratch = [None] * max(n_scratch, 1)
    pc = 0
    while True:
        op, d, a, b, c, e, imm = isa.decode(words[pc])
        pc += 1
        if op == "END":
            break
        if op == "MUL":
            x = _pre(slots, a, b, imm & isa.MUL_B, imm & isa.MUL_BNEG)
            y = _pre(slots, c, e, imm & isa.MUL_E, imm & isa.MUL_ENEG)
            slots[d] = ((x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
        elif op == "SQR":
            x = _pre(slots, a, b, imm & isa.MUL_B, imm & isa.MUL_BNEG)
            slots[d] = ((x[0] * x[0] - x[1] * x[1]) % P, 2 * x[0] * x[1] % P)
        elif op == "MULFP":
            x, s = slots[a], slots[b][imm & 1]
            slots[d] = (x[0] * s % P, x[1] * s % P)
        elif op == "LIN":
            nterms, K = a, imm
            terms, vals = [], []
            for j in range(nterms):
                w = words[pc + j // 2]
                t = (w >> (32 * (j % 2))) & 0xFFFFFFFF
                slot, xi, m0, m1 = isa.decode_term(t)
                terms.append((xi, m0, m1))
                vals.append(slots[slot])
            pc += (nterms + 1) // 2
            assert lin_worst_case_ok(terms, K), "LIN K too small / range too large"
            slots[d] = lin_eval(terms, K, vals)
        elif op == "LDC":
            slots[d] = consts[imm]
        elif op == "LDG":
            slots[d] = (arrays[imm][a], arrays[imm][b])
        elif op == "STG":
            arrays[imm][d], arrays[imm][b] = slots[a]
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
    """Yield (opname, n_slot_moves) per instruction, skipping LIN term words."""
    pc = 0
    while pc < len(words):
        op, d, a, b, c, e, imm = isa.decode(words[pc])
        pc += 1
        if op == "LIN":
            pc += (a + 1) // 2
            yield op, a + 1
        elif op == "MUL":
            yield op, 3 + bool(imm & isa.MUL_B) + bool(imm & isa.MUL_E)
        elif op == "SQR":
            yield op, 2 + bool(imm & isa.MUL_B)
        elif op == "MULFP":
            yield op, 2.5
        elif op == "END":
            yield op, 0
            break
        else:
            yield op, {"LDC": 1, "LDG": 1, "STG": 1, "SPILL": 1, "FILL": 1, "INV": 2, "ADD": 3, "SUB": 3, "DBL": 2,
                       "NEG": 2, "CONJ": 2, "MULXI": 2}[op]


def work(words):
    """Algorithmic work of one program run: dict with Fp products, reductions, and 32x32 MACs
    (one product = 64 MACs, one reduction = 72), plus the opcode histogram and the
    number of 64-byte shared-memory slot moves."""
    prod = red = 0
    hist = {}
    moves = 0
    for op, mv in walk(words):
        hist[op] = hist.get(op, 0) + 1
        moves += mv
        if op in WORK:
            prod += WORK[op][0]
            red += WORK[op][1]
    return {"products": prod, "reductions": red, "macs": 64 * prod + 72 * red, "hist": hist, "slot_moves": moves}
