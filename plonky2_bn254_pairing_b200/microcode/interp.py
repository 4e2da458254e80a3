"""Big-integer interpreter for allocated sequencer programs (build/test tool, never on the product path).

Executes exactly the instruction words the GPU runs - including spills, fills and re-loads - on one
"thread", with plain (non-Montgomery) residues.  Montgomery form is a representation detail of the
kernel: every opcode is a ring operation, so the plain-domain run is isomorphic to the device run.
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


def run(words, consts, arrays, n_slots, n_scratch):
    """arrays: {arr_id: list of Fq ints}; STG writes into arrays[arr_id] (a dict or list)."""
    slots = [None] * n_slots
    scratch = [None] * max(n_scratch, 1)
    for w in words:
        op, d, a, b, imm = isa.decode(w)
        if op == "END":
            break
        if op == "MUL":
            x, y = slots[a], slots[b]
            slots[d] = ((x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
        elif op == "SQR":
            x = slots[a]
            slots[d] = ((x[0] * x[0] - x[1] * x[1]) % P, 2 * x[0] * x[1] % P)
        elif op == "MULFP":
            x, s = slots[a], slots[b][imm]
            slots[d] = (x[0] * s % P, x[1] * s % P)
        elif op == "ADD":
            x, y = slots[a], slots[b]
            slots[d] = ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
        elif op == "SUB":
            x, y = slots[a], slots[b]
            slots[d] = ((x[0] - y[0]) % P, (x[1] - y[1]) % P)
        elif op == "NEG":
            x = slots[a]
            slots[d] = ((-x[0]) % P, (-x[1]) % P)
        elif op == "CONJ":
            x = slots[a]
            slots[d] = (x[0], (-x[1]) % P)
        elif op == "MULXI":
            x = slots[a]
            slots[d] = ((9 * x[0] - x[1]) % P, (x[0] + 9 * x[1]) % P)
        elif op == "MOV":
            slots[d] = slots[a]
        elif op == "DBL":
            x = slots[a]
            slots[d] = (2 * x[0] % P, 2 * x[1] % P)
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
        elif op == "INV":
            x = slots[a]
            n = (x[0] * x[0] + x[1] * x[1]) % P
            ni = pow(n, P - 2, P)  # 0 -> 0, like the device ladder
            slots[d] = (x[0] * ni % P, (-x[1]) * ni % P)
        else:
            raise ValueError(op)
    return arrays


def work(words):
    """Algorithmic work of one program run: dict with Fp products, reductions, and 32x32 MACs
    (one product = 64 MACs, one reduction = 72)."""
    prod = red = 0
    hist = {}
    for w in words:
        op = isa.OPS[w & 0xFF]
        hist[op] = hist.get(op, 0) + 1
        if op in WORK:
            prod += WORK[op][0]
            red += WORK[op][1]
    return {"products": prod, "reductions": red, "macs": 64 * prod + 72 * red, "hist": hist}
