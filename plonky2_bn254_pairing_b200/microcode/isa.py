"""Instruction set of the Fq2 sequencer kernel (csrc/vm.cuh).

The GPU kernel is a per-thread register machine whose registers ("slots") are Fq2 values held in
shared memory; every thread runs the same straight-line program on its own pairing, so control
flow is uniform across the grid.  One instruction word is 64 bits, eight byte-wide fields:

    byte 0  opcode
    byte 1  d     destination slot          (STG: Fq index of the c0 half)
    byte 2  a     source slot               (LDG: Fq index of the c0 half)
    byte 3  b     source slot               (LDG/STG: Fq index of the c1 half)
    byte 4  c     source slot
    byte 5  e     source slot
    byte 6-7 imm  16-bit immediate (constant index, array id, scratch index, flags)

MUL / SQR take optional pre-additions (Karatsuba operands are sums of two slots):
    MUL  d = (a [+-b]) * (c [+-e])     imm bit0: b present, bit1: b subtracted, bit2: e present, bit3: e subtracted
    SQR  d = (a [+-b])^2               imm bit0: b present, bit1: b subtracted
ADD / SUB / DBL / NEG / CONJ / MULXI are the elementary linear opcodes (one modular step each).
LIN is an optional fused form (kept for experiments; measured slower on B200 than the elementary ops,
see DESIGN.md section 6): a variable-length instruction
    d = sum_i diag(m0_i, m1_i) * x_i  +  xi * sum_j diag(m0_j, m1_j) * x_j        (xi = 9 + u)
with small signed integer multipliers per Fq component (this subsumes add, sub, neg, double, conj,
multiplication by xi and by small constants).  Header word: d, a = number of terms, imm = K (the
multiple of p that makes the lazily accumulated value non-negative); each following word carries
two 32-bit term descriptors: [slot:8][flags:8][m0:int8][m1:int8], flags bit0 = term belongs to the xi sum.

`emit_c_defines()` writes the opcode numbers into the generated header so the CUDA side cannot drift.
"""

OPS = [
    "END",    # stop
    "MUL",    # d = (a [+-b]) * (c [+-e])
    "SQR",    # d = (a [+-b])^2
    "MULFP",  # d = a * s, s = c0 (imm=0) or c1 (imm=1) half of slot b, an Fq scalar
    "LIN",    # d = linear combination (variable length, see above)
    "LDC",    # d = const[imm]
    "LDG",    # d = (arr[imm][a], arr[imm][b])   two Fq of this thread's element of global array imm
    "STG",    # arr[imm][d], arr[imm][b] = a.c0, a.c1
    "SPILL",  # scratch[imm] = a
    "FILL",   # d = scratch[imm]
    "INV",    # d = 1 / a   (Fq2)
    "ADD",    # d = a + b            elementary linear ops: canonical in, canonical out
    "SUB",    # d = a - b
    "DBL",    # d = 2 a
    "NEG",    # d = -a
    "CONJ",   # d = conj(a)
    "MULXI",  # d = (9 + u) a
]
OPCODE = {name: i for i, name in enumerate(OPS)}

# global array ids (kernel argument `arr[]`)
ARR_G1 = 0    # [2k Fq][n]   (x, y) of each G1 point; k points per element for multi-pairing programs
ARR_G2 = 1    # [4k Fq][n]   (x.c0, x.c1, y.c0, y.c1)
ARR_F12 = 2   # [12 Fq][n]   MyFq12 input  (coeffs[0..11])
ARR_OUT = 3   # [12 Fq][n]   MyFq12 output
ARR_AUX = 4   # second input array (program specific)

MUL_B, MUL_BNEG, MUL_E, MUL_ENEG = 1, 2, 4, 8
LIN_XI = 1
LIN_MAX_TERMS = 8
LIN_MAX_MULT = 31
LIN_MAX_K = 63      # size of the K*p table in the kernel


def encode(op, d=0, a=0, b=0, c=0, e=0, imm=0):
    for v in (d, a, b, c, e):
        assert 0 <= v <= 0xFF, v
    assert 0 <= imm <= 0xFFFF
    return OPCODE[op] | (d << 8) | (a << 16) | (b << 24) | (c << 32) | (e << 40) | (imm << 48)


def decode(word):
    return (OPS[word & 0xFF], (word >> 8) & 0xFF, (word >> 16) & 0xFF, (word >> 24) & 0xFF, (word >> 32) & 0xFF,
            (word >> 40) & 0xFF, (word >> 48) & 0xFFFF)


def encode_term(slot, xi, m0, m1):
    assert 0 <= slot <= 0xFF and -LIN_MAX_MULT <= m0 <= LIN_MAX_MULT and -LIN_MAX_MULT <= m1 <= LIN_MAX_MULT
    return slot | ((LIN_XI if xi else 0) << 8) | ((m0 & 0xFF) << 16) | ((m1 & 0xFF) << 24)


def decode_term(t):
    def s8(v):
        return v - 256 if v >= 128 else v

    return (t & 0xFF, bool((t >> 8) & LIN_XI), s8((t >> 16) & 0xFF), s8((t >> 24) & 0xFF))


def emit_c_defines():
    s = "".join("#define BNP_OP_%s %d\n" % (name, i) for i, name in enumerate(OPS))
    s += "#define BNP_MUL_B %d\n#define BNP_MUL_BNEG %d\n#define BNP_MUL_E %d\n#define BNP_MUL_ENEG %d\n" % (
        MUL_B, MUL_BNEG, MUL_E, MUL_ENEG)
    s += "#define BNP_LIN_XI %d\n#define BNP_LIN_MAX_K %d\n" % (LIN_XI, LIN_MAX_K)
    return s
