"""Instruction set of the Fq2 sequencer kernel (csrc/vm.cu).

The GPU kernel is a per-thread register machine whose registers ("slots") are Fq2 values held in
shared memory; every thread runs the same straight-line program on its own pairing, so control
flow is uniform across the grid.  One instruction is a 64-bit word:

    bits  0..7   opcode
    bits  8..19  d   (destination slot, or an Fq index for STG)
    bits 20..31  a   (source slot,      or an Fq index for LDG)
    bits 32..43  b   (source slot,      or an Fq index for LDG/STG)
    bits 44..63  imm (constant index, array id, scratch index, flags)

`emit_c_defines()` writes the opcode numbers into the generated header so the CUDA side cannot drift.
"""

OPS = [
    "END",    # stop
    "MUL",    # d = a * b                      (Fq2)
    "SQR",    # d = a^2
    "MULFP",  # d = a * s, s = c0 (imm=0) or c1 (imm=1) half of slot b, an Fq scalar
    "ADD",    # d = a + b
    "SUB",    # d = a - b
    "NEG",    # d = -a
    "CONJ",   # d = conj(a) = (a.c0, -a.c1)
    "MULXI",  # d = a * (9 + u)
    "MOV",    # d = a
    "LDC",    # d = const[imm]
    "LDG",    # d = (arr[imm][a], arr[imm][b])   two Fq of this thread's element of global array imm
    "STG",    # arr[imm][d], arr[imm][b] = a.c0, a.c1
    "SPILL",  # scratch[imm] = a
    "FILL",   # d = scratch[imm]
    "INV",    # d = 1 / a   (Fq2)
    "DBL",    # d = a + a
]
OPCODE = {name: i for i, name in enumerate(OPS)}

# global array ids (kernel argument `arr[]`)
ARR_G1 = 0    # [2k Fq][n]   (x, y) of each G1 point; k points per element for multi-pairing programs
ARR_G2 = 1    # [4k Fq][n]   (x.c0, x.c1, y.c0, y.c1)
ARR_F12 = 2   # [12 Fq][n]   MyFq12 input  (coeffs[0..11])
ARR_OUT = 3   # [12 Fq][n]   MyFq12 output
ARR_AUX = 4   # second output / input array (program specific)

FIELD_MAX = 0xFFF
IMM_MAX = 0xFFFFF


def encode(op, d=0, a=0, b=0, imm=0):
    assert 0 <= d <= FIELD_MAX and 0 <= a <= FIELD_MAX and 0 <= b <= FIELD_MAX and 0 <= imm <= IMM_MAX
    return OPCODE[op] | (d << 8) | (a << 20) | (b << 32) | (imm << 44)


def decode(word):
    return (OPS[word & 0xFF], (word >> 8) & 0xFFF, (word >> 20) & 0xFFF, (word >> 32) & 0xFFF, (word >> 44) & 0xFFFFF)


def emit_c_defines():
    return "".join("#define BNP_OP_%s %d\n" % (name, i) for i, name in enumerate(OPS))
