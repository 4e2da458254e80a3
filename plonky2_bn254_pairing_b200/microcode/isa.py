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
LDG / STG: imm = array id | page << 8; the Fq indices are page * 256 + the byte fields (both halves of a value lie in one
page), so that an array can hold more than 256 Fq per element (the line coefficients of prepared G2 points).

MUL / SQR take optional pre-additions (Karatsuba operands are sums of two slots):
    MUL  d = (a [+-b]) * (c [+-e])     imm bit0: b present, bit1: b subtracted, bit2: e present, bit3: e subtracted
    SQR  d = (a [+-b])^2               imm bit0: b present, bit1: b subtracted
MUL / SQR / MULFP (the "product class") share one epilogue, which is where most linear work of the tower
lives (measured: a stand-alone linear opcode costs ~250 sub-partition cycles, all of it shared-memory
traffic and dispatch, against ~100 cycles of arithmetic):
    T   = wide (512-bit, unreduced) product of the operands
    T  += 2^256 * sum(+-S[h_i])        up to three "hi terms": canonical slots added into the upper half
                                       of T before the Montgomery reduction, so r' picks them up for free
    r'  = canon(redc(T))               imm bits 5-6: number of extra conditional subtractions (bound / 2p)
    S[d]  = r'                         (skipped when only the post value is wanted)
    S[d2] = LIN(r', slots ...)         optional post stage: a LIN expression that may read r' (it is parked in
                                       slot d, or in d2 when it is not wanted for itself)
imm bit 4 (EXT) says an extension word follows:
    byte 0 d2, bytes 1-3 h1 h2 h3, byte 4 hflags (bits0-1 number of hi terms, bits 2-4 negate h1..h3,
    bit 7 store r' to d as well as the post value to d2), byte 5 = number of entry pairs of the post LIN;
then ceil(pairs / 2) words of 16-bit LIN entries.
ADD / SUB / DBL / NEG / CONJ / MULXI are the elementary linear opcodes (one modular step each).
LIN is the general linear instruction: each of the two output components is a lazily accumulated sum
    out_c = sum_j mult_j * z_j,   z_j = one Fq half of a slot, or p minus it,   1 <= mult_j <= 31
(this subsumes add, sub, neg, double, conj, multiplication by xi = 9 + u and by small constants), reduced once by a
quotient estimate.  Every entry is one 8-MAC IMAD.WIDE chain into a 64-bit-column accumulator - the accumulation
costs no ALU instructions.  Entries come in (component 0, component 1) pairs; header word: d, a = number of pairs;
entries follow, two pairs per word: [2 * slot + half : 9][neg:1][mult:6].

`emit_c_defines()` writes the opcode numbers into the generated header so the CUDA side cannot drift.
"""

OPS = [
    "END",    # stop
    "MUL",    # d = (a [+-b]) * (c [+-e])
    "SQR",    # d = (a [+-b])^2
    "MULFP",  # d = a * s, s = c0 or c1 (imm bit 15) half of slot b, an Fq scalar
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
ARR_STATE = 5 # [2 x state values][n]  values handed from one phase of a split program to the next

MUL_B, MUL_BNEG, MUL_E, MUL_ENEG = 1, 2, 4, 8
MUL_EXT = 16            # an extension word follows
MUL_BCANON = 0x200      # MUL: canonicalise the (a +- b) operand (makes room for a third hi term)
MUL_CANON_SHIFT = 5     # imm bits 5-6: extra conditional subtractions (0: r' < 2p, 1: < 4p, 2: < 6p)
EXT_STORE_R = 0x80
MULFP_HALF = 0x8000     # MULFP: the scalar is the c1 half of slot b
PRODUCT_OPS = ("MUL", "SQR", "MULFP")
LIN_MAX_ENT = 15    # entry pairs per LIN
LIN_MAX_MULT = 31
LIN_MAX_SUM = 1000  # sum of the multipliers of one component: the lazy value stays below 1024 p


def encode(op, d=0, a=0, b=0, c=0, e=0, imm=0):
    for v in (d, a, b, c, e):
        assert 0 <= v <= 0xFF, v
    assert 0 <= imm <= 0xFFFF
    return OPCODE[op] | (d << 8) | (a << 16) | (b << 24) | (c << 32) | (e << 40) | (imm << 48)


def decode(word):
    return (OPS[word & 0xFF], (word >> 8) & 0xFF, (word >> 16) & 0xFF, (word >> 24) & 0xFF, (word >> 32) & 0xFF,
            (word >> 40) & 0xFF, (word >> 48) & 0xFFFF)


def encode_ext(d2, hi, store_r, npairs):
    """hi: list of (slot, negate); npairs: LIN entry pairs of the post stage."""
    assert len(hi) <= 3 and 0 <= d2 <= 0xFF and 0 <= npairs <= LIN_MAX_ENT
    w = d2
    hflags = len(hi)
    for i, (slot, neg) in enumerate(hi):
        assert 0 <= slot <= 0xFF
        w |= slot << (8 * (i + 1))
        if neg:
            hflags |= 4 << i
    if store_r:
        hflags |= EXT_STORE_R
    return w | (hflags << 32) | (npairs << 40)


def decode_ext(w):
    hflags = (w >> 32) & 0xFF
    n = hflags & 3
    hi = [((w >> (8 * (i + 1))) & 0xFF, bool(hflags & (4 << i))) for i in range(n)]
    return w & 0xFF, hi, bool(hflags & EXT_STORE_R), (w >> 40) & 0xFF


# LIN entry (16 bits): out_component += (neg ? p - z : z) * mult,  z = half `half` (0: c0, 1: c1) of slot `slot`
def encode_entry(slot, half, mult, neg):
    """bits 0-8: 2 * slot + half - the index of the Fq the entry reads, so that the kernel gets both of its addresses
    (shared memory: index * T uint4, tensor memory: index * 4 columns) with one multiplication each."""
    assert 0 <= slot <= 0xFF and half in (0, 1) and 0 <= mult <= LIN_MAX_MULT
    return (2 * slot + half) | ((1 if neg else 0) << 9) | (mult << 10)


def decode_entry(t):
    return ((t & 0x1FF) >> 1, t & 1, (t >> 10) & 63, bool((t >> 9) & 1))


def pair_entries(ent0, ent1, pad_slot):
    """Interleave the two components' entry lists into (comp0, comp1) pairs; the shorter list is padded with
    multiplier-0 entries (they read `pad_slot`, any valid slot)."""
    n = max(len(ent0), len(ent1))
    pad = encode_entry(pad_slot, 0, 0, False)
    out = []
    for j in range(n):
        out.append(ent0[j] if j < len(ent0) else pad)
        out.append(ent1[j] if j < len(ent1) else pad)
    return out


def pack_entries(ents):
    """16-bit entries -> 64-bit words, four (two pairs) per word."""
    words = []
    for j in range(0, len(ents), 4):
        w = 0
        for k, t in enumerate(ents[j:j + 4]):
            w |= t << (16 * k)
        words.append(w)
    return words


class Ins:
    """One decoded instruction (with its extension / entry words)."""
    __slots__ = ("op", "d", "a", "b", "c", "e", "imm", "d2", "hi", "store_r", "ent0", "ent1", "nwords")

    def has_post(self):
        return bool(self.ent0 or self.ent1)

    def r_slot(self):
        """Slot that holds r' while a post stage runs."""
        return self.d if self.store_r else self.d2

    def slots_read(self):
        """Slot numbers the instruction reads (for range checks)."""
        r = []
        if self.op == "MUL":
            r = [self.a, self.c] + ([self.b] if self.imm & MUL_B else []) + ([self.e] if self.imm & MUL_E else [])
        elif self.op == "SQR":
            r = [self.a] + ([self.b] if self.imm & MUL_B else [])
        elif self.op in ("MULFP", "ADD", "SUB"):
            r = [self.a, self.b]
        elif self.op in ("INV", "DBL", "NEG", "CONJ", "MULXI", "STG", "SPILL"):
            r = [self.a]
        r += [s for s, _ in self.hi]
        r += [t[0] for t in self.ent0 + self.ent1]
        return r

    def slots_written(self):
        if self.op in PRODUCT_OPS:
            w = [self.d] if self.store_r else []
            return w + ([self.d2] if self.has_post() else [])
        if self.op in ("STG", "SPILL", "END"):
            return []
        return [self.d]


def parse(words):
    """Iterate over the instructions of a program (stops after END)."""
    pc = 0
    while pc < len(words):
        i = Ins()
        i.op, i.d, i.a, i.b, i.c, i.e, i.imm = decode(words[pc])
        start = pc
        pc += 1
        i.d2, i.hi, i.store_r, i.ent0, i.ent1 = 0, [], True, [], []
        npairs = 0
        if i.op in PRODUCT_OPS and i.imm & MUL_EXT:
            i.d2, i.hi, st, npairs = decode_ext(words[pc])
            pc += 1
            i.store_r = st or npairs == 0
        elif i.op == "LIN":
            npairs = i.a
        ents = []
        for j in range(2 * npairs):
            ents.append(decode_entry((words[pc + j // 4] >> (16 * (j % 4))) & 0xFFFF))
        i.ent0, i.ent1 = ents[0::2], ents[1::2]
        pc += (npairs + 1) // 2
        i.nwords = pc - start
        yield i
        if i.op == "END":
            return


def emit_c_defines():
    s = "".join("#define BNP_OP_%s %d\n" % (name, i) for i, name in enumerate(OPS))
    s += "#define BNP_MUL_B %d\n#define BNP_MUL_BNEG %d\n#define BNP_MUL_E %d\n#define BNP_MUL_ENEG %d\n" % (
        MUL_B, MUL_BNEG, MUL_E, MUL_ENEG)
    s += "#define BNP_ARR_AUX %d\n" % ARR_AUX
    s += "#define BNP_MUL_EXT %d\n#define BNP_MUL_CANON_SHIFT %d\n#define BNP_EXT_STORE_R %d\n#define BNP_MULFP_HALF %d\n#define BNP_MUL_BCANON %d\n" % (
        MUL_EXT, MUL_CANON_SHIFT, EXT_STORE_R, MULFP_HALF, MUL_BCANON)
    return s


def ldst_fields(f_lo, f_hi, arr):
    """(low byte of f_lo, low byte of f_hi, imm) of an LDG / STG on Fq indices f_lo, f_hi of array `arr`."""
    assert 0 <= arr <= 0xFF and f_lo >> 8 == f_hi >> 8 and f_lo >> 8 <= 0xFF, (f_lo, f_hi)
    return f_lo & 0xFF, f_hi & 0xFF, arr | ((f_lo >> 8) << 8)
