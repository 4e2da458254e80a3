"""Instruction set of the Fq2 sequencer kernel (csrc/vm.cuh).

The GPU kernel is a register machine whose registers ("slots") are Fq2 values held in shared memory.
Every pairing owns one slot file and TWO lanes of a warp ("units" A and B); the program is a sequence of
BUNDLES, one instruction per unit, both of the same opcode and executed by the same handler at the same
time (control flow is uniform across the grid; the units differ only in operand fields and per-unit flags).
A bundle is: instruction word of unit A, instruction word of unit B (op byte bit 7 set: unit B idles),
then - product class with EXT - one extension word per unit, then the LIN entry words (one word per entry
index: unit A's (c0, c1) pair in the low half, unit B's in the high half), padded to an even word count so
that every bundle starts 16-byte aligned.  One instruction word is 64 bits, eight byte-wide fields:

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
imm bit 4 (EXT) says extension words follow (one per unit):
    byte 0 d2, bytes 1-3 h1 h2 h3, byte 4 hflags (bits0-1 number of hi terms of the BUNDLE, bits 2-4 negate h1..h3,
    bit 7 store r' to d as well as the post value to d2), byte 5 = number of entry pairs of the post LIN of the
    BUNDLE, byte 6 = which of h1..h3 this unit really has (the others are skipped by masking), byte 7 = entry
    pairs of this unit's own post LIN (0: the unit has no post stage and only stores r');
then one word of 16-bit LIN entries per entry index.
The two units of a bundle may differ in every slot field, in the presence and sign of pre-additions and hi terms
(handled by masks / selects when the bundle is mixed, by uniform branches when it is not), in store_r and in
the post LIN (the shorter entry list is padded with multiplier-0 entries).
ADD / SUB / DBL / NEG / CONJ / MULXI are the elementary linear opcodes (one modular step each).
LIN is the general linear instruction: each of the two output components is a lazily accumulated sum
    out_c = sum_j mult_j * z_j,   z_j = one Fq half of a slot, or p minus it,   1 <= mult_j <= 31
(this subsumes add, sub, neg, double, conj, multiplication by xi = 9 + u and by small constants), reduced once by a
quotient estimate.  Every entry is one 8-MAC IMAD.WIDE chain into a 64-bit-column accumulator - the accumulation
costs no ALU instructions.  Entries come in (component 0, component 1) pairs; header words: d, a = number of pairs
of the bundle; entry words follow, one per entry index: [slot:8][half:1][neg:1][mult:6] x (c0, c1) x (A, B).

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


UNIT_IDLE = 0x80        # op byte of unit B's word: the unit has no instruction of its own (it shadows unit A, stores nothing)
CONTROL_FLAGS = MUL_EXT | MUL_BCANON | (15 << MUL_CANON_SHIFT)   # imm bits that must agree inside a bundle


def encode_ext(d2, hi, store_r, np_bundle, n_hi_bundle=None, np_own=None):
    """hi: list of (slot, negate) of THIS unit, placed in positions 0..len-1; n_hi_bundle >= len(hi) is the
    bundle's count (absent positions name a valid slot and are masked out)."""
    if n_hi_bundle is None:
        n_hi_bundle = len(hi)
    if np_own is None:
        np_own = np_bundle
    assert len(hi) <= n_hi_bundle <= 3 and 0 <= d2 <= 0xFF and 0 <= np_own <= np_bundle <= LIN_MAX_ENT
    w = d2
    hflags = n_hi_bundle
    present = 0
    pad = hi[0][0] if hi else d2
    for i in range(n_hi_bundle):
        slot, neg = hi[i] if i < len(hi) else (pad, False)
        assert 0 <= slot <= 0xFF
        w |= slot << (8 * (i + 1))
        if i < len(hi):
            present |= 1 << i
            if neg:
                hflags |= 4 << i
    if store_r:
        hflags |= EXT_STORE_R
    return w | (hflags << 32) | (np_bundle << 40) | (present << 48) | (np_own << 56)


def decode_ext(w):
    hflags = (w >> 32) & 0xFF
    n = hflags & 3
    present = (w >> 48) & 0xFF
    hi = [((w >> (8 * (i + 1))) & 0xFF, bool(hflags & (4 << i))) for i in range(n) if present & (1 << i)]
    return w & 0xFF, hi, bool(hflags & EXT_STORE_R), (w >> 56) & 0xFF


# LIN entry (16 bits): out_component += (neg ? p - z : z) * mult,  z = half `half` (0: c0, 1: c1) of slot `slot`
def encode_entry(slot, half, mult, neg):
    assert 0 <= slot <= 0xFF and half in (0, 1) and 0 <= mult <= LIN_MAX_MULT
    return slot | (half << 8) | ((1 if neg else 0) << 9) | (mult << 10)


def decode_entry(t):
    return (t & 0xFF, (t >> 8) & 1, (t >> 10) & 63, bool((t >> 9) & 1))


def entry_words(ent_a, ent_b, pad_a, pad_b):
    """ent_x = (ent0, ent1) of unit x as lists of 16-bit entries (or ([], []) for a unit without a LIN).
    One 64-bit word per entry index: A.c0 | A.c1 << 16 | B.c0 << 32 | B.c1 << 48; short lists are padded
    with multiplier-0 entries that read `pad_x` (any valid slot)."""
    n = max(len(ent_a[0]), len(ent_a[1]), len(ent_b[0]), len(ent_b[1]))
    words = []
    for j in range(n):
        w = 0
        for k, (lst, pad) in enumerate(((ent_a[0], pad_a), (ent_a[1], pad_a), (ent_b[0], pad_b), (ent_b[1], pad_b))):
            t = lst[j] if j < len(lst) else encode_entry(pad, 0, 0, False)
            w |= t << (16 * k)
        words.append(w)
    return words


class Ins:
    """One decoded instruction of one unit (with its extension word / entries)."""
    __slots__ = ("op", "d", "a", "b", "c", "e", "imm", "d2", "hi", "store_r", "ent0", "ent1", "nwords", "unit",
                 "idle", "np_own", "ext")

    def has_post(self):
        return self.np_own > 0 if self.op in PRODUCT_OPS else bool(self.ent0 or self.ent1)

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


def parse_bundles(words):
    """Iterate over the bundles of a program as (ins_a, ins_b) - ins_b.idle says unit B only shadows A (stops after END)."""
    pc = 0
    while pc < len(words):
        start = pc
        units = []
        for u in (0, 1):
            i = Ins()
            w = words[pc + u]
            i.unit, i.idle = u, bool(w & UNIT_IDLE)
            i.op, i.d, i.a, i.b, i.c, i.e, i.imm = decode(w & ~UNIT_IDLE)
            i.d2, i.hi, i.store_r, i.ent0, i.ent1, i.np_own, i.ext = 0, [], True, [], [], 0, False
            units.append(i)
        a, b = units
        assert a.op == b.op and not a.idle
        pc += 2
        npairs = 0
        if a.op in PRODUCT_OPS and a.imm & MUL_EXT:
            for u, i in enumerate(units):
                i.ext = True
                i.d2, i.hi, st, i.np_own = decode_ext(words[pc + u])
                i.store_r = st or i.np_own == 0
            npairs = (words[pc] >> 40) & 0xFF
            assert npairs == (words[pc + 1] >> 40) & 0xFF
            pc += 2
        elif a.op == "LIN":
            npairs = a.a
            for i in units:
                i.np_own = npairs
        for j in range(npairs):
            w = words[pc + j]
            for u, i in enumerate(units):
                i.ent0.append(decode_entry((w >> (32 * u)) & 0xFFFF))
                i.ent1.append(decode_entry((w >> (32 * u + 16)) & 0xFFFF))
        pc += npairs
        pc += (pc - start) & 1
        for i in units:
            i.nwords = pc - start
            if i.op in PRODUCT_OPS and i.np_own == 0:
                i.ent0, i.ent1 = [], []
        yield a, b
        if a.op == "END":
            return


def parse(words):
    """Iterate over the instructions of a program, unit A then unit B of every bundle (idle units are skipped)."""
    for a, b in parse_bundles(words):
        yield a
        if not b.idle and a.op != "END":
            yield b


def emit_c_defines():
    s = "".join("#define BNP_OP_%s %d\n" % (name, i) for i, name in enumerate(OPS))
    s += "#define BNP_MUL_B %d\n#define BNP_MUL_BNEG %d\n#define BNP_MUL_E %d\n#define BNP_MUL_ENEG %d\n" % (
        MUL_B, MUL_BNEG, MUL_E, MUL_ENEG)
    s += "#define BNP_MUL_EXT %d\n#define BNP_MUL_CANON_SHIFT %d\n#define BNP_EXT_STORE_R %d\n#define BNP_MULFP_HALF %d\n#define BNP_MUL_BCANON %d\n" % (
        MUL_EXT, MUL_CANON_SHIFT, EXT_STORE_R, MULFP_HALF, MUL_BCANON)
    return s
