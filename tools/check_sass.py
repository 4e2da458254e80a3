"""Build-time check of the sequencer kernel's SASS (no GPU needed): the properties DESIGN.md argues from.

    python tools/check_sass.py [libbnp.so]

  * registers <= 170: the 12 warps of the 384-thread block must fit an SM's register file (3 warps x 32 x 170 per sub-partition);
  * every 32x32->64 multiply-accumulate is ONE instruction: IMAD.WIDE.U32[.X], or IMAD.HI.U32 for the first column of
    a Montgomery-reduction row (its low word is zero by construction, so ptxas keeps only the high half and the
    carry: 8 per reduction instance) - no mul.lo / mul.hi pairs left un-fused;
  * the kernel stays below 4 608 instructions (72 KB): the code that runs more than 100 times per task must fit the
    32 KB instruction cache;
  * the slot file really uses tensor memory: LDTM / STTM (tcgen05.ld / tcgen05.st) are present.
Prints the mnemonic histogram; exits non-zero if a property fails."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL = "bnp_vm_kernelILi384E"
MAX_REGS = 170
MAX_INSTRS = 4608
MAX_IMAD_HI = 8 * 10   # reduction instances: 2 in the product tail, the rest in the inversion / LIN / MULXI helpers


def main(lib=None):
    lib = lib or os.path.join(ROOT, "plonky2_bn254_pairing_b200", "libbnp.so")
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    regs = None
    lines = res.splitlines()
    for i, ln in enumerate(lines):
        if KERNEL in ln and i + 1 < len(lines):
            m = re.search(r"REG:(\d+)", lines[i + 1])
            regs = int(m.group(1)) if m else None
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    hist = collections.Counter()
    inside = False
    for ln in sass.splitlines():
        if "Function :" in ln:
            inside = KERNEL in ln
            continue
        if not inside:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m:
            hist[m.group(1)] += 1
    n = sum(hist.values())
    wide = sum(v for k, v in hist.items() if k.startswith("IMAD.WIDE.U32"))
    hi = hist.get("IMAD.HI.U32", 0)
    print("kernel %s: %d registers, %d instructions (%.1f KB), %d IMAD.WIDE.U32[.X], %d IMAD.HI.U32" % (
        KERNEL, regs or -1, n, n * 16 / 1024.0, wide, hi))
    print("top mnemonics:", ", ".join("%s %d" % kv for kv in hist.most_common(12)))
    ok = True
    if regs is None or regs > MAX_REGS:
        print("FAIL: registers", regs, ">", MAX_REGS)
        ok = False
    if n > MAX_INSTRS:
        print("FAIL: kernel has", n, "instructions >", MAX_INSTRS)
        ok = False
    if hi > MAX_IMAD_HI:
        print("FAIL:", hi, "IMAD.HI.U32 - un-fused multiply pairs?")
        ok = False
    tm = sum(v for k, v in hist.items() if k.startswith("LDTM") or k.startswith("STTM"))
    print("tensor-memory accesses (LDTM / STTM):", tm)
    if tm == 0:
        print("FAIL: no LDTM / STTM - the tensor-memory half of the slot file is missing")
        ok = False
    if hist.get("IMAD.HI", 0):
        print("FAIL: signed IMAD.HI present")
        ok = False
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else None))
