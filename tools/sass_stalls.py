"""Static issue-cycle estimate per basic block from `cuobjdump -sass`: decodes the control word of every
sm_100 instruction (stall count, yield, scoreboard waits) and prints, per basic block, #instructions,
#IMAD.WIDE and the sum of stall counts (= minimum cycles one warp alone needs to issue the block when no
variable-latency scoreboard wait fires).  usage: sass_stalls.py lib.so 'kernel-substring' [--dump lo hi]"""
import re
import subprocess
import sys


def parse(lib, kern):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
    out, on, cur = [], False, None
    for ln in txt:
        if "Function :" in ln:
            on = kern in ln
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", ln)
        if m:
            cur = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]
            continue
        m = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", ln)
        if m and cur:
            cur[3] = int(m.group(1), 16)
            out.append(cur)
            cur = None
    return out


def ctrl(w1):
    return {"stall": (w1 >> 41) & 0xF, "yield": (w1 >> 45) & 1, "wbar": (w1 >> 46) & 7, "rbar": (w1 >> 49) & 7,
            "wait": (w1 >> 52) & 0x3F}


def main():
    ins = parse(sys.argv[1], sys.argv[2])
    if len(sys.argv) > 3 and sys.argv[3] == "--dump":
        lo, hi = int(sys.argv[4], 16), int(sys.argv[5], 16)
        for a, s, w0, w1 in ins:
            if lo <= a <= hi:
                c = ctrl(w1)
                print("%06x s%-2d %s w%02x wb%d rb%d  %s" % (a, c["stall"], "Y" if c["yield"] else " ", c["wait"], c["wbar"], c["rbar"], s))
        return
    targets = set()
    for a, s, w0, w1 in ins:
        m = re.search(r"(?:BRA|BSSY\S*|CALL\S*|BRX)\s.*?(0x[0-9a-f]+)", s)
        if m:
            targets.add(int(m.group(1), 16))
    blocks, cur = [], []
    for rec in ins:
        if rec[0] in targets and cur:
            blocks.append(cur)
            cur = []
        cur.append(rec)
        if re.match(r"(@!?U?P\d+\s+)?(BRA|EXIT|RET|BRX|BREAK|BSYNC)", rec[1]):
            blocks.append(cur)
            cur = []
    if cur:
        blocks.append(cur)
    print("start    n_ins  wide  imad  stall_sum  waits")
    for b in blocks:
        n = len(b)
        if n < int(sys.argv[3]) if len(sys.argv) > 3 else n < 40:
            continue
        wide = sum("IMAD.WIDE" in r[1] for r in b)
        imad = sum(r[1].split()[0].startswith("IMAD") or (r[1].startswith("@") and "IMAD" in r[1].split()[1]) for r in b)
        st = sum(ctrl(r[3])["stall"] for r in b)
        wt = sum(1 for r in b if ctrl(r[3])["wait"])
        print("%06x  %5d  %4d  %4d  %6d  %5d" % (b[0][0], n, wide, imad, st, wt))


if __name__ == "__main__":
    main()
