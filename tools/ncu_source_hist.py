"""Aggregate the per-SASS-instruction stall samples of an `ncu --page source --csv` dump by mnemonic and
by stall reason (read here, no GPU needed).   python tools/ncu_source_hist.py gpurun_out/prof_vm_src.csv"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = 0
    bymn, exe = collections.Counter(), collections.Counter()
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    st = collections.Counter()
    for r in data:
        s = int(r[ix["# Samples"]])
        tot += s
        mn = r[ix["Source"]].split()
        m = (mn[1] if mn[0].startswith("@") else mn[0]).rstrip(";")
        bymn[m] += s
        exe[m] += int(r[ix["Instructions Executed"]])
        for c in stall_cols:
            st[c] += int(r[ix[c]])
    print("SASS instructions", len(data), " total samples", tot, " executed", sum(exe.values()))
    for m, s in bymn.most_common(28):
        print("%-24s samples %7d (%5.1f%%)  executed %12d (%5.1f%%)" % (m, s, 100 * s / tot, exe[m], 100 * exe[m] / sum(exe.values())))
    print("stall reasons:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in st.most_common(10)))


if __name__ == "__main__":
    main()
