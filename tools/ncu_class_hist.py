"""Attribute the stall samples of an `ncu --page source --csv` dump to code regions, identified by how often
each SASS instruction executed per warp-task (all warps run the same straight-line program, so e.g. the MUL
handler's instructions all executed exactly #MUL times).  usage: ncu_class_hist.py src.csv n_warp_tasks"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    ntask = float(sys.argv[2]) if len(sys.argv) > 2 else 2048.0
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    cls = {}
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    for r in data:
        e, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        t = r[ix["Source"]].split()
        m = (t[1] if t[0].startswith("@") else t[0]).rstrip(";")
        c = cls.setdefault(round(e / ntask), {"n": 0, "s": 0, "mn": collections.Counter()})
        c["n"] += 1
        c["s"] += s
        c["mn"][m] += 1
    print("exec/task  #sass  samples%  instr/task   top mnemonics")
    for k, c in sorted(cls.items(), key=lambda kv: -kv[1]["s"])[:int(sys.argv[3]) if len(sys.argv) > 3 else 24]:
        print("%8d  %5d  %5.1f%%  %9d   %s" % (k, c["n"], 100 * c["s"] / tot, k * c["n"],
                                             " ".join("%s:%d" % kv for kv in c["mn"].most_common(5))))


if __name__ == "__main__":
    main()
