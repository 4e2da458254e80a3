#!/bin/bash
# Per-opcode instruction / pipe counts: runs the opbench_* programs (2048 back-to-back instances of one opcode)
# under ncu and prints warp-instructions per op.   usage: tools/gpu_opcount.sh "mul sqr add" [n]
OPS_LIST=${1:-"mul sqr add"}
N=${2:-65536}
export OPS="$OPS_LIST"
ncu --metrics smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_lsu.sum,gpu__time_duration.sum \
    --clock-control none -k regex:bnp_vm_kernel --csv --log-file gpurun_out/opcount.csv python tools/gpu_opbench.py $N > gpurun_out/opcount.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/opcount.csv") if l.startswith('"')))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[ix["ID"]], {"blk": r[ix["Block Size"]]})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
ops = "$OPS_LIST".split()
n = $N
k = 0
for T in (32, 64, 128):
    for op in ops:
        for rep in range(3):
            m = list(per.values())[k]; k += 1
            if rep == 2 and T == 64:
                w = n / 32 * 2048
                print("%-8s T=%d inst/op %7.1f  alu %6.1f  fma %6.1f  fmaheavy %6.1f  lsu %5.1f  time %.2f ms" % (
                    op, T, m["smsp__inst_executed.sum"] / w, m["sm__inst_executed_pipe_alu.sum"] / w,
                    m["sm__inst_executed_pipe_fma.sum"] / w, m["sm__inst_executed_pipe_fmaheavy.sum"] / w,
                    m["sm__inst_executed_pipe_lsu.sum"] / w, m["gpu__time_duration.sum"] / 1e6))
PY
