"""Per-opcode cost model: time the opbench_* programs (2048 instances of one opcode) at full occupancy.
Prints cycles per warp-op per SMSP-resident-warp and the implied share of a pairing."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from plonky2_bn254_pairing_b200 import native  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
    lib = native.init([0])
    if os.environ.get("BNP_LIB_VARIANT"):
        lib = native.load(os.environ["BNP_LIB_VARIANT"])
        dev0 = (ctypes.c_int * 1)(0)
        native.check(lib.bnp_init(dev0, 1), lib)
    rs = np.random.RandomState(1)
    f = rs.randint(0, 1 << 62, size=(12, 4, n)).astype(np.uint64)
    f[:, 3, :] &= np.uint64((1 << 60) - 1)
    d_in = torch.from_numpy(f.view(np.int64)).cuda()
    d_out = torch.zeros((12, 4, n), dtype=torch.int64, device="cuda")
    stream = torch.cuda.Stream()
    res = {}
    for T in (64,):
        native.check(lib.bnp_set_launch_config(T, 0))
        for op in os.environ.get("OPS", "mul sqr mulfp add sub dbl neg mulxi lin4 lin4xi muls").split():
            prog = ("opbench_" + op).encode()
            best = None
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    e0.record(stream)
                    native.check(lib.bnp_run_program_dev(0, ctypes.c_void_p(stream.cuda_stream), prog, None, None,
                                                         d_in.data_ptr(), None, d_out.data_ptr(), n))
                    e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
            ns_per_op_per_thread = best * 1e6 / 2048 / n  # aggregate: device time per (thread, op)
            # cycles of one SM sub-partition per warp-op: time * clock / (warp-ops per SMSP)
            warp_ops_per_smsp = (n / 16) * 2048 / (148 * 4)   # a warp holds 16 pairings (two lanes each)
            cyc = best * 1e-3 * 1.965e9 / warp_ops_per_smsp
            res["%s_T%d" % (op, T)] = {"ms": best, "cycles_per_warp_op_per_smsp": cyc}
            print("T=%3d %-6s %8.2f ms   %7.1f SMSP-cycles per warp-op" % (T, op, best, cyc))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "opbench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
