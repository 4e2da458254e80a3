"""GPU exploration run (development tool): op-level parity, pairing parity, IMAD peak, and a
(slots x threads-per-block) throughput sweep over prebuilt libbnp_s<slots>.so variants.

    python tools/gpu_explore.py [--n 16384] [--sweep]
Writes gpurun_out/explore.json.
"""
import argparse
import ctypes
import glob
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bn254_oracle as O  # noqa: E402
from plonky2_bn254_pairing_b200 import api, native  # noqa: E402


def dev_tensor(a):
    return torch.from_numpy(a.view(np.int64)).cuda()


def optest(lib):
    import random

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import optest_expect as X
    from plonky2_bn254_pairing_b200.microcode.programs import OPTEST_OUTPUTS

    rows = X.edge_rows(random.Random(7))
    n = len(rows)
    d_in = dev_tensor(api.pack_soa(rows))
    d_out = torch.zeros((2 * OPTEST_OUTPUTS, 4, n), dtype=torch.int64, device="cuda")
    native.check(lib.bnp_run_program_dev(0, None, b"optest", None, None, d_in.data_ptr(), None, d_out.data_ptr(), n))
    torch.cuda.synchronize()
    try:
        out = api.unpack_soa(d_out.cpu().numpy().view(np.uint64))
    except native.BnpError as ex:
        print("optest: non-canonical output", ex)
        return False
    bad = 0
    for e in range(n):
        want = X.expected(rows[e])
        for i, nm in enumerate(X.NAMES):
            got = (out[e][2 * i], out[e][2 * i + 1])
            if got != want[i]:
                bad += 1
                if bad < 12:
                    print("optest mismatch elem", e, nm, rows[e][:2], hex(got[0])[:20], hex(want[i][0])[:20])
    print("optest:", "OK" if bad == 0 else "FAIL %d" % bad)
    return bad == 0


def pairing_parity(npts=12):
    pts = O.seeded_points(0xB2540002, npts)
    Ps = [p for p, _ in pts]
    Qs = [q for _, q in pts]
    ok = True
    got = api.pairing_batch(Ps, Qs)
    ok &= all(g == O.pairing(p, q) for (p, q), g in zip(pts, got))
    print("pairing_v0 parity:", ok)
    got1 = api.pairing_batch(Ps, Qs, variant=1)
    ok1 = all(g == O.final_exp_ark(O.miller_loop_native(q, p)) for (p, q), g in zip(pts, got1))
    print("pairing_v1 parity:", ok1)
    gm = api.miller_loop_native_batch(Qs, Ps)
    okm = all(g == O.miller_loop_native(q, p) for (p, q), g in zip(pts, gm))
    print("miller parity:", okm)
    ms = [O.miller_loop_native(q, p) for (p, q) in pts]
    gf = api.final_exp_native_batch(ms)
    okf = all(g == O.final_exp_native(m) for m, g in zip(ms, gf))
    print("final_exp parity:", okf)
    prods = [pts[0:3], pts[3:6], pts[6:9], pts[9:12]]
    gmm = api.multi_miller_loop_native_batch(prods)
    okmm = all(g == O.multi_miller_loop_native(pr) for pr, g in zip(prods, gmm))
    print("multi miller x3 parity:", okmm)
    gp = api.pairing_product(pts)
    acc = ms[0]
    for m in ms[1:]:
        acc = O.fq12_mul(acc, m)
    okp = gp == O.final_exp_native(acc)
    print("pairing product parity:", okp)
    return ok and ok1 and okm and okf and okmm and okp


def make_inputs(n, pool=256):
    """Pool of valid points by an additive walk; pair i = (P[i % K], Q[(i // K + 7 i) % K])."""
    p0, q0 = O.seeded_points(0xB2540003, 1)[0]
    dp, dq = O.seeded_points(0xB2540004, 1)[0]
    Ps, Qs = [p0], [q0]
    for _ in range(pool - 1):
        Ps.append(O.g1_add(Ps[-1], dp))
        Qs.append(O.g2_add(Qs[-1], dq))
    g1p = api.pack_soa(api.g1_rows(Ps))
    g2p = api.pack_soa(api.g2_rows(Qs))
    idx = np.arange(n)
    i1 = idx % pool
    i2 = (idx // pool + 7 * idx) % pool
    return np.ascontiguousarray(g1p[:, :, i1]), np.ascontiguousarray(g2p[:, :, i2]), Ps, Qs, i1, i2


def time_prog(lib, prog, n, g1, g2, reps=3):
    d1, d2 = dev_tensor(g1), dev_tensor(g2)
    out = torch.zeros((12, 4, n), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    stream = torch.cuda.Stream()  # a real (non-default) stream so the events bracket the kernel
    best = None
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            native.check(lib.bnp_run_program_dev(0, ctypes.c_void_p(stream.cuda_stream), prog.encode(), d1.data_ptr(),
                                                 d2.data_ptr(), None, None, out.data_ptr(), n))
            e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if r > 0:
            best = ms if best is None else min(best, ms)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--threads", type=int, nargs="*", default=[32, 64, 128])
    ap.add_argument("--progs", nargs="*", default=["pairing_v0"])
    args = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    res = {}
    lib = native.init([0])
    if not args.no_parity:
        res["optest"] = optest(lib)
        res["parity"] = pairing_parity()
    peak = ctypes.c_double()
    native.check(lib.bnp_imad_peak(0, ctypes.byref(peak)))
    res["imad_peak_macs_per_s"] = peak.value
    peak32 = ctypes.c_double()
    native.check(lib.bnp_imad32_peak(0, ctypes.byref(peak32)))
    res["imad32_peak_per_s"] = peak32.value
    print("IMAD.WIDE peak: %.3e MAC/s   IMAD(32) peak: %.3e /s" % (peak.value, peak32.value))
    g1, g2, Ps, Qs, i1, i2 = make_inputs(args.n)
    macs = lib.bnp_program_macs(b"pairing_v0")
    runs = []
    libs = [("default", native.LIB_PATH)]
    if args.sweep:
        libs = [(os.path.basename(p), p) for p in sorted(glob.glob(os.path.join(ROOT, "plonky2_bn254_pairing_b200", "libbnp_s*.so")))]
    for tag, path in libs:
        l2 = native.load(path) if path != native.LIB_PATH else lib
        dev0 = (ctypes.c_int * 1)(0)
        native.check(l2.bnp_init(dev0, 1), l2)
        for T in args.threads:
            native.check(l2.bnp_set_launch_config(T, 0), l2)
            for prog in args.progs:
                try:
                    ms, out = time_prog(l2, prog, args.n, g1, g2)
                except native.BnpError as ex:
                    print(tag, T, prog, "ERR", ex)
                    continue
                rate = args.n / (ms * 1e-3)
                frac = rate * l2.bnp_program_macs(prog.encode()) / peak.value
                runs.append({"lib": tag, "T": T, "prog": prog, "ms": ms, "pairings_per_s": rate, "imad_frac": frac})
                print("%-16s T=%3d %-12s %8.2f ms  %10.0f /s  imad %.3f" % (tag, T, prog, ms, rate, frac))
        # spot check of the last output against the oracle
        o = api.unpack_soa(out[:, :, :2].cpu().numpy().view(np.uint64))
        ok = all(o[j] == O.pairing(Ps[i1[j]], Qs[i2[j]]) for j in range(2))
        print(tag, "spot parity:", ok)
        runs.append({"lib": tag, "spot_parity": ok})
    res["runs"] = runs
    res["macs_pairing_v0"] = macs
    with open(os.path.join(ROOT, "gpurun_out", "explore.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
