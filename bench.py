#!/usr/bin/env python
"""bench.py - BN254 pairings/sec on N B200s (one process per GPU), next to the CPU reference port.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload pairing|miller|final_exp|groth16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm's CPU port on all host cores

A step is one pass of the hot path over one batch of 2^16 synthetic pairings per GPU (BASELINE.json
configs[1]; weak scaling: every rank gets its own 2^16).  `value` is device-resident throughput
(inputs already in HBM), `e2e` goes through the host-buffer C ABI (bnp_pairing_batch: H2D + kernel + D2H
inside the timed region, pinned host memory).  The roofline is the integer-multiply pipe, measured live
on the box by the library's IMAD.WIDE microbenchmark (MEASURED_PEAKS.json has no integer peak).

Every number is parity-gated: after the timed region >= 1024 sampled outputs of every rank are compared with the
C oracle (oracle/_build/libbn254_ref.so, the checker - never the thing measured); on a mismatch no line is
printed and the exit code is 3.  `extra` (short runs outside the headline timing, --no-extras skips them) carries
the other BASELINE configurations at their full sizes (2^20 Miller loops, 2^20 final exponentiations, 2^18
Groth16-shaped products, 2^22 pairings STRONG-scaled over the ranks), the one real collective of the path (ONE
product over 2^20 pairs: per-GPU fused Miller loops -> tree product -> 384-byte NCCL all-gather -> one final
exponentiation, with the tail broken out), the latency of a single pairing and e2e from pageable host memory.
Prints exactly one JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bn254_pairings_per_sec"
UNIT = "pairings/s"
BATCH = 1 << 16
WORKLOADS = {
    # name: (program, pairs per element, label)
    "pairing": ("pairing_v0", 1, "2^16 independent BN254 pairings per GPU (Miller loop + final exp fused), BASELINE configs[1]"),
    "miller": ("miller", 1, "2^16 reference-exact Miller loops per GPU"),
    "final_exp": ("final_exp_v0", 1, "2^16 final exponentiations per GPU"),
    "groth16": ("pairing_x4_v0", 4, "2^16 Groth16-shaped 4-way pairing products per GPU (4 pairings each)"),
}


POOL_K = 4096   # pool of K G1 x K G2 subgroup points, all index pairs distinct up to 2^24


def program_stats(prog):
    """Instruction histogram of the shipped program (build-time tables, recomputed here from the same sources):
    how many sequencer instructions, spills and re-loads one pairing costs."""
    try:
        from plonky2_bn254_pairing_b200.microcode import gen

        _, progs = gen.build_all(names={prog}, with_phases=False)
        _, al, w = progs[0]
        h = w["hist"]
        return {"instructions": sum(h.values()), "spill": h.get("SPILL", 0), "fill": h.get("FILL", 0),
                "products": h.get("MUL", 0) + h.get("SQR", 0) + h.get("MULFP", 0), "slots": al.n_slots,
                "scratch_entries": al.n_scratch}
    except Exception as e:  # noqa: BLE001 - diagnostics only
        return {"error": repr(e)}


def reference_chain_macs(prog):
    """MACs of the same program when the hard part raises to BN_X the way the reference does (the 62-squaring,
    23-multiplication NAF walk of pow_native, final_exp_native.rs:56-84) instead of the shipped width-4 windows
    (16 multiplications): the figure round 1 was quoted on, reported next to the shipped schedule's own count."""
    try:
        from plonky2_bn254_pairing_b200.microcode import gen

        old = os.environ.get("BNP_POWX_WINDOW")
        os.environ["BNP_POWX_WINDOW"] = "2"
        try:
            _, progs = gen.build_all(names={prog}, with_phases=False)
        finally:
            if old is None:
                del os.environ["BNP_POWX_WINDOW"]
            else:
                os.environ["BNP_POWX_WINDOW"] = old
        return int(progs[0][2]["macs"])
    except Exception:  # noqa: BLE001 - diagnostics only
        return None


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        mhz, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                mhz.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(mhz) if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz)}


# --------------------------------------------------------------------------------------------- CPU reference port
def load_cref():
    """oracle/_build/libbn254_ref.so - the C port of the reference's native path (checker / CPU baseline)."""
    path = os.path.join(ROOT, "oracle", "_build", "libbn254_ref.so")
    if not os.path.exists(path):
        import __graft_entry__ as g

        g.build_oracle()
    lib = ctypes.CDLL(path)
    vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.bn254_ref_pairing_batch.argtypes = [vp, vp, vp, sz, ci, ci, ci]
    lib.bn254_ref_pairing_batch.restype = None
    lib.bn254_ref_miller_batch.argtypes = [vp, vp, vp, sz, ci, ci, ci]
    lib.bn254_ref_miller_batch.restype = None
    lib.bn254_ref_final_exp_batch.argtypes = [vp, vp, sz, ci, ci]
    lib.bn254_ref_final_exp_batch.restype = None
    lib.bn254_ref_max_threads.restype = ci
    return lib


def cpu_run(cref, workload, g1, g2, f12, n, threads):
    """One pass of the reference port (faithful = as written in the reference) over n elements."""
    import numpy as np

    out = np.empty((12, 4, n), dtype=np.uint64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    prog, k, _ = WORKLOADS[workload]
    t0 = time.perf_counter()
    if workload == "miller":
        cref.bn254_ref_miller_batch(p(g1), p(g2), p(out), n, 1, 1, threads)
    elif workload == "final_exp":
        cref.bn254_ref_final_exp_batch(p(f12), p(out), n, 1, threads)
    else:
        cref.bn254_ref_pairing_batch(p(g1), p(g2), p(out), n, k, 1, threads)
    return time.perf_counter() - t0, out


def cpu_inputs(workload, n):
    import numpy as np

    from plonky2_bn254_pairing_b200 import workload as wl

    k = WORKLOADS[workload][1]
    g1, g2, _ = wl.pairing_inputs(n, K=POOL_K, k=k)
    f12 = None
    if workload == "final_exp":
        rs = np.random.RandomState(5)
        # random canonical Fq12 inputs (final_exp_native.rs:266-286 uses a random element too)
        f12 = rs.randint(0, 1 << 62, size=(12, 4, n)).astype(np.uint64)
        f12[:, 3, :] &= np.uint64((1 << 60) - 1)
    return g1, g2, f12


def cpu_baseline(workload, target_seconds=12.0):
    """Bounded sample of the same workload on all host cores; returns the cpu_baseline object."""
    cref = load_cref()
    cores = cref.bn254_ref_max_threads()
    k = WORKLOADS[workload][1]
    n0 = max(cores * 2, 8)
    g1, g2, f12 = cpu_inputs(workload, n0)
    dt, _ = cpu_run(cref, workload, g1, g2, f12, n0, cores)
    n = int(max(n0, min(BATCH, n0 * target_seconds / max(dt, 1e-3))))
    g1, g2, f12 = cpu_inputs(workload, n)
    dt, _ = cpu_run(cref, workload, g1, g2, f12, n, cores)
    return {"value": n * k / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d elements (%d pairings) of the same workload, oracle/bn254_ref.c reference-faithful mode, "
                      "%d threads, %.1f s" % (n, n * k, cores, dt)}


def run_reference(args):
    """--impl reference: the reference's own algorithm (C port, rayon-style over all host cores)."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    cref = load_cref()
    cores = cref.bn254_ref_max_threads()
    k = WORKLOADS[args.workload][1]
    n0 = max(cores * 2, 8)
    g1, g2, f12 = cpu_inputs(args.workload, n0)
    dt, _ = cpu_run(cref, args.workload, g1, g2, f12, n0, cores)
    per_step_seconds = 4.0
    n = int(max(n0, min(BATCH, n0 * per_step_seconds / max(dt, 1e-3))))
    g1, g2, f12 = cpu_inputs(args.workload, n)
    for _ in range(args.warmup):
        cpu_run(cref, args.workload, g1, g2, f12, n, cores)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_run(cref, args.workload, g1, g2, f12, n, cores)
        t += dt
    value = n * k * args.steps / t
    sample = ("each step = %d elements (%d pairings) of the workload, reference-faithful C port of "
              "miller_loop_native.rs/final_exp_native.rs/pairing.rs, %d host threads" % (n, n * k, cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64x4 Montgomery (CPU)", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][2], "sample_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------- GPU arm
def oracle_outputs(cref, workload, g1, g2, f12):
    """The C oracle (memoised mode, all host threads) on the given columns - the parity checker."""
    import numpy as np

    n = (f12 if workload == "final_exp" else g1).shape[2]
    out = np.empty((12, 4, n), dtype=np.uint64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    k = WORKLOADS[workload][1]
    if workload == "miller":
        cref.bn254_ref_miller_batch(p(g1), p(g2), p(out), n, 1, 0, 0)
    elif workload == "final_exp":
        cref.bn254_ref_final_exp_batch(p(f12), p(out), n, 0, 0)
    else:
        cref.bn254_ref_pairing_batch(p(g1), p(g2), p(out), n, k, 0, 0)
    return out


class Rig:
    """One rank's device state: library handle, stream, helpers to run a workload device-resident."""

    def __init__(self, local, rank, world):
        import torch

        from plonky2_bn254_pairing_b200 import native

        self.torch, self.native = torch, native
        self.local, self.rank, self.world = local, rank, world
        self.dev = torch.device("cuda", local)
        self.lib = native.init([local])
        self.stream = torch.cuda.Stream(device=self.dev)
        self.sp = ctypes.c_void_p(self.stream.cuda_stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.cref = None

    def to_dev(self, a):
        import numpy as np

        return self.torch.from_numpy(a.view(np.int64)).to(self.dev)

    def barrier(self):
        import torch.distributed as dist

        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        import torch.distributed as dist

        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(self, ok):
        import torch.distributed as dist

        if self.world == 1:
            return ok
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def inputs(self, workload, n, offset):
        """Device-resident inputs of `n` elements of the workload (plus the host copies for parity / e2e)."""
        from plonky2_bn254_pairing_b200 import workload as wl

        torch = self.torch
        prog, k, _ = WORKLOADS[workload]
        g1, g2, _ = wl.pairing_inputs(n, K=POOL_K, k=k, offset=offset)
        d = {"g1": g1, "g2": g2, "f12": None, "d_g1": self.to_dev(g1), "d_g2": self.to_dev(g2), "d_f12": None, "n": n, "k": k,
             "prog": prog, "workload": workload}
        d["d_out"] = torch.empty((12, 4, n), dtype=torch.int64, device=self.dev)
        if workload == "final_exp":
            # inputs = GPU Miller outputs of the same indices
            d["d_f12"] = torch.empty((12, 4, n), dtype=torch.int64, device=self.dev)
            self.native.check(self.lib.bnp_run_program_dev(self.local, self.sp, b"miller", d["d_g1"].data_ptr(),
                                                           d["d_g2"].data_ptr(), None, None, d["d_f12"].data_ptr(), n))
            self.torch.cuda.synchronize(self.dev)
        return d

    def launch(self, w):
        self.native.check(self.lib.bnp_run_program_dev(
            self.local, self.sp, w["prog"].encode(), w["d_g1"].data_ptr(), w["d_g2"].data_ptr(),
            w["d_f12"].data_ptr() if w["d_f12"] is not None else None, None, w["d_out"].data_ptr(), w["n"]))

    def timed(self, w, steps, warmup):
        """(total ms over `steps` launches incl. the L2 flushes, [kernel ms per launch]) - CUDA events on the launch stream."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                self.flush.zero_()
                self.launch(w)
        self.barrier()
        k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            t0.record(self.stream)
            for i in range(steps):
                self.flush.zero_()  # L2 flush between timed iterations (inside the timed region, ~0.1% of a step)
                k_ev[i][0].record(self.stream)
                self.launch(w)
                k_ev[i][1].record(self.stream)
            t1.record(self.stream)
        self.barrier()
        return t0.elapsed_time(t1), [a.elapsed_time(b) for a, b in k_ev]

    def parity(self, w, n_check, seed):
        """Compare `n_check` sampled outputs of the last launch with the C oracle.  Returns (checked, ok)."""
        import numpy as np

        if self.cref is None:
            self.cref = load_cref()
        n = w["n"]
        m = min(n_check, n)
        idx = np.sort(np.random.RandomState(seed + 7919 * self.rank).choice(n, m, replace=False))
        got = w["d_out"][:, :, self.torch.from_numpy(idx).to(self.dev)].cpu().numpy().view(np.uint64)
        col = lambda a: None if a is None else np.ascontiguousarray(a[:, :, idx])  # noqa: E731
        f12 = None
        if w["d_f12"] is not None:
            f12 = np.ascontiguousarray(w["d_f12"][:, :, self.torch.from_numpy(idx).to(self.dev)].cpu().numpy().view(np.uint64))
        want = oracle_outputs(self.cref, w["workload"], col(w["g1"]), col(w["g2"]), f12)
        return m, bool(np.array_equal(got, want))


def run_extras(rig, peak, args):
    """Short device-resident runs of the other BASELINE configurations and of the one collective (outside the headline
    timing); every one parity-checked on a sample."""
    import numpy as np
    import torch
    import torch.distributed as dist

    lib, native, world, rank, local = rig.lib, rig.native, rig.world, rig.rank, rig.local
    extra = {"workloads": {}}
    cfgs = [  # (tag, workload, total elements, strong?)
        ("miller_2e20", "miller", 1 << 20, False),
        ("final_exp_2e20", "final_exp", 1 << 20, False),
        ("groth16_2e18", "groth16", 1 << 18, False),
        ("pairing_2e22_strong", "pairing", 1 << 22, True),
    ]
    for tag, wk, n_total, strong in cfgs:
        n = n_total // world if strong else n_total
        w = rig.inputs(wk, n, offset=(1 << 24) + rank * n)
        total_ms, kern = rig.timed(w, steps=2, warmup=1)
        total_ms = rig.max_over_ranks(total_ms)
        checked, ok = rig.parity(w, 256, seed=11)
        ok = rig.all_ok(ok)
        macs = lib.bnp_program_macs(w["prog"].encode())
        kavg = sum(kern) / len(kern)
        extra["workloads"][tag] = {
            "program": w["prog"], "elements_per_gpu": n, "scaling": "strong" if strong else "weak",
            "value": world * n * w["k"] * 2 / (total_ms * 1e-3), "unit": UNIT, "kernel_ms_avg": kavg,
            "frac": n * macs / (kavg * 1e-3) / peak, "parity": {"checked": checked * world, "ok": ok}}
        del w
        torch.cuda.empty_cache()
        if not ok:
            return extra, False

    # ---- Groth16 shape with the verifying key's three G2 points PREPARED: e(A_i, B_i) * prod_j e(C_ij, vk_j)
    # per proof; B_i varies, the vk points are one set of line coefficients shared by the whole batch
    from plonky2_bn254_pairing_b200 import api
    from plonky2_bn254_pairing_b200 import workload as wl

    n = 1 << 18
    g1, g2, _ = wl.pairing_inputs(n, K=POOL_K, k=4, offset=(5 << 24) + rank * n)
    vk = np.ascontiguousarray(g2[4:, :, :1].reshape(3, 4, 4).transpose(1, 2, 0))          # the three fixed points, [4][4][3]
    coeffs = api.g2_prepare_soa(vk)                                                           # [PREP_FQ][4][3]
    prepared = np.ascontiguousarray(coeffs.transpose(2, 0, 1).reshape(3 * api.PREP_FQ, 4, 1))
    g2_fixed = np.ascontiguousarray(g2.copy())
    g2_fixed[4:, :, :] = g2[4:, :, :1]                                                        # the same proofs, unprepared
    d_g1, d_g2v, d_g2f, d_prep = rig.to_dev(g1), rig.to_dev(np.ascontiguousarray(g2[:4])), rig.to_dev(g2_fixed), rig.to_dev(prepared)
    d_o1 = torch.empty((12, 4, n), dtype=torch.int64, device=rig.dev)
    d_o2 = torch.empty((12, 4, n), dtype=torch.int64, device=rig.dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.cuda.stream(rig.stream):
        for rep in range(2):
            if rep:
                ev[0].record()
            native.check(lib.bnp_pairing_prepared_dev(local, rig.sp, d_g1.data_ptr(), d_g2v.data_ptr(), d_prep.data_ptr(),
                                                      d_o1.data_ptr(), n, 1, 3, 0))
            if rep:
                ev[1].record()
        ev[2].record()
        native.check(lib.bnp_run_program_dev(local, rig.sp, b"pairing_x4_v0", d_g1.data_ptr(), d_g2f.data_ptr(), None, None,
                                             d_o2.data_ptr(), n))
        ev[3].record()
    torch.cuda.synchronize(rig.dev)
    same = rig.all_ok(bool(torch.equal(d_o1, d_o2)))
    ms_prep = rig.max_over_ranks(ev[0].elapsed_time(ev[1]))
    extra["groth16_prepared_2e18"] = {
        "program": "pairing_p1_3_v0", "proofs_per_gpu": n, "proofs_per_s": world * n / (ms_prep * 1e-3),
        "pairings_per_s": world * n * 4 / (ms_prep * 1e-3), "kernel_ms": ms_prep,
        "unprepared_kernel_ms": rig.max_over_ranks(ev[2].elapsed_time(ev[3])),
        "macs_per_proof": lib.bnp_program_macs(b"pairing_p1_3_v0"),
        "frac": n * lib.bnp_program_macs(b"pairing_p1_3_v0") / (ms_prep * 1e-3) / peak,
        "bit_equal_to_unprepared_on_every_element": same}
    del d_g1, d_g2v, d_g2f, d_prep, d_o1, d_o2
    torch.cuda.empty_cache()
    if not same:
        return extra, False

    # ---- wire formats: 2^16 compressed ark points decoded on the device (host bytes in, SoA out; the
    # G2 half includes one Fq2 square root and the subgroup test per point), checked against the points they encode
    okw = True
    if rank == 0:
        import time as _time

        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import wire_formats as W

        nw = 1 << 16
        g1w, g2w, _ = wl.pairing_inputs(nw, K=POOL_K, offset=(7 << 24))
        pts1, pts2 = api.unpack_soa(g1w[:, :, :256]), api.unpack_soa(g2w[:, :, :256])
        b1 = b"".join(W.encode_g1((r[0], r[1]), W.COMPRESSED) for r in pts1) * (nw // 256)
        b2 = b"".join(W.encode_g2(((r[0], r[1]), (r[2], r[3])), W.COMPRESSED) for r in pts2) * (nw // 256)
        # steady state: the first full-size call of each kind pays for the library's buffers (cudaMalloc after torch
        # has just released gigabytes is slow and erratic: 10 ms ... 1 s), the second is timed
        for rep in range(2):
            t0 = _time.perf_counter()
            d1, s1 = api.decode_g1_soa(W.COMPRESSED, b1)
            t1 = _time.perf_counter()
        for rep in range(2):
            t1b = _time.perf_counter()
            d2, s2 = api.decode_g2_soa(W.COMPRESSED, b2, check_subgroup=True)
            t2 = _time.perf_counter()
        okw = (not s1.any() and not s2.any() and np.array_equal(d1[:, :, :256], g1w[:, :, :256])
               and np.array_equal(d2[:, :, :256], g2w[:, :, :256]) and np.array_equal(d2[:, :, -256:], g2w[:, :, :256]))
        extra["wire_decode_2e16"] = {
            "g1_compressed_points_per_s": nw / (t1 - t0), "g2_compressed_with_subgroup_check_points_per_s": nw / (t2 - t1b),
            "api": "bnp_decode_g1_batch / bnp_decode_g2_batch (host pointers, copies and status bytes included)",
            "equal_to_the_encoded_points": bool(okw)}
    if not rig.all_ok(bool(okw)):
        return extra, False

    # ---- scalar multiplication: 2^16 random 254-bit scalars on G1 and on G2 through the host-pointer
    # entry, checked through the pairing on a sample: e(k P, Q) == e(P, k Q)
    oks = True
    if rank == 0:
        import time as _time

        ns = 1 << 16
        g1s, g2s, _ = wl.pairing_inputs(ns, K=POOL_K, offset=(9 << 24))
        rng = np.random.default_rng(0xB2540F05)
        ks = rng.integers(0, 1 << 63, size=(4, ns), dtype=np.uint64)
        ks[3] &= np.uint64((1 << 61) - 1)   # below r: the scalars a verifier draws
        for rep in range(2):   # the second call of each kind is timed (see the wire formats above)
            t0 = _time.perf_counter()
            kp, i1 = api.scalar_mul_soa(1, g1s, ks)
            t1 = _time.perf_counter()
        for rep in range(2):
            t1b = _time.perf_counter()
            kq, i2 = api.scalar_mul_soa(2, g2s, ks)
            t2 = _time.perf_counter()
        m = 512
        a = api.pairing_soa(np.ascontiguousarray(kp[:, :, :m]), np.ascontiguousarray(g2s[:, :, :m]))
        b = api.pairing_soa(np.ascontiguousarray(g1s[:, :, :m]), np.ascontiguousarray(kq[:, :, :m]))
        oks = bool(not i1.any() and not i2.any() and np.array_equal(a, b))
        extra["scalar_mul_2e16"] = {
            "g1_points_per_s": ns / (t1 - t0), "g2_points_per_s": ns / (t2 - t1b),
            "api": "bnp_scalar_mul_batch (host pointers, copies included), 254-bit scalars, one thread per point",
            "pairing_commutes_on_512_sampled": oks}
    if not rig.all_ok(bool(oks)):
        return extra, False

    # ---- the one real exchange step: ONE product over 2^20 pairs, sharded by index range over the ranks
    from plonky2_bn254_pairing_b200 import sharding
    from plonky2_bn254_pairing_b200 import workload as wl

    n_total = 1 << 20
    off, cnt = sharding.shard_range(n_total, rank, world)
    g1, g2, _ = wl.pairing_inputs(cnt, K=POOL_K, offset=(3 << 24) + off)
    d_g1, d_g2 = rig.to_dev(g1), rig.to_dev(g2)
    d_ml = torch.empty((12, 4, cnt), dtype=torch.int64, device=rig.dev)
    d_part = torch.zeros((12, 4, 1), dtype=torch.int64, device=rig.dev)
    d_fe = torch.zeros((12, 4, 1), dtype=torch.int64, device=rig.dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    one_stream = ctypes.c_void_p(1)  # cudaStreamLegacy: ordered with torch's default stream and with NCCL
    best = None
    for it in range(3):
        rig.barrier()
        ev[0].record()
        native.check(lib.bnp_miller_loop_fused_dev(local, one_stream, d_g1.data_ptr(), d_g2.data_ptr(), d_ml.data_ptr(), cnt))
        ev[1].record()
        native.check(lib.bnp_fq12_product_dev(local, one_stream, d_ml.data_ptr(), d_part.data_ptr(), cnt))
        ev[2].record()
        if world > 1:
            gathered = sharding.all_gather_fq12(d_part)
            total = torch.zeros((12, 4, 1), dtype=torch.int64, device=rig.dev)
            native.check(lib.bnp_fq12_product_dev(local, one_stream, gathered.data_ptr(), total.data_ptr(), world))
        else:
            total = d_part
        ev[3].record()
        native.check(lib.bnp_final_exp_dev(local, one_stream, total.data_ptr(), d_fe.data_ptr(), 1, 0))
        ev[4].record()
        torch.cuda.synchronize(rig.dev)
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
        if it and (best is None or sum(t) < sum(best)):
            best = t
    tot_ms = rig.max_over_ranks(sum(best))
    same = True
    if world > 1:
        chk = [torch.empty_like(d_fe) for _ in range(world)]
        dist.all_gather(chk, d_fe)
        same = all(torch.equal(c, chk[0]) for c in chk)
    # checker: the same product, from the oracle's Miller values of a 512-pair prefix... the full product is covered by
    # tests/test_gpu_configs.py::test_config5; here: every rank holds the same 384 bytes, and they are a canonical
    # element of the order-r subgroup's image (x^r == 1 is too slow to test here), so only equality is asserted
    extra["global_product_2e20"] = {
        "pairs": n_total, "pairs_per_gpu": cnt, "total_ms": tot_ms, "pairs_per_s": n_total / (tot_ms * 1e-3),
        "miller_ms": best[0], "tree_us": best[1] * 1e3, "allgather_and_combine_us": best[2] * 1e3, "final_exp_us": best[3] * 1e3,
        "collective": "NCCL all_gather of one 384-byte Fq12 per rank" if world > 1 else "none (1 rank)",
        "bit_equal_on_every_rank": bool(same)}
    del d_g1, d_g2, d_ml
    torch.cuda.empty_cache()

    # ---- a single pairing through the host ABI (the reference's scalar `pairing(p, q)` call), and pageable-memory e2e
    if rank == 0:
        g1, g2, _ = wl.pairing_inputs(BATCH, K=POOL_K, offset=0)
        o1 = np.empty((12, 4, 1), dtype=np.uint64)
        a1, b1 = np.ascontiguousarray(g1[:, :, :1]), np.ascontiguousarray(g2[:, :, :1])
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        native.check(lib.bnp_pairing_batch(p(a1), p(b1), p(o1), 1, 0))
        t0 = time.perf_counter()
        for _ in range(5):
            native.check(lib.bnp_pairing_batch(p(a1), p(b1), p(o1), 1, 0))
        extra["latency_one_pairing_us"] = (time.perf_counter() - t0) / 5 * 1e6
        out = np.empty((12, 4, BATCH), dtype=np.uint64)  # plain (pageable) numpy memory, like a Rust Vec
        native.check(lib.bnp_pairing_batch(p(g1), p(g2), p(out), BATCH, 0))
        t0 = time.perf_counter()
        for _ in range(3):
            native.check(lib.bnp_pairing_batch(p(g1), p(g2), p(out), BATCH, 0))
        extra["e2e_pageable"] = {"value": BATCH * 3 / (time.perf_counter() - t0), "unit": UNIT,
                                 "note": "2^16 pairings per call, caller buffers in pageable memory"}
    rig.barrier()
    return extra, bool(same)


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints a "NCCL version ..." banner on STDOUT when the first communicator is created; stdout carries
        # ONE JSON line, so the file descriptor points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize(local)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    rig = Rig(local, rank, world)
    lib, native, dev = rig.lib, rig.native, rig.dev

    prog, k, label = WORKLOADS[args.workload]
    n = args.batch
    w = rig.inputs(args.workload, n, offset=rank * n)

    # roofline denominator, measured live (rank-local; reported from rank 0)
    peak = ctypes.c_double()
    native.check(lib.bnp_imad_peak(local, ctypes.byref(peak)))
    peak32 = ctypes.c_double()
    native.check(lib.bnp_imad32_peak(local, ctypes.byref(peak32)))
    macs = lib.bnp_program_macs(prog.encode())             # algorithmic (Karatsuba Fq2 products), SURVEY 8(d)
    macs_x = lib.bnp_program_macs_executed(prog.encode())  # what the kernel issues (one thread = one whole Fq2 operation)

    sampler = ClockSampler(local)
    warm = max(args.warmup, 3)
    with torch.cuda.stream(rig.stream):
        for _ in range(warm):
            rig.flush.zero_()
            rig.launch(w)
    rig.barrier()
    if rank == 0:
        sampler.start()
    launches0 = lib.bnp_launch_count()
    elapsed_ms, kern_ms = rig.timed(w, args.steps, 0)
    launches = lib.bnp_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = rig.max_over_ranks(elapsed_ms)
    value = world * n * k * args.steps / (elapsed_ms * 1e-3)

    # ---- parity gate: sampled outputs of the timed launches against the C oracle, on every rank
    checked, ok = rig.parity(w, args.parity, seed=1)
    ok = rig.all_ok(ok)
    if not ok:
        sys.stderr.write("bench.py: PARITY FAILURE - device results differ from the oracle; no line printed\n")
        if world > 1:
            dist.destroy_process_group()
        return 3

    # ---- end to end through the host-buffer C ABI (pinned host memory, copies inside the timed region)
    h_g1 = torch.from_numpy(w["g1"].view(np.int64)).pin_memory()
    h_g2 = torch.from_numpy(w["g2"].view(np.int64)).pin_memory()
    h_out = torch.empty((12, 4, n), dtype=torch.int64).pin_memory()
    h_f12 = w["d_f12"].cpu().pin_memory() if w["d_f12"] is not None else None

    def e2e_call():
        if args.workload == "pairing":
            native.check(lib.bnp_pairing_batch(h_g1.data_ptr(), h_g2.data_ptr(), h_out.data_ptr(), n, 0))
        elif args.workload == "miller":
            native.check(lib.bnp_miller_loop_batch(h_g1.data_ptr(), h_g2.data_ptr(), h_out.data_ptr(), n))
        elif args.workload == "final_exp":
            native.check(lib.bnp_final_exp_batch(h_f12.data_ptr(), h_out.data_ptr(), n, 0))
        else:
            native.check(lib.bnp_multi_pairing_batch(h_g1.data_ptr(), h_g2.data_ptr(), h_out.data_ptr(), n, k, 0))

    e2e_steps = max(3, min(args.steps, 10))
    e2e_call()
    rig.barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_call()  # synchronous: returns after the D2H copy completed
    torch.cuda.synchronize(dev)
    e2e_s = rig.max_over_ranks(time.perf_counter() - w0)
    e2e_value = world * n * k * e2e_steps / e2e_s
    in_bytes = (h_f12.numel() * 8) if args.workload == "final_exp" else (h_g1.numel() + h_g2.numel()) * 8
    # the e2e result is the same computation: bit-equal to the device-resident result that was just parity-checked
    e2e_same = rig.all_ok(bool(torch.equal(h_out, w["d_out"].cpu())))
    if not e2e_same:
        sys.stderr.write("bench.py: e2e and device-resident results differ; no line printed\n")
        return 3

    extra, extra_ok = (None, True)
    if not args.no_extras:
        del h_g1, h_g2, h_out
        extra, extra_ok = run_extras(rig, peak.value, args)
        if not extra_ok:
            sys.stderr.write("bench.py: PARITY FAILURE in the extra workloads; no line printed\n")
            return 3

    if rank == 0:
        kavg = sum(kern_ms) / len(kern_ms)
        achieved = n * macs / (kavg * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        alg_bytes = n * (k * 192 + 384) if args.workload != "final_exp" else n * 768
        threads = lib.bnp_threads_per_block()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (8x32-bit-limb Fp254 Montgomery, IMAD.WIDE.U32 carry chains)", "data": "synthetic",
            "config": {"workload": label, "program": prog, "batch_per_gpu": n, "pairings_per_element": k,
                       "parallelism": "index-sharded, %d rank(s), no data-path collective" % world,
                       "l2": "256 MiB memset between timed iterations (inside the timed region)",
                       "inputs": "pool of %d G1 x %d G2 subgroup points, pair i = (P[i%%K], Q[(i/K+7i)%%K])" % (POOL_K, POOL_K)},
            "parity": {"checked": checked * world, "ok": True, "e2e_bit_equal": True,
                       "oracle": "oracle/_build/libbn254_ref.so (C restatement of the reference, memoised mode), "
                                 "%d sampled elements per rank after the timed region" % checked},
            "roofline": {
                "bound": "imad", "achieved": achieved / 1e9, "peak": peak.value / 1e9, "unit": "GMAC/s",
                "frac": achieved / peak.value,
                # DRAM bytes cannot be measured outside a profiler: see profiles/ncu_r2_*.txt for the captured value
                # (the spill scratch lives in L2; algorithmic I/O is 576 B per pairing - not the bound)
                "traffic": None,
                "kernel": "bnp_vm_kernel<%d>" % threads, "kernel_ms_avg": kavg, "macs_per_element": macs,
                "executed": {"macs_per_element": macs_x, "gmacs": n * macs_x / (kavg * 1e-3) / 1e9,
                             "pipe_frac": n * macs_x / (kavg * 1e-3) / peak.value,
                             "note": "one thread runs a whole Karatsuba Fq2 operation: executed == algorithmic MACs"},
                "program": program_stats(prog),
                "reference_chain": (lambda m: None if not m else {
                    "macs_per_element": m, "frac": n * m / (kavg * 1e-3) / peak.value,
                    "note": "the same launch quoted on the MACs the reference's own addition chain for m^x needs (NAF walk, "
                            "23 multiplications per exponentiation; shipped: width-4 windows, 16): what the kernel would "
                            "be credited with had the schedule not been shortened"})(reference_chain_macs(prog)),
                "peak_source": "measured live: bnp_imad_peak (IMAD.WIDE.U32[.X] 4-deep carry chains, 32 MACs/thread/iter); "
                               "MEASURED_PEAKS.json has no integer peak",
                "imad32_peak_gops": peak32.value / 1e9,
                "hbm": {"achieved_gbs": alg_bytes / (kavg * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "frac": alg_bytes / (kavg * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback", "note": "informational; not the bound"},
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": n * 384,
                    "steps": e2e_steps, "api": "bnp_%s_batch (host pointers, synchronous)" % (
                        "pairing" if args.workload == "pairing" else "miller_loop" if args.workload == "miller"
                        else "final_exp" if args.workload == "final_exp" else "multi_pairing")},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if extra is not None:
            line["extra"] = extra
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pairing", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=BATCH, help="elements per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the `extra` workloads (other configs, global product)")
    ap.add_argument("--parity", type=int, default=1024, help="outputs per rank compared with the oracle after the timed region")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
