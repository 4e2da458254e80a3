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


def ncu_traffic_bytes():
    """dram__bytes_read + dram__bytes_write of the sequencer kernel, per launch, from the committed `ncu --set full`
    capture of this same workload (profiles/ncu_r1_split.txt); None if the summary is missing."""
    try:
        tot = 0.0
        for ln in open(os.path.join(ROOT, "profiles", "ncu_r1_split.txt")):
            f = ln.split()
            if f and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[-1]) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[f[-2]]
        return tot or None
    except (OSError, ValueError, KeyError, IndexError):
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
    g1, g2, _ = wl.pairing_inputs(n, k=k)
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
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from plonky2_bn254_pairing_b200 import native
    from plonky2_bn254_pairing_b200 import workload as wl

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
    lib = native.init([local])

    prog, k, label = WORKLOADS[args.workload]
    n = args.batch
    g1, g2, _ = wl.pairing_inputs(n, k=k, offset=rank * n)
    dev = torch.device("cuda", local)

    def to_dev(a):
        return torch.from_numpy(a.view(np.int64)).to(dev)

    d_g1, d_g2 = to_dev(g1), to_dev(g2)
    d_out = torch.empty((12, 4, n), dtype=torch.int64, device=dev)
    d_f12 = None
    stream = torch.cuda.Stream(device=dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    if args.workload == "final_exp":
        # inputs = GPU Miller outputs of the same indices
        d_f12 = torch.empty((12, 4, n), dtype=torch.int64, device=dev)
        native.check(lib.bnp_run_program_dev(local, sp, b"miller", d_g1.data_ptr(), d_g2.data_ptr(), None, None,
                                             d_f12.data_ptr(), n))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def launch():
        native.check(lib.bnp_run_program_dev(local, sp, prog.encode(), d_g1.data_ptr(), d_g2.data_ptr(),
                                             d_f12.data_ptr() if d_f12 is not None else None, None, d_out.data_ptr(), n))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # roofline denominator, measured live (rank-local; reported from rank 0)
    peak = ctypes.c_double()
    native.check(lib.bnp_imad_peak(local, ctypes.byref(peak)))
    peak32 = ctypes.c_double()
    native.check(lib.bnp_imad32_peak(local, ctypes.byref(peak32)))
    macs = lib.bnp_program_macs(prog.encode())             # algorithmic (Karatsuba Fq2 products), SURVEY 8(d)
    macs_x = lib.bnp_program_macs_executed(prog.encode())  # what the component-split kernel issues

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            flush.zero_()
            launch()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.bnp_launch_count()
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        t0.record(stream)
        for i in range(args.steps):
            flush.zero_()  # L2 flush between timed iterations (inside the timed region, ~0.1% of a step)
            k_ev[i][0].record(stream)
            launch()
            k_ev[i][1].record(stream)
        t1.record(stream)
    barrier()
    launches = lib.bnp_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = t0.elapsed_time(t1)
    kern_ms = [a.elapsed_time(b) for a, b in k_ev]
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * n * k * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the host-buffer C ABI (pinned host memory, copies inside the timed region)
    h_g1 = torch.from_numpy(g1.view(np.int64)).pin_memory()
    h_g2 = torch.from_numpy(g2.view(np.int64)).pin_memory()
    h_out = torch.empty((12, 4, n), dtype=torch.int64).pin_memory()
    h_f12 = d_f12.cpu().pin_memory() if d_f12 is not None else None

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
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_call()  # synchronous: returns after the D2H copy completed
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - w0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n * k * e2e_steps / e2e_s
    in_bytes = (h_f12.numel() * 8) if args.workload == "final_exp" else (h_g1.numel() + h_g2.numel()) * 8
    # quick integrity check of the e2e result against the device-resident result
    assert torch.equal(h_out[:, :, :64], d_out[:, :, :64].cpu()), "e2e and device-resident results differ"

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
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (8x32-bit-limb Fp254 Montgomery, IMAD.WIDE.U32 carry chains)", "data": "synthetic",
            "config": {"workload": label, "program": prog, "batch_per_gpu": n, "pairings_per_element": k,
                       "parallelism": "index-sharded, %d rank(s), no data-path collective" % world,
                       "l2": "256 MiB memset between timed iterations (inside the timed region)",
                       "inputs": "pool of 256 G1 x 256 G2 subgroup points, pair i = (P[i%K], Q[(i/K+7i)%K])"},
            "roofline": {
                "bound": "imad", "achieved": achieved / 1e9, "peak": peak.value / 1e9, "unit": "GMAC/s",
                "frac": achieved / peak.value,
                # bytes: one ncu capture at 2^16 pairings (includes the L2-flush write-back and spill traffic that
                # leaves the L2); not a bound - the algorithmic I/O is 576 B per pairing
                "traffic": ncu_traffic_bytes() if (args.workload == "pairing" and n == BATCH) else None,
                "kernel": "bnp_vm_kernel<64>", "kernel_ms_avg": kavg, "macs_per_element": macs,
                "executed": {"macs_per_element": macs_x, "gmacs": n * macs_x / (kavg * 1e-3) / 1e9,
                             "pipe_frac": n * macs_x / (kavg * 1e-3) / peak.value,
                             "note": "the kernel computes an Fq2 product as a two-term dot product per lane "
                                     "(4 Fp products, not Karatsuba's 3); frac above counts only the algorithmic MACs"},
                "peak_source": "measured live: bnp_imad_peak (IMAD.WIDE.U32[.X] 4-deep carry chains, 32 MACs/thread/iter); "
                               "MEASURED_PEAKS.json has no integer peak",
                "imad32_peak_gops": peak32.value / 1e9,
                "hbm": {"achieved_gbs": alg_bytes / (kavg * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "frac": alg_bytes / (kavg * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback", "note": "informational; not the bound"},
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": h_out.numel() * 8,
                    "steps": e2e_steps, "api": "bnp_%s_batch (host pointers, synchronous)" % (
                        "pairing" if args.workload == "pairing" else "miller_loop" if args.workload == "miller"
                        else "final_exp" if args.workload == "final_exp" else "multi_pairing")},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
