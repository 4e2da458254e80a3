"""Multi-GPU checks that need real devices (run with gpurun --gpus 2):
  1. single process, bnp_init on ALL devices: host-pointer batches are split by contiguous index range inside the
     library, and bnp_pairing_product reduces per-device partials over peer copies - both bit-compared with the
     C oracle / the product of single pairings;
  2. (under torchrun) one process per GPU: sharding.pairing_product_distributed with the NCCL all-gather of the
     384-byte partials, every rank bit-compares the result with the single-process value."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bn254_oracle as O  # noqa: E402
from plonky2_bn254_pairing_b200 import api, native, sharding  # noqa: E402


def inputs(n, K=32):
    pts = O.seeded_points(0xB2540042, K)
    Ps, Qs = [p for p, _ in pts], [q for _, q in pts]
    idx = np.arange(n)
    g1 = np.ascontiguousarray(api.pack_soa(api.g1_rows(Ps))[:, :, idx % K])
    g2 = np.ascontiguousarray(api.pack_soa(api.g2_rows(Qs))[:, :, (idx * 5 + idx // K) % K])
    return g1, g2


def single_process():
    ndev = torch.cuda.device_count()
    native.init(list(range(ndev)))
    n = 1000 + ndev  # ragged split
    g1, g2 = inputs(n)
    got = api.pairing_soa(g1, g2)
    import ctypes

    cl = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libbn254_ref.so"))
    cl.bn254_ref_pairing_batch.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_size_t] + [ctypes.c_int] * 3
    cl.bn254_ref_pairing_batch.restype = None
    want = np.zeros((12, 4, n), dtype=np.uint64)
    cl.bn254_ref_pairing_batch(g1.ctypes.data_as(ctypes.c_void_p), g2.ctypes.data_as(ctypes.c_void_p),
                               want.ctypes.data_as(ctypes.c_void_p), n, 1, 0, 0)
    assert np.array_equal(got, want), "multi-device batch differs from the oracle"
    prod = api.pairing_product_soa(np.ascontiguousarray(g1[:, :, :257]), np.ascontiguousarray(g2[:, :, :257]))
    singles = got[:, :, :257]
    acc = np.ascontiguousarray(singles[:, :, :1])
    for i in range(1, 257):
        acc = api.fq12_mul_soa(acc, np.ascontiguousarray(singles[:, :, i:i + 1]))
    assert np.array_equal(prod, acc), "global product over %d devices differs from the product of single pairings" % ndev
    transport = native.lib().bnp_gather_transport().decode()
    print("single-process, %d device(s): sharded batch and global product bit-exact (partials gathered by %s)" % (ndev, transport))
    if ndev > 1 and transport == "nccl":
        # the same product with NCCL switched off must give the same bits
        native.lib().bnp_shutdown()
        os.environ["BNP_NO_NCCL"] = "1"
        native.init(list(range(ndev)))
        assert native.lib().bnp_gather_transport() == b"peer-copy"
        again = api.pairing_product_soa(np.ascontiguousarray(g1[:, :, :257]), np.ascontiguousarray(g2[:, :, :257]))
        assert np.array_equal(prod, again)
        del os.environ["BNP_NO_NCCL"]
        print("  peer-copy gather: the same bits")


def distributed():
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    native.init([local])
    n = 301
    g1, g2 = inputs(n)
    off, cnt = sharding.shard_range(n, rank, world)
    dev = torch.device("cuda", local)
    t1 = torch.from_numpy(np.ascontiguousarray(g1[:, :, off:off + cnt]).view(np.int64)).to(dev)
    t2 = torch.from_numpy(np.ascontiguousarray(g2[:, :, off:off + cnt]).view(np.int64)).to(dev)
    out = sharding.pairing_product_distributed(sharding.DeviceOps(local), t1, t2)
    got = out.cpu().numpy().view(np.uint64)
    want = api.pairing_product_soa(g1, g2)  # this rank's own GPU, all pairs
    assert np.array_equal(got.reshape(12, 4, 1), want), "rank %d: distributed product differs" % rank
    dist.barrier()
    if rank == 0:
        print("torchrun, %d ranks: NCCL-gathered global product bit-exact on every rank" % world)
    # timing of the one-product shape of BASELINE config 4: 2^20 pairs in total, fused Miller loops, per-GPU tree
    # product, 384-byte all-gather, one final exponentiation (device-resident inputs, CUDA events, max over ranks)
    if os.environ.get("BNP_TIME_PRODUCT", "1") != "0":
        from plonky2_bn254_pairing_b200 import workload as wl

        n_loc = (1 << 20) // world
        h1, h2, _ = wl.pairing_inputs(n_loc, k=1, offset=rank * n_loc)
        b1 = torch.from_numpy(h1.view(np.int64)).to(dev)
        b2 = torch.from_numpy(h2.view(np.int64)).to(dev)
        ops = sharding.DeviceOps(local)
        for _ in range(2):
            sharding.pairing_product_distributed(ops, b1, b2)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            res = sharding.pairing_product_distributed(ops, b1, b2)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = float(t.item())
            print("global product of 2^20 pairs on %d GPU(s): %.2f ms per product = %.2f M pairs/s (Miller fused + tree product + "
                  "384 B all-gather + one final exponentiation)" % (world, ms, (1 << 20) / ms / 1e3))
        del res
    dist.destroy_process_group()


if __name__ == "__main__":
    distributed() if "RANK" in os.environ else single_process()
