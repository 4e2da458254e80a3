"""Host-side multi-GPU logic on CPU: index partitioning and the gather-and-combine step of the global
pairing product, run under torch.distributed `gloo` with world_size 2 (the compute steps are the oracle)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bn254_oracle as O
from plonky2_bn254_pairing_b200 import api, sharding


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            parts = [sharding.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0
            assert sum(c for _, c in parts) == n
            for (o0, c0), (o1, _) in zip(parts, parts[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


class OracleOps:
    """`ops` for pairing_product_distributed backed by the Python oracle (CPU tensors)."""

    @staticmethod
    def _to_rows(t):
        return api.unpack_soa(t.numpy().view(np.uint64))

    @staticmethod
    def _from_rows(rows):
        return torch.from_numpy(api.pack_soa(rows).view(np.int64).copy())

    def miller_fused(self, g1, g2):
        ps = self._to_rows(g1)
        qs = self._to_rows(g2)
        # any representative of the Miller value modulo proper-subfield factors will do; use the exact one
        return self._from_rows([O.miller_loop_native(((q[0], q[1]), (q[2], q[3])), (p[0], p[1])) for p, q in zip(ps, qs)])

    def product(self, f):
        rows = self._to_rows(f)
        acc = rows[0]
        for r in rows[1:]:
            acc = O.fq12_mul(acc, r)
        return self._from_rows([acc])

    def final_exp(self, f, variant):
        fn = O.final_exp_native if variant == 0 else O.final_exp_ark
        return self._from_rows([fn(r) for r in self._to_rows(f)])

    def one(self):
        return self._from_rows([[1] + [0] * 11])


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pts = O.seeded_points(0xB2540020, n)
    off, cnt = sharding.shard_range(n, rank, world)
    mine = pts[off:off + cnt]
    if cnt:
        g1 = OracleOps._from_rows(api.g1_rows([p for p, _ in mine]))
        g2 = OracleOps._from_rows(api.g2_rows([qq for _, qq in mine]))
    else:
        g1 = torch.zeros((2, 4, 0), dtype=torch.int64)
        g2 = torch.zeros((4, 4, 0), dtype=torch.int64)
    res = sharding.pairing_product_distributed(OracleOps(), g1, g2, variant=0)
    q.put((rank, OracleOps._to_rows(res)[0]))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_world(n, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_global_product_world2_matches_single_process():
    n = 3  # ragged: rank 0 gets 2 pairs, rank 1 gets 1
    out = _run_world(n)
    pts = O.seeded_points(0xB2540020, n)
    want = O.final_exp_native(O.multi_miller_loop_native(pts))
    assert out[0] == want and out[1] == want


def test_global_product_with_an_empty_rank():
    out = _run_world(1)
    p, qq = O.seeded_points(0xB2540020, 1)[0]
    assert out[0] == out[1] == O.pairing(p, qq)
