"""Multi-GPU plumbing: one process per GPU (torchrun), work partitioned by pairing index.

Independent pairings need no collective at all - each rank runs its contiguous index range
(SURVEY 8(e)).  The only exchange step in the whole path is the ONE-product case (Groth16-style
aggregation over many pairs): every rank reduces its fused Miller values to a single Fq12 (384 bytes),
the partials are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests), and the product of
the partials goes through one final exponentiation.  Field multiplication is exact and commutative,
so the result does not depend on how the pairs were partitioned.

The compute steps are injected (`ops`), so the same orchestration runs on the GPU library
(`DeviceOps`) and, in tests/test_sharding.py, on the CPU oracle under gloo with world_size 2.
"""
import ctypes

import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous partition of range(n): same rule as libbnp's split_range (csrc/bnp.cu)."""
    cnt = n // world + (1 if rank < n % world else 0)
    off = rank * (n // world) + min(rank, n % world)
    return off, cnt


def all_gather_fq12(partial, group=None):
    """partial: int64 tensor [12, 4, 1] (u64 limbs) on this rank -> [12, 4, world] on every rank."""
    world = dist.get_world_size(group)
    flat = partial.reshape(48).contiguous()
    bufs = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(bufs, flat, group=group)
    return torch.stack(bufs, dim=1).reshape(12, 4, world).contiguous()


def pairing_product_distributed(ops, g1_local, g2_local, variant=0, group=None):
    """final_exp(prod over ALL ranks' pairs of miller(Q_i, P_i)); every rank returns the same [12,4,1].

    g1_local / g2_local: this rank's shard, int64 tensors [2,4,n_r] / [4,4,n_r] (n_r may be 0)."""
    n_local = g1_local.shape[2]
    if n_local > 0:
        partial = ops.product(ops.miller_fused(g1_local, g2_local))
    else:
        partial = ops.one()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        gathered = all_gather_fq12(partial, group)
        total = ops.product(gathered)
    else:
        total = partial
    return ops.final_exp(total, variant)


class DeviceOps:
    """`ops` backed by libbnp.so device-pointer entry points on the current CUDA device."""

    def __init__(self, device_index):
        from . import native

        self.native = native
        self.lib = native.lib()
        self.dev = device_index

    def _stream(self):
        # libbnp reads a NULL stream as "the library's own non-blocking stream", which is ordered neither with
        # torch's default stream nor with the NCCL stream.  torch reports its default stream as handle 0, so name
        # it explicitly: cudaStreamLegacy (0x1) IS that stream, and everything stays ordered with the tensors
        # torch made before the call and with the collective that follows (found on 2 GPUs: the partials were
        # gathered before the tree product had run).
        h = torch.cuda.current_stream(self.dev).cuda_stream
        return ctypes.c_void_p(h if h else 1)

    def _out(self, n):
        return torch.empty((12, 4, n), dtype=torch.int64, device="cuda:%d" % self.dev)

    def miller_fused(self, g1, g2):
        out = self._out(g1.shape[2])
        self.native.check(self.lib.bnp_miller_loop_fused_dev(self.dev, self._stream(), g1.data_ptr(), g2.data_ptr(),
                                                             out.data_ptr(), g1.shape[2]))
        return out

    def product(self, f):
        out = self._out(1)
        buf = f.clone()  # the tree product works in place
        self.native.check(self.lib.bnp_fq12_product_dev(self.dev, self._stream(), buf.data_ptr(), out.data_ptr(),
                                                        f.shape[2]))
        return out

    def final_exp(self, f, variant):
        out = self._out(f.shape[2])
        self.native.check(self.lib.bnp_final_exp_dev(self.dev, self._stream(), f.data_ptr(), out.data_ptr(),
                                                     f.shape[2], variant))
        return out

    def one(self):
        import numpy as np

        from . import api

        return torch.from_numpy(api.pack_soa([[1] + [0] * 11]).view(np.int64)).to("cuda:%d" % self.dev)
