"""The global product over 2, 3, 4 and 8 devices of ONE process against the single-device value, three calls each, with the
partials gathered by NCCL and by peer copies (BNP_NO_NCCL=1).  Found the race of the round-1 peer-copy gather (a plain
cudaMemcpyPeer on the legacy stream next to a non-blocking stream: wrong products on 8 devices, two calls in three)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from plonky2_bn254_pairing_b200 import api, native
from conftest import point_pool
Ps, Qs = point_pool(32)
n = 257
idx = np.arange(n)
g1 = np.ascontiguousarray(api.pack_soa(api.g1_rows(Ps))[:, :, idx % 32])
g2 = np.ascontiguousarray(api.pack_soa(api.g2_rows(Qs))[:, :, (idx * 5 + idx // 32) % 32])
native.init([0])
want = api.pairing_product_soa(g1, g2)
for nd in (2, 3, 4, 8):
    if nd > torch.cuda.device_count(): break
    for no_nccl in ("", "1"):
        native.lib().bnp_shutdown()
        if no_nccl: os.environ["BNP_NO_NCCL"] = "1"
        else: os.environ.pop("BNP_NO_NCCL", None)
        native.init(list(range(nd)))
        res = [np.array_equal(api.pairing_product_soa(g1, g2), want) for _ in range(3)]
        print(nd, "devices", native.lib().bnp_gather_transport().decode(), res, flush=True)
