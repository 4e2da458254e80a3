import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a host without a GPU skips the gpu-marked tests instead of erroring in their
    fixtures (the product path has no CPU fallback: bnp_init fails with BNP_ENODEV there)."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    d = os.path.join(ROOT, "tests", "golden")
    return {
        "survey": json.load(open(os.path.join(d, "survey_kat.json"))),
        "oracle": json.load(open(os.path.join(d, "oracle_vectors.json"))),
    }


@pytest.fixture(scope="session")
def built():
    """Make sure the native artefacts exist (compiles on first use; nvcc/gcc need no GPU)."""
    import __graft_entry__ as g

    g.build_lib()
    g.build_oracle()
    return g


@pytest.fixture(scope="session")
def cref(built):
    """ctypes handle of the C oracle (oracle/_build/libbn254_ref.so) with numpy helpers."""
    import ctypes

    import numpy as np

    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libbn254_ref.so"))
    vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.bn254_ref_miller_batch.argtypes = [vp, vp, vp, sz, ci, ci, ci]
    lib.bn254_ref_final_exp_batch.argtypes = [vp, vp, sz, ci, ci]
    lib.bn254_ref_pairing_batch.argtypes = [vp, vp, vp, sz, ci, ci, ci]
    lib.bn254_ref_frobenius_batch.argtypes = [vp, vp, sz, sz, ci, ci]
    lib.bn254_ref_fq12_mul_batch.argtypes = [vp, vp, vp, sz, ci]
    lib.bn254_ref_pow_batch.argtypes = [vp, vp, sz, ctypes.c_uint64, ci]
    for f in ("miller_batch", "final_exp_batch", "pairing_batch", "frobenius_batch", "fq12_mul_batch", "pow_batch"):
        getattr(lib, "bn254_ref_" + f).restype = None

    class C:
        pass

    c = C()
    c.lib = lib

    def ptr(a):
        assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(vp)

    def miller(g1, g2, k=1, faithful=0, threads=0):
        n = g1.shape[2]
        out = np.zeros((12, 4, n), dtype=np.uint64)
        lib.bn254_ref_miller_batch(ptr(g1), ptr(g2), ptr(out), n, k, faithful, threads)
        return out

    def final_exp(f, faithful=0, threads=0):
        n = f.shape[2]
        out = np.zeros((12, 4, n), dtype=np.uint64)
        lib.bn254_ref_final_exp_batch(ptr(f), ptr(out), n, faithful, threads)
        return out

    def pairing(g1, g2, k=1, faithful=0, threads=0):
        n = g1.shape[2]
        out = np.zeros((12, 4, n), dtype=np.uint64)
        lib.bn254_ref_pairing_batch(ptr(g1), ptr(g2), ptr(out), n, k, faithful, threads)
        return out

    def frobenius(f, power, faithful=0, threads=0):
        n = f.shape[2]
        out = np.zeros((12, 4, n), dtype=np.uint64)
        lib.bn254_ref_frobenius_batch(ptr(f), ptr(out), n, power, faithful, threads)
        return out

    def fq12_mul(a, b, threads=0):
        n = a.shape[2]
        out = np.zeros((12, 4, n), dtype=np.uint64)
        lib.bn254_ref_fq12_mul_batch(ptr(a), ptr(b), ptr(out), n, threads)
        return out

    def pow_u64(f, e, threads=0):
        n = f.shape[2]
        out = np.zeros((12, 4, n), dtype=np.uint64)
        lib.bn254_ref_pow_batch(ptr(f), ptr(out), n, e, threads)
        return out

    c.miller, c.final_exp, c.pairing, c.frobenius, c.fq12_mul, c.pow_u64 = miller, final_exp, pairing, frobenius, fq12_mul, pow_u64
    return c


def point_pool(n, seed=0xB2540003, step_seed=0xB2540004):
    """n valid (P, Q) pairs by an additive walk from seeded subgroup points (cheap, all distinct)."""
    import bn254_oracle as O

    p0, q0 = O.seeded_points(seed, 1)[0]
    dp, dq = O.seeded_points(step_seed, 1)[0]
    Ps, Qs = [p0], [q0]
    for _ in range(n - 1):
        Ps.append(O.g1_add(Ps[-1], dp))
        Qs.append(O.g2_add(Qs[-1], dq))
    return Ps, Qs
