"""The C-ABI boundary: header, exported symbols, host-side marshalling, and the no-fallback rule."""
import ctypes
import os
import re

import numpy as np
import pytest

import bn254_oracle as O
from plonky2_bn254_pairing_b200 import api, native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "bnp.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bnp_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(native.SIGNATURES)


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(native.LIB_PATH)
    for name in header_symbols():
        assert getattr(lib, name) is not None
    native.load()


def test_no_cpu_fallback_without_device(built):
    """Without a CUDA device the product path must fail loudly, never compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = native.load()
    assert lib.bnp_init(None, 0) == -2  # BNP_ENODEV
    g1 = np.zeros((2, 4, 1), dtype=np.uint64)
    g2 = np.zeros((4, 4, 1), dtype=np.uint64)
    out = np.zeros((12, 4, 1), dtype=np.uint64)
    rc = lib.bnp_pairing_batch(g1.ctypes.data_as(ctypes.c_void_p), g2.ctypes.data_as(ctypes.c_void_p),
                               out.ctypes.data_as(ctypes.c_void_p), 1, 0)
    assert rc == -2
    assert not out.any()
    with pytest.raises(native.BnpError):
        api.pairing(O.G1_GEN, O.G2_GEN)


def test_program_work_is_exported(built):
    lib = native.load()
    assert 1.9e6 < lib.bnp_program_macs(b"pairing_v0") <= 2.04e6
    assert lib.bnp_program_macs(b"no_such_program") == 0
    # one thread runs a whole Karatsuba Fq2 product: the kernel issues exactly the algorithmic count
    assert lib.bnp_program_macs_executed(b"pairing_v0") == lib.bnp_program_macs(b"pairing_v0")
    assert lib.bnp_strerror(-2).decode().startswith("no CUDA device")


def test_soa_layout_is_ark_montgomery_limbs():
    """buf[(k*4 + j)*n + e] = limb j of value k of element e, Montgomery R = 2^256 (include/bnp.h)."""
    rows = [[1, 2], [O.P - 1, 12345678901234567890123]]
    a = api.pack_soa(rows)
    assert a.shape == (2, 4, 2) and a.dtype == np.uint64
    for e, r in enumerate(rows):
        for k, v in enumerate(r):
            assert [int(a[k, j, e]) for j in range(4)] == O.to_mont_limbs(v)
    assert api.unpack_soa(a) == rows
    bad = a.copy()
    bad[0, :, 0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    with pytest.raises(native.BnpError):
        api.unpack_soa(bad)


def test_myfq12_to_ark_order():
    x = list(range(12))
    assert api.myfq12_to_ark(x) == O.myfq12_to_ark(x)


def test_scalar_limbs_and_window_digits():
    """host side of bnp_scalar_mul_batch: scalars travel as four plain little-endian 64-bit limbs, SoA; and the signed
    4-bit window recoding csrc/scalar.cuh performs per thread (restated here) represents every 256-bit scalar with
    digits in [-8, 8) plus a carry digit, so the table P .. 8P suffices"""
    ks = [0, 1, 7, 8, 15, 16, O.R_ORDER - 1, O.R_ORDER, (1 << 256) - 1, 1 << 255, 0x8888888888888888, 0x7777777777777777 << 64]
    rows = api._scalar_rows(ks)
    assert rows.shape == (4, len(ks)) and rows.dtype == np.uint64
    for i, k in enumerate(ks):
        assert sum(int(rows[j, i]) << (64 * j) for j in range(4)) == k
        limbs = [int(rows[j, i]) for j in range(4)]
        d, carry = [], 0
        for n in range(64):
            v = ((limbs[n >> 4] >> ((n & 15) * 4)) & 15) + carry
            carry = 1 if v >= 8 else 0
            d.append(v - (carry << 4))
        d.append(carry)
        assert all(-8 <= x <= 7 for x in d[:64]) and d[64] in (0, 1)
        assert sum(x << (4 * n) for n, x in enumerate(d)) == k
