"""Parity tests proper: the CUDA path (through the C ABI, host buffers) against the oracle.
Bit-exact is the bar - all of this is integer arithmetic.  Run on the B200 box: pytest -m gpu."""
import ctypes
import random

import numpy as np
import pytest

import bn254_oracle as O
from conftest import point_pool
from plonky2_bn254_pairing_b200 import api, native

pytestmark = pytest.mark.gpu


def ints(hexes):
    return [int(h, 16) for h in hexes]


@pytest.fixture(scope="module")
def lib(built):
    return native.init([0])


def _dev(a):
    import torch

    return torch.from_numpy(a.view(np.int64)).cuda()


# ----------------------------------------------------------------------------- op level
def test_every_opcode_against_plain_integer_semantics(lib):
    import torch

    import optest_expect as X
    from plonky2_bn254_pairing_b200.microcode.programs import OPTEST_OUTPUTS

    rows = X.edge_rows(random.Random(7))
    n = len(rows)
    d_in = _dev(api.pack_soa(rows))
    d_out = torch.zeros((2 * OPTEST_OUTPUTS, 4, n), dtype=torch.int64, device="cuda")
    native.check(lib.bnp_run_program_dev(0, None, b"optest", None, None, d_in.data_ptr(), None, d_out.data_ptr(), n))
    torch.cuda.synchronize()
    out = api.unpack_soa(d_out.cpu().numpy().view(np.uint64))  # also asserts every limb is canonical
    for e in range(n):
        want = X.expected(rows[e])
        for i, nm in enumerate(X.NAMES):
            assert (out[e][2 * i], out[e][2 * i + 1]) == want[i], (e, nm)


# ----------------------------------------------------------------------------- golden fixtures
def test_survey_kat_on_gpu(lib, golden):
    assert api.miller_loop_native(O.G2_GEN, O.G1_GEN) == ints(golden["survey"]["kat1_miller"])
    assert api.pairing(O.G1_GEN, O.G2_GEN) == ints(golden["survey"]["kat1_pairing"])
    assert api.pairing_batch([O.G1_GEN], [O.G2_GEN], variant=1)[0] == ints(golden["survey"]["kat1_pairing_ark"])
    p5, q6 = O.g1_mul(O.G1_GEN, 5), O.g2_mul(O.G2_GEN, 6)
    assert api.miller_loop_native(q6, p5) == ints(golden["survey"]["kat2_miller"])
    assert api.pairing(p5, q6) == ints(golden["survey"]["kat2_pairing"])


def test_oracle_vectors_on_gpu(lib, golden):
    cases = golden["oracle"]["cases"]
    Ps = [tuple(ints(c["g1"])) for c in cases]
    Qs = [((ints(c["g2"])[0], ints(c["g2"])[1]), (ints(c["g2"])[2], ints(c["g2"])[3])) for c in cases]
    assert api.miller_loop_native_batch(Qs, Ps) == [ints(c["miller"]) for c in cases]
    assert api.final_exp_native_batch([ints(c["miller"]) for c in cases]) == [ints(c["pairing"]) for c in cases]
    assert api.pairing_batch(Ps, Qs) == [ints(c["pairing"]) for c in cases]
    assert api.pairing_batch(Ps, Qs, variant=1) == [ints(c["pairing_ark"]) for c in cases]
    m3 = golden["oracle"]["multi3"]
    pairs = list(zip(Ps[:3], Qs[:3]))
    assert api.multi_miller_loop_native(pairs) == ints(m3["miller"])
    assert api.multi_pairing_batch([pairs])[0] == ints(m3["pairing"])
    r = golden["oracle"]["random_fq12"]
    assert api.final_exp_native(ints(r["x"])) == ints(r["final_exp"])
    for k, v in r["frobenius"].items():
        assert api.frobenius_map_native(ints(r["x"]), int(k)) == ints(v)


# ----------------------------------------------------------------------------- reference's own test relations, on the GPU
def test_multi_miller_equals_product_of_singles(lib):
    """miller_loop_native.rs:336-348"""
    (p0, q0), (p1, q1) = O.seeded_points(31, 2)
    r0, r1 = api.miller_loop_native(q0, p0), api.miller_loop_native(q1, p1)
    assert api.multi_miller_loop_native([(p0, q0), (p1, q1)]) == api.fq12_mul_batch([r0], [r1])[0] == O.fq12_mul(r0, r1)


def test_to_one(lib):
    """final_exp_native.rs:240-264"""
    p0, q0 = O.g1_mul(O.G1_GEN, 5), O.g2_mul(O.G2_GEN, 6)
    p1, q1 = O.g1_mul(O.G1_GEN, 30), O.g2_neg(O.G2_GEN)
    m = api.multi_miller_loop_native([(p0, q0), (p1, q1)])
    m0, m1 = api.miller_loop_native(q0, p0), api.miller_loop_native(q1, p1)
    assert m == O.fq12_mul(m0, m1)
    r_sep = O.fq12_mul(api.final_exp_native(m0), api.final_exp_native(m1))
    assert r_sep == api.final_exp_native(m) == O.FQ12_ONE
    assert api.multi_pairing_batch([[(p0, q0), (p1, q1)]])[0] == O.FQ12_ONE
    assert api.pairing_product([(p0, q0), (p1, q1)]) == O.FQ12_ONE


def test_final_exp_is_the_exact_exponent_on_random_input(lib):
    """final_exp_native.rs:274-285"""
    rnd = random.Random(6)
    xs = [[rnd.randrange(O.P) for _ in range(12)] for _ in range(3)]
    exp = (O.P ** 12 - 1) // O.R_ORDER
    assert api.final_exp_native_batch(xs) == [O.fq12_pow(x, exp) for x in xs]
    got1 = api.final_exp_native_batch(xs, variant=1)
    assert got1 == [O.fq12_pow(O.fq12_pow(x, exp), O.ARK_LAMBDA) for x in xs]


def test_final_exp_witness_values(lib):
    """The circuit-witness entry point: m, m^x, m^(x^2), m^(x^3), final_exp_native per input."""
    pts = O.seeded_points(0xB2540009, 3)
    ins = [O.miller_loop_native(q, p) for p, q in pts]
    got = api.final_exp_witness_batch(ins)
    assert api.final_exp_witness_batch([]) == []
    for a, w in zip(ins, got):
        m = O.easy_part(a)
        mx = O.pow_native(m, [O.BN_X])
        mx2 = O.pow_native(mx, [O.BN_X])
        mx3 = O.pow_native(mx2, [O.BN_X])
        assert w == {"m": m, "mx": mx, "mx2": mx2, "mx3": mx3, "out": O.final_exp_native(a)}
    assert [w["out"] for w in got] == api.final_exp_native_batch(ins)


# ----------------------------------------------------------------------------- batches vs the C oracle (every element, raw limbs)
@pytest.mark.parametrize("n", [1, 15, 16, 17, 31, 33, 65, 1000])   # a warp holds 16 pairings (two lanes each)
def test_ragged_batches_bit_exact(lib, cref, n):
    Ps, Qs = point_pool(min(n, 64))
    idx = np.arange(n)
    g1 = np.ascontiguousarray(api.pack_soa(api.g1_rows(Ps))[:, :, idx % len(Ps)])
    g2 = np.ascontiguousarray(api.pack_soa(api.g2_rows(Qs))[:, :, (idx * 7 + idx // len(Ps)) % len(Qs)])
    assert np.array_equal(api.pairing_soa(g1, g2), cref.pairing(g1, g2))
    m = api.miller_loop_soa(g1, g2)
    assert np.array_equal(m, cref.miller(g1, g2))
    assert np.array_equal(api.final_exp_soa(m), cref.final_exp(m))


@pytest.mark.parametrize("threads", [32, 64, 128, 256])
@pytest.mark.parametrize("phase_mode", [2, 3])
def test_every_launch_configuration_bit_exact(lib, cref, threads, phase_mode):
    """Block size (>= 128 threads: the warps of a block run in lockstep behind a block barrier, idle warps shadow the
    last chunk) and phase splitting (2: never, 3: always) change the schedule, never the bits."""
    n = 300
    Ps, Qs = point_pool(64)
    idx = np.arange(n)
    g1 = np.ascontiguousarray(api.pack_soa(api.g1_rows(Ps))[:, :, idx % len(Ps)])
    g2 = np.ascontiguousarray(api.pack_soa(api.g2_rows(Qs))[:, :, (idx * 5 + 1) % len(Qs)])
    try:
        assert lib.bnp_set_launch_config(threads, phase_mode) == 0
        got = api.pairing_soa(g1, g2)
        got_m = api.miller_loop_soa(g1, g2)
    finally:
        assert lib.bnp_set_launch_config(384, 1) == 0
    assert np.array_equal(got, cref.pairing(g1, g2))
    assert np.array_equal(got_m, cref.miller(g1, g2))


def test_empty_batch_and_bad_arguments(lib):
    assert api.pairing_batch([], []) == []
    z1 = np.zeros((2, 4, 0), dtype=np.uint64)
    z2 = np.zeros((4, 4, 0), dtype=np.uint64)
    zo = np.zeros((12, 4, 0), dtype=np.uint64)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    assert lib.bnp_pairing_batch(vp(z1), vp(z2), vp(zo), 0, 0) == 0
    g1 = np.zeros((2, 4, 1), dtype=np.uint64)
    g2 = np.zeros((4, 4, 1), dtype=np.uint64)
    out = np.zeros((12, 4, 1), dtype=np.uint64)
    assert lib.bnp_pairing_batch(vp(g1), vp(g2), vp(out), 1, 7) == -1  # BNP_EINVAL: unknown variant
    assert lib.bnp_multi_pairing_batch(vp(g1), vp(g2), vp(out), 1, 5, 0) == -5  # BNP_EUNSUPPORTED: k > 4
    assert lib.bnp_pairing_batch(None, vp(g2), vp(out), 1, 0) == -1
    with pytest.raises(ValueError):
        api.pairing_batch([O.G1_GEN], [])


@pytest.mark.parametrize("k", [2, 3, 4])
def test_groth16_shaped_products(lib, cref, k):
    n = 96
    Ps, Qs = point_pool(32)
    rnd = random.Random(k)
    g1 = np.concatenate([api.pack_soa(api.g1_rows([Ps[rnd.randrange(32)] for _ in range(n)])) for _ in range(k)])
    g2 = np.concatenate([api.pack_soa(api.g2_rows([Qs[rnd.randrange(32)] for _ in range(n)])) for _ in range(k)])
    assert np.array_equal(api.miller_loop_soa(g1, g2, k=k), cref.miller(g1, g2, k=k))
    got = api.pairing_soa(g1, g2, k=k)
    assert np.array_equal(got, cref.pairing(g1, g2, k=k))
    # == product of the k single pairings (final_exp_native.rs:258-263)
    acc = None
    for j in range(k):
        single = api.pairing_soa(np.ascontiguousarray(g1[2 * j:2 * j + 2]), np.ascontiguousarray(g2[4 * j:4 * j + 4]))
        acc = single if acc is None else api.fq12_mul_soa(acc, single)
    assert np.array_equal(got, acc)


@pytest.mark.parametrize("k", [2, 4])
def test_ark_variant_products(lib, k):
    """The ark-compatible exponent (variant 1) through the k-way product programs: equal to the product of the
    single variant-1 pairings, and to the reference value raised to 2x(6x^2+3x+1)."""
    n = 5
    Ps, Qs = point_pool(16)
    rnd = random.Random(100 + k)
    g1 = np.concatenate([api.pack_soa(api.g1_rows([Ps[rnd.randrange(16)] for _ in range(n)])) for _ in range(k)])
    g2 = np.concatenate([api.pack_soa(api.g2_rows([Qs[rnd.randrange(16)] for _ in range(n)])) for _ in range(k)])
    got = api.pairing_soa(g1, g2, variant=api.VARIANT_ARK, k=k)
    acc = None
    for j in range(k):
        single = api.pairing_soa(np.ascontiguousarray(g1[2 * j:2 * j + 2]), np.ascontiguousarray(g2[4 * j:4 * j + 4]),
                                 variant=api.VARIANT_ARK)
        acc = single if acc is None else api.fq12_mul_soa(acc, single)
    assert np.array_equal(got, acc)
    ref = api.unpack_soa(api.pairing_soa(g1, g2, k=k))
    assert api.unpack_soa(got) == [O.fq12_pow(x, O.ARK_LAMBDA) for x in ref]


def test_frobenius_all_powers(lib, cref):
    rnd = random.Random(12)
    x = api.pack_soa([[rnd.randrange(O.P) for _ in range(12)] for _ in range(40)])
    for power in list(range(12)) + [12, 25, 10 ** 9 + 7]:
        assert np.array_equal(api.frobenius_soa(x, power), cref.frobenius(x, power)), power


def test_global_product_any_count(lib):
    Ps, Qs = point_pool(37)
    pairs = list(zip(Ps, Qs))
    want = O.final_exp_native(O.multi_miller_loop_native(pairs[:5]))
    assert api.pairing_product(pairs[:5]) == want
    ms = api.miller_loop_native_batch(Qs, Ps)
    acc = ms[0]
    for m in ms[1:]:
        acc = O.fq12_mul(acc, m)
    assert api.pairing_product(pairs) == O.final_exp_native(acc)


def test_device_ops_on_torch_default_stream(lib):
    """sharding.DeviceOps enqueues on torch's CURRENT stream.  On the default stream that used to mean handle 0 =
    the library's own non-blocking stream: unordered with torch's copies (and with NCCL), found on 2 GPUs."""
    import torch

    from plonky2_bn254_pairing_b200 import sharding

    n = 257
    Ps, Qs = point_pool(32)
    idx = np.arange(n)
    g1 = np.ascontiguousarray(api.pack_soa(api.g1_rows(Ps))[:, :, idx % 32])
    g2 = np.ascontiguousarray(api.pack_soa(api.g2_rows(Qs))[:, :, (idx * 3 + idx // 32) % 32])
    want = api.pairing_product_soa(g1, g2)
    for _ in range(3):  # fresh tensors each time: the host-to-device copies are still in flight when the kernels are enqueued
        t1 = torch.from_numpy(g1.view(np.int64)).pin_memory().to("cuda:0", non_blocking=True)
        t2 = torch.from_numpy(g2.view(np.int64)).pin_memory().to("cuda:0", non_blocking=True)
        out = sharding.pairing_product_distributed(sharding.DeviceOps(0), t1, t2)
        assert np.array_equal(out.cpu().numpy().view(np.uint64).reshape(12, 4, 1), want)


def test_all_devices_in_one_process(lib):
    """bnp_init on every visible device: host batches are split by index range inside the library and the global
    product reduces per-device partials over peer copies (skipped on a single-GPU box; tools/gpu_multi_check.py
    runs the same check plus the torchrun/NCCL variant)."""
    import torch

    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs at least two GPUs")
    try:
        native.init(list(range(ndev)))
        n = 500 + ndev
        Ps, Qs = point_pool(32)
        idx = np.arange(n)
        g1 = np.ascontiguousarray(api.pack_soa(api.g1_rows(Ps))[:, :, idx % 32])
        g2 = np.ascontiguousarray(api.pack_soa(api.g2_rows(Qs))[:, :, (idx * 3 + idx // 32) % 32])
        got = api.pairing_soa(g1, g2)
        prod = api.pairing_product_soa(g1, g2)
        assert lib.bnp_gather_transport() in (b"nccl", b"peer-copy")
        ks = api._scalar_rows(O.seeded_scalars(0xB2540F06, n))
        kp_multi, kq_multi = api.scalar_mul_soa(1, g1, ks), api.scalar_mul_soa(2, g2, ks)
        ok_multi = api.validate_soa(g1, g2)
    finally:
        lib.bnp_shutdown()
        native.init([0])
    assert np.array_equal(prod, api.pairing_product_soa(g1, g2))
    # the index-range split of the other host-pointer entries: the same bits as one device
    kp_one, kq_one = api.scalar_mul_soa(1, g1, ks), api.scalar_mul_soa(2, g2, ks)
    assert np.array_equal(kp_multi[0], kp_one[0]) and np.array_equal(kp_multi[1], kp_one[1])
    assert np.array_equal(kq_multi[0], kq_one[0]) and np.array_equal(kq_multi[1], kq_one[1])
    assert ok_multi.all() and np.array_equal(got, api.pairing_soa(g1, g2))
    acc = np.ascontiguousarray(got[:, :, :1])
    for i in range(1, 64):
        acc = api.fq12_mul_soa(acc, np.ascontiguousarray(got[:, :, i:i + 1]))
    assert np.array_equal(acc, api.pairing_product_soa(np.ascontiguousarray(g1[:, :, :64]), np.ascontiguousarray(g2[:, :, :64])))


# ----------------------------------------------------------------------------- 2^16 pairings: whole-batch properties
# (every one of the 65 536 results is compared with the oracle in tests/test_gpu_configs.py::test_config2_*)
def test_full_size_batch_bit_exact_and_properties(lib, cref):
    n = 1 << 16
    K = 256
    Ps, Qs = point_pool(K)
    idx = np.arange(n)
    i1, i2 = idx % K, (idx // K + 7 * idx) % K
    g1 = np.ascontiguousarray(api.pack_soa(api.g1_rows(Ps))[:, :, i1])
    g2 = np.ascontiguousarray(api.pack_soa(api.g2_rows(Qs))[:, :, i2])
    got = api.pairing_soa(g1, g2)
    # all 65 536 index pairs are distinct combinations of the pool; a sample of 4 096 against the C oracle here
    pairs, inv = np.unique(np.stack([i1, i2]), axis=1, return_inverse=True)
    assert pairs.shape[1] == n
    sample = np.random.RandomState(1).choice(n, 4096, replace=False)
    want = cref.pairing(np.ascontiguousarray(g1[:, :, sample]), np.ascontiguousarray(g2[:, :, sample]))
    assert np.array_equal(got[:, :, sample], want)
    # size-independent properties over the WHOLE batch:
    # (1) canonical residues everywhere
    top = got[:, 3, :]
    assert (top <= np.uint64(0x30644e72e131a029)).all()
    # (2) e(P,Q) * e(P,-Q) == 1 for every element: negate Q.y on the host, multiply on the GPU
    g2n = g2.copy()
    ys = api.unpack_soa(np.ascontiguousarray(api.pack_soa(api.g2_rows(Qs))[2:4]))
    neg = api.pack_soa([[(-y[0]) % O.P, (-y[1]) % O.P] for y in ys])
    g2n[2:4] = neg[:, :, i2]
    prod = api.fq12_mul_soa(got, api.pairing_soa(g1, g2n))
    one = api.pack_soa([[1] + [0] * 11])
    assert np.array_equal(prod, np.broadcast_to(one, prod.shape))
