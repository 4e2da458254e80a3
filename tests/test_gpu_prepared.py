"""Prepared G2 points: the line coefficients program g2_prepare writes, and the pairings that read them,
against the programs that do the point arithmetic themselves - bit for bit - and against the oracle."""
import numpy as np
import pytest

import bn254_oracle as O
from conftest import point_pool
from plonky2_bn254_pairing_b200 import api, native

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib(built):
    return native.init([0])


def _shared(coeffs, cols):
    """[PREP_FQ][4][m] -> the columns `cols` as ONE shared element [len(cols) * PREP_FQ][4][1]"""
    return np.ascontiguousarray(coeffs[:, :, cols].transpose(2, 0, 1).reshape(len(cols) * api.PREP_FQ, 4, 1))


def test_single_prepared_pairing(lib):
    Ps, Qs = point_pool(48)
    g1 = api.pack_soa(api.g1_rows(Ps))
    g2 = api.pack_soa(api.g2_rows(Qs[:3]))
    coeffs = api.g2_prepare_soa(g2)
    assert coeffs.shape == (api.PREP_FQ, 4, 3)
    for j in range(3):
        got = api.pairing_prepared_soa(g1, None, _shared(coeffs, [j]), 0, 1)
        g2j = np.ascontiguousarray(np.repeat(g2[:, :, j:j + 1], len(Ps), axis=2))
        assert np.array_equal(got, api.pairing_soa(g1, g2j))
    assert api.unpack_soa(got)[5] == O.pairing(Ps[5], Qs[2])


@pytest.mark.parametrize("variant", [0, 1])
def test_groth16_shape_with_a_prepared_verifying_key(lib, cref, variant):
    """1 live + 3 prepared pairs per proof == the 4-way product with the key's points spelled out for every proof"""
    n = 1000
    Ps, Qs = point_pool(64)
    idx = np.arange(n)
    g1 = np.concatenate([api.pack_soa(api.g1_rows(Ps))[:, :, (idx * (j + 1) + j) % len(Ps)] for j in range(4)], axis=0)
    g2v = np.ascontiguousarray(api.pack_soa(api.g2_rows(Qs))[:, :, (idx * 7 + 3) % len(Qs)])
    vk = api.pack_soa(api.g2_rows(Qs[10:13]))
    prepared = _shared(api.g2_prepare_soa(vk), [0, 1, 2])
    got = api.pairing_prepared_soa(np.ascontiguousarray(g1), g2v, prepared, 1, 3, variant=variant)
    g2_all = np.concatenate([g2v] + [np.repeat(vk[:, :, j:j + 1], n, axis=2) for j in range(3)], axis=0)
    want = api.pairing_soa(np.ascontiguousarray(g1), np.ascontiguousarray(g2_all), variant=variant, k=4)
    assert np.array_equal(got, want)
    if variant == 0:
        assert np.array_equal(got[:, :, :64], cref.pairing(np.ascontiguousarray(g1[:, :, :64]),
                                                            np.ascontiguousarray(g2_all[:, :, :64]), k=4))


def test_unknown_shape_is_refused(lib):
    Ps, Qs = point_pool(4)
    g1 = api.pack_soa([r + r + r + r + r for r in api.g1_rows(Ps)])
    prepared = np.zeros((5 * api.PREP_FQ, 4, 1), dtype=np.uint64)
    with pytest.raises(native.BnpError):
        api.pairing_prepared_soa(g1, None, prepared, 0, 5)
