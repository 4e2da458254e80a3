"""The sequencer programs, executed by the big-integer interpreter (the exact instruction words the GPU
runs, including spills/fills), against the oracle - for several shared-memory slot budgets."""
import random

import pytest

import bn254_oracle as O
from plonky2_bn254_pairing_b200.microcode import alloc, gen, interp, isa, programs
from plonky2_bn254_pairing_b200.microcode.builder import ConstPool, naf_digits


def run(name, arrays, n_slots):
    pool = ConstPool()
    b = programs.build_program(name, pool)
    from plonky2_bn254_pairing_b200.microcode import fuse
    al = alloc.allocate(fuse.fuse(b.ops, max_srcs=min(fuse.MAX_SRCS, n_slots - 2)), n_slots)   # gen.build_all's rule
    for ins in isa.parse(al.words):  # every operand field within the slot budget
        for sl in ins.slots_read() + ins.slots_written():
            assert sl < n_slots, ins.op
        if ins.op == "SPILL":
            assert ins.imm < max(al.n_scratch, 1)
        elif ins.op == "FILL":
            assert ins.imm < al.n_scratch
        assert len(ins.ent0) <= isa.LIN_MAX_ENT and len(ins.ent1) <= isa.LIN_MAX_ENT
        if ins.op in isa.PRODUCT_OPS and ins.has_post():
            # r' is parked in d (or d2) before the post LIN runs: neither may alias a slot the LIN still reads
            ent_slots = {t[0] for t in ins.ent0 + ins.ent1}
            assert ins.d2 not in ent_slots - {ins.r_slot()}
            if ins.store_r:
                assert ins.d != ins.d2 and ins.d not in ent_slots - {ins.d}
    arrays = dict(arrays)
    arrays[isa.ARR_OUT] = {}
    interp.run(al.words, pool.values, arrays, al.n_slots, al.n_scratch)
    return [arrays[isa.ARR_OUT][i] for i in range(len(arrays[isa.ARR_OUT]))], al


PTS = O.seeded_points(0xB2540001, 4)


def g1g2(pairs):
    g1, g2 = [], []
    for p, q in pairs:
        g1 += [p[0], p[1]]
        g2 += [q[0][0], q[0][1], q[1][0], q[1][1]]
    return {isa.ARR_G1: g1, isa.ARR_G2: g2}


@pytest.mark.parametrize("n_slots", [10, 14, 16, 24])
def test_miller_exact(n_slots):
    p, q = PTS[0]
    got, _ = run("miller", g1g2([PTS[0]]), n_slots)
    assert got == O.miller_loop_native(q, p)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n_slots", [10, 16])
def test_pairing_and_final_exp(variant, n_slots):
    p, q = PTS[1]
    m = O.miller_loop_native(q, p)
    want = O.final_exp_native(m) if variant == 0 else O.final_exp_ark(m)
    got, _ = run("pairing_v%d" % variant, g1g2([PTS[1]]), n_slots)
    assert got == want
    got, _ = run("final_exp_v%d" % variant, {isa.ARR_F12: m}, n_slots)
    assert got == want


@pytest.mark.parametrize("n_slots", [13, 16])
def test_final_exp_witness_program(n_slots):
    """the values final_exp_target.rs takes from the CPU - easy part, its three BN_X powers (the
    outputs of the Fq12ExpU64 starks, final_exp_target.rs:89-117) and the final result - from one program."""
    p, q = PTS[2]
    a = O.miller_loop_native(q, p)
    got, _ = run("final_exp_witness", {isa.ARR_F12: a}, n_slots)
    assert len(got) == programs.WITNESS_FQ
    m = O.easy_part(a)
    mx = O.pow_native(m, [O.BN_X])
    mx2 = O.pow_native(mx, [O.BN_X])
    mx3 = O.pow_native(mx2, [O.BN_X])
    want = {"m": m, "mx": mx, "mx2": mx2, "mx3": mx3, "out": O.final_exp_native(a)}
    for name, off in programs.WITNESS_LAYOUT.items():
        assert got[off:off + 12] == want[name], name


def test_final_exp_on_non_miller_input():
    """final_exp_native.rs:266-286 feeds a uniformly random Fq12."""
    rnd = random.Random(4)
    x = [rnd.randrange(O.P) for _ in range(12)]
    got, _ = run("final_exp_v0", {isa.ARR_F12: x}, 14)
    assert got == O.final_exp_native(x) == O.fq12_pow(x, (O.P ** 12 - 1) // O.R_ORDER)


@pytest.mark.parametrize("k", [2, 3, 4])
def test_multi_miller_and_product(k):
    got, _ = run("miller_x%d" % k, g1g2(PTS[:k]), 16)
    want = O.multi_miller_loop_native(PTS[:k])
    assert got == want
    got, _ = run("pairing_x%d_v0" % k, g1g2(PTS[:k]), 16)
    assert got == O.final_exp_native(want)


def test_fused_miller_differs_only_by_a_subfield_factor():
    p, q = PTS[2]
    got, _ = run("miller_fused", g1g2([PTS[2]]), 16)
    assert got != O.miller_loop_native(q, p)
    assert O.final_exp_native(got) == O.pairing(p, q)


@pytest.mark.parametrize("power", [0, 1, 2, 3, 6, 11])
def test_frobenius(power):
    rnd = random.Random(power)
    x = [rnd.randrange(O.P) for _ in range(12)]
    got, _ = run("frobenius_%d" % power, {isa.ARR_F12: x}, 16)
    assert got == O.frobenius_map_native(x, power)


def test_fq12_mul_program():
    rnd = random.Random(8)
    x = [rnd.randrange(O.P) for _ in range(12)]
    y = [rnd.randrange(O.P) for _ in range(12)]
    got, _ = run("fq12_mul", {isa.ARR_F12: x, isa.ARR_AUX: y}, 16)
    assert got == O.fq12_mul(x, y)


def test_work_counts_not_above_survey_canonical():
    """the shipped schedules must not do more algorithmic work than the canonical table."""
    pool, progs = gen.build_all(16, names={"pairing_v0", "miller", "miller_fused", "final_exp_v0", "pairing_x4_v0"})
    macs = {name: w["macs"] for name, _, w in progs}
    assert macs["miller_fused"] <= 1.000e6
    assert macs["miller"] <= 1.10e6
    assert macs["final_exp_v0"] <= 1.05e6
    assert macs["pairing_v0"] <= 2.04e6
    assert macs["pairing_x4_v0"] <= 4.28e6


def test_encode_decode_roundtrip_and_naf():
    w = isa.encode("MUL", d=5, a=255, b=77, c=3, e=9, imm=isa.MUL_B | isa.MUL_ENEG)
    assert isa.decode(w) == ("MUL", 5, 255, 77, 3, 9, isa.MUL_B | isa.MUL_ENEG)
    assert isa.decode_entry(isa.encode_entry(13, 1, 27, True)) == (13, 1, 27, True)
    assert isa.decode_ext(isa.encode_ext(7, [(3, True), (250, False)], True, 11)) == (7, [(3, True), (250, False)], True, 11)
    assert naf_digits(O.BN_X) == O.get_naf([O.BN_X])[:len(naf_digits(O.BN_X))]
    assert sum(d << i for i, d in enumerate(programs.SIX_U_PLUS_2_NAF)) == 6 * O.BN_X + 2
    assert programs.SIX_U_PLUS_2_NAF == O.SIX_U_PLUS_2_NAF


@pytest.mark.parametrize("mode", [{"enable": False}, {"attach": False}, {}])
def test_optest_program_matches_expectations(mode):
    """The op-level GPU test's expected values, checked here against the interpreter (incl. edge values),
    for all lowerings of the linear operations (elementary opcodes / LIN trees / product epilogues)."""
    import optest_expect as X

    rows = X.edge_rows(random.Random(7), n_random=24)
    for r in rows:
        pool = ConstPool()
        b = programs.build_program("optest", pool)
        from plonky2_bn254_pairing_b200.microcode import fuse
        al = alloc.allocate(fuse.fuse(b.ops, **mode), 14)
        arrays = {isa.ARR_F12: r, isa.ARR_OUT: {}}
        interp.run(al.words, pool.values, arrays, al.n_slots, al.n_scratch)
        want = X.expected(r)
        for i in range(programs.OPTEST_OUTPUTS):
            assert (arrays[isa.ARR_OUT][2 * i], arrays[isa.ARR_OUT][2 * i + 1]) == want[i], X.NAMES[i]


def test_fusion_preserves_results_and_cuts_slot_moves():
    from plonky2_bn254_pairing_b200.microcode import fuse

    p, q = PTS[3]
    pool = ConstPool()
    b = programs.build_program("pairing_v0", pool)
    res = {}
    for en in (False, True):
        al = alloc.allocate(fuse.fuse(b.ops, enable=en), 14)
        arrays = g1g2([PTS[3]])
        arrays[isa.ARR_OUT] = {}
        interp.run(al.words, pool.values, arrays, al.n_slots, al.n_scratch)
        res[en] = ([arrays[isa.ARR_OUT][i] for i in range(12)], interp.work(al.words))
    assert res[False][0] == res[True][0] == O.pairing(p, q)
    assert res[True][1]["macs"] == res[False][1]["macs"]
    assert res[True][1]["slot_moves"] < 0.7 * res[False][1]["slot_moves"]
    lin = sum(res[True][1]["hist"].get(k, 0) for k in ("ADD", "SUB", "DBL", "NEG", "CONJ", "MULXI", "LIN"))
    assert lin < 2000  # the elementary lowering has ~16 500 stand-alone linear instructions


@pytest.mark.parametrize("name,k", [("pairing_v0", 4), ("pairing_v1", 3), ("miller", 4), ("final_exp_v0", 2)])
def test_phase_split_programs_equal_the_monolithic_program(name, k):
    """microcode/phases.py: K separately allocated phase programs, run in order and handing their live
    values over through the state array, give the monolithic program's output bit for bit."""
    from plonky2_bn254_pairing_b200.microcode import fuse, phases

    p, q = PTS[0]
    pool = ConstPool()
    fops = fuse.fuse(programs.build_program(name, pool).ops)
    segs, n_state = phases.split(fops, k)
    assert len(segs) == k and 0 < n_state <= 64
    arrays = g1g2([PTS[0]])
    arrays[isa.ARR_F12] = O.miller_loop_native(q, p)
    arrays[isa.ARR_OUT] = {}
    arrays[isa.ARR_STATE] = {}
    works = []
    for seg in segs:
        al = alloc.allocate(seg, 14)
        interp.run(al.words, pool.values, arrays, al.n_slots, al.n_scratch)
        works.append(interp.work(al.words)["macs"])
    got = [arrays[isa.ARR_OUT][i] for i in range(12)]
    m = O.miller_loop_native(q, p)
    want = {"pairing_v0": O.final_exp_native(m), "pairing_v1": O.final_exp_ark(m), "miller": m,
            "final_exp_v0": O.final_exp_native(m)}[name]
    assert got == want
    assert max(works) < 1.25 * sum(works) / k  # balanced: the longest phase bounds the partial last round


def test_split_phases_taper_off():
    """the shipped split: 16 phases, the last three about 1/2, 1/4 and 1/8 of the others (phases.TAIL) - the warps
    that run out of tasks at the end of a launch wait for half a LAST-phase task on average"""
    from plonky2_bn254_pairing_b200.microcode import fuse, phases

    for name in ("pairing_v0", "pairing_v1", "final_exp_v0", "pairing_x4_v0"):
        fops = fuse.fuse(programs.build_program(name, ConstPool()).ops)
        segs, _ = phases.split(fops, 16)
        assert len(segs) == 16
        w = [sum(phases.COST.get(o.op, 20) for o in seg) for seg in segs]
        body = sum(w[:13]) / 13
        assert max(w[:13]) < 1.25 * body
        assert 0.35 * body < w[13] < 0.65 * body and 0.15 * body < w[14] < 0.35 * body and w[15] < 0.2 * body, (name, w)


def test_validation_programs():
    """the residuals are all zero exactly for points `G1Affine::new` / `G2Affine::new` accept."""
    p, q = PTS[0]
    got, _ = run("validate_g1", {isa.ARR_G1: [p[0], p[1]]}, 9)
    assert got[:2] == [0, 0]
    got, _ = run("validate_g1", {isa.ARR_G1: [p[0], (p[1] + 5) % O.P]}, 9)
    assert got[0] != 0 and got[1] != 0
    flat = [q[0][0], q[0][1], q[1][0], q[1][1]]
    got, al = run("validate_g2", {isa.ARR_G2: flat}, 9)
    assert got[:6] == [0] * 6
    bad = list(flat)
    bad[3] = (bad[3] + 1) % O.P
    got, _ = run("validate_g2", {isa.ARR_G2: bad}, 9)
    assert any(got[:2]) and any(got[2:6])
    # 3 * Q' for a point Q' of the twist outside the subgroup: cofactor-cleared points pass, Q' itself fails
    # (Q' = a fixed on-twist point found by solving y^2 = x^3 + 3/(9+u) for x = (1, 0) ... (5, 0))
    b_tw = O.fq2_mul((3, 0), O.fq2_inv((9, 1)))
    for x0 in range(1, 40):
        x = (x0, 0)
        rhs = O.fq2_mul(O.fq2_mul(x, x), x)
        rhs = ((rhs[0] + b_tw[0]) % O.P, (rhs[1] + b_tw[1]) % O.P)
        # Fq2 square root by the norm trick (p = 3 mod 4)
        n = (rhs[0] * rhs[0] + rhs[1] * rhs[1]) % O.P
        s = pow(n, (O.P + 1) // 4, O.P)
        if s * s % O.P != n:
            continue
        for sgn in (1, -1):
            t = (rhs[0] + sgn * s) * pow(2, O.P - 2, O.P) % O.P
            c = pow(t, (O.P + 1) // 4, O.P)
            if c * c % O.P == t and c:
                y = (c, rhs[1] * pow(2 * c, O.P - 2, O.P) % O.P)
                if O.fq2_mul(y, y) == rhs:
                    got, _ = run("validate_g2", {isa.ARR_G2: [x[0], x[1], y[0], y[1]]}, 9)
                    assert got[:2] == [0, 0] and any(got[2:6])   # on the twist, not in the subgroup
                    cleared = O.g2_mul((x, y), 2 * O.P - O.R_ORDER)   # times the cofactor: now in the subgroup
                    got, _ = run("validate_g2", {isa.ARR_G2: [cleared[0][0], cleared[0][1], cleared[1][0], cleared[1][1]]}, 9)
                    assert got[:6] == [0] * 6
                    return
    raise AssertionError("no twist point found")


def test_prepared_g2_programs():
    """g2_prepare's line coefficients fed to the prepared-pair programs give the same values as the
    programs that do the point arithmetic themselves (a single pairing, and the Groth16 shape 1 live + 3 prepared)."""
    coeffs = []
    for _, q in PTS[:4]:
        got, _ = run("g2_prepare", {isa.ARR_G2: [q[0][0], q[0][1], q[1][0], q[1][1]]}, 18)
        assert len(got) == programs.PREP_FQ
        coeffs.append(got)
    p, q = PTS[0]
    got, _ = run("pairing_p0_1_v0", {isa.ARR_G1: [p[0], p[1]], isa.ARR_AUX: coeffs[0]}, 18)
    assert got == O.pairing(p, q)
    # 1 live + 3 prepared = the product of four pairings
    arrays = g1g2(PTS[:4])
    arrays[isa.ARR_G2] = arrays[isa.ARR_G2][:4]
    arrays[isa.ARR_AUX] = coeffs[1] + coeffs[2] + coeffs[3]
    want = O.final_exp_native(O.multi_miller_loop_native(PTS[:4]))
    got, _ = run("pairing_p1_3_v0", arrays, 18)
    assert got == want
    got, _ = run("pairing_p1_3_v1", arrays, 18)
    assert got == O.final_exp_ark(O.multi_miller_loop_native(PTS[:4]))


def test_windowed_exponent_digits():
    """builder.wnaf_digits: the digits the hard part's m^x walks (width 4) rebuild BN_X, are odd and below 8, keep at
    least three zeros between non-zero digits, and width 2 is the reference's NAF (get_naf, final_exp_native.rs:86-128)."""
    from plonky2_bn254_pairing_b200.microcode.builder import BN_X, wnaf_digits

    d = wnaf_digits(BN_X, 4)
    assert sum(z << i for i, z in enumerate(d)) == BN_X
    nz = [i for i, z in enumerate(d) if z]
    assert all(d[i] % 2 and abs(d[i]) < 8 for i in nz) and all(b - a >= 4 for a, b in zip(nz, nz[1:]))
    assert d[-1] > 0 and len(nz) == 14
    assert wnaf_digits(BN_X, 2) == naf_digits(BN_X) == [int(z) for z in O.get_naf([BN_X])][:len(naf_digits(BN_X))]
