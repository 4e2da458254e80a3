"""Expected outputs of the `optest` sequencer program (microcode/programs.py::prog_optest), spelled out
with the oracle's plain-integer field arithmetic."""
import bn254_oracle as O

NAMES = ["MUL", "SQR", "MULFP0", "MULFP1", "ADD", "SUB", "NEG", "CONJ", "MULXI", "DBL", "INV", "MULC",
         "(a+b)(c+e)", "(a-b)(c-e)", "(a+b)c", "a(c-e)", "(a+b)^2", "(a-b)^2",
         "a-b-c+d", "a+xi(b-c-d)", "27a-xi b", "3a-2b", "conj(a)+xi b-2c", "12a-conj(b)",
         "ab-c", "(a+b)(c+d)-e-f", "(a+b)(c+d)-e-f+a", "cd+a-b+f", "e+xi(af-b-c)", "(b+c)(d-e)-a-f+xi c",
         "3d^2-2e", "3((a+b)^2-c-d)+2f", "b*c1-d+e", "(a-b)c-d-e-f", "be-a-c", "d+xi(be-a-c)"]


def times(x, k):
    return (x[0] * k % O.P, x[1] * k % O.P)


def expected(row):
    """row: 12 Fq ints (MyFq12 order) -> list of OPTEST_OUTPUTS Fq2 values."""
    x = [(row[i], row[i + 6]) for i in range(6)]
    add, sub, mul, xi = O.fq2_add, O.fq2_sub, O.fq2_mul, lambda v: O.fq2_mul(v, O.XI)
    c = O._expected_c()
    c3 = mul(mul(c, c), c)
    return [
        mul(x[0], x[1]), O.fq2_sqr(x[2]), mul(x[3], (x[4][0], 0)), mul(x[3], (x[4][1], 0)),
        add(x[0], x[5]), sub(x[1], x[2]), O.fq2_neg(x[3]), O.conjugate_fp2(x[4]), xi(x[5]), add(x[0], x[0]),
        O.fq2_inv(x[1]) if x[1] != (0, 0) else (0, 0), mul(x[2], c3),
        mul(add(x[0], x[1]), add(x[2], x[3])), mul(sub(x[0], x[1]), sub(x[2], x[3])),
        mul(add(x[0], x[1]), x[2]), mul(x[0], sub(x[2], x[3])),
        O.fq2_sqr(add(x[4], x[5])), O.fq2_sqr(sub(x[4], x[5])),
        add(sub(sub(x[0], x[1]), x[2]), x[3]),
        add(x[0], xi(sub(sub(x[1], x[2]), x[3]))),
        sub(times(x[0], 27), xi(x[1])),
        sub(times(x[2], 3), times(x[3], 2)),
        sub(add(O.conjugate_fp2(x[0]), xi(x[1])), times(x[2], 2)),
        sub(times(x[5], 12), O.conjugate_fp2(x[4])),
        sub(mul(x[0], x[1]), x[2]),
        sub(sub(mul(add(x[0], x[1]), add(x[2], x[3])), x[4]), x[5]),
        add(sub(sub(mul(add(x[0], x[1]), add(x[2], x[3])), x[4]), x[5]), x[0]),
        add(sub(add(mul(x[2], x[3]), x[0]), x[1]), x[5]),
        add(x[4], xi(sub(sub(mul(x[0], x[5]), x[1]), x[2]))),
        add(sub(sub(mul(add(x[1], x[2]), sub(x[3], x[4])), x[0]), x[5]), xi(x[2])),
        sub(times(O.fq2_sqr(x[3]), 3), times(x[4], 2)),
        add(times(sub(sub(O.fq2_sqr(add(x[0], x[1])), x[2]), x[3]), 3), times(x[5], 2)),
        add(sub(mul(x[1], (x[2][1], 0)), x[3]), x[4]),
        sub(sub(sub(mul(sub(x[0], x[1]), x[2]), x[3]), x[4]), x[5]),
        sub(sub(mul(x[1], x[4]), x[0]), x[2]),
        add(x[3], xi(sub(sub(mul(x[1], x[4]), x[0]), x[2]))),
    ]


def edge_rows(rnd, n_random=96):
    """Input rows hitting the boundaries of the lazy-reduction ranges: 0, 1, p-1, p-2, (p+-1)/2, ..."""
    edge = [0, 1, 2, O.P - 1, O.P - 2, (O.P - 1) // 2, (O.P + 1) // 2, 9, 2 ** 253, 2 ** 224 - 1]
    rows = [[rnd.randrange(O.P) for _ in range(12)] for _ in range(n_random)]
    for i, v in enumerate(edge):
        rows[i] = [v] * 12
        rows[len(edge) + i] = [rnd.choice(edge) for _ in range(12)]
    rows[2 * len(edge)] = [O.P - 1 if k % 2 == 0 else 0 for k in range(12)]
    rows[2 * len(edge) + 1] = [0 if k % 2 == 0 else O.P - 1 for k in range(12)]
    for r in rows:  # slot 1 = (r[1], r[7]) is inverted by the program: keep it non-zero
        if r[1] == 0 and r[7] == 0:
            r[1] = 5
    return rows
