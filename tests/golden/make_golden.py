"""Regenerates tests/golden/*.json.

  survey_kat.json   - the known-answer vectors of SURVEY.md Appendix C, copied verbatim (they were
                      produced by an independent transcription of the reference during the survey and
                      are what pins oracle/bn254_oracle.py; the reference itself ships no vectors).
  oracle_vectors.json - seeded inputs with oracle outputs (miller / final exp / pairing / ark variant /
                      3-way multi miller), used by the GPU parity tests and by the C-oracle tests.

    python tests/golden/make_golden.py
The reference cannot be imported or built in this environment (Rust, un-vendored git deps), so these
fixtures come from the oracle, after the oracle was checked against survey_kat.json.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bn254_oracle as O  # noqa: E402


def parse_survey():
    txt = open(os.path.join(ROOT, "SURVEY.md")).read()
    app = txt[txt.index("## Appendix C"):txt.index("## Appendix D")]
    blocks = {}
    cur = None
    for line in app.splitlines():
        m = re.match(r"^(KAT\d)\s+(.*)$", line)
        if m and "[" not in line.split("  ")[0]:
            cur = m.group(1) + " " + m.group(2).strip()
            blocks[cur] = {}
            continue
        if cur:
            for idx, val in re.findall(r"\[\s*(\d+)\]\s+([0-9a-f]{64})", line):
                blocks[cur][int(idx)] = val
    names = ["kat1_miller", "kat1_pairing", "kat1_pairing_ark", "kat2_miller", "kat2_pairing"]
    full = [v for v in blocks.values() if len(v) == 12]
    assert len(full) == len(names)
    return {nm: [v[i] for i in range(12)] for nm, v in zip(names, full)}


def hexs(xs):
    return ["%064x" % x for x in xs]


def main():
    kat = parse_survey()
    assert len(kat) == 5, list(kat)
    with open(os.path.join(HERE, "survey_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    vec = {"seed": 0xB2540000, "cases": []}
    pts = O.seeded_points(0xB2540000, 6)
    for (p, q) in pts:
        m = O.miller_loop_native(q, p)
        vec["cases"].append({
            "g1": hexs([p[0], p[1]]),
            "g2": hexs([q[0][0], q[0][1], q[1][0], q[1][1]]),
            "miller": hexs(m),
            "pairing": hexs(O.final_exp_native(m)),
            "pairing_ark": hexs(O.final_exp_ark(m)),
        })
    vec["multi3"] = {
        "pairs": [0, 1, 2],
        "miller": hexs(O.multi_miller_loop_native(pts[0:3])),
        "pairing": hexs(O.final_exp_native(O.multi_miller_loop_native(pts[0:3]))),
    }
    import random
    rnd = random.Random(0xB254)
    x = [rnd.randrange(O.P) for _ in range(12)]
    vec["random_fq12"] = {
        "x": hexs(x),
        "final_exp": hexs(O.final_exp_native(x)),
        "pow_x": hexs(O.pow_native(x, [O.BN_X])),
        "frobenius": {str(k): hexs(O.frobenius_map_native(x, k)) for k in (0, 1, 2, 3, 6, 7, 13)},
    }
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as f:
        json.dump(vec, f, indent=1)
    print("wrote", len(kat), "survey KAT blocks and", len(vec["cases"]), "oracle cases")


if __name__ == "__main__":
    main()
