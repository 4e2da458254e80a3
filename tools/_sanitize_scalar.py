"""compute-sanitizer target: scalar multiplication on a handful of points (edge scalars included) and one phase-split launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bn254_oracle as O
from plonky2_bn254_pairing_b200 import api, native

native.init([0])
pts = O.seeded_points(0xB2540000, 4)
R = O.R_ORDER
ks = [0, 1, R, (1 << 256) - 1]
assert api.g1_scalar_mul_batch([p for p, _ in pts], ks) == [O.g1_mul(p, k) for (p, _), k in zip(pts, ks)]
assert api.g2_scalar_mul_batch([q for _, q in pts], ks) == [O.g2_mul(q, k) for (_, q), k in zip(pts, ks)]
print("scalar ok")
