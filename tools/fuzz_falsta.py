"""Differential fuzz of the device FALSTA renderer against the oracle writer: random record sets (ties at the 7th
decimal, subnormals, huge and tiny magnitudes, NaN / +-inf / +-0), duplicate and out-of-region positions, regions
that start below 1 or end before they start.  usage: python tools/fuzz_falsta.py [n_cases]"""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ferromic_b200 import falsta  # noqa: E402
from oracle import falsta as ofa  # noqa: E402


def values(rng, n):
    kind = rng.integers(0, 8, size=n)
    v = rng.random(n)
    v = np.where(kind == 1, np.exp(rng.uniform(-30, 30, n)) * rng.choice([-1.0, 1.0], n), v)
    v = np.where(kind == 2, rng.integers(1, 4096, n) / 128.0, v)                 # exact ties
    v = np.where(kind == 3, (rng.integers(0, 10 ** 6, n) + 0.5) / 1e6, v)         # near rounding boundaries
    v = np.where(kind == 4, rng.choice([0.0, -0.0, np.nan, np.inf, -np.inf, 5e-324, 1e-7, 4.9999995e-7], n), v)
    near = kind == 5
    if near.any():
        b = (rng.integers(0, 10 ** 7, near.sum()) + 0.5) / 1e6
        u = b.view(np.int64) + rng.integers(-2, 3, near.sum())
        v[near] = u.view(np.float64)
    return v


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    rng = np.random.default_rng(77)
    bad = 0
    for case in range(n_cases):
        rs = int(rng.integers(-20, 3000))
        re_ = rs + int(rng.integers(-5, 4000))
        n = int(rng.choice([0, 1, 3, 50, 700, 5000]))
        pos = rng.integers(rs - 30, max(re_, rs) + 30, size=n)
        wc = [(int(p), *map(float, values(rng, 6))) for p in pos]
        hud = [(int(p), *map(float, values(rng, 3))) for p in pos[::2]]
        div = [(int(p), float(a), float(b), int(g), bool(f)) for p, a, b, g, f in
               zip(pos, np.nan_to_num(values(rng, n), posinf=1.5, neginf=-1.5), np.nan_to_num(values(rng, n), posinf=2.5, neginf=-2.5),
                   rng.integers(0, 3, n), rng.random(n) < 0.5)]
        try:
            assert falsta.fst_falsta_text("c", rs, re_, wc, hud).decode() == ofa.fst_falsta_text("c", rs, re_, wc, hud)
            assert falsta.diversity_falsta_text("c", rs, re_, div).decode() == ofa.diversity_falsta_text("c", rs, re_, div)
        except AssertionError:
            bad += 1
            print("MISMATCH case", case, rs, re_, n)
    print("falsta fuzz done:", n_cases, "cases,", bad, "mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
