"""Differential fuzz of K1 (u8 rows -> group bitplanes / counts) against the oracle's dense summary: random sample
counts (row strides that are and are not multiples of 16: direct and staged packing), ploidy 1/2, the three
missingness modes (none, bitmap, in-band int8), random overlapping / empty / single-haplotype groups created
together (compress plans) and alone, W&C partitions (count-only groups).  usage: python tools/fuzz_repack.py [n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ferromic_b200.api import _Matrix  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(2718)
    bad = 0
    for case in range(n_cases):
        S = int(rng.choice([1, 2, 7, 8, 16, 33, 64, 100, 256, 500, 1000, 2504]))
        P = int(rng.choice([1, 2, 2, 2]))
        V = int(rng.choice([1, 31, 32, 33, 200, 1500]))
        mode = int(rng.integers(0, 3))
        g = rng.binomial(1, rng.beta(0.5, 0.5, size=V)[:, None, None], size=(V, S, P)).astype(np.int8)
        miss = (rng.random(g.shape) < rng.choice([0.0, 0.02, 0.3])) if mode else np.zeros(g.shape, bool)
        pos = np.cumsum(rng.integers(1, 9, size=V)).astype(np.int64)
        if mode == 2:
            gi = g.copy()
            gi[miss] = -1
            m = _Matrix.from_int8(gi, pos, 1)
        else:
            m = _Matrix(g.astype(np.uint8), miss if mode == 1 else None, pos, max_allele=1, always_bitmap=mode == 1)
        d = orc.Dense(g.astype(np.uint8).reshape(-1), orc.pack_missing_bits(miss.reshape(-1)) if mode else None, V, S, P, 1)
        lists = []
        for _ in range(int(rng.integers(1, 6))):
            k = int(rng.choice([0, 1, max(1, S // 3), S]))
            samples = rng.choice(S, size=min(k, S), replace=False) if k else []
            haps = [(int(s), int(rng.integers(0, 2))) for s in samples]
            if rng.random() < 0.5:
                haps += [(int(s), 1 - side) for s, side in haps[: len(haps) // 2]]
            if haps and rng.random() < 0.3:
                haps.append(haps[0])  # a duplicate is counted once
            lists.append(haps)
        groups = m.groups(lists) if rng.random() < 0.7 else [m.group(h) for h in lists]
        for haps, grp in zip(lists, groups):
            s, so = grp.summary(True), orc.build_summary(d, haps)
            if not (np.array_equal(s["alt"], so.alt) and np.array_equal(s["called"], so.called)
                    and s["segregating_sites"] == so.seg):
                bad += 1
                print("MISMATCH case", case, dict(S=S, P=P, V=V, mode=mode, n=len(haps)))
                break
    print("repack fuzz done:", n_cases, "cases,", bad, "mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
