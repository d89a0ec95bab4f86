"""Differential fuzz of the device VCF parser against the oracle (tests/test_vcf.py's generator and checker) over
many seeds and shapes.  usage: python tools/fuzz_vcf.py [n_seeds]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_vcf import check_against_oracle, make_vcf  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    bad = 0
    for seed in range(1000, 1000 + n):
        rng = np.random.default_rng(seed)
        n_cols = int(rng.choice([1, 2, 3, 9, 33, 120, 600, 1500]))
        n_lines = int(rng.choice([1, 5, 40, 200])) if n_cols > 500 else int(rng.choice([1, 7, 90, 700]))
        odd = float(rng.choice([0.0, 0.01, 0.05, 0.2]))
        text = make_vcf(rng, n_lines, n_cols, odd=odd, crlf=bool(rng.integers(0, 2)))
        k = max(1, int(n_cols * rng.uniform(0.3, 1.0)))
        kept = sorted(rng.choice(np.arange(9, 9 + n_cols), size=k, replace=False).tolist())
        regions = [(950, 1200), (1200, 1300), (1500, 2100)] if rng.integers(0, 2) else [(0, 10 ** 9)]
        allow = None if rng.integers(0, 2) else {"1": [(int(a), int(a + rng.integers(1, 300))) for a in rng.integers(900, 2200, 6)]}
        mask = None if rng.integers(0, 2) else {"1": [(int(a), int(a + rng.integers(1, 30))) for a in rng.integers(900, 2200, 9)]}
        try:
            check_against_oracle(text, "chr1", regions, kept, int(rng.choice([0, 30, 99])), allow, mask,
                                 max_ploidy=int(rng.choice([2, 3, 4])))
        except AssertionError as e:
            bad += 1
            print("MISMATCH seed", seed, n_lines, n_cols, odd, str(e)[:300])
    print("fuzz done:", n, "cases,", bad, "mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
