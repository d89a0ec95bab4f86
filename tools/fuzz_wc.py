"""Differential fuzz of the Weir & Cockerham path (K1 count-only groups -> fm_k_wc_pairs_pc / fm_k_wc_overall ->
fm_k_wc_fold) against the C oracle: random numbers of populations (2 .. 45: up to three chunks of pair slots), unequal
sizes, samples without a group or with their haplotypes in two groups, missing calls, monomorphic / empty sites, and
regions that cut segments at odd places.  Per-site (a, b) of every pair must match to 1e-12, region sums to 1e-9
relative (or, for a sum that cancels to ~0, to 1e-13 of the sum of its summands' magnitudes: the reference adds site
after site, the device adds 1024-site segment partials), states and site counts exactly.  usage: python tools/fuzz_wc.py [n_cases]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as orc  # noqa: E402
from tests.test_gpu_parity import _to_python_variants, _wc_compare, fm  # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(20261019)
    F = fm()
    bad = cancel = 0
    for case in range(n_cases):
        G = int(rng.choice([2, 3, 5, 9, 17, 27, 28, 33, 45]))
        per = rng.integers(1, 7, size=G)                       # diploid samples per population
        S = int(per.sum()) + int(rng.integers(0, 4))           # a few samples without a group
        V = int(rng.choice([1, 31, 33, 200, 1100, 2300]))      # 1024-site segments: up to three per region
        base = rng.beta(0.6, 0.6, size=V)
        g = (rng.random((V, S, 2)) < base[:, None, None]).astype(np.int8)
        miss = rng.random((V, S)) < rng.choice([0.0, 0.02, 0.3])
        g[miss] = -1
        if V > 6:
            g[2] = 0
            g[3] = -1
            g[4, : S // 2] = -1
        pos = np.cumsum(rng.integers(1, 30, size=V, dtype=np.int64))
        left = np.full(S, 0xFFFF, dtype=np.uint16)
        s = 0
        for p, k in enumerate(per):
            left[s:s + k] = p
            s += k
        right = left.copy()
        swap = rng.choice(S, size=min(S, 3), replace=False)
        right[swap] = rng.integers(0, G, size=swap.size)
        if S > 4:
            left[rng.integers(0, S)] = 0xFFFF
        labels = sorted(str(i) for i in range(G))
        lo = int(pos[int(rng.integers(0, max(1, V // 3)))]) - int(rng.integers(0, 2))
        hi = int(pos[int(rng.integers(2 * V // 3, V))]) + int(rng.integers(0, 2))
        vs, _ = orc.from_numpy(g, pos)
        try:
            _wc_compare(F, _to_python_variants(g, pos), vs, left, right, labels, (lo, hi))
        except AssertionError as e:  # noqa: PERF203
            # A pair's REGION sums may fail the relative test when the sum cancels to ~0: the reference adds the sites
            # one after the other, the device adds 1024-site segment partials, and the two differ by rounding noise of
            # the SUMMANDS (~1e-16 * sum |a_site|).  That is not a mismatch; anything larger, or any other assertion, is.
            ref = orc.wc_fst(vs, left, right, G, (lo, hi))
            got = F.wc_fst_from_membership(_to_python_variants(g, pos), labels, left, right, (lo, hi))
            keys = [f"{labels[i]}_vs_{labels[j]}" for i in range(G) for j in range(i + 1, G)]
            real = str(e) not in keys
            for k, key in enumerate(keys):
                if not ref["pair_present"][k] or key not in got.pairwise_fst:
                    continue
                e2, r = got.pairwise_fst[key], ref["pairs"][k]
                mag = {"sum_a": sum(abs(float(ref["pair_a"][i][k])) for i in range(ref["n_sites"])
                                    if ref["has_maps"][i] and ref["pair_a"][i][k] == ref["pair_a"][i][k]),
                       "sum_b": sum(abs(float(ref["pair_b"][i][k])) for i in range(ref["n_sites"])
                                    if ref["has_maps"][i] and ref["pair_b"][i][k] == ref["pair_b"][i][k])}
                if e2.state != r["state"] or e2.sites != r["sites"]:
                    real = True
                for name in ("sum_a", "sum_b"):
                    x, y = getattr(e2, name), r[name]
                    if abs(x - y) > 1e-9 * max(abs(x), abs(y)) and abs(x - y) > 1e-13 * mag[name]:
                        real = True
                        print(f"   {key} {name}: got {x!r} ref {y!r} (sum of |site values| {mag[name]:.6g})")
            if real:
                bad += 1
                print(f"case {case}: G={G} S={S} V={V} region=({lo}, {hi}) MISMATCH {e}")
            else:
                cancel += 1
    print(f"wc fuzz done: {n_cases} cases, {bad} mismatches ({cancel} cases with a pair sum that cancels to ~0 and agrees "
          "to rounding noise of its summands)")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
