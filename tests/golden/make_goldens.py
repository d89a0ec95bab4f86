#!/usr/bin/env python
"""Writes tests/golden/reference_goldens.json: the golden vectors / known-answer tests that the
reference's OWN tests hold for the per-site estimator path, transcribed by hand with their
file:line under /root/reference/src (the reference is Rust and cannot be executed here, so the
expected values are the literals asserted by those tests, not outputs of a run).
Weir & Cockerham entries are NOT reference goldens (no reference test pins W&C values): they are
the hand-derived KATs of SURVEY.md §8c and are marked "pinned": false."""
import json
import os


def V(pos, gts):
    return {"position": pos, "genotypes": gts}


BOTH = lambda samples: [[s, side] for s in samples for side in (0, 1)]  # noqa: E731
PI43 = (4.0 / 3.0) * 0.625

CASES = [
    # ---- segregating sites
    {"id": "seg_mixed", "kind": "seg_sites", "source": "tests/stats_tests.rs:240-272", "pinned": True,
     "variants": [V(1, [[0, 0], [0, 1], [1, 1]]), V(2, [[0, 0], [0, 0], [0, 0]]), V(3, [[0, 1], [0, 1], [0, 1]]),
                  V(4, [[0, 0], [1, 1], [0, 1]])], "expect": 3},
    {"id": "seg_fixed", "kind": "seg_sites", "source": "tests/stats_tests.rs:240-272", "pinned": True,
     "variants": [V(1, [[0, 0]] * 3), V(2, [[1, 1]] * 3)], "expect": 0},
    {"id": "seg_missing", "kind": "seg_sites", "source": "tests/stats_tests.rs:240-272", "pinned": True,
     "variants": [V(1, [[0, 0], None, [1, 1]]), V(2, [[0, 1], [0, 1], None])], "expect": 2},
    {"id": "seg_pytest", "kind": "seg_sites", "source": "pytests/test_ferromic.py:14-25", "pinned": True,
     "variants": [V(100, [[0, 0], [0, 1]]), V(150, [[0, 0], [0, 0]]), V(200, [[0, 1], [1, 1]])], "expect": 2},
    # ---- Watterson theta (S, n, L)
    {"id": "theta_1", "kind": "theta", "source": "tests/stats_tests.rs:473-506", "pinned": True,
     "args": [10, 5, 1000], "expect": 0.0048, "abs": 1e-6},
    {"id": "theta_2", "kind": "theta", "source": "tests/stats_tests.rs:473-506", "pinned": True,
     "args": [5, 2, 1000], "expect": 0.005, "abs": 1e-6},
    {"id": "theta_3", "kind": "theta", "source": "tests/stats_tests.rs:473-506", "pinned": True,
     "args": [100, 10, 1000000], "expect": 0.00003534, "abs": 1e-6},
    {"id": "theta_4", "kind": "theta", "source": "tests/stats_tests.rs:1360-1420", "pinned": True,
     "args": [2, 4, 100], "expect": 12.0 / 11.0 / 100.0, "abs": 1e-10},
    {"id": "theta_5", "kind": "theta", "source": "pytests/test_ferromic.py:27-33", "pinned": True,
     "args": [3, 4, 100], "expect": 3 / (1 + 1 / 2 + 1 / 3) / 100, "rel": 1e-12},
    # ---- nucleotide diversity (variants, haplotypes, L)
    {"id": "pi_fixed", "kind": "pi", "source": "tests/stats_tests.rs:520-539", "pinned": True,
     "variants": [V(100, [[0, 0], [0, 0]]), V(200, [[1, 1], [1, 1]])], "haplotypes": BOTH([0, 1]), "L": 1000,
     "expect": 0.0, "abs": 0.0},
    {"id": "pi_uncallable_site", "kind": "pi", "source": "tests/stats_tests.rs:607-623", "pinned": True,
     "variants": [V(10, [[0, 0], [1, 1]]), V(20, [None, None])], "haplotypes": BOTH([0, 1]), "L": 2,
     "expect": 2.0 / 3.0, "abs": 1e-9},
    {"id": "pi_single_haplotype", "kind": "pi", "source": "tests/stats_tests.rs:651-665", "pinned": True,
     "variants": [V(100, [[0, 1]])], "haplotypes": [[0, 0]], "L": 1000, "expect": "nan"},
    # ---- Hudson per-site + regional (4 samples, pops {0,1} vs {2,3}, region, L)
    {"id": "hudson_ratio_of_sums", "kind": "hudson_sites", "source": "tests/hudson_fst_tests.rs:363-513",
     "pinned": True, "variants": [V(100, [[0, 0], [0, 0], [1, 1], [1, 1]]), V(200, [[0, 1]] * 4)],
     "pop1": BOTH([0, 1]), "pop2": BOTH([2, 3]), "region": [100, 200], "L": 2,
     "expect": {"fst": 5.0 / 9.0, "site_fst": [1.0, -1.0 / 3.0], "site_num": [1.0, -1.0 / 6.0],
                "site_den": [1.0, 0.5]}, "abs": 1e-12},
    {"id": "hudson_uneven_missingness", "kind": "hudson_sites", "source": "tests/hudson_fst_tests.rs:516-665",
     "pinned": True, "variants": [V(100, [[0, 0], [0, 0], [1, 1], [1, 1]]), V(200, [None, [0, 1], None, [0, 1]])],
     "pop1": BOTH([0, 1]), "pop2": BOTH([2, 3]), "region": [100, 200], "L": 2,
     "expect": {"fst": 1.0 / 3.0, "site_fst": [1.0, -1.0], "site_num": [1.0, -0.5], "site_den": [1.0, 0.5]},
     "abs": 1e-12},
    {"id": "hudson_tri_allelic", "kind": "hudson_sites", "source": "tests/hudson_fst_tests.rs:877-1006",
     "pinned": True, "variants": [V(100, [[0, 0], [1, 2], [0, 1], [2, 2]])],
     "pop1": BOTH([0, 1]), "pop2": BOTH([2, 3]), "region": [100, 100], "L": 1,
     "expect": {"site_dxy": [0.6875], "site_pi1": [PI43], "site_pi2": [PI43],
                "site_fst": [(0.6875 - PI43) / 0.6875]}, "abs": 1e-12},
    {"id": "hudson_falsta_tracks", "kind": "hudson_sites", "source": "tests/stats_tests.rs:1860-2034",
     "pinned": True, "variants": [V(1, [[0, 0], [1, 1]]), V(2, [[0, 1], [0, 1]]), V(3, [[1, 1], [0, 0]])],
     "pop1": BOTH([0]), "pop2": BOTH([1]), "region": [0, 10], "L": 3,
     "expect": {"site_fst": [1.0, -1.0, 1.0], "site_num": [1.0, -0.5, 1.0], "site_den": [1.0, 0.5, 1.0]},
     "abs": 1e-12},
    # ---- adjusted sequence length (start, end, allow, mask)
    {"id": "ladj_mask", "kind": "adjusted_length", "source": "tests/stats_tests.rs:1829-1858", "pinned": True,
     "args": [100, 200, None, [[100, 101]]], "expect": 100},
    {"id": "ladj_allow_mask", "kind": "adjusted_length",
     "source": "pytests/test_ferromic.py:49-60 (stale literal 25; current stats.rs:3644-3736 yields 24, SURVEY §4)",
     "pinned": True, "args": [1, 100, [[11, 20], [40, 60]], [[45, 50]]], "expect": 24},
    # ---- Weir & Cockerham (derived KATs, NOT pinned by the reference)
    {"id": "wc_perfect", "kind": "wc_site", "source": "SURVEY.md §8c (derived from stats.rs:1814-2127)",
     "pinned": False, "genotypes": [[0, 0], [0, 0], [1, 1], [1, 1]], "groups": [0, 0, 1, 1],
     "expect": {"a": 1.0, "b": 0.0, "fst": 1.0}, "abs": 1e-12},
    {"id": "wc_all_het", "kind": "wc_site", "source": "SURVEY.md §8c", "pinned": False,
     "genotypes": [[0, 1]] * 4, "groups": [0, 0, 1, 1],
     "expect": {"a": -1.0 / 6.0, "b": 2.0 / 3.0, "fst": -1.0 / 3.0}, "abs": 1e-12},
    {"id": "wc_mixed", "kind": "wc_site", "source": "SURVEY.md §8c", "pinned": False,
     "genotypes": [[0, 0], [0, 1], [1, 1], [0, 1]], "groups": [0, 0, 1, 1],
     "expect": {"a": 0.125, "b": 0.5, "fst": 0.2}, "abs": 1e-12},
    {"id": "wc_unequal", "kind": "wc_site", "source": "SURVEY.md §8c", "pinned": False,
     "genotypes": [[0, 0], [0, 1], [0, 0], [1, 1]], "groups": [0, 0, 0, 1],
     "expect": {"a": 0.601851851851852, "b": 0.277777777777778, "fst": 0.684210526315789}, "abs": 1e-12},
]

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")
    with open(out, "w") as f:
        json.dump({"reference": "SauersML/ferromic (src/ paths below are relative to /root/reference/src)",
                   "cases": CASES}, f, indent=1)
    print(f"wrote {len(CASES)} cases to {out}")
