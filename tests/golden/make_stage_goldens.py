#!/usr/bin/env python
"""Writes tests/golden/stage_goldens.json: fixtures for the two adjacent stages (VCF parse/filter, FALSTA writers).
The `reference` cases carry the literals the reference's own tests assert (file:line cited); the `generated` cases
are inputs from tests/test_vcf.py's generator with the outputs of the line-by-line restatement (oracle/vcf.py,
oracle/falsta.py) -- the reference is Rust and cannot be executed here.  Run from the repo root."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import falsta as ofa  # noqa: E402
from oracle import vcf as ov  # noqa: E402
from tests.test_vcf import make_vcf  # noqa: E402


def vcf_case(cid, text, chr_, regions, kept, min_gq, allow=None, mask=None, source="generated"):
    out, miss, stats, errors = ov.process_lines(ov.split_lines(text), chr_, regions, kept, min_gq, allow, mask)
    return {"id": cid, "source": source, "text": text, "chr": chr_, "regions": regions, "kept": kept, "min_gq": min_gq,
            "allow": allow, "mask": mask,
            "expect": {"positions": [v[0] for v in out], "flags": [v[2] for v in out],
                       "genotypes": [v[1] for v in out], "allele_info": [[v[3][0], v[3][1]] for v in out],
                       "errors": [[l, m] for l, m in errors],
                       "stats": {"total_variants": stats.total_variants, "filtered_variants": stats.filtered_variants,
                                 "filtered_due_to_mask": stats.filtered_due_to_mask,
                                 "filtered_due_to_allow": stats.filtered_due_to_allow,
                                 "missing_data_variants": stats.missing_data_variants,
                                 "low_gq_variants": stats.low_gq_variants, "mnp_variants": stats.mnp_variants,
                                 "total_data_points": miss.total_data_points,
                                 "missing_data_points": miss.missing_data_points},
                       "positions_with_missing": sorted(miss.positions_with_missing),
                       "filtered_positions": sorted(stats.filtered_positions)}}


def main():
    region = [[999, 2000]]
    vcf = [
        vcf_case("filter_tests_unit", "chr1\t1000\t.\tA\tT\t.\tPASS\t.\tGT:GQ\t0|0:50\t0|1:60\n"
                 "chr1\t1001\t.\tA\tT\t.\tPASS\t.\tGT:GQ\t0|0:20\t0|1:25\n", "1", region, [9, 10], 30,
                 source="src/tests/filter_tests.rs:8-78 (flags 0 then non-zero, positions 999 / 1000, low_gq > 0)"),
        vcf_case("mnp_mixed", "chr1\t1000\t.\tA\tG,TT\t.\tPASS\t.\tGT:GQ\t1|2:40\n", "1", region, [9], 30,
                 source="src/tests/mnp_test.rs:8-43 (None, mnp_variants == 1)"),
        vcf_case("gq_low_and_valid", "chr1\t1000\t.\tA\tT\t.\tPASS\t.\tGT:GQ\t0|0:20\t0|1:40\n"
                 "chr1\t1000\t.\tA\tT\t.\tPASS\t.\tGT:GQ\t0|0:35\t0|1:40\n", "1", region, [9, 10], 30,
                 source="src/tests/stats_tests.rs:882-975 (genotypes [0,0],[0,1]; allele info (999,'A',['T']))"),
    ]
    rng = np.random.default_rng(2026)
    allow = {"1": [[1000, 1100], [1050, 1250], [1600, 1900], [-5, 3], [7, 2]]}
    mask = {"1": [[1020, 1030], [1700, 1705], [-1, 5], [1990, -1]]}
    vcf.append(vcf_case("generated_small", make_vcf(rng, 120, 6, odd=0.1), "chr1",
                        [[950, 1200], [1200, 1300], [1500, 2100]], [9, 10, 12, 14], 30, allow, mask))
    vcf.append(vcf_case("generated_crlf_wide", make_vcf(rng, 40, 90, odd=0.02, crlf=True), " 1",
                        [[0, 10 ** 9]], list(range(9, 99, 2)), 0, None, {"2": [[0, 10]]}))
    # the reference asserts mnp/flags literals: keep them next to the computed expectations
    assert vcf[0]["expect"]["positions"] == [999, 1000] and vcf[0]["expect"]["flags"][0] == 0
    assert vcf[0]["expect"]["flags"][1] != 0 and vcf[0]["expect"]["stats"]["low_gq_variants"] > 0
    assert vcf[1]["expect"]["positions"] == [] and vcf[1]["expect"]["stats"]["mnp_variants"] == 1
    assert vcf[2]["expect"]["genotypes"][0] == [[0, 0], [0, 1]] and vcf[2]["expect"]["allele_info"][0] == ["A", ["T"]]

    def nan_to_str(recs):
        return [[("nan" if isinstance(x, float) and x != x else ("inf" if x == float("inf") else
                 ("-inf" if x == float("-inf") else x))) for x in r] for r in recs]

    fal = []
    div = [(3, 1.0, 1.0, 0, False)]
    fal.append({"id": "missing_sites_default_to_zero", "kind": "diversity", "seqname": "1", "start": 1, "end": 5,
                "source": "src/tests/stats_tests.rs:82-241 ('0,0,<val>,0,0' under >unfiltered_pi_chr_1_start_1_end_5_group_0)",
                "per_site": nan_to_str(div), "expect": ofa.diversity_falsta_text("1", 1, 5, div)})
    hud = [(1, 1.0, 1.0, 1.0), (2, -1.0, -0.5, 0.5), (3, 1.0, 1.0, 1.0)]
    fal.append({"id": "hudson_components", "kind": "fst", "seqname": "1", "start": 1, "end": 3,
                "source": "src/tests/stats_tests.rs:1861-2034 (1, -1, 1 / 1, -0.5, 1 / 1, 0.5, 1)",
                "wc": [], "hudson": nan_to_str(hud), "expect": ofa.fst_falsta_text("1", 1, 3, [], hud)})
    n = 300
    pos = rng.integers(90, 700, size=n)
    v = rng.random((n, 6))
    v[rng.random((n, 6)) < 0.1] = np.nan
    v[rng.random((n, 6)) < 0.1] = 0.0
    v[rng.random((n, 6)) < 0.03] = np.inf
    v[::9, 0] = rng.integers(1, 256, size=len(v[::9, 0])) / 128.0
    wc = [(int(p), *map(float, row)) for p, row in zip(pos, v)]
    hud = [(int(p), float(r[0]), float(r[1]), float(r[2])) for p, r in zip(pos[::3], v[::3])]
    fal.append({"id": "generated_fst", "kind": "fst", "seqname": "chr7", "start": 100, "end": 650, "source": "generated",
                "wc": nan_to_str(wc), "hudson": nan_to_str(hud), "expect": ofa.fst_falsta_text("chr7", 100, 650, wc, hud)})
    dv = [(int(p), float(a) if a == a and abs(a) != float("inf") else 0.25, float(b) if b == b and abs(b) != float("inf") else 0.5,
           int(g), bool(f)) for p, a, b, g, f in zip(pos, v[:, 0], v[:, 1], rng.integers(0, 2, n), rng.random(n) < 0.4)]
    fal.append({"id": "generated_diversity", "kind": "diversity", "seqname": "X", "start": 100, "end": 650,
                "source": "generated", "per_site": nan_to_str(dv), "expect": ofa.diversity_falsta_text("X", 100, 650, dv)})
    path = os.path.join(ROOT, "tests", "golden", "stage_goldens.json")
    json.dump({"vcf": vcf, "falsta": fal}, open(path, "w"), indent=0)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
