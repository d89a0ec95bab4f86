"""Committed fixtures of the two adjacent stages (tests/golden/stage_goldens.json, written by
tests/golden/make_stage_goldens.py): reference-test literals + generated inputs with the restatement's outputs.
CPU: both oracles (Python, compiled C) reproduce the fixture.  GPU (-m gpu): the device parser / renderer do."""
import json
import os

import numpy as np
import pytest

from oracle import falsta as ofa
from oracle import vcf as ov

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "stage_goldens.json")))


def _num(x):
    return {"nan": float("nan"), "inf": float("inf"), "-inf": float("-inf")}.get(x, x) if isinstance(x, str) else x


def _recs(rows):
    return [tuple(_num(x) for x in r) for r in rows]


def _maps(c):
    conv = lambda m: None if m is None else {k: [tuple(iv) for iv in v] for k, v in m.items()}  # noqa: E731
    return conv(c["allow"]), conv(c["mask"])


@pytest.mark.parametrize("case", G["vcf"], ids=[c["id"] for c in G["vcf"]])
def test_vcf_fixture_oracles(case):
    allow, mask = _maps(case)
    regions = [tuple(r) for r in case["regions"]]
    e = case["expect"]
    out, miss, stats, errors = ov.process_lines(ov.split_lines(case["text"]), case["chr"], regions, case["kept"],
                                                case["min_gq"], allow, mask)
    assert [v[0] for v in out] == e["positions"] and [v[2] for v in out] == e["flags"]
    assert [v[1] for v in out] == e["genotypes"] and [[l, m] for l, m in errors] == e["errors"]
    assert stats.total_variants == e["stats"]["total_variants"] and stats.mnp_variants == e["stats"]["mnp_variants"]
    c = ov.c_process_lines(case["text"].encode(), case["chr"], regions, case["kept"], case["min_gq"], allow, mask,
                           max_ploidy=4, threads=2)
    assert list(c["positions"]) == e["positions"] and list(c["flags"]) == e["flags"]
    assert [int(l) for l in c["err_line"]] == [l for l, _ in e["errors"]]
    assert list(map(int, c["counters"])) == [e["stats"][k] for k in (
        "total_variants", "filtered_variants", "filtered_due_to_mask", "filtered_due_to_allow", "missing_data_variants",
        "low_gq_variants", "mnp_variants", "total_data_points", "missing_data_points")]


@pytest.mark.parametrize("case", G["falsta"], ids=[c["id"] for c in G["falsta"]])
def test_falsta_fixture_oracle(case):
    if case["kind"] == "diversity":
        got = ofa.diversity_falsta_text(case["seqname"], case["start"], case["end"], _recs(case["per_site"]))
    else:
        got = ofa.fst_falsta_text(case["seqname"], case["start"], case["end"], _recs(case["wc"]), _recs(case["hudson"]))
    assert got == case["expect"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", G["vcf"], ids=[c["id"] for c in G["vcf"]])
def test_vcf_fixture_device(case):
    from ferromic_b200 import vcf
    allow, mask = _maps(case)
    e = case["expect"]
    b = vcf.process_lines(case["text"].encode(), case["chr"], [tuple(r) for r in case["regions"]], case["kept"],
                          case["min_gq"], allow, mask, max_ploidy=4)
    assert list(b.positions) == e["positions"] and list(b.flags) == e["flags"]
    assert [[l, m] for l, m in b.errors] == e["errors"]
    assert b.stats() == e["stats"]
    assert b.positions_with_missing().tolist() == e["positions_with_missing"]
    assert b.filtered_positions().tolist() == e["filtered_positions"]
    assert [[r, a] for r, a in b.allele_info()] == e["allele_info"]
    gt = b.genotypes()
    for i, gts in enumerate(e["genotypes"]):
        data, stride = ov.compressed(gts)
        assert int(b.stride[i]) == stride and gt[i, :, :stride].tobytes() == data


@pytest.mark.gpu
@pytest.mark.parametrize("case", G["falsta"], ids=[c["id"] for c in G["falsta"]])
def test_falsta_fixture_device(case):
    from ferromic_b200 import falsta
    if case["kind"] == "diversity":
        got = falsta.diversity_falsta_text(case["seqname"], case["start"], case["end"], _recs(case["per_site"]))
    else:
        got = falsta.fst_falsta_text(case["seqname"], case["start"], case["end"], _recs(case["wc"]), _recs(case["hudson"]))
    assert got.decode() == case["expect"]
