"""The committed golden fixtures (tests/golden/reference_goldens.json, written by
tests/golden/make_goldens.py from the reference's own tests) against (a) the CPU oracle and
(b) the CUDA path through the Python mirror of the reference API."""
import json
import math
import os

import numpy as np
import pytest

from oracle import pyoracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "reference_goldens.json")))["cases"]


def check(got, case, expect=None):
    exp = case["expect"] if expect is None else expect
    if exp == "nan":
        assert isinstance(got, float) and math.isnan(got), case["id"]
    elif "rel" in case:
        assert got == pytest.approx(exp, rel=case["rel"]), case["id"]
    elif "abs" in case:
        assert abs(got - exp) <= case["abs"], (case["id"], got, exp)
    else:
        assert got == exp, (case["id"], got, exp)


def tup(h):
    return [tuple(x) for x in h]


def run_case(case, api):
    """api: dict of callables implemented by the oracle or by the GPU package."""
    k = case["kind"]
    if k == "seg_sites":
        check(api["seg"](case["variants"]), case)
    elif k == "theta":
        check(api["theta"](*case["args"]), case)
    elif k == "pi":
        check(api["pi"](case["variants"], tup(case["haplotypes"]), case["L"]), case)
    elif k == "adjusted_length":
        s, e, allow, mask = case["args"]
        check(api["ladj"](s, e, None if allow is None else tup(allow), None if mask is None else tup(mask)), case)
    elif k == "hudson_sites":
        out_fst, sites = api["hudson"](case["variants"], tup(case["pop1"]), tup(case["pop2"]), tuple(case["region"]),
                                       case["L"])
        e = case["expect"]
        if "fst" in e:
            check(out_fst, case, e["fst"])
        for key, attr in (("site_fst", "fst"), ("site_num", "numerator_component"),
                          ("site_den", "denominator_component"), ("site_dxy", "d_xy"), ("site_pi1", "pi_pop1"),
                          ("site_pi2", "pi_pop2")):
            if key in e:
                assert len(sites) == len(e[key]), case["id"]
                for s, x in zip(sites, e[key]):
                    check(s[attr], case, x)
    elif k == "wc_site":
        a, b, fst = api["wc"](case["genotypes"], case["groups"])
        check(a, case, case["expect"]["a"])
        check(b, case, case["expect"]["b"])
        check(fst, case, case["expect"]["fst"])
    else:
        raise AssertionError(k)


# ------------------------------------------------------------------ (a) oracle, CPU
def _oracle_api():
    def hudson(variants, p1, p2, region, L):
        vs = orc.variants_from_python(variants)
        n = vs.n_samples
        rc, out, sites = orc.hudson_pair(orc.Pop(p1, vs, n, L), orc.Pop(p2, vs, n, L), region=region)
        assert rc == 0
        return out["fst"], sites

    def wc(gts, groups):
        vs = orc.variants_from_python([{"position": 10, "genotypes": gts}])
        g = np.array(groups, dtype=np.uint16)
        r = orc.wc_fst(vs, g, g, int(max(groups)) + 1, (0, 100))
        return float(r["a"][0]), float(r["b"][0]), r["overall"]["value"]

    return {"seg": lambda v: orc.count_segregating_sites(orc.variants_from_python(v)),
            "theta": orc.watterson_theta,
            "pi": lambda v, h, L: orc.pi_sparse(orc.variants_from_python(v), h, L),
            "ladj": orc.adjusted_sequence_length, "hudson": hudson, "wc": wc}


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_oracle_matches_reference_goldens(case):
    run_case(case, _oracle_api())


# ------------------------------------------------------------------ (b) CUDA path
def _gpu_api():
    import ferromic_b200 as F

    def hudson(variants, p1, p2, region, L):
        names = [f"s{i}" for i in range(len(variants[0]["genotypes"]))]
        a = {"id": 0, "haplotypes": p1, "variants": variants, "sequence_length": L, "sample_names": names}
        b = {"id": 1, "haplotypes": p2, "variants": variants, "sequence_length": L, "sample_names": names}
        out, sites = F.hudson_fst_with_sites(a, b, region)
        return out.fst, [{k: getattr(s, k) for k in ("fst", "numerator_component", "denominator_component", "d_xy",
                                                     "pi_pop1", "pi_pop2")} for s in sites]

    def wc(gts, groups):
        names = [f"s{i}" for i in range(len(gts))]
        res = F.wc_fst([{"position": 10, "genotypes": gts}], names,
                       {n: (g, g) for n, g in zip(names, groups)}, (0, 100))
        s = res.site_fst[0]
        return s.variance_components_a, s.variance_components_b, res.overall_fst.value

    return {"seg": F.segregating_sites, "theta": F.watterson_theta, "pi": F.nucleotide_diversity,
            "ladj": F.adjusted_sequence_length, "hudson": hudson, "wc": wc}


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_gpu_path_matches_reference_goldens(case):
    run_case(case, _gpu_api())
