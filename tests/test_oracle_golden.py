"""Pins the CPU oracle (oracle/ferromic_oracle.c) against every golden vector the reference's
own tests hold for the per-site estimator path.  Each test cites the reference test it
transcribes (paths relative to /root/reference/src)."""
import math

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests import allel_formulas as allel

L, R = 0, 1


def V(pos, gts):
    return {"position": pos, "genotypes": gts}


def both_sides(samples):
    return [(s, side) for s in samples for side in (0, 1)]


# ------------------------------------------------------------------ segregating sites
def test_seg_sites_goldens():
    # tests/stats_tests.rs:240-272
    vs = orc.variants_from_python([
        V(1, [[0, 0], [0, 1], [1, 1]]), V(2, [[0, 0], [0, 0], [0, 0]]),
        V(3, [[0, 1], [0, 1], [0, 1]]), V(4, [[0, 0], [1, 1], [0, 1]])])
    assert orc.count_segregating_sites(vs) == 3
    assert orc.count_segregating_sites(orc.variants_from_python([], n_samples=0)) == 0
    vs = orc.variants_from_python([V(1, [[0, 0]] * 3), V(2, [[1, 1]] * 3)])
    assert orc.count_segregating_sites(vs) == 0
    vs = orc.variants_from_python([V(1, [[0, 0], None, [1, 1]]), V(2, [[0, 1], [0, 1], None])])
    assert orc.count_segregating_sites(vs) == 2
    # pytests/test_ferromic.py:14-25
    vs = orc.variants_from_python([V(100, [[0, 0], [0, 1]]), V(150, [[0, 0], [0, 0]]),
                                   V(200, [[0, 1], [1, 1]])])
    assert orc.count_segregating_sites(vs) == 2


@pytest.mark.parametrize("dense", [False, True])
def test_seg_sites_population_dense_sparse_parity(dense):
    # tests/stats_tests.rs:35-80
    pop_a = [(0, L), (0, R)]
    for gts, expected in (([[0, 0], [1, 1]], 0), ([[0, 1], [1, 1]], 1)):
        vs = orc.variants_from_python([V(100, gts)])
        d = orc.dense_from_variants(vs, 2) if dense else None
        p = orc.Pop(pop_a, vs, 2, 1, dense=d)
        assert orc.count_segregating_sites_for_population(p) == expected


# ------------------------------------------------------------------ theta / harmonic
def test_harmonic_goldens():
    # tests/stats_tests.rs:346-365
    assert orc.harmonic(1) == 1.0
    assert abs(orc.harmonic(2) - 1.5) < 1e-10
    assert abs(orc.harmonic(3) - (1.0 + 0.5 + 1.0 / 3.0)) < 1e-10
    assert abs(orc.harmonic(10) - 2.9289682539682538) < 1e-10


def test_watterson_theta_goldens():
    # tests/stats_tests.rs:473-506
    assert abs(orc.watterson_theta(10, 5, 1000) - 0.0048) < 1e-6
    assert abs(orc.watterson_theta(5, 2, 1000) - 0.005) < 1e-6
    assert abs(orc.watterson_theta(100, 10, 1_000_000) - 0.00003534) < 1e-6
    assert math.isinf(orc.watterson_theta(100, 1, 1000))
    assert math.isnan(orc.watterson_theta(0, 0, 1000))
    assert math.isinf(orc.watterson_theta(10, 5, 0))
    # tests/stats_tests.rs:1360-1420 (S=2, n=4, L=100)
    assert abs(orc.watterson_theta(2, 4, 100) - 12.0 / 11.0 / 100.0) < 1e-10
    # pytests/test_ferromic.py:27-33
    assert math.isclose(orc.watterson_theta(3, 4, 100), 3 / (1 + 1 / 2 + 1 / 3) / 100, rel_tol=1e-12)


# ------------------------------------------------------------------ pi
def test_pi_goldens():
    four = both_sides([0, 1])
    # tests/stats_tests.rs:520-539
    vs = orc.variants_from_python([V(100, [[0, 0], [0, 0]]), V(200, [[1, 1], [1, 1]])])
    assert orc.pi_sparse(vs, four, 1000) == 0.0
    assert orc.pi_sparse(orc.variants_from_python([], n_samples=0), [(0, L), (0, R)], 1000) == 0.0
    # tests/stats_tests.rs:607-623: uncallable site leaves the denominator
    vs = orc.variants_from_python([V(10, [[0, 0], [1, 1]]), V(20, [None, None])])
    assert abs(orc.pi_sparse(vs, four, 2) - 2.0 / 3.0) < 1e-9
    # tests/stats_tests.rs:651-665
    vs = orc.variants_from_python([V(100, [[0, 1]])])
    assert math.isnan(orc.pi_sparse(vs, [(0, L)], 1000))
    assert math.isnan(orc.pi_sparse(vs, [], 1000))
    # tests/stats_tests.rs:641-648
    vs = orc.variants_from_python([V(100, [[0, 1], [1, 0]])])
    pi = orc.pi_sparse(vs, four, 1_000_000_000)
    assert 0.0 < pi < 0.001


def _diversity_panel():
    # pytests/test_diversity_integration.py:27-68
    return [V(0, [[0, 0], [0, 1], [1, 1], [1, 1]]), V(3, [[0, 1], [0, 0], [0, 1], [0, 0]]),
            V(5, [[0, 0], [0, 1], [0, 1], [1, 1]]), V(7, [[0, 1], [1, 1], None, [0, 1]])]


def _panel_array(variants):
    return np.array([[[-1, -1] if g is None else g for g in v["genotypes"]] for v in variants],
                    dtype=np.int16)


def test_diversity_panel_matches_allel_formulas():
    # pytests/test_diversity_integration.py:130-213 (rel 1e-12 gates)
    variants = _diversity_panel()
    vs = orc.variants_from_python(variants)
    g = _panel_array(variants)
    Lseq = 10
    for samples in ([0, 1], [2, 3], [0, 1, 2, 3]):
        ac = allel.count_alleles(g, subpop=samples, max_allele=1)
        expected = float(np.nansum(allel.mean_pairwise_difference(ac)) / Lseq)
        assert orc.pi_sparse(vs, both_sides(samples), Lseq) == pytest.approx(expected, rel=1e-12)
    pos, pi, _ = orc.per_site_diversity(vs, both_sides([0, 1]), (0, Lseq - 1))
    ac1 = allel.count_alleles(g, subpop=[0, 1], max_allele=1)
    exp = np.nan_to_num(allel.mean_pairwise_difference(ac1), nan=0.0)
    assert list(pos) == [v["position"] + 1 for v in variants]
    assert pi == pytest.approx(exp, rel=1e-12)
    # hudson_dxy
    ac2 = allel.count_alleles(g, subpop=[2, 3], max_allele=1)
    p1 = orc.Pop(both_sides([0, 1]), vs, 4, Lseq)
    p2 = orc.Pop(both_sides([2, 3]), vs, 4, Lseq)
    rc, dxy = orc.dxy_hudson(p1, p2)
    exp_dxy = float(np.nansum(allel.mean_pairwise_difference_between(ac1, ac2)) / Lseq)
    assert rc == 0 and dxy == pytest.approx(exp_dxy, rel=1e-12)


# ------------------------------------------------------------------ Hudson
def _two_pops(vs, Lseq, **kw):
    return (orc.Pop(both_sides([0, 1]), vs, 4, Lseq, **kw), orc.Pop(both_sides([2, 3]), vs, 4, Lseq, **kw))


def test_hudson_ratio_of_sums_no_missingness():
    # tests/hudson_fst_tests.rs:363-513
    vs = orc.variants_from_python([V(100, [[0, 0], [0, 0], [1, 1], [1, 1]]), V(200, [[0, 1]] * 4)])
    p1, p2 = _two_pops(vs, 2)
    rc, out, sites = orc.hudson_pair(p1, p2, region=(100, 200))
    assert rc == 0 and len(sites) == 2
    a, b = sites
    assert a["position"] == 101 and b["position"] == 201
    assert abs(a["fst"] - 1.0) < 1e-12 and abs(a["numerator_component"] - 1.0) < 1e-12
    assert abs(a["denominator_component"] - 1.0) < 1e-12
    assert abs(b["fst"] + 1.0 / 3.0) < 1e-12 and abs(b["numerator_component"] + 1.0 / 6.0) < 1e-12
    assert abs(b["denominator_component"] - 0.5) < 1e-12
    assert abs(out["fst"] - 5.0 / 9.0) < 1e-12


def test_hudson_ratio_of_sums_uneven_missingness():
    # tests/hudson_fst_tests.rs:516-665
    vs = orc.variants_from_python([V(100, [[0, 0], [0, 0], [1, 1], [1, 1]]),
                                   V(200, [None, [0, 1], None, [0, 1]])])
    p1, p2 = _two_pops(vs, 2)
    rc, out, sites = orc.hudson_pair(p1, p2, region=(100, 200))
    assert rc == 0
    b = sites[1]
    assert abs(b["fst"] + 1.0) < 1e-12 and abs(b["numerator_component"] + 0.5) < 1e-12
    assert abs(b["denominator_component"] - 0.5) < 1e-12
    assert abs(out["fst"] - 1.0 / 3.0) < 1e-12


def test_hudson_multi_allelic_site():
    # tests/hudson_fst_tests.rs:877-1006
    vs = orc.variants_from_python([V(100, [[0, 0], [1, 2], [0, 1], [2, 2]])])
    p1, p2 = _two_pops(vs, 1)
    rc, _, sites = orc.hudson_pair(p1, p2, region=(100, 100))
    s = sites[0]
    exp_pi = (4.0 / 3.0) * 0.625
    exp_num = 0.6875 - 0.5 * (exp_pi + exp_pi)
    assert abs(s["d_xy"] - 0.6875) < 1e-12
    assert abs(s["pi_pop1"] - exp_pi) < 1e-12 and abs(s["pi_pop2"] - exp_pi) < 1e-12
    assert abs(s["fst"] - exp_num / 0.6875) < 1e-12


def test_hudson_no_variants_and_incompatible():
    # tests/hudson_fst_tests.rs:301-360, 668-745: no variants -> fst None, no sites
    empty = orc.variants_from_python([], n_samples=4)
    p1, p2 = _two_pops(empty, 1000)
    rc, out, sites = orc.hudson_pair(p1, p2, region=(0, 999))
    assert rc == 0 and out["fst"] is None and sites == []
    # tests/hudson_fst_tests.rs:1103-1188: incompatible variants -> empty / Err
    g = [[0, 0], [0, 1], [1, 1], [1, 0]]
    v1 = orc.variants_from_python([V(100, g)])
    v2 = orc.variants_from_python([V(200, g)])
    q1 = orc.Pop(both_sides([0, 1]), v1, 4, 2)
    q2 = orc.Pop(both_sides([2, 3]), v2, 4, 2)
    assert orc.hudson_per_site(q1, q2, (100, 200)) == []
    rc, _, _ = orc.hudson_pair(q1, q2, region=(100, 200))
    assert rc != 0


def test_hudson_per_site_sum_over_length_consistency():
    # tests/hudson_fst_tests.rs:747-874: d_xy outcome == sum(per-site dxy)/L on a sparse context
    vs = orc.variants_from_python([V(10, [[0, 0], [0, 1], [1, 1], [1, 0]]),
                                   V(20, [[0, 1], [0, 1], [0, 0], [1, 1]]),
                                   V(30, [[1, 1], [0, 1], [0, 0], [0, 1]])])
    p1, p2 = _two_pops(vs, 50)
    rc, out, sites = orc.hudson_pair(p1, p2, region=(0, 49))
    assert rc == 0 and len(sites) == 3
    assert out["d_xy"] == pytest.approx(sum(s["d_xy"] for s in sites) / 50, rel=1e-12)
    assert out["pi_pop1"] == pytest.approx(sum(s["pi_pop1"] for s in sites) / 50, rel=1e-12)


def test_hudson_dxy_from_summaries():
    # tests/hudson_fst_tests.rs:1271-1418
    s1 = orc.Summary([0, 1, 0, 0], [2, 2, 2, 2], 2, 1, 1.0)
    s2 = orc.Summary([2, 1, 0, 0], [2, 2, 0, 0], 2, 1, 1.0)
    haps = [(0, L), (0, R)]
    p1 = orc.Pop(haps, None, 1, 4, summary=s1)
    p2 = orc.Pop(haps, None, 1, 4, summary=s2)
    rc, out, _ = orc.hudson_pair(p1, p2)
    assert rc == 0 and abs(out["d_xy"] - 0.75) < 1e-12
    rc, dxy = orc.dxy_hudson(p1, p2)
    assert rc == 0 and abs(dxy - 0.75) < 1e-12
    s1 = orc.Summary([0, 1], [2, 2], 1, 1, 1.0)
    s2 = orc.Summary([0, 0], [0, 0], 0, 0, 1.0)
    p1 = orc.Pop(haps, None, 1, 2, summary=s1)
    p2 = orc.Pop(haps, None, 1, 2, summary=s2)
    rc, dxy = orc.dxy_hudson(p1, p2)
    assert rc == 0 and dxy is None


def test_hudson_matches_allel_formulas():
    # pytests/test_hudson_fst_integration.py:20-152 (rel 1e-12 gates)
    variants = [V(0, [[0, 0], [0, 0], [1, 1], [1, 1]]), V(1, [[0, 1], [0, 0], [0, 1], [0, 1]]),
                V(2, [[0, 0], [0, 1], [0, 1], [1, 1]])]
    g = np.array([v["genotypes"] for v in variants])
    num, den = allel.hudson_fst(allel.count_alleles(g, [0, 1]), allel.count_alleles(g, [2, 3]))
    vs = orc.variants_from_python(variants)
    p1, p2 = _two_pops(vs, 3)
    rc, out, _ = orc.hudson_pair(p1, p2)
    assert out["fst"] == pytest.approx(float(num.sum() / den.sum()), rel=1e-12)
    assert out["d_xy"] == pytest.approx(float(den.sum() / 3), rel=1e-12)
    rc, out, sites = orc.hudson_pair(p1, p2, region=(0, 2))
    assert out["fst"] == pytest.approx(float(num.sum() / den.sum()), rel=1e-12)
    for i, s in enumerate(sites):
        assert s["position"] == i + 1
        assert s["numerator_component"] == pytest.approx(float(num[i]), rel=1e-12)
        assert s["denominator_component"] == pytest.approx(float(den[i]), rel=1e-12)
        assert s["fst"] == pytest.approx(float(num[i] / den[i]), rel=1e-12)


def test_hudson_falsta_track_values():
    # tests/stats_tests.rs:1860-2034 asserts per-site Hudson tracks fst=[1,-1,1], num=[1,-.5,1],
    # den=[1,.5,1]; the same three site patterns (perfect structure / one haplotype pair each /
    # perfect structure) evaluated by the sparse per-site path.
    vs = orc.variants_from_python([V(1, [[0, 0], [1, 1]]), V(2, [[0, 1], [0, 1]]), V(3, [[1, 1], [0, 0]])])
    p1 = orc.Pop(both_sides([0]), vs, 2, 3)
    p2 = orc.Pop(both_sides([1]), vs, 2, 3)
    rc, _, sites = orc.hudson_pair(p1, p2, region=(0, 10))
    assert [s["fst"] for s in sites] == pytest.approx([1.0, -1.0, 1.0], abs=1e-12)
    assert [s["numerator_component"] for s in sites] == pytest.approx([1.0, -0.5, 1.0], abs=1e-12)
    assert [s["denominator_component"] for s in sites] == pytest.approx([1.0, 0.5, 1.0], abs=1e-12)


# ------------------------------------------------------------------ adjusted length
def test_adjusted_sequence_length_goldens():
    # tests/stats_tests.rs:1829-1858
    assert orc.adjusted_sequence_length(100, 200, None, [(100, 101)]) == 100
    # pytests/test_ferromic.py:49-60 is stale (expects 25); current Rust yields 24
    assert orc.adjusted_sequence_length(1, 100, [(11, 20), (40, 60)], [(45, 50)]) == 24
    assert orc.adjusted_sequence_length(1, 100) == 100


# ------------------------------------------------------------------ Weir & Cockerham
# No reference test pins W&C numbers ("parity unpinned"); these KATs were derived by hand from
# stats.rs:1814-2127 and guard the restatement against regressions.
def _wc(gts, left, right, G=2):
    vs = orc.variants_from_python([V(10, gts)])
    return orc.wc_fst(vs, np.array(left), np.array(right), G, (0, 100))


def test_wc_derived_kats():
    lr = [0, 0, 1, 1]
    r = _wc([[0, 0], [0, 0], [1, 1], [1, 1]], lr, lr)
    assert r["a"][0] == pytest.approx(1.0, abs=1e-12) and r["b"][0] == pytest.approx(0.0, abs=1e-12)
    assert r["overall"]["value"] == pytest.approx(1.0, abs=1e-12)
    r = _wc([[0, 1]] * 4, lr, lr)
    assert r["a"][0] == pytest.approx(-1.0 / 6.0, abs=1e-12) and r["b"][0] == pytest.approx(2.0 / 3.0, abs=1e-12)
    assert r["overall"]["value"] == pytest.approx(-1.0 / 3.0, abs=1e-12)
    r = _wc([[0, 0], [0, 1], [1, 1], [0, 1]], lr, lr)
    assert r["a"][0] == pytest.approx(0.125, abs=1e-12) and r["b"][0] == pytest.approx(0.5, abs=1e-12)
    assert r["overall"]["value"] == pytest.approx(0.2, abs=1e-12)
    lr = [0, 0, 0, 1]
    r = _wc([[0, 0], [0, 1], [0, 0], [1, 1]], lr, lr)
    assert r["a"][0] == pytest.approx(0.601851851851852, abs=1e-12)
    assert r["b"][0] == pytest.approx(0.277777777777778, abs=1e-12)
    assert r["overall"]["value"] == pytest.approx(0.684210526315789, abs=1e-12)
    assert r["overall"]["state"] == "calculable" and r["overall"]["sites"] == 1
    assert list(r["pop_sizes"][0]) == [6, 2]


def test_wc_matches_published_weir_cockerham_1984():
    """Independent check of the restatement: the reference's (a, b) are Weir & Cockerham (1984) eqs. 2-3
    with h-bar = 0 summed over alleles (stats.rs:2033-2127 comments; `1 - c^2/(r-1)` with c^2 over r equals
    the published n_c / n-bar).  tests/allel_formulas.py evaluates the published form the way scikit-allel's
    weir_cockerham_fst does; unequal group sizes, missing calls, a third allele.  1e-9 relative."""
    from tests import allel_formulas as af
    rng = np.random.default_rng(1984)
    Vn, S, G = 400, 37, 3
    freq = rng.beta(0.7, 0.7, size=Vn)
    g = rng.binomial(1, freq[:, None, None], size=(Vn, S, 2)).astype(np.int64)
    g[(rng.random(g.shape) < 0.03) & (g == 1)] = 2  # some tri-allelic sites
    missing = rng.random((Vn, S)) < 0.08  # whole-sample missing calls (sparse semantics)
    label = np.array([0] * 9 + [1] * 17 + [2] * 11)
    variants = [V(10 + 3 * v, [None if missing[v, s] else [int(g[v, s, 0]), int(g[v, s, 1])] for s in range(S)])
                for v in range(Vn)]
    vs = orc.variants_from_python(variants)
    r = orc.wc_fst(vs, label, label, G, (0, 10 + 3 * Vn))
    assert r["n_sites"] == Vn
    acs = []
    for k in range(G):
        ac = np.zeros((Vn, 3))
        for al in range(3):
            ac[:, al] = ((g[:, label == k, :] == al) & ~missing[:, label == k, None]).sum(axis=(1, 2))
        acs.append(ac)
    a, b = af.weir_cockerham_ab_haploid(acs)
    scale = np.maximum(np.abs(a) + np.abs(b), 1e-300)
    assert np.max(np.abs(r["a"] - a) / scale) < 1e-9
    assert np.max(np.abs(r["b"] - b) / scale) < 1e-9
    assert r["overall"]["state"] == "calculable"
    assert r["overall"]["value"] == pytest.approx(a.sum() / (a.sum() + b.sum()), rel=1e-9)
    # every pair is the same estimator restricted to two populations
    k = 0
    for i in range(G):
        for j in range(i + 1, G):
            pa, pb = af.weir_cockerham_ab_haploid([acs[i], acs[j]])
            assert np.allclose(r["pair_a"][:, k], pa, rtol=1e-9, atol=1e-12)
            assert np.allclose(r["pair_b"][:, k], pb, rtol=1e-9, atol=1e-12)
            assert r["pairs"][k]["value"] == pytest.approx(pa.sum() / (pa.sum() + pb.sum()), rel=1e-9)
            k += 1


def test_wc_state_rules():
    lr = [0, 0, 1, 1]
    # all samples missing -> InsufficientData{sites_attempted: 1}, empty maps (stats.rs:1987-2001)
    r = _wc([None, None, None, None], lr, lr)
    assert r["state"][0] == 3 and r["has_maps"][0] == 0
    assert r["overall"]["state"] == "insufficient_data_for_estimation" and r["overall"]["sites"] == 1
    assert not r["pair_present"][0]
    # data in one group only -> NoInterPopulationVariance(0,0) but still informative
    r = _wc([[0, 1], [0, 0], None, None], lr, lr)
    assert r["state"][0] == 2 and r["overall"]["state"] == "no_inter_population_variance"
    assert r["overall"]["sites"] == 1 and r["pair_present"][0]
    assert r["pairs"][0]["state"] == "insufficient_data_for_estimation" and r["pairs"][0]["sites"] == 1
    # empty input (stats.rs:2152-2159)
    vs = orc.variants_from_python([], n_samples=4)
    r = orc.wc_fst(vs, np.array(lr), np.array(lr), 2, (0, 100))
    assert r["overall"]["state"] == "insufficient_data_for_estimation" and r["overall"]["sites"] == 0
    # threshold ladder (stats.rs:1785-1811)
    assert orc.fst_estimate_from_components(1e-13, -1e-13)["state"] == "no_inter_population_variance"
    assert orc.fst_estimate_from_components(-1.0, 0.5)["state"] == "components_yield_indeterminate_ratio"
    e = orc.fst_estimate_from_components(1.0, -1.0)
    assert e["state"] == "calculable" and math.isinf(e["value"])
