"""The reference's own Python tests (src/pytests/test_ferromic.py, test_diversity_integration.py,
test_hudson_fst_integration.py and the equivalence checks of
src/pybenches/test_population_statistics_benchmarks.py:509-606) pointed at the GPU package:
`import ferromic_b200 as fm`.  scikit-allel is replaced by tests/allel_formulas.py (a numpy
restatement of the same published formulas)."""
import copy
import math

import numpy as np
import pytest

from tests import allel_formulas as allel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fm():
    import ferromic_b200
    return ferromic_b200


def build_variant(position, genotypes):
    return {"position": position, "genotypes": genotypes}


# ---- src/pytests/test_ferromic.py
def test_segregating_sites_counts_polymorphic_sites(fm):
    variants = [build_variant(100, [[0, 0], [0, 1]]), build_variant(150, [[0, 0], [0, 0]]),
                build_variant(200, [[0, 1], [1, 1]])]
    assert fm.segregating_sites(variants) == 2


def test_watterson_theta(fm):
    assert math.isclose(fm.watterson_theta(3, 4, 100), 3 / (1 + 1 / 2 + 1 / 3) / 100, rel_tol=1e-12)
    with pytest.raises(ValueError) as excinfo:
        fm.watterson_theta(1, 1, 100)
    assert "sample_count" in str(excinfo.value)


def test_adjusted_sequence_length(fm):
    # the reference's pytest expects 25 but is stale; the current Rust code yields 24
    assert fm.adjusted_sequence_length(1, 100, allow=[(11, 20), (40, 60)], mask=[(45, 50)]) == 24
    assert fm.adjusted_sequence_length(100, 200, mask=[(100, 101)]) == 100  # stats_tests.rs:1829-1858


def test_population_rejects_non_positive_sequence_length(fm):
    with pytest.raises(ValueError) as excinfo:
        fm.Population("demo", [], [], 0)
    assert "sequence_length" in str(excinfo.value)


def test_inversion_allele_frequency(fm):
    assert fm.inversion_allele_frequency({"sampleA": (0, 1), "sampleB": (1, 1), "sampleC": (2, 255)}) == \
        pytest.approx(0.75)


def test_population_from_numpy_accepts_python_positions(fm):
    genotypes = np.array([[[0, 0], [0, 1]]], dtype=np.uint8)
    population = fm.Population.from_numpy("demo", genotypes=genotypes, positions=[101],
                                          haplotypes=[(0, 0), (0, 1)], sequence_length=500,
                                          sample_names=["sampleA", "sampleB"])
    assert population.variant_count == 1
    assert population.sample_names == ["sampleA", "sampleB"]
    assert population.haplotypes == [(0, 0), (0, 1)]


# ---- src/pytests/test_diversity_integration.py
SAMPLE_NAMES = ["pop1_individual_1", "pop1_individual_2", "pop2_individual_1", "pop2_individual_2"]
POP1, POP2 = [0, 1], [2, 3]


def build_haplotypes(samples):
    return [(s, side) for s in samples for side in (0, 1)]


def diversity_variants():
    return [build_variant(0, [[0, 0], [0, 1], [1, 1], [1, 1]]), build_variant(3, [[0, 1], [0, 0], [0, 1], [0, 0]]),
            build_variant(5, [[0, 0], [0, 1], [0, 1], [1, 1]]), build_variant(7, [[0, 1], [1, 1], None, [0, 1]])]


def genotype_array(variants):
    return np.array([[[-1, -1] if g is None else list(g) for g in v["genotypes"]] for v in variants], dtype=np.int16)


def build_population(pop_id, samples, variants, L):
    return {"id": pop_id, "haplotypes": build_haplotypes(samples), "variants": copy.deepcopy(variants),
            "sequence_length": L, "sample_names": SAMPLE_NAMES}


def test_diversity_integration(fm):
    variants = diversity_variants()
    g = genotype_array(variants)
    L = 10
    ac = {"pop1": allel.count_alleles(g, POP1, 1), "pop2": allel.count_alleles(g, POP2, 1),
          "combined": allel.count_alleles(g, None, 1)}
    haps = {"pop1": build_haplotypes(POP1), "pop2": build_haplotypes(POP2), "combined": build_haplotypes(POP1 + POP2)}
    for k in ac:
        expected = float(np.nansum(allel.mean_pairwise_difference(ac[k])) / L)
        assert fm.nucleotide_diversity(variants, haps[k], L) == pytest.approx(expected, rel=1e-12)
    sites = fm.per_site_diversity(variants, haps["pop1"], (0, L - 1))
    by_pos = {s.position: s for s in sites}
    per_variant = np.nan_to_num(allel.mean_pairwise_difference(ac["pop1"]), nan=0.0)
    for v, exp in zip(variants, per_variant):
        assert by_pos[v["position"] + 1].pi == pytest.approx(exp, rel=1e-12)
    result = fm.hudson_dxy(build_population("pop1", POP1, variants, L), build_population("pop2", POP2, variants, L))
    expected = float(np.nansum(allel.mean_pairwise_difference_between(ac["pop1"], ac["pop2"])) / L)
    assert result.d_xy == pytest.approx(expected, rel=1e-12)


# ---- src/pytests/test_hudson_fst_integration.py
def test_hudson_integration(fm):
    variants = [build_variant(0, [[0, 0], [0, 0], [1, 1], [1, 1]]), build_variant(1, [[0, 1], [0, 0], [0, 1], [0, 1]]),
                build_variant(2, [[0, 0], [0, 1], [0, 1], [1, 1]])]
    g = np.array([v["genotypes"] for v in variants])
    num, den = allel.hudson_fst(allel.count_alleles(g, POP1), allel.count_alleles(g, POP2))
    p1, p2 = build_population("pop1", POP1, variants, 3), build_population("pop2", POP2, variants, 3)
    result = fm.hudson_fst(p1, p2)
    assert result.fst == pytest.approx(float(num.sum() / den.sum()), rel=1e-12)
    assert result.d_xy == pytest.approx(float(den.sum() / 3), rel=1e-12)
    result, sites = fm.hudson_fst_with_sites(p1, p2, (0, 2))
    assert result.fst == pytest.approx(float(num.sum() / den.sum()), rel=1e-12)
    informative = [s for s in sites if s.numerator_component is not None and s.denominator_component is not None]
    assert len(informative) == 3
    for i, s in enumerate(informative):
        assert s.position == i + 1
        assert s.numerator_component == pytest.approx(float(num[i]), rel=1e-12)
        assert s.denominator_component == pytest.approx(float(den[i]), rel=1e-12)
        assert s.fst == pytest.approx(float(num[i] / den[i]), rel=1e-12)


# ---- Rust goldens through the GPU API (src/tests/hudson_fst_tests.rs:363-665, 877-1006)
def test_rust_hudson_goldens(fm):
    names = ["sample0", "sample1", "sample2", "sample3"]

    def pops(variants, L):
        return ({"id": 0, "haplotypes": build_haplotypes(POP1), "variants": variants, "sequence_length": L,
                 "sample_names": names},
                {"id": 1, "haplotypes": build_haplotypes(POP2), "variants": variants, "sequence_length": L,
                 "sample_names": names})
    va = build_variant(100, [[0, 0], [0, 0], [1, 1], [1, 1]])
    out, sites = fm.hudson_fst_with_sites(*pops([va, build_variant(200, [[0, 1]] * 4)], 2), (100, 200))
    assert abs(out.fst - 5 / 9) < 1e-12
    assert abs(sites[1].fst + 1 / 3) < 1e-12 and abs(sites[1].numerator_component + 1 / 6) < 1e-12
    out, sites = fm.hudson_fst_with_sites(*pops([va, build_variant(200, [None, [0, 1], None, [0, 1]])], 2), (100, 200))
    assert abs(out.fst - 1 / 3) < 1e-12 and abs(sites[1].fst + 1.0) < 1e-12
    assert sites[1].n1_called == 2 and sites[1].n2_called == 2
    assert out.population1_label == "haplotype_group_0" and out.population2_haplotype_group == 1


# ---- src/pybenches/test_population_statistics_benchmarks.py:113-261, 509-606 (abs 1e-12 gates)
@pytest.mark.parametrize("variant_count,sample_count,scale", [(512, 48, 0.02), (4096, 96, 0.05), (16384, 128, 0.08)])
def test_benchmark_panels_equivalence(fm, variant_count, sample_count, scale):
    rng = np.random.default_rng(seed=variant_count + sample_count)
    half = sample_count // 2
    base = rng.beta(0.8, 0.8, size=variant_count)
    div = rng.normal(0.0, scale, size=variant_count)
    f1, f2 = np.clip(base + div, 0.001, 0.999), np.clip(base - div, 0.001, 0.999)
    h1 = rng.binomial(1, f1[:, None], size=(variant_count, half * 2)).astype(np.int8)
    h2 = rng.binomial(1, f2[:, None], size=(variant_count, half * 2)).astype(np.int8)
    genotypes = np.concatenate([h1.reshape(variant_count, half, 2), h2.reshape(variant_count, half, 2)], axis=1)
    genotypes[0, :half, :] = 0
    genotypes[0, half:, :] = 1
    genotypes[1, :half, 0] = 0
    genotypes[1, :half, 1] = 1
    genotypes[1, half:, :] = 1
    positions = np.cumsum(rng.integers(1, 50, size=variant_count, dtype=np.int64), dtype=np.int64)
    L = int(positions[-1]) + 1 - int(positions[0])
    haplotypes = [(s, side) for s in range(sample_count) for side in (0, 1)]
    population = fm.Population.from_numpy("all_samples", genotypes, positions, haplotypes, L,
                                          sample_names=[f"sample_{i}" for i in range(sample_count)])
    ac = allel.count_alleles(genotypes, None, 2)
    ac1 = allel.count_alleles(genotypes, range(half), 2)
    ac2 = allel.count_alleles(genotypes, range(half, sample_count), 2)
    seg = int(((ac > 0).sum(axis=1) > 1).sum())
    assert population.segregating_sites() == seg
    pi = float(np.nansum(allel.mean_pairwise_difference(ac)) / L)
    assert math.isclose(population.nucleotide_diversity(), pi, rel_tol=0.0, abs_tol=1e-12)
    a1 = sum(1.0 / i for i in range(1, sample_count * 2))
    assert math.isclose(fm.watterson_theta(seg, sample_count * 2, L), seg / a1 / L, rel_tol=0.0, abs_tol=1e-12)
    lookup1 = {(s, side) for s in range(half) for side in (0, 1)}
    pop1 = population.with_haplotypes("population_1", [h for h in haplotypes if h in lookup1])
    pop2 = population.with_haplotypes("population_2", [h for h in haplotypes if h not in lookup1])
    num, den = allel.hudson_fst(ac1, ac2)
    result = fm.hudson_fst(pop1, pop2)
    assert math.isclose(result.fst, float(num.sum() / den.sum()), rel_tol=0.0, abs_tol=1e-12)
    assert math.isclose(result.d_xy, float(den.sum() / L), rel_tol=0.0, abs_tol=1e-12)
