"""Edge cases the reference's tests and data model call out: empty and
ragged inputs, haploid calls, single sites, unsorted positions, empty groups, out-of-range
haplotypes -- CUDA path against the oracle."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.synth import both_sides
from tests.test_gpu_parity import assert_arrays_close, close

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("both_ingest_modes")]


def F():
    import ferromic_b200 as m
    return m


def test_empty_inputs():
    f = F()
    assert f.segregating_sites([]) == 0
    # calculate_pi with no variants: 0.0 (stats_tests.rs:520-539); < 2 haplotypes: NaN
    assert f.nucleotide_diversity([], [(0, 0), (0, 1)], 1000) == 0.0
    assert math.isnan(f.nucleotide_diversity([], [(0, 0)], 1000))
    assert f.per_site_diversity([], [(0, 0), (0, 1)], (0, 10)) == []
    names = ["a", "b"]
    pa = {"id": 0, "haplotypes": [(0, 0), (0, 1)], "variants": [], "sequence_length": 10, "sample_names": names}
    pb = {"id": 1, "haplotypes": [(1, 0), (1, 1)], "variants": [], "sequence_length": 10, "sample_names": names}
    out = f.hudson_fst(pa, pb)
    assert out.fst is None
    out, sites = f.hudson_fst_with_sites(pa, pb, (0, 9))
    assert out.fst is None and sites == []
    res = f.wc_fst([], names, {"a": (0, 0), "b": (1, 1)}, (0, 9))
    assert res.overall_fst.state == "insufficient_data_for_estimation" and res.overall_fst.sites == 0
    # a zero-row matrix through the C ABI
    from ferromic_b200 import _lib
    from ferromic_b200.api import _Matrix
    m = _Matrix(np.zeros((0, 3, 2), dtype=np.uint8), None, np.zeros(0, dtype=np.int64), max_allele=0)
    g = m.group(both_sides(range(3)))
    s = g.summary(True)
    assert s["segregating_sites"] == 0 and s["pi_sum"] == 0.0 and len(s["alt"]) == 0
    assert g.pi(100, _lib.FM_PI_SUMMARY) == 0.0


def test_ragged_and_haploid_genotypes_follow_from_variants_semantics():
    """ploidy = longest genotype over all variants; shorter genotypes leave trailing sides missing;
    a None sample contributes nothing (stats.rs:349-359, 452-460)."""
    f = F()
    variants = [{"position": 5, "genotypes": [[0], [1], [1, 0], None, [0, 1]]},
                {"position": 9, "genotypes": [[1], [1], [0, 0], [1, 1], None]},
                {"position": 12, "genotypes": [[0], [0], [0], [0], [0]]},
                {"position": 20, "genotypes": [None, None, None, None, [1]]}]
    vs = orc.variants_from_python(variants, 5)
    haps = both_sides(range(5))
    assert f.segregating_sites(variants) == orc.count_segregating_sites(vs)
    assert close(f.nucleotide_diversity(variants, haps, 30), orc.pi_sparse(vs, haps, 30), 1e-12)
    gp, gpi, gth = f.per_site_diversity_arrays(variants, haps, (0, 29))
    rp, rpi, rth = orc.per_site_diversity(vs, haps, (0, 29))
    assert np.array_equal(gp, rp)
    assert_arrays_close(gpi, rpi, 1e-12)
    assert_arrays_close(gth, rth, 1e-12)
    names = [f"s{i}" for i in range(5)]
    h1, h2 = both_sides([0, 1, 2]), both_sides([3, 4])
    a = {"id": 0, "haplotypes": h1, "variants": variants, "sequence_length": 30, "sample_names": names}
    b = {"id": 1, "haplotypes": h2, "variants": variants, "sequence_length": 30, "sample_names": names}
    rc, ref, rsites = orc.hudson_pair(orc.Pop(h1, vs, 5, 30), orc.Pop(h2, vs, 5, 30), region=(0, 29))
    out, sites = f.hudson_fst_with_sites(a, b, (0, 29))
    assert rc == 0 and len(sites) == len(rsites)
    for k in ("fst", "d_xy", "pi_pop1", "pi_pop2"):
        assert close(getattr(out, k), ref[k], 1e-12), k
    for s, r in zip(sites, rsites):
        assert s.n1_called == r["n1_called"] and s.n2_called == r["n2_called"]
        assert close(s.fst, r["fst"], 1e-12) and close(s.d_xy, r["d_xy"], 1e-12)
    left = np.array([0, 0, 0, 1, 1], dtype=np.uint16)
    ref = orc.wc_fst(vs, left, left, 2, (0, 29))
    got = f.wc_fst(variants, names, {n: (int(g), int(g)) for n, g in zip(names, left)}, (0, 29))
    assert got.overall_fst.sites == ref["overall"]["sites"]
    assert close(got.overall_fst.sum_a, ref["overall"]["sum_a"], 1e-12)
    assert close(got.overall_fst.sum_b, ref["overall"]["sum_b"], 1e-12)
    assert [s.overall_fst.state for s in got.site_fst] == [orc.STATE_NAMES[x] for x in ref["state"]]


@pytest.mark.parametrize("V", [1, 31, 32, 33, 8191, 8193])
def test_site_counts_around_batch_and_superbatch_boundaries(V):
    from ferromic_b200.api import _Matrix
    rng = np.random.default_rng(V)
    g = rng.integers(0, 2, size=(V, 9, 2)).astype(np.int8)
    g[rng.random(g.shape) < 0.1] = -1
    pos = np.arange(V, dtype=np.int64) * 3 + 7
    vs, d = orc.from_numpy(g, pos)
    m = _Matrix(np.where(g < 0, 0, g).astype(np.uint8), g < 0, pos, max_allele=1)
    for haps in (both_sides(range(9)), [(0, 0)], [], [(4, 1), (4, 1), (100, 0)]):
        ref = orc.build_summary(d, haps)
        got = m.group(haps).summary(True)
        assert np.array_equal(got["alt"], ref.alt) and np.array_equal(got["called"], ref.called)
        assert got["segregating_sites"] == ref.seg and close(got["pi_sum"], ref.pi_sum, 1e-12)


def test_unsorted_positions_are_rejected_for_region_queries_not_for_summaries():
    from ferromic_b200 import _lib
    from ferromic_b200.api import _Matrix
    rng = np.random.default_rng(0)
    g = rng.integers(0, 2, size=(50, 4, 2)).astype(np.uint8)
    pos = rng.permutation(50).astype(np.int64)
    m = _Matrix(g, None, pos, max_allele=1)
    grp = m.group(both_sides(range(4)))
    assert grp.summary()["segregating_sites"] >= 0  # whole-matrix statistics do not need an order
    out = np.zeros(50)
    n = C.c_size_t()
    st = _lib.lib().fm_per_site_diversity(grp.handle, 8, 0, 49, None, 0, None, 0,
                                          np.zeros(50, dtype=np.int64).ctypes.data, out.ctypes.data, out.ctypes.data,
                                          50, C.byref(n))
    assert st == _lib.FM_ERR_UNSUPPORTED and b"sorted" in _lib.lib().fm_last_error()


def test_c_abi_argument_errors_do_not_crash():
    from ferromic_b200 import _lib
    L = _lib.lib()
    assert L.fm_group_summary(None, None, None, None, None, None) == _lib.FM_ERR_INVALID_ARG
    assert L.fm_matrix_release(None) == 0 and L.fm_group_release(None) == 0 and L.fm_partition_release(None) == 0
    h = C.c_void_p()
    assert L.fm_matrix_create(None, None, 5, 5, 2, 1, None, C.byref(h)) == _lib.FM_ERR_INVALID_ARG
    assert L.fm_ingest_rows(None, None, None, 0, 0) == _lib.FM_ERR_INVALID_ARG
    ih = C.c_void_p()
    _lib.check(L.fm_ingest_begin(10, 2, 2, 0, 1, None, 0, C.byref(ih)))
    data = np.zeros(10 * 4, dtype=np.uint8)
    assert L.fm_ingest_rows(ih, data.ctypes.data, None, 8, 5) == _lib.FM_ERR_INVALID_ARG  # rows beyond V
    mh = C.c_void_p()
    assert L.fm_ingest_finish(ih, C.byref(mh), None, None) == _lib.FM_ERR_INVALID_ARG     # not all rows pushed
    assert L.fm_ingest_abort(ih) == 0


@pytest.mark.parametrize("S,missing", [(37, True), (37, False), (64, True), (200, True)])
def test_groups_created_together_equal_groups_created_alone(S, missing):
    """fm_groups_create (one pass over the u8 rows, compress plans) against fm_group_create per group and the
    oracle's counts: row strides that are / are not multiples of 16 (both row-packing paths), overlapping,
    empty and single-haplotype groups."""
    from ferromic_b200.api import _Matrix
    rng = np.random.default_rng(S)
    V = 300
    g = rng.binomial(1, rng.beta(0.6, 0.6, size=V)[:, None, None], size=(V, S, 2)).astype(np.uint8)
    miss = (rng.random(g.shape) < 0.07) if missing else None
    pos = np.arange(V, dtype=np.int64) * 3
    lists = [[(s, int(rng.integers(0, 2))) for s in range(S)], both_sides(range(0, S, 3)), [], [(S - 1, 1)],
             both_sides(range(S // 2, S))]
    m1 = _Matrix(g, miss, pos, max_allele=1)
    together = m1.groups(lists)
    m2 = _Matrix(g, miss, pos, max_allele=1)
    d = orc.Dense(g.reshape(-1), None if miss is None else orc.pack_missing_bits(miss.reshape(-1)), V, S, 2, 1)
    for haps, gt in zip(lists, together):
        a, b = gt.summary(True), m2.group(haps).summary(True)
        so = orc.build_summary(d, haps)
        assert np.array_equal(a["alt"], b["alt"]) and np.array_equal(a["called"], b["called"])
        assert np.array_equal(a["alt"], so.alt) and np.array_equal(a["called"], so.called)
        assert a["segregating_sites"] == b["segregating_sites"] == so.seg

