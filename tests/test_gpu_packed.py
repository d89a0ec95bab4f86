"""Packed (2 bits per genotype) ingest, SURVEY 8 f1: fm_pack_rows -> fm_matrix_create_packed /
fm_ingest_rows_packed -> K1p (groups compressed out of the resident packed rows) must give exactly the counts
of the u8 path and of the CPU oracle, for every missingness mode, ragged strides, streaming in several calls,
partitions declared at ingest time and groups created afterwards."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.synth import both_sides, make_cohort

pytestmark = pytest.mark.gpu


def _mats(g, pos, always_bitmap=False):
    from ferromic_b200.api import _Matrix
    miss = g < 0
    alle = np.where(miss, 0, g).astype(np.uint8)
    packed = _Matrix(alle, miss, pos, max_allele=1, always_bitmap=always_bitmap, ingest="packed-dense")
    u8 = _Matrix(alle, miss, pos, max_allele=1, always_bitmap=always_bitmap, ingest="u8")
    assert packed.ingest_mode == "packed-dense" and u8.ingest_mode == "u8"
    return packed, u8


def _same_summary(a, b):
    assert np.array_equal(a["alt"], b["alt"]) and np.array_equal(a["called"], b["called"])
    assert a["segregating_sites"] == b["segregating_sites"] and a["uncallable_lt2"] == b["uncallable_lt2"]
    assert a["pi_sum"] == b["pi_sum"]


@pytest.mark.parametrize("n_samples,missing", [(1, 0.0), (3, 0.2), (16, 0.1), (17, 0.3), (64, 0.0), (65, 0.05),
                                               (700, 0.02), (2504, 0.0), (2504, 0.01)])
def test_packed_counts_equal_u8_and_oracle(n_samples, missing):
    V = 3000 if n_samples < 1000 else 600
    g, pos, _ = make_cohort(V, n_samples, missing_rate=missing, seed=100 + n_samples)
    packed, u8 = _mats(g, pos)
    vs, d = orc.from_numpy(g, pos)
    rng = np.random.default_rng(n_samples)
    groups = [both_sides(range(n_samples)),
              [(int(s), int(rng.integers(0, 2))) for s in rng.choice(n_samples, max(1, n_samples // 2), replace=False)],
              both_sides(range(n_samples))[::3] + [(0, 0), (0, 0), (n_samples + 5, 1)],
              [(0, 1)], []]
    got_all = [x.summary(want_arrays=True) for x in packed.groups(groups)]  # one K1p launch for all five
    for haps, got in zip(groups, got_all):
        ref = orc.build_summary(d, haps)
        assert np.array_equal(got["alt"], ref.alt) and np.array_equal(got["called"], ref.called)
        assert got["segregating_sites"] == ref.seg
        _same_summary(got, u8.group(haps).summary(want_arrays=True))


def test_packed_wide_rows():
    """52,000 haplotypes: the packed row words (13 KB) take the large shared-memory slice of K1p and the plane pass
    runs in column-chunked mode."""
    g, pos, _ = make_cohort(60, 26000, missing_rate=0.01, seed=3)
    packed, _u8 = _mats(g[:, :, :], pos)
    vs, d = orc.from_numpy(g, pos)
    for haps in (both_sides(range(26000)), both_sides(range(0, 26000, 2))):
        ref = orc.build_summary(d, haps)
        got = packed.group(haps).summary(want_arrays=True)
        assert np.array_equal(got["alt"], ref.alt) and np.array_equal(got["called"], ref.called)
        assert got["segregating_sites"] == ref.seg


def test_packed_in_band_int8():
    from ferromic_b200.api import _Matrix
    g, pos, pops = make_cohort(2000, 37, missing_rate=0.15, seed=9)
    m = _Matrix.from_int8(g.astype(np.int8), pos, 1, ingest="packed")
    vs, d = orc.from_numpy(g, pos)
    for haps in (both_sides(range(37)), both_sides(pops[0])):
        ref = orc.build_summary(d, haps)
        got = m.group(haps).summary(want_arrays=True)
        assert np.array_equal(got["alt"], ref.alt) and np.array_equal(got["called"], ref.called)


@pytest.mark.parametrize("packed", [True, "library"])
@pytest.mark.parametrize("calls,missing", [(1, 0.1), (3, 0.1), (7, 0.0), (2, 0.3)])
def test_packed_streaming_ingest(calls, missing, packed):
    """fm_ingest_rows_packed (rows packed by the caller) / fm_ingest_rows_pack (u8 rows packed inside the call) in
    several calls, groups and a W&C partition declared up front; afterwards the packed rows are resident, so groups
    and partitions can still be created (Population.with_haplotypes, lib.rs:622)."""
    from ferromic_b200 import _lib
    from ferromic_b200.api import _Matrix
    g, pos, pops = make_cohort(1500, 45, n_pops=3, missing_rate=missing, seed=17 + calls)
    miss = g < 0
    alle = np.where(miss, 0, g).astype(np.uint8)
    hap_lists = [both_sides(pops[0]), both_sides(pops[1]) + [(pops[2][0], 1)], [(s, 0) for s in pops[2]]]
    left = np.full(45, 0xFFFF, dtype=np.uint16)
    for p, members in enumerate(pops):
        left[members] = p
    resident = _Matrix(alle, miss, pos, max_allele=1, ingest="u8")
    streamed = _Matrix.ingest(alle, miss, pos, hap_lists, partitions=[(left, left, 3)], calls=calls, packed=packed,
                              always_bitmap=packed == "library")
    for haps in hap_lists:
        _same_summary(resident.group(haps).summary(True), streamed.group(haps).summary(True))
    late = [(0, 0), (1, 1), (44, 0)]
    _same_summary(resident.group(late).summary(True), streamed.group(late).summary(True))
    L = _lib.lib()
    w = np.array([int(pos[0]), int(pos[-1])], dtype=np.int64)

    def totals(handle):
        oa, ob = np.zeros(1), np.zeros(1)
        pa, pb = np.zeros(3), np.zeros(3)
        pn, osz, nv = np.zeros(3, dtype=np.uint64), np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.uint64)
        _lib.check(L.fm_wc_window_sums(handle, w.ctypes.data, 1, nv.ctypes.data, oa.ctypes.data, ob.ctypes.data,
                                       osz.ctypes.data, pa.ctypes.data, pb.ctypes.data, pn.ctypes.data))
        return (oa[0], ob[0], int(osz[0]), pa.tolist(), pb.tolist(), pn.tolist())

    ph, ph2 = C.c_void_p(), C.c_void_p()
    _lib.check(L.fm_partition_create(resident.handle, left.ctypes.data, left.ctypes.data, 45, 3, C.byref(ph)))
    _lib.check(L.fm_partition_create(streamed.handle, left.ctypes.data, left.ctypes.data, 45, 3, C.byref(ph2)))
    assert totals(ph) == totals(streamed.partitions[0]) == totals(ph2)
    L.fm_partition_release(ph)
    L.fm_partition_release(ph2)


def test_packed_per_site_and_hudson_through_the_population_api(monkeypatch):
    import ferromic_b200 as F
    g, pos, pops = make_cohort(2500, 24, missing_rate=0.1, seed=5)
    L = int(pos[-1] - pos[0] + 1)
    h1, h2 = both_sides(pops[0]), both_sides(pops[1])
    out = {}
    for mode in ("packed", "u8"):
        monkeypatch.setenv("FERROMIC_GPU_INGEST", mode)
        base = F.Population.from_numpy("all", g, pos, h1 + h2, L, sample_names=[f"s{i}" for i in range(24)])
        p1, p2 = base.with_haplotypes("p1", h1), base.with_haplotypes("p2", h2)
        r = F.hudson_fst(p1, p2)
        out[mode] = (p1.segregating_sites(), p1.nucleotide_diversity(), r.fst, r.d_xy, r.pi_pop1, r.pi_pop2)
    assert out["packed"] == out["u8"]


def test_packed_error_behaviour():
    from ferromic_b200 import _lib
    L = _lib.lib()
    pos = np.arange(10, dtype=np.int64)
    rw = 1
    bits = np.zeros((10, rw), dtype=np.uint32)
    u8 = np.zeros((10, 4), dtype=np.uint8)
    # multi-allelic matrices cannot be packed (one allele bit per cell)
    ih = C.c_void_p()
    _lib.check(L.fm_ingest_begin(10, 2, 2, 0, 3, pos.ctypes.data, 0, C.byref(ih)))
    assert L.fm_ingest_rows_packed(ih, bits.ctypes.data, None, 0, 10) == _lib.FM_ERR_UNSUPPORTED
    L.fm_ingest_abort(ih)
    # called_bits must match the declared missingness; u8 and packed rows do not mix
    _lib.check(L.fm_ingest_begin(10, 2, 2, 1, 1, pos.ctypes.data, 0, C.byref(ih)))
    assert L.fm_ingest_rows_packed(ih, bits.ctypes.data, None, 0, 5) == _lib.FM_ERR_INVALID_ARG
    _lib.check(L.fm_ingest_rows_packed(ih, bits.ctypes.data, bits.ctypes.data, 0, 5))
    bm = np.zeros(2, dtype=np.uint64)
    assert L.fm_ingest_rows(ih, u8.ctypes.data, bm.ctypes.data, 5, 5) == _lib.FM_ERR_INVALID_ARG
    assert L.fm_ingest_rows_packed(ih, bits.ctypes.data, bits.ctypes.data, 8, 5) == _lib.FM_ERR_INVALID_ARG  # beyond V
    mh = C.c_void_p()
    assert L.fm_ingest_finish(ih, C.byref(mh), None, None) == _lib.FM_ERR_INVALID_ARG  # rows 5..9 still missing
    L.fm_ingest_abort(ih)
    _lib.check(L.fm_ingest_begin(10, 2, 2, 0, 1, pos.ctypes.data, 0, C.byref(ih)))
    assert L.fm_ingest_rows_packed(ih, bits.ctypes.data, bits.ctypes.data, 0, 10) == _lib.FM_ERR_INVALID_ARG
    _lib.check(L.fm_ingest_rows(ih, u8.ctypes.data, None, 0, 5))
    assert L.fm_ingest_rows_packed(ih, bits.ctypes.data, None, 5, 5) == _lib.FM_ERR_INVALID_ARG
    L.fm_ingest_abort(ih)
    _lib.check(L.fm_ingest_begin(10, 2, 2, 2, 1, pos.ctypes.data, 0, C.byref(ih)))  # in-band is a u8 notion
    assert L.fm_ingest_rows_packed(ih, bits.ctypes.data, bits.ctypes.data, 0, 10) == _lib.FM_ERR_INVALID_ARG
    L.fm_ingest_abort(ih)
    # empty matrices
    m = C.c_void_p()
    _lib.check(L.fm_matrix_create_packed(None, None, 0, 3, 2, None, C.byref(m)))
    L.fm_matrix_release(m)


@pytest.mark.parametrize("gap_code", [True, False])
@pytest.mark.parametrize("n_samples,missing", [(3, 0.02), (65, 0.01), (700, 0.02), (2504, 0.01), (40000, 0.005)])
def test_sparse_missing_list_equals_called_plane(n_samples, missing, gap_code, monkeypatch):
    """Packed rows whose missingness arrives as a sparse CSR list (fm_pack_rows_sparse -> fm_matrix_create_packed_sparse
    -> fm_k_expand_called), as column indices or as the one-byte gap code, give the counts of the called-plane form,
    of the u8 path and of the oracle."""
    from ferromic_b200 import api
    from ferromic_b200.api import _Matrix
    monkeypatch.setattr(api, "SPARSE_GAP_CODE", gap_code)
    V = 2500 if n_samples < 1000 else (500 if n_samples < 10000 else 40)
    g, pos, _ = make_cohort(V, n_samples, missing_rate=missing, seed=900 + n_samples)
    g[7] = -1  # a row without any call
    miss = g < 0
    alle = np.where(miss, 0, g).astype(np.uint8)
    sparse = _Matrix(alle, miss, pos, max_allele=1, always_bitmap=True, ingest="packed-sparse")
    assert sparse.ingest_mode == "packed-sparse"
    u8 = _Matrix(alle, miss, pos, max_allele=1, always_bitmap=True, ingest="u8")
    vs, d = orc.from_numpy(g, pos)
    for haps in (both_sides(range(n_samples)), both_sides(range(0, n_samples, 3)) + [(1, 1)]):
        ref = orc.build_summary(d, haps)
        got = sparse.group(haps).summary(want_arrays=True)
        assert np.array_equal(got["alt"], ref.alt) and np.array_equal(got["called"], ref.called)
        _same_summary(got, u8.group(haps).summary(want_arrays=True))


@pytest.mark.parametrize("gap_code", [True, False])
@pytest.mark.parametrize("calls", [1, 4])
def test_sparse_missing_streaming_ingest(calls, gap_code, monkeypatch):
    from ferromic_b200 import api
    from ferromic_b200.api import _Matrix
    monkeypatch.setattr(api, "SPARSE_GAP_CODE", gap_code)
    g, pos, pops = make_cohort(1800, 45, n_pops=3, missing_rate=0.02, seed=55 + calls)
    miss = g < 0
    alle = np.where(miss, 0, g).astype(np.uint8)
    hap_lists = [both_sides(pops[0]), both_sides(pops[1]) + [(pops[2][0], 1)], [(s, 0) for s in pops[2]]]
    left = np.full(45, 0xFFFF, dtype=np.uint16)
    for p, members in enumerate(pops):
        left[members] = p
    resident = _Matrix(alle, miss, pos, max_allele=1, ingest="u8")
    streamed = _Matrix.ingest(alle, miss, pos, hap_lists, partitions=[(left, left, 3)], calls=calls, packed=True,
                              sparse=True, always_bitmap=True)
    for haps in hap_lists + [[(0, 0), (1, 1), (44, 0)]]:
        _same_summary(resident.group(haps).summary(True), streamed.group(haps).summary(True))
