"""fm_ingest_request_tracks: per-site pi / theta tracks computed while the rows of a streaming ingest arrive must be
bit-identical to fm_per_site_diversity_multi on the finished groups -- for every row format of the ingest, for rows
pushed in several calls with odd boundaries, for page-locked (direct stores) and pageable (copied at finish) outputs,
with a region inside the matrix, a mask, filtered positions and a group with fewer than two haplotypes."""
import ctypes as C

import numpy as np
import pytest

from tests.synth import both_sides, make_cohort

pytestmark = pytest.mark.gpu


def _reference(L, _lib, m, lists, raw_n, region, mask, filt, cap):
    groups = [m.group(h) for h in lists]
    gh = (C.c_void_p * len(groups))(*[g.handle for g in groups])
    rn = (C.c_size_t * len(groups))(*raw_n)
    pos = np.full(cap, -7, dtype=np.int64)
    pi, th = np.full((len(groups), cap), -7.0), np.full((len(groups), cap), -7.0)
    n = C.c_size_t()
    _lib.check(L.fm_per_site_diversity_multi(gh, rn, len(groups), region[0], region[1], mask.ctypes.data, mask.size // 2,
                                             filt.ctypes.data, filt.size, pos.ctypes.data, pi.ctypes.data, th.ctypes.data,
                                             cap, C.byref(n)))
    return n.value, pos, pi, th


def _same(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("mode", ["u8", "packed", "sparse", "library"])
@pytest.mark.parametrize("V,calls", [(5000, 1), (5000, 3), (70001, 4)])
def test_streamed_tracks_equal_the_call_after_finish(V, calls, mode, pinned):
    import torch
    from ferromic_b200 import _lib
    from ferromic_b200.api import _Matrix
    L = _lib.lib()
    g, pos, pops = make_cohort(V, 36, n_pops=2, missing_rate=0.03, seed=31 * V + calls)
    miss = g < 0
    alle = np.where(miss, 0, g).astype(np.uint8)
    lists = [both_sides(pops[0]), both_sides(pops[1]) + [(pops[0][0], 1)], [(3, 0)]]  # the last one: a single haplotype
    raw_n = [len(h) for h in lists]
    region = (int(pos[37]), int(pos[V - 11]))
    mask = np.array([int(pos[100]), int(pos[160]), int(pos[V // 2]), int(pos[V // 2 + 333])], dtype=np.int64)
    filt = np.array([int(pos[41]), int(pos[V // 3]), 5], dtype=np.int64)
    cap = V + 5
    if pinned:
        t_pos = torch.full((cap,), -7, dtype=torch.int64).pin_memory()
        t_pi = torch.full((3, cap), -7.0, dtype=torch.float64).pin_memory()
        t_th = torch.full((3, cap), -7.0, dtype=torch.float64).pin_memory()
        o_pos, o_pi, o_th = t_pos.numpy(), t_pi.numpy(), t_th.numpy()
        a_pos, a_pi, a_th = t_pos.data_ptr(), t_pi.data_ptr(), t_th.data_ptr()
    else:
        o_pos = np.full(cap, -7, dtype=np.int64)
        o_pi, o_th = np.full((3, cap), -7.0), np.full((3, cap), -7.0)
        a_pos, a_pi, a_th = o_pos, o_pi, o_th
    req = {"groups": [0, 1, 2], "raw_n": raw_n, "region": region, "mask": mask, "filtered": filt, "pos": a_pos,
           "pi": a_pi, "theta": a_th, "capacity": cap}
    m = _Matrix.ingest(alle, miss, pos, lists, calls=calls, always_bitmap=True,
                       packed="library" if mode == "library" else mode != "u8", sparse=mode == "sparse", tracks=req)
    n_ref, r_pos, r_pi, r_th = _reference(L, _lib, m, lists, raw_n, region, mask, filt, cap)
    assert m.track_sites == n_ref == V - 11 - 37 + 1
    n = n_ref
    assert np.array_equal(o_pos[:n], r_pos[:n]) and np.all(o_pos[n:] == -7)
    for k in range(3):
        assert _same(o_pi[k, :n], r_pi[k, :n]) and _same(o_th[k, :n], r_th[k, :n])
        assert np.all(o_pi[k, n:] == -7.0) and np.all(o_th[k, n:] == -7.0)
    assert np.all(np.isnan(o_pi[2, :n]))                      # fewer than two haplotypes: no sites in the reference
    assert np.isnan(o_pi[0, 100 - 37]) and not np.all(np.isnan(o_pi[0, :n]))   # masked site


def test_request_errors_and_empty_region():
    from ferromic_b200 import _lib
    L = _lib.lib()
    g, pos, pops = make_cohort(300, 10, n_pops=2, seed=5)
    alle = g.astype(np.uint8)
    ih = C.c_void_p()
    _lib.check(L.fm_ingest_begin(300, 10, 2, 0, 1, pos.ctypes.data, 0, C.byref(ih)))
    idx = np.asarray([0, 1, 2], dtype=np.uint64)
    side = np.zeros(3, dtype=np.uint8)
    _lib.check(L.fm_ingest_add_group(ih, idx.ctypes.data, side.ctypes.data, 3, None))
    gi = np.zeros(1, dtype=np.uint64)
    rn = np.array([3], dtype=np.uint64)
    pi, th = np.zeros(300), np.zeros(300)
    n = C.c_size_t(99)
    bad = np.array([4], dtype=np.uint64)
    assert L.fm_ingest_request_tracks(ih, bad.ctypes.data, rn.ctypes.data, 1, 0, 10 ** 9, None, 0, None, 0, None,
                                      pi.ctypes.data, th.ctypes.data, 300, C.byref(n)) == _lib.FM_ERR_INVALID_ARG
    assert L.fm_ingest_request_tracks(ih, gi.ctypes.data, rn.ctypes.data, 1, 0, 10 ** 9, None, 0, None, 0, None,
                                      pi.ctypes.data, th.ctypes.data, 10, C.byref(n)) == _lib.FM_ERR_INVALID_ARG  # capacity
    # an empty region: accepted, no sites (stats.rs:4656-4666)
    _lib.check(L.fm_ingest_request_tracks(ih, gi.ctypes.data, rn.ctypes.data, 1, 50, 10, None, 0, None, 0, None,
                                          pi.ctypes.data, th.ctypes.data, 300, C.byref(n)))
    assert n.value == 0
    assert L.fm_ingest_request_tracks(ih, gi.ctypes.data, rn.ctypes.data, 1, 0, 10 ** 9, None, 0, None, 0, None,
                                      pi.ctypes.data, th.ctypes.data, 300, C.byref(n)) == _lib.FM_ERR_INVALID_ARG  # twice
    _lib.check(L.fm_ingest_rows(ih, alle.ctypes.data, None, 0, 300))
    mh, gh = C.c_void_p(), (C.c_void_p * 1)()
    _lib.check(L.fm_ingest_finish(ih, C.byref(mh), gh, None))
    L.fm_group_release(C.c_void_p(gh[0]))
    L.fm_matrix_release(mh)
    # after the first rows call a request is refused
    ih = C.c_void_p()
    _lib.check(L.fm_ingest_begin(300, 10, 2, 0, 1, pos.ctypes.data, 0, C.byref(ih)))
    _lib.check(L.fm_ingest_add_group(ih, idx.ctypes.data, side.ctypes.data, 3, None))
    _lib.check(L.fm_ingest_rows(ih, alle.ctypes.data, None, 0, 100))
    assert L.fm_ingest_request_tracks(ih, gi.ctypes.data, rn.ctypes.data, 1, 0, 10 ** 9, None, 0, None, 0, None,
                                      pi.ctypes.data, th.ctypes.data, 300, C.byref(n)) == _lib.FM_ERR_INVALID_ARG
    L.fm_ingest_abort(ih)
