"""Thread safety of the boundary: PyO3 methods run under allow_threads, so concurrent calls from several
host threads on ONE Population / group handle are legal in the reference (Arc + OnceLock, lib.rs:637-650, 738,
777-789).  Here several threads hammer the same fm_group / fm_partition handles -- including the first, lazily
caching call -- and every result must equal the single-threaded one bit for bit."""
import ctypes as C
import threading

import numpy as np
import pytest

from tests.synth import both_sides, make_cohort

pytestmark = pytest.mark.gpu


def _run_threads(n, fn):
    errs, out = [], [None] * n

    def w(i):
        try:
            out[i] = fn(i)
        except Exception as e:  # pragma: no cover - reported below
            errs.append(e)

    ts = [threading.Thread(target=w, args=(i,)) for i in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errs, errs
    return out


@pytest.mark.parametrize("ingest", ["packed", "u8"])
def test_concurrent_calls_on_one_group(ingest):
    from ferromic_b200 import _lib
    from ferromic_b200.api import _Matrix
    L = _lib.lib()
    g, pos, pops = make_cohort(6000, 80, n_pops=2, missing_rate=0.05, seed=2024)
    miss = g < 0
    alle = np.where(miss, 0, g).astype(np.uint8)
    h1, h2 = both_sides(pops[0]), both_sides(pops[1])
    Lseq = int(pos[-1] - pos[0] + 1)

    def everything(m, rounds):
        a, b = m.group(h1), m.group(h2)
        res = []
        for _ in range(rounds):
            s = a.summary(want_arrays=True)
            pi = a.pi(Lseq, _lib.FM_PI_SUMMARY)
            out = _lib.HudsonOutcome()
            n = C.c_size_t()
            _lib.check(L.fm_hudson_pair(a.handle, b.handle, Lseq, Lseq, _lib.FM_HUDSON_SUMMARIES, 0, 0, 0, len(h1),
                                        len(h2), C.byref(out), None, C.byref(n)))
            V = m.V
            p_out = np.zeros(V, dtype=np.int64)
            pis, ths = np.zeros(V), np.zeros(V)
            k = C.c_size_t()
            mask = np.array([int(pos[100]), int(pos[400])], dtype=np.int64)
            _lib.check(L.fm_per_site_diversity(b.handle, len(h2), int(pos[0]), int(pos[-1]), mask.ctypes.data, 1, None, 0,
                                               p_out.ctypes.data, pis.ctypes.data, ths.ctypes.data, V, C.byref(k)))
            w = np.array([int(pos[10]), int(pos[3000])], dtype=np.int64)
            nv, sg, un = (np.zeros(1, dtype=np.uint64) for _ in range(3))
            ps = np.zeros(1)
            _lib.check(L.fm_group_window_sums(a.handle, w.ctypes.data, 1, nv.ctypes.data, sg.ctypes.data, ps.ctypes.data,
                                              un.ctypes.data))
            res.append((s["alt"].tobytes(), s["called"].tobytes(), s["segregating_sites"], s["pi_sum"], pi, out.fst,
                        out.d_xy, out.pi_pop1, out.pi_pop2, pis.tobytes(), ths.tobytes(), int(sg[0]), float(ps[0])))
        return res

    ref = everything(_Matrix(alle, miss, pos, max_allele=1, ingest=ingest), 1)[0]
    shared = _Matrix(alle, miss, pos, max_allele=1, ingest=ingest)  # nothing cached yet: the threads race on the first call
    shared.group(h1)
    shared.group(h2)
    for res in _run_threads(6, lambda i: everything(shared, 4)):
        for r in res:
            assert r == ref


def test_concurrent_wc_and_group_creation():
    from ferromic_b200 import _lib
    from ferromic_b200.api import _Matrix
    L = _lib.lib()
    g, pos, pops = make_cohort(3000, 60, n_pops=3, missing_rate=0.0, seed=7)
    alle = g.astype(np.uint8)
    left = np.full(60, 0xFFFF, dtype=np.uint16)
    for p, members in enumerate(pops):
        left[members] = p
    m = _Matrix(alle, None, pos, max_allele=1)
    ph = C.c_void_p()
    _lib.check(L.fm_partition_create(m.handle, left.ctypes.data, left.ctypes.data, 60, 3, C.byref(ph)))
    w = np.array([int(pos[0]), int(pos[-1])], dtype=np.int64)

    def wc(_i):
        oa, ob = np.zeros(1), np.zeros(1)
        pa, pb = np.zeros(3), np.zeros(3)
        pn, osz, nv = np.zeros(3, dtype=np.uint64), np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.uint64)
        outs = []
        for _ in range(5):
            _lib.check(L.fm_wc_window_sums(ph, w.ctypes.data, 1, nv.ctypes.data, oa.ctypes.data, ob.ctypes.data,
                                           osz.ctypes.data, pa.ctypes.data, pb.ctypes.data, pn.ctypes.data))
            outs.append((oa[0], ob[0], int(osz[0]), pa.tobytes(), pb.tobytes(), pn.tobytes()))
            # new groups of the same matrix from several threads at once (Population.with_haplotypes, lib.rs:622)
            haps = both_sides(pops[_i % 3])
            idx = np.asarray([h[0] for h in haps], dtype=np.uint64)
            side = np.asarray([h[1] for h in haps], dtype=np.uint8)
            gh = C.c_void_p()
            _lib.check(L.fm_group_create(m.handle, idx.ctypes.data, side.ctypes.data, len(haps), C.byref(gh)))
            seg = C.c_uint64()
            _lib.check(L.fm_group_segregating_sites(gh, C.byref(seg)))
            outs.append(seg.value)
            L.fm_group_release(gh)
        return outs

    res = _run_threads(6, wc)
    first = res[0][0]
    for i, r in enumerate(res):
        assert all(x == first for x in r[0::2])
        assert r[1::2] == res[i % 3][1::2]
    L.fm_partition_release(ph)
