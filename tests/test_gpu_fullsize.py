"""Parity at BASELINE.json's full sizes: configs[0] site by site, configs 2-5 through size-independent properties:
  * the cohort is generated on the device by the library's counter-based generator
    (fm_synth_fill) -- the host never holds the 50 GB matrix;
  * random site slices are re-evaluated on the CPU (tests/synth.py::synth_rows, integer-exact),
    fed to the oracle, and compared with the device results for those sites (ints bit-exact,
    FP64 <= 1e-9 relative);
  * window / shard totals must add up to the whole-region totals (ints exact, FP 1e-12).
FM_FULLSIZE_SCALE (default 1.0) scales the site counts for quick runs."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.synth import synth_rows
from tests.test_sharded_gloo import _close, oracle_hudson_totals

pytestmark = pytest.mark.gpu
SCALE = float(os.environ.get("FM_FULLSIZE_SCALE", "1.0"))


def _lib():
    from ferromic_b200 import _lib as L
    return L


class DeviceCohort:
    """u8 matrix (+ bitmap) filled on the device, wrapped in an fm_matrix handle."""

    def __init__(self, V, S, seed, pop_of_sample, sigma, missing_rate):
        import torch
        L = _lib()
        self.V, self.S, self.seed, self.sigma, self.missing_rate = V, S, seed, sigma, missing_rate
        self.pop_of_sample = np.ascontiguousarray(pop_of_sample, dtype=np.uint16)
        total = V * S * 2
        self.d_data = torch.empty(total, dtype=torch.uint8, device="cuda:0")
        self.d_miss = torch.empty((total + 63) // 64, dtype=torch.int64, device="cuda:0") if missing_rate > 0 else None
        L.check(L.lib().fm_set_device(0))
        L.check(L.lib().fm_synth_fill(self.d_data.data_ptr(), self.d_miss.data_ptr() if self.d_miss is not None else None,
                                      V, S, 2, 0, seed, self.pop_of_sample.ctypes.data, sigma, missing_rate))
        rng = np.random.default_rng(seed)
        self.pos = np.cumsum(rng.integers(1, 50, size=V, dtype=np.int64))
        self.m = C.c_void_p()
        L.check(L.lib().fm_matrix_create_device(self.d_data.data_ptr(),
                                                self.d_miss.data_ptr() if self.d_miss is not None else None, V, S, 2, 1,
                                                self.pos.ctypes.data, C.byref(self.m)))
        self.groups = {}

    def group(self, haps):
        L = _lib()
        key = (haps[0], haps[-1], len(haps))
        if key not in self.groups:
            idx = np.asarray([h[0] for h in haps], dtype=np.uint64)
            side = np.asarray([h[1] for h in haps], dtype=np.uint8)
            h = C.c_void_p()
            L.check(L.lib().fm_group_create(self.m, idx.ctypes.data, side.ctypes.data, len(haps), C.byref(h)))
            self.groups[key] = h
        return self.groups[key]

    def counts(self, g):
        L = _lib()
        alt = np.zeros(self.V, dtype=np.uint32)
        called = np.zeros(self.V, dtype=np.uint32)
        seg, unc, pi = C.c_uint64(), C.c_uint64(), C.c_double()
        L.check(L.lib().fm_group_summary(g, alt.ctypes.data, called.ctypes.data, C.byref(seg), C.byref(pi), C.byref(unc)))
        return alt, called, seg.value, pi.value, unc.value

    def slice_rows(self, lo, hi):
        return synth_rows(self.seed, lo, hi, self.S, 2, self.pop_of_sample, self.sigma, self.missing_rate)

    def close(self):
        import torch
        L = _lib()
        for h in self.groups.values():
            L.lib().fm_group_release(h)
        L.lib().fm_matrix_release(self.m)
        self.d_data = self.d_miss = None
        torch.cuda.empty_cache()
        L.lib().fm_trim_pool()


def hudson_windows(L, g1, g2, windows):
    n = len(windows)
    f = [np.zeros(n) for _ in range(5)]
    sk = np.zeros(n, dtype=np.uint64)
    w = np.ascontiguousarray(windows, dtype=np.int64)
    L.check(L.lib().fm_hudson_window_sums(g1, g2, w.ctypes.data, n, f[0].ctypes.data, f[1].ctypes.data,
                                          f[2].ctypes.data, sk.ctypes.data, f[3].ctypes.data, f[4].ctypes.data))
    return np.stack(f, axis=1), sk


def shard_windows(pos, n):
    from ferromic_b200.sharded import shard_bounds
    b = shard_bounds(len(pos), n)
    return np.array([(int(pos[a]), int(pos[e - 1])) for a, e in zip(b[:-1], b[1:]) if e > a], dtype=np.int64)


def halves(S):
    return ([(s, k) for s in range(S // 2) for k in (0, 1)], [(s, k) for s in range(S // 2, S) for k in (0, 1)])


def check_count_slices(co, groups_haps, n_slices, width, rng):
    """alt/called of random site slices against the oracle on CPU-regenerated genotypes."""
    dev = [co.counts(co.group(h)) for h in groups_haps]
    slices = []
    for _ in range(n_slices):
        lo = int(rng.integers(0, max(1, co.V - width)))
        hi = min(co.V, lo + width)
        g = co.slice_rows(lo, hi)
        vs, d = orc.from_numpy(g, co.pos[lo:hi])
        sums = []
        for k, haps in enumerate(groups_haps):
            s = orc.build_summary(d, haps)
            assert np.array_equal(dev[k][0][lo:hi], s.alt), "alt counts differ"
            assert np.array_equal(dev[k][1][lo:hi], s.called), "called counts differ"
            sums.append(s)
        slices.append((lo, hi, g, vs, d, sums))
    return dev, slices


def test_config1_100k_sites_one_group_every_site_bit_exact():
    """BASELINE configs[0] at its full size: 100k sites x 2504 diploid samples, ONE group of all 5008 haplotypes, no
    missing data.  Every site's counts are compared with the oracle (the whole matrix is regenerated on the CPU in
    chunks), then S, sum pi, pi and Watterson's theta of the region."""
    L = _lib()
    V, S = int(100_000 * SCALE), 2504
    co = DeviceCohort(V, S, 102_504, np.zeros(S, dtype=np.uint16), 0.0, 0.0)
    try:
        haps = [(s, k) for s in range(S) for k in (0, 1)]
        g = co.group(haps)
        alt, called, seg, pi_sum, unc = co.counts(g)
        ref_seg, ref_pi = 0, 0.0
        step = 10_000
        for lo in range(0, V, step):
            hi = min(V, lo + step)
            rows = co.slice_rows(lo, hi)
            _, d = orc.from_numpy(rows, co.pos[lo:hi])
            s = orc.build_summary(d, haps)
            assert np.array_equal(alt[lo:hi], s.alt), f"alt counts differ in sites [{lo}, {hi})"
            assert np.array_equal(called[lo:hi], s.called), f"called counts differ in sites [{lo}, {hi})"
            ref_seg += s.seg
            ref_pi += s.pi_sum
        assert seg == ref_seg and unc == 0
        assert _close(pi_sum, ref_pi, 1e-9)
        Lr = int(co.pos[-1] - co.pos[0] + 1)
        whole = orc.Summary(alt, called, len(haps), ref_seg, ref_pi)
        ref = orc.pi_for_population(orc.Pop(haps, None, S, Lr, summary=whole))
        for path in (L.FM_PI_SUMMARY, L.FM_PI_DENSE):
            out = C.c_double()
            L.check(L.lib().fm_group_pi(g, Lr, path, len(haps), C.byref(out)))
            assert _close(out.value, ref, 1e-9)
        n_seg = C.c_uint64()
        L.check(L.lib().fm_group_segregating_sites(g, C.byref(n_seg)))
        assert n_seg.value == ref_seg
        th = C.c_double()
        L.check(L.lib().fm_watterson_theta(ref_seg, len(haps), Lr, C.byref(th)))
        assert _close(th.value, orc.watterson_theta(ref_seg, len(haps), Lr), 1e-12)
    finally:
        co.close()


def test_config3_hudson_10M_sites_two_populations():
    L = _lib()
    V, S = int(10_000_000 * SCALE), 2504
    pop = (np.arange(S) >= S // 2).astype(np.uint16)
    co = DeviceCohort(V, S, 10_002_504, pop, 0.05, 0.01)
    try:
        h1, h2 = halves(S)
        g1, g2 = co.group(h1), co.group(h2)
        Lr = int(co.pos[-1] - co.pos[0] + 1)
        out = L.HudsonOutcome()
        n = C.c_size_t()
        L.check(L.lib().fm_hudson_pair(g1, g2, Lr, Lr, L.FM_HUDSON_SUMMARIES, 0, 0, 0, len(h1), len(h2), C.byref(out),
                                       None, C.byref(n)))
        assert out.some == 31 and 0.0 < out.fst < 0.2 and 0.0 < out.d_xy < 1.0
        whole, wsk = hudson_windows(L, g1, g2, np.array([[int(co.pos[0]), int(co.pos[-1])]]))
        parts, psk = hudson_windows(L, g1, g2, shard_windows(co.pos, 8))
        assert psk.sum() == wsk[0]
        assert np.allclose(parts.sum(axis=0), whole[0], rtol=1e-12, atol=0)
        assert _close(out.fst, whole[0, 0] / whole[0, 1], 1e-12)
        assert _close(out.d_xy, whole[0, 2] / (Lr - int(wsk[0])), 1e-12)
        rng = np.random.default_rng(3)
        dev, slices = check_count_slices(co, (h1, h2), 3, 10_000, rng)
        for lo, hi, g, vs, d, (s1, s2) in slices:
            win = np.array([[int(co.pos[lo]), int(co.pos[hi - 1])]])
            got, gsk = hudson_windows(L, g1, g2, win)
            ref = oracle_hudson_totals(s1, s2, co.pos[lo:hi], win)
            assert int(gsk[0]) == int(ref.u[0, 0])
            assert np.allclose(got[0], ref.f[0], rtol=1e-9, atol=0)
            # and the regional outcome of the slice through the oracle's own Hudson entry point
            Ls = int(win[0, 1] - win[0, 0] + 1)
            rc, o, _ = orc.hudson_pair(orc.Pop(h1, None, S, Ls, summary=s1), orc.Pop(h2, None, S, Ls, summary=s2))
            assert rc == 0 and _close(got[0, 0] / got[0, 1], o["fst"], 1e-9)
        # summary scalars: S and sum of pi over all 10M sites vs the per-slice oracle values summed is
        # not possible without the whole matrix; check internal consistency instead
        alt, called, seg, pis, unc = dev[0]
        assert seg == int(((called >= 2) & (alt > 0) & (alt < called)).sum())
        assert unc == int((called < 2).sum())
    finally:
        co.close()


def test_config4_wc_26_populations_10M_sites():
    from tools.wc_timing import POP_SIZES, membership
    L = _lib()
    V, S = int(10_000_000 * SCALE), 2504
    left, right = membership(S)
    co = DeviceCohort(V, S, 10_002_504, left, 0.08, 0.0)  # no missing data: dense == sparse semantics
    ph = C.c_void_p()
    try:
        L.check(L.lib().fm_partition_create(co.m, left.ctypes.data, right.ctypes.data, S, 26, C.byref(ph)))
        NP = 325

        def wc(windows):
            w = np.ascontiguousarray(windows, dtype=np.int64)
            n = len(w)
            nv, osz = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
            oa, ob = np.zeros(n), np.zeros(n)
            pa, pb, pn = np.zeros((n, NP)), np.zeros((n, NP)), np.zeros((n, NP), dtype=np.uint64)
            L.check(L.lib().fm_wc_window_sums(ph, w.ctypes.data, n, nv.ctypes.data, oa.ctypes.data, ob.ctypes.data,
                                              osz.ctypes.data, pa.ctypes.data, pb.ctypes.data, pn.ctypes.data))
            return nv, oa, ob, osz, pa, pb, pn

        whole = wc(np.array([[int(co.pos[0]), int(co.pos[-1])]]))
        parts = wc(shard_windows(co.pos, 8))
        assert parts[0].sum() == whole[0][0] == V and parts[3].sum() == whole[3][0]
        assert np.array_equal(parts[6].sum(axis=0), whole[6][0])
        assert _close(parts[1].sum(), whole[1][0], 1e-11) and _close(parts[2].sum(), whole[2][0], 1e-11)
        assert np.allclose(parts[4].sum(axis=0), whole[4][0], rtol=1e-10, atol=1e-9)
        assert np.allclose(parts[5].sum(axis=0), whole[5][0], rtol=1e-11, atol=0)
        fst = whole[1][0] / (whole[1][0] + whole[2][0])
        assert 0.0 < fst < 0.2
        rng = np.random.default_rng(4)
        for _ in range(3):
            lo = int(rng.integers(0, max(1, V - 1500)))
            hi = min(V, lo + 1500)
            g = co.slice_rows(lo, hi)
            vs, _d = orc.from_numpy(g, co.pos[lo:hi])
            region = (int(co.pos[lo]), int(co.pos[hi - 1]))
            ref = orc.wc_fst(vs, left, right, 26, region, want_pairs=False)
            got = wc(np.array([region]))
            assert int(got[0][0]) == ref["n_sites"] and int(got[3][0]) == ref["overall"]["sites"]
            assert _close(got[1][0], ref["overall"]["sum_a"], 1e-9) and _close(got[2][0], ref["overall"]["sum_b"], 1e-9)
            for k in range(NP):
                e = ref["pairs"][k]
                assert int(got[6][0, k]) == e["sites"]
                assert _close(got[4][0, k], e["sum_a"], 1e-9) and _close(got[5][0, k], e["sum_b"], 1e-9), k
        assert len(POP_SIZES) == 26
    finally:
        if ph:
            L.lib().fm_partition_release(ph)
        co.close()


def test_config2_per_site_tracks_1M_sites_against_count_formulas():
    """1M sites, two orientation groups, bitmap + mask: the device pi/theta tracks against the
    per-site formulas evaluated in numpy from the oracle-checked counts (stats.rs:2723-2733,
    4716-4722), NaN exactly where called < 2 or the mask covers the site."""
    L = _lib()
    V, S = int(1_000_000 * SCALE), 2504
    co = DeviceCohort(V, S, 1_002_504, np.zeros(S, dtype=np.uint16), 0.0, 0.01)
    try:
        rng = np.random.default_rng(1_002_506)
        orient = rng.random((S, 2)) < 0.3
        g0 = [(s, k) for s in range(S) for k in (0, 1) if not orient[s, k]]
        g1 = [(s, k) for s in range(S) for k in (0, 1) if orient[s, k]]
        starts = rng.integers(int(co.pos[0]), int(co.pos[-1]), size=2000)
        mask = np.stack([starts, starts + rng.integers(1000, 5000, size=2000)], axis=1).astype(np.int64)
        dev, _ = check_count_slices(co, (g0, g1), 2, 5_000, np.random.default_rng(5))
        masked = np.zeros(V, dtype=bool)
        for s, e in mask:
            masked[np.searchsorted(co.pos, s, "left"):np.searchsorted(co.pos, e, "left")] = True
        for k, haps in enumerate((g0, g1)):
            pos_out = np.zeros(V, dtype=np.int64)
            pi, th = np.zeros(V), np.zeros(V)
            n = C.c_size_t()
            L.check(L.lib().fm_per_site_diversity(co.group(haps), len(haps), int(co.pos[0]), int(co.pos[-1]),
                                                  mask.ctypes.data, len(mask), None, 0, pos_out.ctypes.data,
                                                  pi.ctypes.data, th.ctypes.data, V, C.byref(n)))
            assert n.value == V and np.array_equal(pos_out, co.pos + 1)
            alt, called = dev[k][0].astype(np.float64), dev[k][1].astype(np.float64)
            bad = masked | (called < 2)
            assert np.array_equal(np.isnan(pi), bad) and np.array_equal(np.isnan(th), bad)
            ok = ~bad
            nn, a = called[ok], alt[ok]
            inv = 1.0 / nn
            ref_pi = nn / (nn - 1.0) * (1.0 - ((nn - a) * (nn - a) + a * a) * inv * inv)
            assert np.allclose(pi[ok], ref_pi, rtol=1e-12, atol=0)
            H = np.concatenate([[0.0], np.cumsum(1.0 / np.arange(1, len(haps) + 1))])  # H[k] = sum_{i<=k} 1/i
            poly = (a > 0) & (a < nn)
            ref_th = np.where(poly, 1.0 / H[(nn - 1).astype(np.int64)], 0.0)
            assert np.allclose(th[ok], ref_th, rtol=1e-12, atol=0)
    finally:
        co.close()


def test_config5_biobank_shard_200k_haplotypes_windows():
    """One GPU's share of config 5 (2M sites over 8 GPUs = 250k sites x 200k haplotypes, 1 %
    missing): the column-chunked plane pass, 100 kb windows, slices against the oracle."""
    L = _lib()
    V, S = int(250_000 * SCALE), 100_000
    pop = (np.arange(S) >= S // 2).astype(np.uint16)
    co = DeviceCohort(V, S, 2_100_000, pop, 0.05, 0.01)
    try:
        h1, h2 = halves(S)
        g1, g2 = co.group(h1), co.group(h2)
        edges = np.arange(int(co.pos[0]), int(co.pos[-1]) + 1, 100_000, dtype=np.int64)
        windows = np.stack([edges, edges + 99_999], axis=1)
        parts, psk = hudson_windows(L, g1, g2, windows)
        whole, wsk = hudson_windows(L, g1, g2, np.array([[int(co.pos[0]), int(co.pos[-1])]]))
        assert psk.sum() == wsk[0]
        assert np.allclose(parts.sum(axis=0), whole[0], rtol=1e-12, atol=0)
        dev, slices = check_count_slices(co, (h1, h2), 2, 200, np.random.default_rng(6))
        for lo, hi, g, vs, d, (s1, s2) in slices:
            win = np.array([[int(co.pos[lo]), int(co.pos[hi - 1])]])
            got, gsk = hudson_windows(L, g1, g2, win)
            ref = oracle_hudson_totals(s1, s2, co.pos[lo:hi], win)
            assert np.allclose(got[0], ref.f[0], rtol=1e-9, atol=0)
        # window diversity of population 1 adds up as well
        nw = len(windows)
        nv, seg, unc = (np.zeros(nw, dtype=np.uint64) for _ in range(3))
        pis = np.zeros(nw)
        w = np.ascontiguousarray(windows)
        L.check(L.lib().fm_group_window_sums(g1, w.ctypes.data, nw, nv.ctypes.data, seg.ctypes.data, pis.ctypes.data,
                                             unc.ctypes.data))
        alt, called, s_all, pi_all, unc_all = dev[0]
        assert nv.sum() == V and seg.sum() == s_all and unc.sum() == unc_all
        assert _close(float(pis.sum()), pi_all, 1e-11)
    finally:
        co.close()
