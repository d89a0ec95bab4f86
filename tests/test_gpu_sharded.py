"""GPU tests of the site-sharded path: two shards of one cohort (as two ranks would hold them)
evaluated by the kernels, merged in rank order, against the single-shard run and the oracle."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.synth import both_sides, make_cohort
from tests.test_sharded_gloo import _close, oracle_diversity_totals, oracle_hudson_totals

pytestmark = pytest.mark.gpu


def _cohort(V=20000, S=30, n_pops=3, missing=0.05, seed=4242):
    g, pos, pops = make_cohort(V, S, n_pops=n_pops, sigma=0.08, missing_rate=missing, seed=seed)
    g[:, :, 1][g[:, :, 0] < 0] = -1
    g[:, :, 0][g[:, :, 1] < 0] = -1
    g[8000:8300] = -1  # no data on either side of the shard cut
    left = np.full(S, 0xFFFF, dtype=np.uint16)
    for p, members in enumerate(pops):
        left[members] = p
    return g, pos, pops, left


def _windows(pos):
    edges = [0, 1500, 8191, 8192, 9000, 16384, 16500, len(pos) - 1]
    w = [(int(pos[a]), int(pos[b])) for a, b in zip(edges[:-1], edges[1:])]
    w.append((int(pos[0]), int(pos[-1])))          # the whole region
    w.append((int(pos[-1]) + 5, int(pos[-1]) + 50))  # no variants
    return np.array(w, dtype=np.int64)


def test_two_shards_merge_to_the_single_shard_result():
    from ferromic_b200 import sharded
    g, pos, pops, left = _cohort()
    windows = _windows(pos)
    h1, h2 = both_sides(pops[0]), both_sides(pops[1])
    whole = sharded.CohortShard(g, pos)
    cuts = sharded.shard_bounds(len(pos), 2)
    assert cuts[1] == 8192
    shards = [sharded.CohortShard(g[a:b], pos[a:b], rank=r, world=2) for r, (a, b) in
              enumerate(zip(cuts[:-1], cuts[1:]))]
    for name, fn in (("div", lambda c: c.diversity_totals(h1, windows)),
                     ("hudson", lambda c: c.hudson_totals(h1, h2, windows)),
                     ("wc", lambda c: c.wc_totals(left, left, 3, windows))):
        ref = fn(whole)
        got = sharded.merge_in_rank_order([fn(s) for s in shards])
        assert np.array_equal(ref.u, got.u), name
        assert np.allclose(ref.f, got.f, rtol=1e-12, atol=1e-300), name
    # against the oracle, window by window
    vs, d = orc.from_numpy(g, pos)
    s1, s2 = orc.build_summary(d, h1), orc.build_summary(d, h2)
    lengths = [int(we - ws + 1) for ws, we in windows]
    div = sharded.merge_in_rank_order([s.diversity_totals(h1, windows) for s in shards])
    odiv = oracle_diversity_totals(s1, pos, windows)
    assert np.array_equal(div.u, odiv.u)
    assert np.allclose(div.f, odiv.f, rtol=1e-11, atol=0)
    hud = sharded.merge_in_rank_order([s.hudson_totals(h1, h2, windows) for s in shards])
    ohud = oracle_hudson_totals(s1, s2, pos, windows)
    assert np.array_equal(hud.u, ohud.u)
    assert np.allclose(hud.f, ohud.f, rtol=1e-11, atol=1e-300)
    res = sharded.finish_hudson(hud, lengths, len(h1), len(h2))
    for w, (ws, we) in enumerate(windows):
        sel = np.nonzero((pos >= ws) & (pos <= we))[0]
        sub1 = orc.Summary(s1.alt[sel], s1.called[sel], s1.capacity, 0, 0.0)
        sub2 = orc.Summary(s2.alt[sel], s2.called[sel], s2.capacity, 0, 0.0)
        rc, ref, _ = orc.hudson_pair(orc.Pop(h1, None, g.shape[1], lengths[w], summary=sub1),
                                     orc.Pop(h2, None, g.shape[1], lengths[w], summary=sub2))
        for k in ("fst", "d_xy", "pi_pop1", "pi_pop2", "pi_xy_avg"):
            assert _close(res[w][k], ref[k], 1e-9), (w, k)
    wc = sharded.merge_in_rank_order([s.wc_totals(left, left, 3, windows) for s in shards])
    ov, prs = sharded.finish_wc(wc, 3)
    for w, (ws, we) in enumerate(windows):
        wref = orc.wc_fst(vs, left, left, 3, (int(ws), int(we)), want_pairs=False)
        e = ov[w]
        assert orc.STATE_NAMES[e.state] == wref["overall"]["state"] and e.sites == wref["overall"]["sites"], w
        assert _close(e.sum_a, wref["overall"]["sum_a"], 1e-9) and _close(e.sum_b, wref["overall"]["sum_b"], 1e-9)
        for k in range(3):
            pe, r = prs[w][k], wref["pairs"][k]
            if not wref["pair_present"][k]:
                assert pe.state == 3 and pe.sites == 0
                continue
            assert orc.STATE_NAMES[pe.state] == r["state"] and pe.sites == r["sites"], (w, k)
            assert _close(pe.sum_a, r["sum_a"], 1e-9) and _close(pe.sum_b, r["sum_b"], 1e-9), (w, k)
            if pe.state == 0:
                assert _close(pe.value, r["value"], 1e-9)
    for s in shards + [whole]:
        s.close()


def test_wc_window_sums_26_populations_match_region_calls():
    """fm_wc_window_sums over many windows == fm_wc_fst on each window's region (same kernel,
    different segmentation), and the windows add up to the whole region."""
    import ferromic_b200 as F
    from ferromic_b200 import sharded
    g, pos, pops, left = _cohort(V=6000, S=104, n_pops=26, missing=0.02, seed=99)
    labels = sorted(str(i) for i in range(26))
    cohort = sharded.CohortShard(g, pos)
    edges = np.linspace(0, len(pos), 8).astype(int)
    windows = np.array([(int(pos[a]), int(pos[b - 1])) for a, b in zip(edges[:-1], edges[1:])], dtype=np.int64)
    tot = cohort.wc_totals(left, left, 26, windows)
    whole = cohort.wc_totals(left, left, 26, np.array([[int(pos[0]), int(pos[-1])]], dtype=np.int64))
    assert np.array_equal(tot.u.sum(axis=0), whole.u[0])
    assert np.allclose(tot.f.sum(axis=0), whole.f[0], rtol=1e-11, atol=1e-300)
    vs, _ = orc.from_numpy(g, pos)
    ref = orc.wc_fst(vs, left, left, 26, (int(pos[0]), int(pos[-1])), want_pairs=False)
    ov, prs = sharded.finish_wc(whole, 325)
    assert ov[0].sites == ref["overall"]["sites"]
    assert _close(ov[0].sum_a, ref["overall"]["sum_a"], 1e-9) and _close(ov[0].value, ref["overall"]["value"], 1e-9)
    for k in range(325):
        assert prs[0][k].sites == ref["pairs"][k]["sites"]
        assert _close(prs[0][k].sum_a, ref["pairs"][k]["sum_a"], 1e-9), k
        assert _close(prs[0][k].sum_b, ref["pairs"][k]["sum_b"], 1e-9), k
    cohort.close()


def test_peer_mailbox_exchange_two_ranks_in_one_process():
    """fm_comm_* (P2P mailbox all-gather, csrc/fm_comm.cuh): two communicators on this GPU wired in
    process, driven from two host threads like two ranks would; gathered words are exact and the
    rank-ordered FP64 sum has identical bits on both ranks."""
    import threading
    from ferromic_b200 import sharded
    comms = [sharded.PeerComm._bare(r, 2) for r in range(2)]
    sharded.PeerComm.connect_in_process(comms)
    rng = np.random.default_rng(11)
    out = {}

    def run(r):
        res = []
        for step in range(6):
            n = [7, 2048, 1, 300, 64, 5][step]
            f = np.random.default_rng(100 * step + r).normal(size=n)
            res.append((f,) + comms[r].allgather_words(f, n_double=n // 2))
        out[r] = res

    ts = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=60)
    assert set(out) == {0, 1}
    for step in range(6):
        f0, g0, m0 = out[0][step]
        f1, g1, m1 = out[1][step]
        assert np.array_equal(g0, g1) and np.array_equal(m0, m1)
        assert np.array_equal(g0[0], f0.view(np.uint64)) and np.array_equal(g0[1], f1.view(np.uint64))
        nd = len(f0) // 2
        assert np.array_equal(m0[:nd].view(np.float64), f0[:nd] + f1[:nd])
        with np.errstate(over="ignore"):
            assert np.array_equal(m0[nd:], f0[nd:].view(np.uint64) + f1[nd:].view(np.uint64))
    # totals of two cohort shards through the mailbox == the rank-ordered merge on the host
    g, pos, pops, left = _cohort(V=17000)
    windows = _windows(pos)[:4]
    h1 = both_sides(pops[0])
    cuts = sharded.shard_bounds(len(pos), 2)
    shards = [sharded.CohortShard(g[a:b], pos[a:b], rank=r, world=2, comm=comms[r])
              for r, (a, b) in enumerate(zip(cuts[:-1], cuts[1:]))]
    local = [s.diversity_totals(h1, windows) for s in shards]
    ref = sharded.merge_in_rank_order(local)
    got = {}

    def gather(r):
        got[r] = sharded.peer_gather_totals(local[r], comms[r])

    ts = [threading.Thread(target=gather, args=(r,)) for r in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=60)
    for r in range(2):
        assert np.array_equal(got[r].f, ref.f) and np.array_equal(got[r].u, ref.u)
    for s in shards:
        s.close()
    for c in comms:
        c.close()


def test_peer_mailbox_single_rank_and_timeout():
    from ferromic_b200 import FerromicGpuError, sharded
    c = sharded.PeerComm(0, 1)
    w = np.arange(10, dtype=np.float64)
    g, m = c.allgather_words(w, n_double=10)
    assert np.array_equal(g[0].view(np.float64), w) and np.array_equal(m.view(np.float64), w)
    c.close()
    # a peer that never shows up: the exchange gives up instead of hanging the GPU
    a, b = sharded.PeerComm._bare(0, 2), sharded.PeerComm._bare(1, 2)
    sharded.PeerComm.connect_in_process([a, b])
    import ctypes as C
    from ferromic_b200 import _lib
    assert _lib.lib().fm_comm_set_timeout_ms(a.handle, 200) == 0
    with pytest.raises(FerromicGpuError):
        a.allgather_words(w)
    a.close()
    b.close()


@pytest.mark.parametrize("max_allele,missing", [(2, 0.0), (5, 0.1)])
def test_multi_allelic_window_totals_match_region_calls_and_merge(max_allele, missing):
    """Window totals of a multi-allelic cohort (general dense forms over per-allele counts): every window equals
    the oracle's dense region calls on that window's sub-cohort (calculate_pi_dense / count_segregating_sites_dense /
    Hudson from the dense matrix), and two site shards merge to the single-shard totals."""
    from ferromic_b200 import sharded
    from tests.test_gpu_parity import make_multi_cohort
    S = 20
    g, pos = make_multi_cohort(16384 + 700, S, max_allele, missing, seed=300 + max_allele)
    h1, h2 = both_sides(range(0, S // 2)), both_sides(range(S // 2, S)) + [(0, 1)]
    cutsw = [0, 300, 301, 5000, 9000, len(pos) - 1]
    windows = np.array([(int(pos[a]), int(pos[b]) - (1 if b < len(pos) - 1 else 0)) for a, b in zip(cutsw, cutsw[1:])] +
                       [(int(pos[-1]) + 5, int(pos[-1]) + 50)], dtype=np.int64)
    lengths = [int(we - ws + 1) for ws, we in windows]
    whole = sharded.CohortShard(g, pos)
    div = whole.diversity_totals(h1, windows)
    hud = whole.hudson_totals(h1, h2, windows)
    pi, theta = sharded.finish_diversity(div, lengths, len(h1))
    hres = sharded.finish_hudson(hud, lengths, len(h1), len(set(h2)))
    for w, (ws, we) in enumerate(windows):
        sel = np.nonzero((pos >= ws) & (pos <= we))[0]
        assert int(div.col("n_variants")[w]) == len(sel)
        if len(sel) == 0:
            continue
        vs, d = orc.from_numpy(g[sel], pos[sel])
        o1 = orc.Pop(h1, vs, S, lengths[w], dense=d)
        o2 = orc.Pop(h2, vs, S, lengths[w], dense=d)
        assert int(div.col("seg_sites")[w]) == orc.count_segregating_sites_for_population(o1)
        assert _close(float(pi[w]), orc.pi_for_population(o1), 1e-9), w
        rc, ref, _ = orc.hudson_pair(o1, o2)
        assert rc == 0
        for k in ("fst", "d_xy", "pi_pop1", "pi_pop2", "pi_xy_avg"):
            assert _close(hres[w][k], ref[k], 1e-9), (w, k, hres[w][k], ref[k])
    cuts = sharded.shard_bounds(len(pos), 2)
    shards = [sharded.CohortShard(g[a:b], pos[a:b], rank=r, world=2) for r, (a, b) in enumerate(zip(cuts[:-1], cuts[1:]))]
    # a shard whose rows happen to miss the largest allele still has to use the same number of allele planes
    for name, fn in (("div", lambda c: c.diversity_totals(h1, windows)), ("hudson", lambda c: c.hudson_totals(h1, h2, windows))):
        ref = fn(whole)
        got = sharded.merge_in_rank_order([fn(s) for s in shards])
        assert np.array_equal(ref.u, got.u), name
        assert np.allclose(ref.f, got.f, rtol=1e-12, atol=1e-300), name
    for s in shards + [whole]:
        s.close()
