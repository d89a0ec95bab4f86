"""World-size-2 `gloo` test of the multi-GPU host logic (ferromic_b200/sharded.py) on CPU:
site-range sharding, ONE all_gather of the packed window totals, rank-ordered merge, and the
library's host-side finishing entry points.  Local shard totals come from the CPU oracle here
(no device in this container); on the GPU box tests/test_gpu_sharded.py feeds the same merge
with totals computed by the kernels."""
import os
import socket

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.synth import both_sides, make_cohort

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def oracle_diversity_totals(summary, pos, windows):
    """Window totals of one group from the oracle's summary arrays (stats.rs:1367-1470)."""
    from ferromic_b200.sharded import WindowTotals
    n = len(windows)
    f = np.zeros((n, 1))
    u = np.zeros((n, 3), dtype=np.uint64)
    for w, (ws, we) in enumerate(windows):
        sel = (pos >= ws) & (pos <= we)
        a, c = summary.alt[sel].astype(np.int64), summary.called[sel].astype(np.int64)
        ok = c >= 2
        nn, al = c[ok].astype(np.float64), a[ok].astype(np.float64)
        rf = nn - al
        terms = nn / (nn - 1.0) * (1.0 - (rf * rf + al * al) / (nn * nn))
        f[w, 0] = float(np.sum(terms)) if terms.size else 0.0
        u[w] = (sel.sum(), int(((c >= 2) & (a > 0) & (a < c)).sum()), int((c < 2).sum()))
    return WindowTotals(f, u, ("pi_sum",), ("n_variants", "seg_sites", "uncallable"))


def oracle_hudson_totals(s1, s2, pos, windows):
    """aggregate_hudson_components_from_summaries (stats.rs:1554-1623) per window, in numpy."""
    from ferromic_b200.sharded import WindowTotals
    n = len(windows)
    f = np.zeros((n, 5))
    u = np.zeros((n, 3), dtype=np.uint64)
    for w, (ws, we) in enumerate(windows):
        for v in np.nonzero((pos >= ws) & (pos <= we))[0]:
            n1, a1, n2, a2 = int(s1.called[v]), int(s1.alt[v]), int(s2.called[v]), int(s2.alt[v])
            u[w, 1] += n1 < 2
            u[w, 2] += n2 < 2
            if n1 == 0 or n2 == 0:
                u[w, 0] += 1
                continue
            dxy = min(max((a1 * (n2 - a2) + (n1 - a1) * a2) / float(n1 * n2), 0.0), 1.0)
            f[w, 2] += dxy
            if n1 < 2 or n2 < 2:
                continue
            p1 = 2.0 * a1 * (n1 - a1) / float(n1 * (n1 - 1))
            p2 = 2.0 * a2 * (n2 - a2) / float(n2 * (n2 - 1))
            f[w, 3] += p1
            f[w, 4] += p2
            if dxy > 1e-12:
                f[w, 0] += dxy - 0.5 * (p1 + p2)
                f[w, 1] += dxy
    return WindowTotals(f, u, ("num", "den", "dxy", "pi1", "pi2"), ("dxy_uncallable", "unc1", "unc2"))


def oracle_wc_totals(vs, left, right, G, windows):
    from ferromic_b200.sharded import WindowTotals
    npairs = G * (G - 1) // 2
    n = len(windows)
    f = np.zeros((n, 2 + 2 * npairs))
    u = np.zeros((n, 2 + npairs), dtype=np.uint64)
    for w, (ws, we) in enumerate(windows):
        r = orc.wc_fst(vs, left, right, G, (int(ws), int(we)))
        f[w, 0], f[w, 1] = r["overall"]["sum_a"], r["overall"]["sum_b"]
        u[w, 0] = r["n_sites"]
        u[w, 1] = r["overall"]["sites"] if r["overall"]["state"] != orc.STATE_NAMES[3] else 0
        for k in range(npairs):
            e = r["pairs"][k]
            informative = r["pair_present"][k] and e["state"] != orc.STATE_NAMES[3]
            f[w, 2 + k] = e["sum_a"] if informative else 0.0
            f[w, 2 + npairs + k] = e["sum_b"] if informative else 0.0
            u[w, 2 + k] = e["sites"] if informative else 0
    f_names = ("overall_a", "overall_b") + tuple(f"pair_a_{k}" for k in range(npairs)) + \
        tuple(f"pair_b_{k}" for k in range(npairs))
    u_names = ("n_variants", "overall_sites") + tuple(f"pair_sites_{k}" for k in range(npairs))
    return WindowTotals(f, u, f_names, u_names)


def _cohort():
    g, pos, pops = make_cohort(3000, 24, n_pops=3, sigma=0.08, missing_rate=0.06, seed=777)
    g[:, :, 1][g[:, :, 0] < 0] = -1  # whole-sample missingness: dense and sparse semantics agree
    g[:, :, 0][g[:, :, 1] < 0] = -1
    g[100:140] = -1                  # a stretch without data (uncallable sites, InsufficientData)
    windows = np.array([(int(pos[0]), int(pos[999])), (int(pos[1000]), int(pos[2500])),
                        (int(pos[2501]), int(pos[-1])), (int(pos[-1]) + 10, int(pos[-1]) + 500)], dtype=np.int64)
    left = np.full(24, 0xFFFF, dtype=np.uint16)
    for p, members in enumerate(pops):
        left[members] = p
    return g, pos, pops, windows, left, left.copy()


def _worker(rank, port, q):
    import torch.distributed as dist

    from ferromic_b200 import sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        g, pos, pops, windows, left, right = _cohort()
        lo, hi = sharded.shard_range(len(pos), WORLD, rank, align=512)
        gs, ps = g[lo:hi], pos[lo:hi]
        vs, d = orc.from_numpy(gs, ps)
        h1, h2 = both_sides(pops[0]), both_sides(pops[1])
        s1, s2 = orc.build_summary(d, h1), orc.build_summary(d, h2)
        div = sharded.all_gather_totals(oracle_diversity_totals(s1, ps, windows))
        hud = sharded.all_gather_totals(oracle_hudson_totals(s1, s2, ps, windows))
        wc = sharded.all_gather_totals(oracle_wc_totals(vs, left, right, 3, windows))
        lengths = [int(we - ws + 1) for ws, we in windows]
        pi, theta = sharded.finish_diversity(div, lengths, len(h1))
        hres = sharded.finish_hudson(hud, lengths, len(h1), len(h2))
        ov, prs = sharded.finish_wc(wc, 3)
        q.put((rank, dict(pi=pi.tolist(), theta=theta.tolist(), seg=div.col("seg_sites").tolist(),
                          hud=hres, div_f=div.f.tolist(), wc_f=wc.f.tolist(),
                          wc=[(e.state, e.value, e.sum_a, e.sum_b, e.sites) for e in ov],
                          wcp=[[(e.state, e.value, e.sum_a, e.sum_b, e.sites) for e in row] for row in prs])))
    finally:
        dist.destroy_process_group()


def _close(a, b, rel=1e-11):
    if a is None or b is None:
        return a is None and b is None
    if a != a or b != b:
        return a != a and b != b
    return abs(a - b) <= rel * max(abs(a), abs(b), 1e-300)


def test_shard_bounds_are_aligned_and_cover():
    from ferromic_b200.sharded import shard_bounds
    for n in (0, 1, 8191, 8192, 100_000, 10_000_000):
        for world in (1, 2, 4, 8):
            b = shard_bounds(n, world)
            assert b[0] == 0 and b[-1] == n and len(b) == world + 1
            assert all(x <= y for x, y in zip(b, b[1:]))
            assert all(x % 8192 == 0 or x == n for x in b[1:-1])
    b = shard_bounds(10_000_000, 8)
    sizes = [y - x for x, y in zip(b, b[1:])]
    assert max(sizes) - min(sizes) <= 8192


def test_two_rank_gather_matches_whole_cohort_oracle():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=90) for _ in range(WORLD))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank ends with bit-identical merged results
    assert got[0] == got[1] or repr(got[0]) == repr(got[1])
    res = got[0]

    g, pos, pops, windows, left, right = _cohort()
    vs, d = orc.from_numpy(g, pos)
    h1, h2 = both_sides(pops[0]), both_sides(pops[1])
    s1, s2 = orc.build_summary(d, h1), orc.build_summary(d, h2)
    for w, (ws, we) in enumerate(windows):
        sel = np.nonzero((pos >= ws) & (pos <= we))[0]
        L = int(we - ws + 1)
        sub1 = orc.Summary(s1.alt[sel], s1.called[sel], s1.capacity, 0, 0.0)
        sub2 = orc.Summary(s2.alt[sel], s2.called[sel], s2.capacity, 0, 0.0)
        # re-derive the summary scalars the way build_dense_population_summary does
        a, c = sub1.alt.astype(np.int64), sub1.called.astype(np.int64)
        seg = int(((c >= 2) & (a > 0) & (a < c)).sum())
        assert res["seg"][w] == seg
        assert _close(res["theta"][w], orc.watterson_theta(seg, len(h1), L))
        rc, ref, _ = orc.hudson_pair(orc.Pop(h1, None, 24, L, summary=sub1), orc.Pop(h2, None, 24, L, summary=sub2))
        assert rc == 0
        for k in ("fst", "d_xy", "pi_pop1", "pi_pop2", "pi_xy_avg"):
            assert _close(res["hud"][w][k], ref[k]), (w, k, res["hud"][w][k], ref[k])
        wref = orc.wc_fst(vs, left, right, 3, (int(ws), int(we)))
        st, val, sa, sb, sites = res["wc"][w]
        assert orc.STATE_NAMES[st] == wref["overall"]["state"] and sites == wref["overall"]["sites"]
        assert _close(sa, wref["overall"]["sum_a"]) and _close(sb, wref["overall"]["sum_b"])
        if st == 0:
            assert _close(val, wref["overall"]["value"])
        for k in range(3):
            pst, pval, pa, pb, psites = res["wcp"][w][k]
            if not wref["pair_present"][k]:
                assert pst == 3 and psites == 0
                continue
            e = wref["pairs"][k]
            assert orc.STATE_NAMES[pst] == e["state"] and psites == e["sites"], (w, k)
            assert _close(pa, e["sum_a"]) and _close(pb, e["sum_b"])
