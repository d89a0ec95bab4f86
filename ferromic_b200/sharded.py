"""Multi-GPU layer of the per-site estimator path: one process per GPU, sites sharded by
contiguous range, window/region totals merged by ONE gather of a tiny buffer.

The path shards by site range: every per-site quantity depends only on its own
row, region/window results are sums and integer counts over sites.  Each rank holds the
bitplanes of its own site range and evaluates window totals for the windows its shard touches
(`fm_group_window_sums`, `fm_hudson_window_sums`, `fm_wc_window_sums`).  The totals of all ranks
are exchanged with a single `all_gather` (NCCL over NVLink on GPUs, gloo in the CPU tests) and
added **in rank order** on every rank -- not `all_reduce`, whose association depends on the
algorithm NCCL picks -- so every rank ends with bit-identical results that are independent of
the collective implementation.  Shard boundaries are multiples of 8192 sites (the kernels'
super-batch) so per-batch partials are the same ones a single GPU would produce.

Finishing (pi = sum / (L - uncallable), theta, FST ladders) runs through the library's host
entry points (`fm_pi_from_sums`, `fm_watterson_theta`, `fm_hudson_outcome_from_sums`,
`fm_fst_estimate_from_sums`), i.e. the same code the single-GPU calls use.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check, lib

SHARD_ALIGN = 8192  # sites; == kSuperBatches * 32 in csrc/fm_kernels.cuh and a multiple of kWcSegSites


def shard_bounds(n_sites: int, world: int, align: int = SHARD_ALIGN) -> List[int]:
    """Cut points of `world` contiguous site shards, balanced and rounded to `align` sites."""
    cuts = [0]
    for r in range(1, world):
        c = (n_sites * r // world + align // 2) // align * align
        cuts.append(min(max(c, cuts[-1]), n_sites))
    cuts.append(n_sites)
    return cuts


def shard_range(n_sites: int, world: int, rank: int, align: int = SHARD_ALIGN) -> Tuple[int, int]:
    b = shard_bounds(n_sites, world, align)
    return b[rank], b[rank + 1]


# ------------------------------------------------------------------------------- totals
@dataclass
class WindowTotals:
    """Shard-mergeable totals of a list of windows: every field adds across site shards.
    f: float64 [n_windows, n_f]; u: uint64 [n_windows, n_u]; names index the columns."""
    f: np.ndarray
    u: np.ndarray
    f_names: Tuple[str, ...]
    u_names: Tuple[str, ...]

    def col(self, name: str) -> np.ndarray:
        if name in self.f_names:
            return self.f[:, self.f_names.index(name)]
        return self.u[:, self.u_names.index(name)]


def merge_in_rank_order(parts: Sequence[WindowTotals]) -> WindowTotals:
    """Sequential sum over ranks 0..N-1 (fixed association; integers are exact anyway)."""
    f = parts[0].f.copy()
    u = parts[0].u.copy()
    for p in parts[1:]:
        f = f + p.f
        u = u + p.u
    return WindowTotals(f, u, parts[0].f_names, parts[0].u_names)


class PeerComm:
    """The library's NVLink mailbox exchange (include/ferromic_gpu.h fm_comm_*, csrc/fm_comm.cuh):
    one small kernel stores this rank's words into every peer's mailbox with P2P stores and
    returns everybody's words.  Handles are swapped once through `exchange_handles`, a callable
    bytes -> list[bytes] over all ranks (torch.distributed.all_gather_object, MPI, ...)."""

    def __init__(self, rank: int, world: int, exchange_handles=None, device: Optional[int] = None):
        L = lib()
        if device is not None:
            check(L.fm_set_device(device))
        self.rank, self.world = rank, world
        self.handle = C.c_void_p()
        check(L.fm_comm_create(rank, world, C.byref(self.handle)))
        if world > 1:
            if exchange_handles is None:
                raise ValueError("exchange_handles is required for world > 1")
            hb = (C.c_uint8 * 64)()
            check(L.fm_comm_export(self.handle, hb))
            allh = exchange_handles(bytes(hb))
            buf = np.frombuffer(b"".join(allh), dtype=np.uint8).copy()
            check(L.fm_comm_connect(self.handle, buf.ctypes.data))

    @classmethod
    def connect_in_process(cls, comms: Sequence["PeerComm"]):
        """Wire communicators that live in one process (one host thread per GPU)."""
        arr = (C.c_void_p * len(comms))(*[c.handle.value for c in comms])
        for c in comms:
            check(lib().fm_comm_connect_local(c.handle, arr))

    @classmethod
    def _bare(cls, rank: int, world: int) -> "PeerComm":
        self = cls.__new__(cls)
        self.rank, self.world = rank, world
        self.handle = C.c_void_p()
        check(lib().fm_comm_create(rank, world, C.byref(self.handle)))
        return self

    def allgather_words(self, words: np.ndarray, n_double: int = 0):
        """words: 8-byte elements (float64 / uint64 view).  Returns (gathered [world, n], merged [n])
        as uint64 bit patterns; the first n_double words of `merged` are FP64 sums in rank order."""
        w = np.ascontiguousarray(words).view(np.uint64).reshape(-1)
        gathered = np.empty((self.world, w.size), dtype=np.uint64)
        merged = np.empty(w.size, dtype=np.uint64)
        check(lib().fm_comm_allgather(self.handle, w.ctypes.data, w.size, n_double, gathered.ctypes.data,
                                      merged.ctypes.data))
        return gathered, merged

    def close(self):
        if getattr(self, "handle", None):
            lib().fm_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def peer_gather_totals(local: WindowTotals, comm: PeerComm) -> WindowTotals:
    """all_gather_totals over the NVLink mailbox instead of torch.distributed (chunks of
    FM_COMM_MAX_WORDS words); the rank-ordered sum is the same merge_in_rank_order."""
    nf, nu = local.f.size, local.u.size
    buf = np.empty(nf + nu, dtype=np.uint64)
    buf[:nf] = local.f.reshape(-1).view(np.uint64)
    buf[nf:] = local.u.reshape(-1)
    rows = []
    for o in range(0, max(buf.size, 1), 2048):
        g, _ = comm.allgather_words(buf[o:o + 2048])
        rows.append(g)
    allw = np.concatenate(rows, axis=1) if rows else np.zeros((comm.world, 0), dtype=np.uint64)
    parts = [WindowTotals(allw[r, :nf].view(np.float64).reshape(local.f.shape).copy(),
                          allw[r, nf:].reshape(local.u.shape).copy(), local.f_names, local.u_names)
             for r in range(comm.world)]
    return merge_in_rank_order(parts)


def all_gather_totals(local: WindowTotals, group=None, device=None) -> WindowTotals:
    """One all_gather of the packed totals, then the rank-ordered sum.  `device` is the CUDA
    device of this rank for the NCCL backend (None: CPU tensors, gloo)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    nf, nu = local.f.size, local.u.size
    buf = np.empty(nf + nu, dtype=np.float64)
    buf[:nf] = local.f.reshape(-1)
    buf[nf:] = local.u.reshape(-1).view(np.float64)  # bit-cast, never interpreted as FP
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device, non_blocking=True)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    parts = []
    for o in out:
        a = o.cpu().numpy()
        parts.append(WindowTotals(a[:nf].reshape(local.f.shape).copy(),
                                  a[nf:].view(np.uint64).reshape(local.u.shape).copy(),
                                  local.f_names, local.u_names))
    return merge_in_rank_order(parts)


# ------------------------------------------------------------------------------- finishing
def finish_diversity(tot: WindowTotals, lengths: Sequence[int], haplotype_capacity: int):
    """pi and Watterson theta per window from merged totals (stats.rs:1480-1542, 4243-4307)."""
    L = lib()
    n = tot.f.shape[0]
    pi = np.empty(n)
    theta = np.empty(n)
    out = C.c_double()
    for w in range(n):
        check(L.fm_pi_from_sums(float(tot.col("pi_sum")[w]), int(tot.col("uncallable")[w]), int(lengths[w]),
                                haplotype_capacity, C.byref(out)))
        pi[w] = out.value
        check(L.fm_watterson_theta(int(tot.col("seg_sites")[w]), haplotype_capacity, int(lengths[w]), C.byref(out)))
        theta[w] = out.value
    return pi, theta


def finish_hudson(tot: WindowTotals, lengths: Sequence[int], cap1: int, cap2: int):
    """HudsonFSTOutcome per window from merged totals (summaries path, stats.rs:3476-3566).
    Returns a list of dicts with None for absent values."""
    L = lib()
    res = []
    for w in range(tot.f.shape[0]):
        s = _lib.HudsonSums(float(tot.col("num")[w]), float(tot.col("den")[w]), float(tot.col("dxy")[w]),
                            float(tot.col("pi1")[w]), float(tot.col("pi2")[w]),
                            int(tot.col("dxy_uncallable")[w]), int(tot.col("unc1")[w]), int(tot.col("unc2")[w]))
        o = _lib.HudsonOutcome()
        check(L.fm_hudson_outcome_from_sums(C.byref(s), int(lengths[w]), cap1, cap2, C.byref(o)))
        bit = lambda v, b: v if (o.some >> b) & 1 else None  # noqa: E731
        res.append(dict(fst=bit(o.fst, 0), d_xy=bit(o.d_xy, 1), pi_pop1=bit(o.pi_pop1, 2),
                        pi_pop2=bit(o.pi_pop2, 3), pi_xy_avg=bit(o.pi_xy_avg, 4)))
    return res


def finish_wc(tot: WindowTotals, n_pairs: int):
    """FstEstimate (overall, per pair) per window from merged W&C totals (stats.rs:2231-2356).
    Returns (overall[n_w], pairs[n_w][n_pairs]) of _lib.FstEstimateC."""
    L = lib()
    overall, pairs = [], []
    for w in range(tot.f.shape[0]):
        n_inf = int(tot.col("overall_sites")[w])
        n_var = int(tot.col("n_variants")[w])
        e = _lib.FstEstimateC()
        check(L.fm_fst_estimate_from_sums(float(tot.col("overall_a")[w]), float(tot.col("overall_b")[w]), n_inf,
                                          n_var, C.byref(e)))
        overall.append(e)
        row = []
        for k in range(n_pairs):
            pe = _lib.FstEstimateC()
            pn = int(tot.col(f"pair_sites_{k}")[w])
            # a pair that was never informative reports sites_attempted = #sites with maps
            check(L.fm_fst_estimate_from_sums(float(tot.col(f"pair_a_{k}")[w]), float(tot.col(f"pair_b_{k}")[w]),
                                              pn, n_inf, C.byref(pe)))
            row.append(pe)
        pairs.append(row)
    return overall, pairs


# ------------------------------------------------------------------------------- local shard
class CohortShard:
    """The site range [v_lo, v_hi) of a cohort resident on this rank's GPU.

    genotypes: u8/int8 array [v_hi - v_lo, S, ploidy] of the shard (negative = missing) or a
    pre-built api._Matrix; positions: 0-based positions of the shard's sites (ascending)."""

    def __init__(self, genotypes, positions, rank: int = 0, world: int = 1, device: Optional[int] = None,
                 comm: Optional[PeerComm] = None, max_allele: Optional[int] = None, reduce_max=None):
        """max_allele: the WHOLE cohort's largest allele index.  The reference picks its estimator forms
        (biallelic / general) from the whole matrix's max_allele (stats.rs:490, lib.rs:779); a shard that happens
        to hold only alleles 0/1 of a multi-allelic cohort must take the same forms as the others, or the merged
        totals would mix roundings.  When it is not given, `reduce_max` (a callable int -> int returning the maximum
        over all ranks, e.g. an all-reduce) is applied to the shard's own maximum; with neither, the shard's own
        maximum is used (single-shard use)."""
        from .api import _Matrix

        self.rank, self.world = rank, world
        self.comm = comm
        if device is not None:
            check(lib().fm_set_device(device))
        self.device = device
        if isinstance(genotypes, _Matrix):
            self.matrix = genotypes
        else:
            g = np.asarray(genotypes)
            miss = g < 0 if g.dtype.kind == "i" else None
            alle = np.where(miss, 0, g).astype(np.uint8) if miss is not None else g.astype(np.uint8)
            # max_allele > 1: the multi-allelic general forms (per-allele counts) serve the same window totals
            if max_allele is None:
                max_allele = int(alle.max()) if alle.size else 0
                if reduce_max is not None:
                    max_allele = int(reduce_max(max_allele))
            self.matrix = _Matrix(alle, miss, np.asarray(positions, dtype=np.int64), max_allele=int(max_allele))
        self.positions = np.asarray(positions, dtype=np.int64)
        self._partitions: Dict[int, C.c_void_p] = {}

    # ---- local totals (device work) ------------------------------------------------------
    def diversity_totals(self, haplotypes, windows: np.ndarray) -> WindowTotals:
        g = self.matrix.group(haplotypes)
        w = np.ascontiguousarray(windows, dtype=np.int64).reshape(-1, 2)
        n = len(w)
        nv = np.zeros(n, dtype=np.uint64)
        seg = np.zeros(n, dtype=np.uint64)
        unc = np.zeros(n, dtype=np.uint64)
        pis = np.zeros(n)
        if n:
            check(lib().fm_group_window_sums(g.handle, w.ctypes.data, n, nv.ctypes.data, seg.ctypes.data,
                                             pis.ctypes.data, unc.ctypes.data))
        return WindowTotals(pis.reshape(n, 1), np.stack([nv, seg, unc], axis=1), ("pi_sum",),
                            ("n_variants", "seg_sites", "uncallable"))

    def hudson_totals(self, haps1, haps2, windows: np.ndarray) -> WindowTotals:
        g1, g2 = self.matrix.groups([haps1, haps2])
        w = np.ascontiguousarray(windows, dtype=np.int64).reshape(-1, 2)
        n = len(w)
        f = np.zeros((5, n))
        sk = np.zeros(n, dtype=np.uint64)
        u1 = self.diversity_totals(haps1, w).col("uncallable")
        u2 = self.diversity_totals(haps2, w).col("uncallable")
        if n:
            check(lib().fm_hudson_window_sums(g1.handle, g2.handle, w.ctypes.data, n, f[0].ctypes.data,
                                              f[1].ctypes.data, f[2].ctypes.data, sk.ctypes.data,
                                              f[3].ctypes.data, f[4].ctypes.data))
        return WindowTotals(np.ascontiguousarray(f.T), np.stack([sk, u1, u2], axis=1),
                            ("num", "den", "dxy", "pi1", "pi2"), ("dxy_uncallable", "unc1", "unc2"))

    def wc_totals(self, left: np.ndarray, right: np.ndarray, n_groups: int, windows: np.ndarray) -> WindowTotals:
        key = hash((left.tobytes(), right.tobytes(), n_groups))
        ph = self._partitions.get(key)
        if ph is None:
            ph = C.c_void_p()
            lft = np.ascontiguousarray(left, dtype=np.uint16)
            rgt = np.ascontiguousarray(right, dtype=np.uint16)
            check(lib().fm_partition_create(self.matrix.handle, lft.ctypes.data, rgt.ctypes.data, len(lft),
                                            n_groups, C.byref(ph)))
            self._partitions[key] = ph
        w = np.ascontiguousarray(windows, dtype=np.int64).reshape(-1, 2)
        n = len(w)
        npairs = n_groups * (n_groups - 1) // 2
        nv = np.zeros(n, dtype=np.uint64)
        oa, ob = np.zeros(n), np.zeros(n)
        os_ = np.zeros(n, dtype=np.uint64)
        pa, pb = np.zeros((n, max(npairs, 1))), np.zeros((n, max(npairs, 1)))
        pn = np.zeros((n, max(npairs, 1)), dtype=np.uint64)
        if n:
            check(lib().fm_wc_window_sums(ph, w.ctypes.data, n, nv.ctypes.data, oa.ctypes.data, ob.ctypes.data,
                                          os_.ctypes.data, pa.ctypes.data, pb.ctypes.data, pn.ctypes.data))
        f = np.concatenate([oa[:, None], ob[:, None], pa[:, :npairs], pb[:, :npairs]], axis=1)
        u = np.concatenate([nv[:, None], os_[:, None], pn[:, :npairs]], axis=1)
        f_names = ("overall_a", "overall_b") + tuple(f"pair_a_{k}" for k in range(npairs)) + \
            tuple(f"pair_b_{k}" for k in range(npairs))
        u_names = ("n_variants", "overall_sites") + tuple(f"pair_sites_{k}" for k in range(npairs))
        return WindowTotals(np.ascontiguousarray(f), np.ascontiguousarray(u), f_names, u_names)

    # ---- merged results (one gather each) ------------------------------------------------
    def _gather(self, t: WindowTotals, group=None) -> WindowTotals:
        if self.comm is not None:
            return peer_gather_totals(t, self.comm)
        dev = None
        if self.device is not None:
            import torch
            dev = torch.device("cuda", self.device)
        return all_gather_totals(t, group=group, device=dev)

    def window_diversity(self, haplotypes, windows, lengths, group=None):
        t = self._gather(self.diversity_totals(haplotypes, windows), group)
        cap = self.matrix.group(haplotypes).capacity
        pi, theta = finish_diversity(t, lengths, cap)
        return dict(pi=pi, watterson_theta=theta, segregating_sites=t.col("seg_sites").copy(),
                    n_variants=t.col("n_variants").copy(), totals=t)

    def window_hudson(self, haps1, haps2, windows, lengths, group=None):
        t = self._gather(self.hudson_totals(haps1, haps2, windows), group)
        return finish_hudson(t, lengths, self.matrix.group(haps1).capacity, self.matrix.group(haps2).capacity), t

    def window_wc(self, left, right, n_groups, windows, group=None):
        t = self._gather(self.wc_totals(left, right, n_groups, windows), group)
        return finish_wc(t, n_groups * (n_groups - 1) // 2) + (t,)

    def close(self):
        for ph in self._partitions.values():
            lib().fm_partition_release(ph)
        self._partitions.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
