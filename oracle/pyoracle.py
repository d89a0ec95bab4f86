"""ctypes front-end of the CPU ORACLE (test infrastructure, NOT product code).

Wraps oracle/_build/libferromic_oracle.so (built by oracle/Makefile from
ferromic_oracle.c, a cited restatement of the reference's src/stats.rs).  Only
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libferromic_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ferromic_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _Variants(C.Structure):
    _fields_ = [("n_variants", C.c_size_t), ("n_samples", C.c_size_t), ("stride", C.c_size_t),
                ("positions", C.c_void_p), ("gt", C.c_void_p)]


class _Dense(C.Structure):
    _fields_ = [("data", C.c_void_p), ("missing", C.c_void_p), ("n_variants", C.c_size_t),
                ("n_samples", C.c_size_t), ("ploidy", C.c_size_t), ("max_allele", C.c_uint8)]


class _Haps(C.Structure):
    _fields_ = [("sample", C.c_void_p), ("side", C.c_void_p), ("n", C.c_size_t)]


class _Summary(C.Structure):
    _fields_ = [("alt", C.c_void_p), ("called", C.c_void_p), ("len", C.c_size_t),
                ("capacity", C.c_size_t), ("seg", C.c_size_t), ("pi_sum", C.c_double)]


class _Pop(C.Structure):
    _fields_ = [("haps", _Haps), ("variants", C.POINTER(_Variants)), ("n_sample_names", C.c_size_t),
                ("L", C.c_int64), ("dense", C.POINTER(_Dense)), ("summary", C.POINTER(_Summary))]


class _Opt(C.Structure):
    _fields_ = [("v", C.c_double), ("some", C.c_int)]

    def get(self):
        return self.v if self.some else None


class _HudsonSite(C.Structure):
    _fields_ = [("position", C.c_int64), ("fst", _Opt), ("d_xy", _Opt), ("pi1", _Opt), ("pi2", _Opt),
                ("num", _Opt), ("den", _Opt), ("n1", C.c_uint64), ("n2", C.c_uint64)]


class _HudsonOutcome(C.Structure):
    _fields_ = [("fst", _Opt), ("d_xy", _Opt), ("pi1", _Opt), ("pi2", _Opt), ("pi_xy_avg", _Opt)]


class _FstEstimate(C.Structure):
    _fields_ = [("state", C.c_int), ("value", C.c_double), ("sum_a", C.c_double),
                ("sum_b", C.c_double), ("sites", C.c_uint64)]


STATE_NAMES = ("calculable", "components_yield_indeterminate_ratio",
               "no_inter_population_variance", "insufficient_data_for_estimation")

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_harmonic.restype = C.c_double
        L.orc_harmonic.argtypes = [C.c_size_t]
        L.orc_watterson_theta.restype = C.c_double
        L.orc_watterson_theta.argtypes = [C.c_size_t, C.c_size_t, C.c_int64]
        L.orc_pi_sparse.restype = C.c_double
        L.orc_pi_sparse.argtypes = [C.POINTER(_Variants), C.POINTER(_Haps), C.c_int64]
        L.orc_pi_for_population.restype = C.c_double
        L.orc_pi_for_population.argtypes = [C.POINTER(_Pop)]
        L.orc_pi_from_summary.restype = C.c_double
        L.orc_pi_from_summary.argtypes = [C.POINTER(_Summary), C.c_int64, C.c_int, C.c_double]
        L.orc_count_segregating_sites.restype = C.c_size_t
        L.orc_count_segregating_sites.argtypes = [C.POINTER(_Variants)]
        L.orc_count_segregating_sites_for_population.restype = C.c_size_t
        L.orc_count_segregating_sites_for_population.argtypes = [C.POINTER(_Pop)]
        L.orc_dense_from_variants.restype = C.c_int
        L.orc_dense_membership.restype = C.c_size_t
        L.orc_per_site_diversity.restype = C.c_size_t
        L.orc_hudson_pair.restype = C.c_int
        L.orc_hudson_per_site.restype = C.c_size_t
        L.orc_dxy_hudson.restype = C.c_int
        L.orc_aggregate_hudson_from_sites.restype = _Opt
        L.orc_adjusted_sequence_length.restype = C.c_int64
        L.orc_fst_estimate_from_components.restype = _FstEstimate
        L.orc_fst_estimate_from_components.argtypes = [C.c_double, C.c_double]
        L.orc_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- data model
@dataclass
class Variants:
    """Sparse variants (process.rs:430-536): gt[V,S,stride] with 0xFF sentinel."""
    positions: np.ndarray
    gt: np.ndarray

    def __post_init__(self):
        self.positions = np.ascontiguousarray(self.positions, dtype=np.int64)
        self.gt = np.ascontiguousarray(self.gt, dtype=np.uint8)
        assert self.gt.ndim == 3 and self.gt.shape[0] == self.positions.shape[0]
        self._c = _Variants(self.gt.shape[0], self.gt.shape[1], self.gt.shape[2],
                            _ptr(self.positions), _ptr(self.gt))

    @property
    def n_variants(self):
        return self.gt.shape[0]

    @property
    def n_samples(self):
        return self.gt.shape[1]

    def c(self):
        return C.byref(self._c)

    def cptr(self):
        return C.pointer(self._c)


def variants_from_python(variants: Sequence, n_samples: Optional[int] = None) -> Variants:
    """Python variant mappings/tuples -> Variants (lib.rs:834-873, 1301-1332)."""
    pos, rows = [], []
    for v in variants:
        if isinstance(v, tuple):
            p, g = v
        else:
            p = v.get("position", v.get("pos", v.get("site")))
            g = v.get("genotypes", v.get("calls"))
        pos.append(int(p))
        rows.append(list(g))
    S = n_samples if n_samples is not None else (max((len(r) for r in rows), default=0))
    stride = 1
    for r in rows:
        for g in r:
            if g is not None and not isinstance(g, (int, np.integer)):
                stride = max(stride, len(g))
    gt = np.full((len(rows), S, stride), 0xFF, dtype=np.uint8)
    for vi, r in enumerate(rows):
        for si, g in enumerate(r):
            if g is None:
                continue
            if isinstance(g, (int, np.integer)):
                g = [g]
            for k, a in enumerate(g):
                gt[vi, si, k] = a
    return Variants(np.asarray(pos, dtype=np.int64), gt)


@dataclass
class Dense:
    """DenseGenotypeMatrix (stats.rs:249-331)."""
    data: np.ndarray
    missing: Optional[np.ndarray]
    n_variants: int
    n_samples: int
    ploidy: int
    max_allele: int

    def __post_init__(self):
        self.data = np.ascontiguousarray(self.data, dtype=np.uint8).reshape(-1)
        assert self.data.size == self.n_variants * self.n_samples * self.ploidy
        if self.missing is not None:
            self.missing = np.ascontiguousarray(self.missing, dtype=np.uint64)
        self._c = _Dense(_ptr(self.data), _ptr(self.missing), self.n_variants, self.n_samples,
                         self.ploidy, self.max_allele)

    def cptr(self):
        return C.pointer(self._c)


def pack_missing_bits(mask_flat: np.ndarray) -> np.ndarray:
    """bool[total] -> u64 bitmap, LSB-first (stats.rs:476-486 / lib.rs:1188-1190)."""
    total = mask_flat.size
    words = (total + 63) // 64
    padded = np.zeros(words * 64, dtype=np.uint8)
    padded[:total] = mask_flat.astype(np.uint8)
    return np.packbits(padded.reshape(words, 64), axis=1, bitorder="little").view(np.uint64).reshape(-1)


def from_numpy(genotypes: np.ndarray, positions) -> tuple[Variants, Optional[Dense]]:
    """convert_numeric_array (lib.rs:1135-1227): negative => missing.  Dense keeps per-allele
    missingness (bitmap only if anything is missing, dense only if ploidy==2); the sparse
    genotype of a sample is None when ANY allele is missing (lib.rs:1195-1199)."""
    g = np.asarray(genotypes)
    assert g.ndim == 3
    V, S, P = g.shape
    if g.dtype == np.uint8:
        miss = np.zeros(g.shape, dtype=bool)
    else:
        miss = g < 0
    alle = np.where(miss, 0, g).astype(np.uint8)
    gt = alle.copy()
    gt[np.broadcast_to(miss.any(axis=2, keepdims=True), g.shape)] = 0xFF
    if P == 0:
        gt = np.full((V, S, 1), 0xFF, dtype=np.uint8)
    variants = Variants(np.asarray(positions, dtype=np.int64), gt)
    dense = None
    if P == 2:
        missing = pack_missing_bits(miss.reshape(-1)) if miss.any() else None
        max_allele = int(alle.max()) if alle.size else 0
        dense = Dense(alle, missing, V, S, P, max_allele)
    return variants, dense


def dense_from_variants(vs: Variants, sample_count: int) -> Optional[Dense]:
    """DenseGenotypeMatrix::from_variants (stats.rs:339-500)."""
    data, missing = C.c_void_p(), C.c_void_p()
    ploidy, max_allele = C.c_size_t(), C.c_uint8()
    rc = lib().orc_dense_from_variants(vs.c(), C.c_size_t(sample_count), C.byref(data), C.byref(missing),
                                       C.byref(ploidy), C.byref(max_allele))
    if rc != 0:
        return None
    total = vs.n_variants * sample_count * ploidy.value
    d = np.ctypeslib.as_array(C.cast(data, C.POINTER(C.c_uint8)), shape=(max(total, 1),))[:total].copy()
    words = (total + 63) // 64
    m = np.ctypeslib.as_array(C.cast(missing, C.POINTER(C.c_uint64)), shape=(max(words, 1),))[:words].copy()
    lib().orc_free(data)
    lib().orc_free(missing)
    return Dense(d, m, vs.n_variants, sample_count, ploidy.value, max_allele.value)


class Haps:
    def __init__(self, haplotypes: Iterable):
        hs = list(haplotypes)
        self.sample = np.asarray([h[0] for h in hs], dtype=np.uint64)
        self.side = np.asarray([h[1] for h in hs], dtype=np.uint8)
        self._c = _Haps(_ptr(self.sample), _ptr(self.side), len(hs))

    def c(self):
        return C.byref(self._c)


@dataclass
class Summary:
    alt: np.ndarray
    called: np.ndarray
    capacity: int
    seg: int
    pi_sum: float

    def __post_init__(self):
        self.alt = np.ascontiguousarray(self.alt, dtype=np.uint32)
        self.called = np.ascontiguousarray(self.called, dtype=np.uint32)
        self._c = _Summary(_ptr(self.alt), _ptr(self.called), self.alt.size, self.capacity, self.seg,
                           self.pi_sum)

    def cptr(self):
        return C.pointer(self._c)


def dense_membership(dense: Dense, haplotypes) -> np.ndarray:
    h = Haps(haplotypes)
    out = np.zeros(max(len(h.sample), 1), dtype=np.uint64)
    n = lib().orc_dense_membership(dense.cptr(), h.c(), _ptr(out))
    return out[:n].copy()


def build_summary(dense: Dense, haplotypes, threads: int = 0) -> Summary:
    h = Haps(haplotypes)
    alt = np.zeros(dense.n_variants, dtype=np.uint32)
    called = np.zeros(dense.n_variants, dtype=np.uint32)
    s = _Summary(_ptr(alt), _ptr(called), dense.n_variants, 0, 0, 0.0)
    if threads > 0:
        lib().orc_build_summary_mt(dense.cptr(), h.c(), C.byref(s), C.c_int(threads))
    else:
        lib().orc_build_summary(dense.cptr(), h.c(), C.byref(s))
    return Summary(alt, called, s.capacity, s.seg, s.pi_sum)


class Pop:
    """PopulationContext (stats.rs:230-247)."""

    def __init__(self, haplotypes, variants: Optional[Variants], n_sample_names: int, L: int,
                 dense: Optional[Dense] = None, summary: Optional[Summary] = None):
        self.h = Haps(haplotypes)
        self.variants, self.dense, self.summary = variants, dense, summary
        self._c = _Pop(self.h._c, variants.cptr() if variants is not None else None, n_sample_names, L,
                       dense.cptr() if dense is not None else None,
                       summary.cptr() if summary is not None else None)

    def c(self):
        return C.byref(self._c)


# --------------------------------------------------------------------------- estimators
def harmonic(n: int) -> float:
    return lib().orc_harmonic(n)


def watterson_theta(seg: int, n: int, L: int) -> float:
    return lib().orc_watterson_theta(seg, n, L)


def count_segregating_sites(vs: Variants) -> int:
    return lib().orc_count_segregating_sites(vs.c())


def count_segregating_sites_for_population(p: Pop) -> int:
    return lib().orc_count_segregating_sites_for_population(p.c())


def pi_sparse(vs: Variants, haplotypes, L: int) -> float:
    h = Haps(haplotypes)  # keep the arrays alive across the call (a temporary would be freed before it)
    return lib().orc_pi_sparse(vs.c(), h.c(), L)


def pi_for_population(p: Pop) -> float:
    return lib().orc_pi_for_population(p.c())


def per_site_diversity(vs: Variants, haplotypes, region, filtered=(), mask=None):
    n = max(vs.n_variants, 1)
    pos = np.zeros(n, dtype=np.int64)
    pi = np.zeros(n, dtype=np.float64)
    th = np.zeros(n, dtype=np.float64)
    filt = np.asarray(list(filtered), dtype=np.int64)
    miv = np.asarray(mask if mask is not None else [], dtype=np.int64).reshape(-1)
    h = Haps(haplotypes)  # keep the arrays alive across the call
    k = lib().orc_per_site_diversity(vs.c(), h.c(), C.c_int64(region[0]),
                                     C.c_int64(region[1]), _ptr(filt), C.c_size_t(filt.size), _ptr(miv),
                                     C.c_size_t(miv.size // 2), C.c_int(mask is not None), _ptr(pos),
                                     _ptr(pi), _ptr(th))
    return pos[:k].copy(), pi[:k].copy(), th[:k].copy()


def _sites_to_dicts(arr, n):
    out = []
    for i in range(n):
        s = arr[i]
        out.append(dict(position=s.position, fst=s.fst.get(), d_xy=s.d_xy.get(), pi_pop1=s.pi1.get(),
                        pi_pop2=s.pi2.get(), n1_called=s.n1, n2_called=s.n2,
                        numerator_component=s.num.get(), denominator_component=s.den.get()))
    return out


def hudson_pair(p1: Pop, p2: Pop, region=None):
    """Returns (rc, outcome dict, sites list)."""
    V = p1.variants.n_variants if p1.variants is not None else 0
    sites = (_HudsonSite * max(V, 1))()
    out = _HudsonOutcome()
    n = C.c_size_t()
    rs, re = region if region is not None else (0, 0)
    rc = lib().orc_hudson_pair(p1.c(), p2.c(), C.c_int(region is not None), C.c_int64(rs), C.c_int64(re),
                               C.byref(out), sites, C.byref(n))
    outcome = dict(fst=out.fst.get(), d_xy=out.d_xy.get(), pi_pop1=out.pi1.get(), pi_pop2=out.pi2.get(),
                   pi_xy_avg=out.pi_xy_avg.get())
    return rc, outcome, _sites_to_dicts(sites, n.value)


def hudson_per_site(p1: Pop, p2: Pop, region):
    V = p1.variants.n_variants if p1.variants is not None else 0
    sites = (_HudsonSite * max(V, 1))()
    n = lib().orc_hudson_per_site(p1.c(), p2.c(), C.c_int64(region[0]), C.c_int64(region[1]), sites)
    return _sites_to_dicts(sites, n)


def dxy_hudson(p1: Pop, p2: Pop):
    o = _Opt()
    rc = lib().orc_dxy_hudson(p1.c(), p2.c(), C.byref(o))
    return rc, o.get()


def _est(e: _FstEstimate):
    return dict(state=STATE_NAMES[e.state], value=(e.value if e.state == 0 else None), sum_a=e.sum_a,
                sum_b=e.sum_b, sites=int(e.sites))


def fst_estimate_from_components(a: float, b: float):
    return _est(lib().orc_fst_estimate_from_components(a, b))


def variance_components(n, p, global_p):
    n = np.asarray(n, dtype=np.uint64)
    p = np.asarray(p, dtype=np.float64)
    a, b = C.c_double(), C.c_double()
    lib().orc_variance_components(_ptr(n), _ptr(p), C.c_size_t(n.size), C.c_double(global_p), C.byref(a),
                                  C.byref(b))
    return a.value, b.value


def wc_fst(vs: Variants, left: np.ndarray, right: np.ndarray, G: int, region, want_pairs: bool = True):
    """Returns dict with per-site arrays and region estimates (pair order i<j)."""
    left = np.ascontiguousarray(left, dtype=np.uint16)
    right = np.ascontiguousarray(right, dtype=np.uint16)
    V = max(vs.n_variants, 1)
    npairs = G * (G - 1) // 2 if G > 0 else 0
    pos = np.zeros(V, dtype=np.int64)
    state = np.zeros(V, dtype=np.int32)
    sa = np.zeros(V, dtype=np.float64)
    sb = np.zeros(V, dtype=np.float64)
    sizes = np.zeros((V, max(G, 1)), dtype=np.uint64)
    has_maps = np.zeros(V, dtype=np.uint8)
    pa = np.zeros((V, max(npairs, 1)), dtype=np.float64) if want_pairs else None
    pb = np.zeros((V, max(npairs, 1)), dtype=np.float64) if want_pairs else None
    ps = np.zeros((V, max(npairs, 1)), dtype=np.int32) if want_pairs else None
    overall = _FstEstimate()
    pairs = (_FstEstimate * max(npairs, 1))()
    present = np.zeros(max(npairs, 1), dtype=np.uint8)
    n = C.c_size_t()
    lib().orc_wc_fst(vs.c(), _ptr(left), _ptr(right), C.c_size_t(G), C.c_int64(region[0]),
                     C.c_int64(region[1]), C.byref(n), _ptr(pos), _ptr(state), _ptr(sa), _ptr(sb),
                     _ptr(sizes), _ptr(has_maps), _ptr(pa), _ptr(pb), _ptr(ps), C.byref(overall), pairs,
                     _ptr(present))
    k = n.value
    res = dict(n_sites=k, position=pos[:k], state=state[:k], a=sa[:k], b=sb[:k], pop_sizes=sizes[:k, :G],
               has_maps=has_maps[:k], overall=_est(overall),
               pairs=[_est(pairs[i]) for i in range(npairs)], pair_present=present[:npairs].astype(bool))
    if want_pairs:
        res.update(pair_a=pa[:k, :npairs], pair_b=pb[:k, :npairs], pair_state=ps[:k, :npairs])
    return res


def adjusted_sequence_length(start: int, end: int, allow=None, mask=None) -> int:
    a = np.asarray(allow if allow is not None else [], dtype=np.int64).reshape(-1)
    m = np.asarray(mask if mask is not None else [], dtype=np.int64).reshape(-1)
    return lib().orc_adjusted_sequence_length(C.c_int64(start), C.c_int64(end), _ptr(a),
                                              C.c_size_t(a.size // 2), C.c_int(allow is not None), _ptr(m),
                                              C.c_size_t(m.size // 2), C.c_int(mask is not None))
