"""Synthetic cohorts modelled on the reference's own generator
(src/pybenches/test_population_statistics_benchmarks.py:113-156): per-site base frequency
Beta(0.8, 0.8), per-population divergence N(0, sigma), alleles Bernoulli(p), positions = cumsum of
U{1..49} gaps; seed = variants + samples."""
import numpy as np


def make_cohort(n_variants, n_samples, n_pops=2, sigma=0.05, missing_rate=0.0, seed=None, dtype=np.int8):
    seed = n_variants + n_samples if seed is None else seed
    rng = np.random.default_rng(seed)
    base = rng.beta(0.8, 0.8, size=n_variants)
    bounds = np.linspace(0, n_samples, n_pops + 1).astype(int)
    g = np.zeros((n_variants, n_samples, 2), dtype=dtype)
    for p in range(n_pops):
        f = np.clip(base + rng.normal(0.0, sigma, size=n_variants), 0.001, 0.999)
        k = bounds[p + 1] - bounds[p]
        g[:, bounds[p]:bounds[p + 1], :] = rng.binomial(1, f[:, None, None], size=(n_variants, k, 2))
    if missing_rate > 0:
        g[rng.random(g.shape) < missing_rate] = -1
    positions = np.cumsum(rng.integers(1, 50, size=n_variants, dtype=np.int64)) if n_variants else \
        np.zeros(0, dtype=np.int64)
    pops = [list(range(bounds[p], bounds[p + 1])) for p in range(n_pops)]
    return g, positions, pops


def both_sides(samples):
    return [(int(s), side) for s in samples for side in (0, 1)]


# ---- CPU re-evaluation of the library's counter-based generator (csrc/fm_kernels.cuh: fm_k_synth)
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _mix(seed, a, b):
    with np.errstate(over="ignore"):
        return _splitmix64(_splitmix64(np.uint64(seed) + a.astype(np.uint64)) + b.astype(np.uint64))


def synth_rows(seed, v_lo, v_hi, n_samples, ploidy=2, pop_of_sample=None, sigma=0.05, missing_rate=0.0):
    """int8 genotypes [v_hi - v_lo, n_samples, ploidy] (-1 = missing) of sites [v_lo, v_hi) of the
    cohort fm_synth_fill(seed, ...) generates -- integer arithmetic only, bit-identical."""
    sigma_q = int(min(65536.0, max(0.0, sigma * 65536.0)))
    miss_q = int(min(65536.0, max(0.0, missing_rate * 65536.0 + 0.5)))
    stride = n_samples * ploidy
    v = np.arange(v_lo, v_hi, dtype=np.uint64)
    pops = np.zeros(n_samples, dtype=np.int64) if pop_of_sample is None else np.asarray(pop_of_sample, dtype=np.int64)
    hs = _mix(seed, v, np.zeros_like(v))
    x = (hs & np.uint64(0xFFFF)).astype(np.int64)
    y = (x * x) >> 16
    y = np.where(((hs >> np.uint64(16)) & np.uint64(1)) == 1, 65535 - y, y)
    thr = np.zeros((len(v), int(pops.max()) + 1 if len(pops) else 1), dtype=np.int64)
    for p in range(thr.shape[1]):
        hp = _mix(seed, v, np.full_like(v, 1 + p))
        num = ((hp & np.uint64(0x1FFF)).astype(np.int64) - 4096) * sigma_q
        d = np.where(num >= 0, num // 4096, -((-num) // 4096))  # C division truncates toward zero
        thr[:, p] = np.clip(y + d, 66, 65470)
    cols = np.arange(stride, dtype=np.uint64)
    out = np.zeros((len(v), stride), dtype=np.int8)
    step = max(1, (1 << 22) // max(stride, 1))
    col_pop = pops[(cols // np.uint64(ploidy)).astype(np.int64)]
    for r0 in range(0, len(v), step):
        r1 = min(len(v), r0 + step)
        he = _mix(seed, np.repeat(v[r0:r1], stride), np.tile(cols + np.uint64(0x10000), r1 - r0)).reshape(r1 - r0, stride)
        miss = ((he >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int64) < miss_q
        allele = (he & np.uint64(0xFFFF)).astype(np.int64) < thr[r0:r1][:, col_pop]
        out[r0:r1] = np.where(miss, -1, allele.astype(np.int8))
    return out.reshape(len(v), n_samples, ploidy)
