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
