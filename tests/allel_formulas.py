"""numpy restatement of the scikit-allel formulas the reference's Python tests use as an
independent oracle (scikit-allel is not installed here).  Published definitions:
allel.mean_pairwise_difference, allel.mean_pairwise_difference_between, allel.hudson_fst."""
import numpy as np


def count_alleles(genotypes, subpop=None, max_allele=None):
    g = np.asarray(genotypes)
    if subpop is not None:
        g = g[:, list(subpop)]
    V = g.shape[0]
    flat = g.reshape(V, -1)
    if max_allele is None:
        max_allele = int(flat.max()) if flat.size else 0
    ac = np.zeros((V, max_allele + 1), dtype=np.int64)
    for a in range(max_allele + 1):
        ac[:, a] = (flat == a).sum(axis=1)
    return ac


def mean_pairwise_difference(ac):
    ac = np.asarray(ac, dtype=np.float64)
    an = ac.sum(axis=1)
    n_pairs = an * (an - 1) / 2
    n_same = (ac * (ac - 1) / 2).sum(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(n_pairs > 0, (n_pairs - n_same) / n_pairs, np.nan)


def mean_pairwise_difference_between(ac1, ac2):
    ac1 = np.asarray(ac1, dtype=np.float64)
    ac2 = np.asarray(ac2, dtype=np.float64)
    n_pairs = ac1.sum(axis=1) * ac2.sum(axis=1)
    n_same = (ac1 * ac2).sum(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(n_pairs > 0, (n_pairs - n_same) / n_pairs, np.nan)


def hudson_fst(ac1, ac2):
    within = (mean_pairwise_difference(ac1) + mean_pairwise_difference(ac2)) / 2
    between = mean_pairwise_difference_between(ac1, ac2)
    return between - within, between


def weir_cockerham_ab_haploid(ac_by_pop):
    """Weir & Cockerham (1984) eqs. 2-3 as scikit-allel's `weir_cockerham_fst` evaluates them
    (a, b summed over alleles), for haploid observations: observed heterozygosity h-bar = 0, so
    c = 0.  `ac_by_pop` = list over populations of allele-count arrays [V, n_alleles].
    Populations without called haplotypes at a site are left out of that site (r counts the rest).
    Returns per-site (a, b) summed over alleles.  Independent of the oracle's code."""
    ac = np.stack([np.asarray(x, dtype=np.float64) for x in ac_by_pop])  # [P, V, A]
    n = ac.sum(axis=2)  # [P, V]
    P, V, A = ac.shape
    a_out, b_out = np.zeros(V), np.zeros(V)
    for v in range(V):
        use = n[:, v] > 0
        r = int(use.sum())
        if r < 2:
            continue
        ni = n[use, v]
        n_bar = ni.sum() / r
        if n_bar <= 1:
            continue
        n_c = (r * n_bar - (ni ** 2).sum() / (r * n_bar)) / (r - 1)
        for al in range(A):
            p = ac[use, v, al] / ni
            p_bar = (ni * p).sum() / (r * n_bar)
            s2 = (ni * (p - p_bar) ** 2).sum() / ((r - 1) * n_bar)
            h_bar = 0.0
            a = (n_bar / n_c) * (s2 - (1 / (n_bar - 1)) * (p_bar * (1 - p_bar) - (r - 1) / r * s2 - h_bar / 4))
            b = (n_bar / (n_bar - 1)) * (p_bar * (1 - p_bar) - (r - 1) / r * s2 - (2 * n_bar - 1) / (4 * n_bar) * h_bar)
            a_out[v] += a
            b_out[v] += b
    return a_out, b_out
