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
