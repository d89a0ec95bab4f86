"""The FP64 building blocks that let K4 (csrc/fm_wc.cuh) drop the per-pair divisions of
calculate_variance_components (stats.rs:2034-2127) without changing a rounding: fm_div_recip (a / b from the
correctly rounded reciprocal, two residual corrections) and fm_recip_rn (branch-free reciprocal) must equal
IEEE division bit for bit on the divisors the kernel meets."""
import numpy as np
import pytest

from ferromic_b200 import _lib

pytestmark = pytest.mark.gpu


def probe(a, b):
    L = _lib.lib()
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    y, q, q3 = np.empty_like(a), np.empty_like(a), np.empty_like(a)
    _lib.check(L.fm_wc_arith_probe(a.ctypes.data, b.ctypes.data, y.ctypes.data, q.ctypes.data, q3.ctypes.data, a.size))
    probe.q_int = q3
    return y, q


def test_division_by_table_reciprocal_is_ieee_division():
    rng = np.random.default_rng(5)
    n = 1 << 24
    k = rng.integers(1, 400_000, size=n).astype(np.float64)
    b = np.where(np.arange(n) % 3 == 0, k, np.where(np.arange(n) % 3 == 1, k * 0.5, k * k * 0.5))
    a = np.concatenate([rng.integers(0, 400_000, size=n // 2).astype(np.float64), rng.random(n - n // 2) * 3.0 - 1.0])
    _, q = probe(a, b)
    assert np.array_equal(q.view(np.uint64), (a / b).view(np.uint64))
    # the one-correction form K4 uses for integer-like divisors (n, n/2, n^2/2 with n < 2^19): provably exact there
    assert np.array_equal(probe.q_int.view(np.uint64), (a / b).view(np.uint64))


def test_branch_free_reciprocal_is_ieee_reciprocal():
    rng = np.random.default_rng(6)
    n = 1 << 24
    # the a-denominator 1 - c^2 lies in (0, 1]; cover [2^-20, 2) densely plus exact powers of two and 1 - ulp
    b = np.concatenate([rng.random(n) * 0.999 + 0.001, np.ldexp(rng.random(n // 4) + 1.0, rng.integers(-20, 1, n // 4)),
                        np.array([1.0, 0.5, 0.25, np.nextafter(0.5, 1.0), 0.75, 1.0 - 2.0 ** -30, 1.0 - 1.0 / 452.0 ** 2])])
    y, q = probe(b, b)
    assert np.array_equal(y.view(np.uint64), (1.0 / b).view(np.uint64))
    assert np.all(q == 1.0)
    # the documented exception (Markstein): a divisor whose significand is all ones; K4's divisor 1 - c^2 is either 1
    # or at most 1 - 1/(n_i + n_j)^2 and never has that form
    ones = np.array([np.nextafter(1.0, 0.0)])
    y1, _ = probe(ones, ones)
    assert y1[0] == 1.0 and 1.0 / ones[0] == np.nextafter(1.0, 2.0)


def test_wc_per_site_values_are_bit_identical_to_the_oracle():
    """Per-site overall and pairwise (a, b) of the GPU path against the C oracle, which divides with IEEE `/` in
    the reference's operation order (stats.rs:2034-2127): not a single bit may differ."""
    import ferromic_b200 as F
    from oracle import pyoracle as orc
    from tests.synth import make_cohort
    from tests.test_gpu_parity import _to_python_variants
    for n_pops, S, V, missing in ((2, 40, 1500, 0.1), (5, 60, 900, 0.05), (26, 130, 300, 0.08)):
        g, pos, pops = make_cohort(V, S, n_pops=n_pops, sigma=0.08, missing_rate=missing, seed=300 + n_pops)
        g[:, :, 1][g[:, :, 0] < 0] = -1
        g[:, :, 0][g[:, :, 1] < 0] = -1
        left = np.full(S, 0xFFFF, dtype=np.uint16)
        for p, members in enumerate(pops):
            left[members] = p
        labels = sorted(str(i) for i in range(n_pops))
        vs, _ = orc.from_numpy(g, pos)
        region = (int(pos[0]), int(pos[-1]))
        ref = orc.wc_fst(vs, left, left, n_pops, region)
        got = F.wc_fst_from_membership(_to_python_variants(g, pos), labels, left, left, region)
        keys = [f"{labels[i]}_vs_{labels[j]}" for i in range(n_pops) for j in range(i + 1, n_pops)]
        n_checked = 0
        for i, s in enumerate(got.site_fst):
            assert s.variance_components_a == float(ref["a"][i]) and s.variance_components_b == float(ref["b"][i])
            if not ref["has_maps"][i]:
                continue
            for k, key in enumerate(keys):
                a, b = s.pairwise_variance_components[key]
                ra, rb = float(ref["pair_a"][i][k]), float(ref["pair_b"][i][k])
                assert (a == ra or (np.isnan(a) and np.isnan(ra))) and (b == rb or (np.isnan(b) and np.isnan(rb))), \
                    (i, key, a, ra, b, rb)
                n_checked += 1
        assert n_checked > 0
