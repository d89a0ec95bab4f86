"""FALSTA per-site track writers (process.rs:3731-4002).

CPU: the oracle restatement against the two reference tests that read the files back, and the
library's token routine (host instance of the device code) against the oracle's formatter.
GPU (-m gpu): track bodies rendered by the device, byte-exact against the oracle."""
import gzip
import math
import struct

import numpy as np
import pytest

from oracle import falsta as ofa


# ----------------------------------------------------------------------------- oracle, pinned
def test_oracle_missing_sites_default_to_zero_diversity():
    """src/tests/stats_tests.rs:82-241: region 1..5, one variant at position 3, one sample 0|1 with
    both haplotypes in group 0 -> pi = theta = 1 at position 3, "0" elsewhere."""
    text = ofa.diversity_falsta_text("1", 1, 5, [(3, 1.0, 1.0, 0, False)])
    lines = text.splitlines()
    i = lines.index(">unfiltered_pi_chr_1_start_1_end_5_group_0")
    vals = lines[i + 1].split(",")
    assert len(vals) == 5 and vals[0] == vals[1] == vals[3] == vals[4] == "0" and vals[2] != "0"
    assert vals[2] == "1.000000"
    j = lines.index(">unfiltered_theta_chr_1_start_1_end_5_group_0")
    tv = lines[j + 1].split(",")
    assert tv == ["0", "0", "1.000000", "0", "0"]
    assert not any(l.startswith(">filtered_") for l in lines)  # tracks without a record are omitted


def test_oracle_per_site_falsta_includes_hudson_components():
    """src/tests/stats_tests.rs:1861-2034: FST 1, -1, 1; numerators 1, -0.5, 1; denominators 1, .5, 1."""
    hud = [(1, 1.0, 1.0, 1.0), (2, -1.0, -0.5, 0.5), (3, 1.0, 1.0, 1.0)]
    lines = ofa.fst_falsta_text("1", 1, 3, [], hud).splitlines()
    i = lines.index(">hudson_pairwise_fst_hap_0v1_chr_1_start_1_end_3")
    assert [float(x) for x in lines[i + 1].split(",")] == [1.0, -1.0, 1.0]
    i = lines.index(">hudson_pairwise_fst_hap_0v1_numerator_chr_1_start_1_end_3")
    assert [float(x) for x in lines[i + 1].split(",")] == [1.0, -0.5, 1.0]
    i = lines.index(">hudson_pairwise_fst_hap_0v1_denominator_chr_1_start_1_end_3")
    assert lines[i + 1] == "1.000000,0.500000,1.000000"


def test_oracle_region_clamping_and_overwrite():
    r = ofa.ZeroBasedHalfOpen.from_1based_inclusive(-5, 3)
    assert (r.start, r.end) == (0, 3)
    r = ofa.ZeroBasedHalfOpen.from_1based_inclusive(10, 4)
    assert (r.start, r.end) == (9, 10)
    assert r.relative_position_1based_inclusive(10) == 1 and r.relative_position_1based_inclusive(0) is None
    # the last record at a position wins
    t = ofa.fst_falsta_text("x", 1, 2, [], [(1, 0.25, 1.0, 4.0), (1, 0.5, 1.0, 2.0)])
    assert t.splitlines()[1] == "0.500000,NA"


# ------------------------------------------------------------------ token routine (host instance)
def _cases():
    rng = np.random.default_rng(7)
    xs = [0.0, -0.0, 1.0, -1.0, 0.5, 1e-7, 4.9e-7, 5e-7, 5.1e-7, 1e-6, 1.5e-6, 2.5e-6, 0.1, 0.2, 0.3, 1 / 3,
          2 / 3, 123456.789, 1e15, 1e20, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308 / 2 ** 930,
          0.9999995, 0.99999949999, 0.9999994999999999, 9.9999995, 99999.9999995, float("nan"), float("inf"),
          float("-inf")]
    # exact decimal ties at the 7th place: odd multiples of 2^-7 ... 2^-1 (k/128 = x.xxxxxx5 exactly)
    for k in range(1, 257):
        xs.append(k / 128.0)
        xs.append(-k / 128.0 - 3.0)
    xs += list(rng.random(4000))
    xs += list(rng.random(2000) * 1e-5)
    xs += list(np.exp(rng.uniform(-40, 40, 3000)) * rng.choice([-1.0, 1.0], 3000))
    # values one ulp either side of a 6-decimal rounding boundary
    for _ in range(2000):
        b = (rng.integers(0, 10 ** 7) + 0.5) / 1e6
        u = struct.unpack("<q", struct.pack("<d", b))[0]
        for d in (-1, 0, 1):
            xs.append(struct.unpack("<d", struct.pack("<q", u + d))[0])
    return xs


def test_token_routine_matches_correctly_rounded_fixed6():
    from ferromic_b200 import falsta
    for v in _cases():
        assert falsta.format_value(v, falsta.FST) == ofa.fst_token(v), repr(v)
        if not math.isinf(v):
            assert falsta.format_value(v, falsta.DIVERSITY) == ofa.diversity_token(v), repr(v)


def test_hudson_tsv_rows_host_formatting():
    """append_hudson_tsv (process.rs:4006-4041): host formatting only, through the library's token routine."""
    from ferromic_b200 import falsta
    rows = [("chr1", 100, 2000, 0, 1, 0.0123456789, 0.0, None, float("nan"), -0.25),
            ("7", 1, 5, "EUR", "AFR\tx", 1.0, -0.0, 2.5e-7, float("inf"), 1 / 128),
            ("X", 10, 20, None, 'q"uote', 0.5, 1e-7, 123456.7890125, -1e-9, 0.9999995)]
    assert falsta.hudson_tsv_text(rows).decode() == ofa.hudson_tsv_text(rows)
    first = ofa.hudson_tsv_text(rows).splitlines()[0].split("\t")
    assert first == ["chr1", "100", "2000", "HaplotypeGroup", "0", "HaplotypeGroup", "1", "0.012346", "0.000000", "NA",
                     "NA", "-0.250000"]
    for v in _cases():
        assert falsta.format_value(v, falsta.TSV) == ofa.format_optional_float(v), repr(v)


# ----------------------------------------------------------------------------- device rendering
def _records(rng, n, lo, hi):
    pos = rng.integers(lo, hi, size=n)
    v = rng.random((n, 6))
    v[rng.random((n, 6)) < 0.1] = np.nan
    v[rng.random((n, 6)) < 0.1] = 0.0
    v[rng.random((n, 6)) < 0.03] = np.inf
    v[rng.random((n, 6)) < 0.03] = -np.inf
    v[rng.random((n, 6)) < 0.05] *= -1
    v[:: 7, 0] = rng.integers(1, 256, size=len(v[:: 7, 0])) / 128.0  # exact ties
    return pos, v


@pytest.mark.gpu
def test_device_fst_tracks_byte_exact():
    from ferromic_b200 import falsta
    rng = np.random.default_rng(11)
    for n, (rs, re) in ((0, (1, 9)), (1, (5, 5)), (300, (100, 900)), (5000, (-3, 2500)), (40000, (1000, 200000))):
        pos, v = _records(rng, n, rs - 50, re + 50)  # some records outside, many duplicates
        wc = [(int(p), *row) for p, row in zip(pos, v)]
        hud = [(int(p), row[0], row[1], row[2]) for p, row in zip(pos[::2], v[::2])]
        got = falsta.fst_falsta_text("7", rs, re, wc, hud)
        assert got.decode() == ofa.fst_falsta_text("7", rs, re, wc, hud)


@pytest.mark.gpu
def test_device_diversity_tracks_byte_exact(tmp_path):
    from ferromic_b200 import falsta
    rng = np.random.default_rng(12)
    n = 20000
    pos, v = _records(rng, n, 1, 60000)
    v = np.where(np.isinf(v), 0.25, v)
    gid = rng.integers(0, 2, size=n)
    flt = rng.random(n) < 0.4
    recs = [(int(p), float(a), float(b), int(g), bool(f)) for p, a, b, g, f in zip(pos, v[:, 0], v[:, 1], gid, flt)]
    for rs, re in ((1, 60000), (20000, 20100), (70000, 70010)):
        assert falsta.diversity_falsta_text("chr2", rs, re, recs).decode() == \
            ofa.diversity_falsta_text("chr2", rs, re, recs)
    # only group 1 / filtered has records inside -> exactly two tracks
    one = [(10, 0.5, 0.25, 1, True), (999, 0.1, 0.1, 0, False)]
    assert falsta.diversity_falsta_text("1", 1, 20, one).decode() == ofa.diversity_falsta_text("1", 1, 20, one)
    # the gzip-append file convention: one member per call, concatenated members read back as one text
    path = tmp_path / "per_site_diversity_output.falsta.gz"
    falsta.append_diversity_falsta(path, "1", 1, 5, [(3, 1.0, 1.0, 0, False)])
    falsta.append_diversity_falsta(path, "1", 1, 20, one)
    falsta.append_diversity_falsta(path, "1", 1, 20, [])
    with gzip.open(path, "rt") as f:
        text = f.read()
    assert text == ofa.diversity_falsta_text("1", 1, 5, [(3, 1.0, 1.0, 0, False)]) + \
        ofa.diversity_falsta_text("1", 1, 20, one)
    lines = text.splitlines()
    assert lines[lines.index(">unfiltered_pi_chr_1_start_1_end_5_group_0") + 1] == "0,0,1.000000,0,0"


@pytest.mark.gpu
def test_device_tracks_from_gpu_estimators():
    """End to end: per-site pi/theta and Hudson per-site records computed by the CUDA path feed the
    device writer; the text equals the oracle writer over the oracle's records."""
    import ferromic_b200 as fm
    from ferromic_b200 import falsta
    from oracle import pyoracle as orc
    from tests.synth import both_sides
    rng = np.random.default_rng(13)
    V, S = 3000, 40
    g = rng.binomial(1, rng.beta(0.5, 0.5, size=V)[:, None, None], size=(V, S, 2)).astype(np.int8)
    g[rng.random(g.shape) < 0.02] = -1
    pos = np.cumsum(rng.integers(1, 30, size=V, dtype=np.int64))
    vs, _ = orc.from_numpy(g, pos)
    from ferromic_b200.api import _Variants
    h0, h1 = both_sides(range(S // 2)), both_sides(range(S // 2, S))
    region = (int(pos[0]), int(pos[-1]))
    recs_gpu, recs_cpu = [], []
    for gid, haps in ((0, h0), (1, h1)):
        gp, gpi, gth = fm.per_site_diversity_arrays(_Variants(vs.positions, vs.gt), haps, region)
        rp, rpi, rth = orc.per_site_diversity(vs, haps, region)
        recs_gpu += [(int(p), float(a), float(b), gid, False) for p, a, b in zip(gp, gpi, gth)]
        recs_cpu += [(int(p), float(a), float(b), gid, False) for p, a, b in zip(rp, rpi, rth)]
    rs, re = region[0] + 1, region[1] + 1
    assert falsta.diversity_falsta_text("9", rs, re, recs_gpu).decode() == \
        ofa.diversity_falsta_text("9", rs, re, recs_cpu)


def test_format_value_argument_checks():
    import ctypes as C
    from ferromic_b200 import _lib
    L = _lib.lib()
    buf, n = C.create_string_buffer(64), C.c_size_t()
    assert L.fm_falsta_format_value(1.5, 7, buf, 64, C.byref(n)) == _lib.FM_ERR_INVALID_ARG  # unknown mode
    assert L.fm_falsta_format_value(1.5, 1, buf, 8, C.byref(n)) == _lib.FM_ERR_INVALID_ARG   # buffer below 56 bytes
    assert L.fm_falsta_format_value(1.5, 1, buf, 64, C.byref(n)) == 0 and buf.raw[: n.value] == b"1.500000"


@pytest.mark.gpu
def test_track_capacity_and_argument_errors():
    import ctypes as C
    from ferromic_b200 import _lib
    L = _lib.lib()
    pos = np.array([2, 4], dtype=np.int64)
    val = np.array([0.5, np.nan])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    total = C.c_size_t()
    out = C.create_string_buffer(64)
    # length query, then a buffer that is too small: the required length is still reported
    assert L.fm_falsta_tracks(p(pos), p(val), 2, 1, 1, 5, 1, None, 0, None, C.byref(total)) == 0
    assert total.value == len("NA,0.500000,NA,NA,NA")
    assert L.fm_falsta_tracks(p(pos), p(val), 2, 1, 1, 5, 1, out, 5, None, C.byref(total)) == _lib.FM_ERR_INVALID_ARG
    assert total.value == len("NA,0.500000,NA,NA,NA")
    assert L.fm_falsta_tracks(p(pos), p(val), 2, 1, 1, 5, 1, out, 64, None, C.byref(total)) == 0
    assert out.raw[: total.value] == b"NA,0.500000,NA,NA,NA"
    assert L.fm_falsta_tracks(p(pos), p(val), 2, 1, 1, 5, 9, out, 64, None, C.byref(total)) == _lib.FM_ERR_INVALID_ARG
    assert L.fm_falsta_tracks(None, None, 2, 1, 1, 5, 1, out, 64, None, C.byref(total)) == _lib.FM_ERR_INVALID_ARG
    # region clamping (from_1based_inclusive): end < start gives one position; start < 1 starts at 1
    assert L.fm_falsta_tracks(p(pos), p(val), 2, 1, 4, 2, 0, out, 64, None, C.byref(total)) == 0
    assert out.raw[: total.value] == b"NA"
    assert L.fm_falsta_tracks(p(pos), p(val), 2, 1, -3, 3, 0, out, 64, None, C.byref(total)) == 0
    assert out.raw[: total.value] == b"0,0.500000,0"


@pytest.mark.gpu
def test_track_body_longer_than_4_gib_is_measured_in_64_bits(monkeypatch):
    """The token offsets are a prefix sum of u32 lengths: it has to run in 64 bits, a body beyond 4 GiB (a 250 Mb
    region with a dozen tracks) must not wrap.  FM_FALSTA_TEST_INFLATE pads every token length (length query only)
    so the test needs no multi-gigabyte input."""
    import ctypes as C
    from ferromic_b200 import _lib
    L = _lib.lib()
    pos = np.array([2, 4], dtype=np.int64)
    val = np.array([0.5, np.nan, 0.25, 0.125])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    total = C.c_size_t()
    lens = (C.c_size_t * 2)()
    region = 3_000_000
    assert L.fm_falsta_tracks(p(pos), p(val), 2, 2, 1, region, 1, None, 0, lens, C.byref(total)) == 0
    plain = total.value
    assert plain == lens[0] + lens[1] + 1
    k = 2000
    monkeypatch.setenv("FM_FALSTA_TEST_INFLATE", str(k))
    assert L.fm_falsta_tracks(p(pos), p(val), 2, 2, 1, region, 1, None, 0, lens, C.byref(total)) == 0
    assert total.value == plain + 2 * region * k and total.value > 2 ** 32
    assert lens[0] + lens[1] + 1 == total.value
    out = C.create_string_buffer(16)
    assert L.fm_falsta_tracks(p(pos), p(val), 2, 2, 1, 4, 1, out, 16, None, C.byref(total)) == _lib.FM_ERR_INVALID_ARG
