"""VCF parse/filter stage (process.rs:4092-4768).

CPU: the oracle restatement of process_variant against the reference's own tests of it.
GPU (-m gpu): the device parser (fm_vcf_parse) against the oracle on generated VCF text that
exercises every branch -- variants, flags, allele info, statistics, position sets, error lines,
output order -- and from_variants on the device feeding the estimators."""
import numpy as np
import pytest

from oracle import vcf as ov

REGION = [(999, 2000)]


def _pv(line, indices, min_gq=30, regions=REGION, chr_="1", allow=None, mask=None):
    miss, stats = ov.MissingDataInfo(), ov.FilteringStats()
    r = ov.process_variant(line, chr_, regions, miss, indices, min_gq, stats, allow, mask)
    return r, miss, stats


# ----------------------------------------------------------------------------- oracle, pinned
def test_oracle_variant_filtering_unit():
    """src/tests/filter_tests.rs:8-78."""
    r, _, _ = _pv("chr1\t1000\t.\tA\tT\t.\tPASS\t.\tGT:GQ\t0|0:50\t0|1:60", [9, 10])
    assert r[2] == 0 and r[0] == 999
    r, _, stats = _pv("chr1\t1001\t.\tA\tT\t.\tPASS\t.\tGT:GQ\t0|0:20\t0|1:25", [9, 10])
    assert r[2] != 0 and r[0] == 1000 and stats.low_gq_variants > 0


def test_oracle_mnp_filtering_mixed_snp_mnp():
    """src/tests/mnp_test.rs:8-43."""
    r, _, stats = _pv("chr1\t1000\t.\tA\tG,TT\t.\tPASS\t.\tGT:GQ\t1|2:40", [9])
    assert r is None and stats.mnp_variants == 1


def test_oracle_gq_filtering():
    """src/tests/stats_tests.rs:882-975."""
    r, _, _ = _pv("chr1\t1000\t.\tA\tT\t.\tPASS\t.\tGT:GQ\t0|0:20\t0|1:40", [9, 10])
    assert r[0] == 999 and r[1] == [[0, 0], [0, 1]] and r[2] != 0 and r[3] == (999, "A", ["T"])
    r, _, _ = _pv("chr1\t1000\t.\tA\tT\t.\tPASS\t.\tGT:GQ\t0|0:35\t0|1:40", [9, 10])
    assert r[2] == 0 and r[1] == [[0, 0], [0, 1]] and r[3] == (999, "A", ["T"])
    assert ov.compressed(r[1]) == (bytes([0, 0, 0, 1]), 2)


def test_oracle_rules():
    # missing genotypes skip the GQ check but flag the variant (process.rs:4682-4690, 4738-4742)
    r, miss, stats = _pv("1\t1500\t.\tC\tG\t.\t.\t.\tGT:GQ\t./.\t0/1:99\t.:.\t1|x:99", [9, 10, 11, 12])
    assert r[1] == [None, [0, 1], None, None] and r[2] == ov.FLAG_MISSING
    assert (miss.total_data_points, miss.missing_data_points, miss.positions_with_missing) == (4, 3, {1499})
    assert stats.missing_data_variants == 1 and stats.filtered_positions == {1499}
    # the line keeps its '\n': a GT-only last field fails u8 parsing, GQ strings are trimmed
    r, _, _ = _pv("1\t1500\t.\tC\tG\t.\t.\t.\tGT:GQ\t0|1:40\t1\n", [9, 10])
    assert r[1] == [[0, 1], None]
    # the genotype is always the FIRST ':' part, whatever FORMAT says (process.rs:4659)
    r, _, _ = _pv("1\t1500\t.\tC\tG\t.\t.\t.\tGQ:GT\t40:0|1\t50:1\n", [9, 10])
    assert r[1] == [[40], [50]] and r[2] == 0
    r, _, _ = _pv("1\t1500\t.\tC\tG\t.\t.\t.\tGT:GQ\t0|1:40\t1|1:50\r\n", [9, 10])
    assert r[1] == [[0, 1], [1, 1]] and r[2] == 0
    # '.', '' and unparsable GQ count as 0; a '+' sign is accepted by from_str
    r, _, _ = _pv("1\t1500\t.\tC\tG\t.\t.\t.\tGT:GQ\t+0|+1:+30\t1:.", [9, 10], min_gq=1)
    assert r[1] == [[0, 1], [1]] and r[2] == ov.FLAG_LOW_GQ
    # errors
    for line, idx, msg in (("1\t1500\t.\tC\tG\t.\t.\t.", [9], "expected at least 9 fixed fields, found 8"),
                           ("1\t1500\t.\tC\tG\t.\t.\t.\tGT:GQ\t0|1:9", [9, 10], "column 11, found 10 columns"),
                           ("1\tx\t.\tC\tG\t.\t.\t.\tGT:GQ\t0|1:9", [9], "Invalid position"),
                           ("1\t0\t.\tC\tG\t.\t.\t.\tGT:GQ\t0|1:9", [9], "Invalid 1-based pos: 0"),
                           ("1\t1500\t.\tC\tG\t.\t.\t.\tGT\t0|1", [9], "GQ field not found in FORMAT"),
                           ("1\t1500\t.\tC\tG\t.\t.\t.\tGT:GQ\t0|1", [9], "GQ value missing in sample genotype field at chr1:1500")):
        with pytest.raises(ov.VcfParseError) as e:
            _pv(line, idx)
        assert msg in str(e.value)
    # other chromosome / outside the regions: None before any statistic
    r, _, stats = _pv("chr2\t1500\t.\tC\tG\t.\t.\t.\tGT:GQ\t0|1:99", [9])
    assert r is None and stats.total_variants == 0
    r, _, stats = _pv("Chr1\t2001\t.\tC\tG\t.\t.\t.\tGT:GQ\t0|1:99", [9])
    assert r is None and stats.total_variants == 0
    # allow / mask
    r, _, stats = _pv("CHR1\t1500\t.\tC\tG\t.\t.\t.\tGT:GQ\t0|1:99", [9], allow={"1": [(0, 10)]}, mask={"1": [(1499, 1500)]})
    assert r[2] == ov.FLAG_ALLOW | ov.FLAG_MASK and stats.filtered_due_to_allow == stats.filtered_due_to_mask == 1
    r, _, _ = _pv("1\t1500\t.\tC\tG\t.\t.\t.\tGT:GQ\t0|1:99", [9], allow={"2": []}, mask={"2": []})
    assert r[2] == ov.FLAG_ALLOW  # chromosome absent from both maps: not allowed, not masked


def test_compiled_oracle_agrees_with_the_python_restatement():
    """oracle/vcf_oracle.c (pthreads; the CPU baseline of tools/bench_vcf.py) against oracle/vcf.py on generated
    text that reaches every branch; 1 and 3 threads."""
    for seed, n_lines, n_cols, odd in ((1, 80, 4, 0.15), (2, 400, 9, 0.05), (3, 60, 300, 0.01)):
        rng = np.random.default_rng(seed)
        text = make_vcf(rng, n_lines, n_cols, odd=odd, crlf=bool(seed & 1))
        kept = sorted(rng.choice(np.arange(9, 9 + n_cols), size=max(1, n_cols * 3 // 4), replace=False).tolist())
        regions = [(950, 1200), (1200, 1300), (1500, 2100)]
        allow = {"1": [(1000, 1100), (1050, 1250), (1600, 1900), (-5, 3), (7, 2)]}
        mask = {"1": [(1020, 1030), (1700, 1705), (-1, 5), (1990, -1)]}
        for threads in (1, 3):
            c = ov.c_process_lines(text.encode(), "chr1", regions, kept, 30, allow, mask, max_ploidy=4, threads=threads)
            skip = {int(l) for l, code in zip(c["err_line"], c["err_code"]) if code == 16}  # genotype longer than 4
            out, miss, stats, errors = ov.process_lines(ov.split_lines(text), "chr1", regions, kept, 30, allow, mask,
                                                        skip=skip)
            assert c["n_lines"] == len(ov.split_lines(text))
            assert [int(l) for l, code in zip(c["err_line"], c["err_code"]) if code != 16] == [l for l, _ in errors]
            assert list(c["positions"]) == [v[0] for v in out] and list(c["flags"]) == [v[2] for v in out]
            for i, v in enumerate(out):
                data, stride = ov.compressed(v[1])
                assert int(c["stride"][i]) == stride and c["gt"][i, :, :stride].tobytes() == data
            assert list(map(int, c["counters"])) == [stats.total_variants, stats.filtered_variants,
                                                     stats.filtered_due_to_mask, stats.filtered_due_to_allow,
                                                     stats.missing_data_variants, stats.low_gq_variants,
                                                     stats.mnp_variants, miss.total_data_points, miss.missing_data_points]


# ----------------------------------------------------------------------------- generated VCF text
GT_POOL = ["0|0", "0|1", "1|0", "1|1", "0/1", "1/1", "0|0", "0|0", "0|1", "1|1"]
ODD_GT = [".", "./.", ".|.", "0|.", ".|1", "", "x", "0|x", "2|1", "10|3", "255|0", "0|255", "256|0", "+1|0", "1|+0",
          "-1|0", "0", "1", "0|1|", "|1", "01|001", "0 |1", "0|1 ", "+", "1|2"]
ODD_GQ = [".", "", " 40", "40 ", "+35", "x", "-5", "65535", "65536", "99999", "0", "29", "30", "3.5"]


def make_vcf(rng, n_lines, n_cols, odd=0.05, chrs=("chr1", "1", "Chr1", "CHR1", "chr2", " chr1 ", "chr11"),
             pos_lo=900, pos_hi=2200, formats=("GT:GQ", "GT:GQ:DP", "GT:DP:GQ", "GQ:GT"), crlf=False, sort=False, odd_gt=None):
    odd_gt = ODD_GT if odd_gt is None else odd_gt
    lines = []
    positions = rng.integers(pos_lo, pos_hi, size=n_lines)
    if sort:
        positions.sort()
    for li in range(n_lines):
        u = rng.random()
        chrom = chrs[0] if u > odd * 3 else chrs[rng.integers(len(chrs))]
        pos = str(int(positions[li]))
        if rng.random() < odd:
            pos = ["0", "-5", "abc", "", "+1200", "12x", "99999999999999999999", "1500"][rng.integers(8)]
        ref = "ACGTacgtN"[rng.integers(9)]
        alt = "ACGT"[rng.integers(4)]
        v = rng.random()
        if v < odd:
            ref = ["AT", "", "ACG"][rng.integers(3)]
        elif v < 2 * odd:
            alt = ["G,TT", "TT", "", "A,", "<DEL>", "A,C,G", "*", ".", "a,c", "A,C,G,T,N,*,.,X", "A,C,G,T,N,*,."][rng.integers(11)]
        fmt = formats[rng.integers(len(formats))]
        if rng.random() < odd:
            fmt = ["GT", "GQ", "GT:GQX", "GT:gq", ""][rng.integers(5)]
        keys = fmt.split(":")
        cols = []
        for _ in range(n_cols):
            gt = GT_POOL[rng.integers(len(GT_POOL))] if rng.random() > odd else odd_gt[rng.integers(len(odd_gt))]
            gq = str(int(rng.integers(25, 99))) if rng.random() > odd else ODD_GQ[rng.integers(len(ODD_GQ))]
            parts = []
            for k in keys:
                parts.append(gt if k == "GT" else gq if k == "GQ" else "7")
            if rng.random() < odd / 2 and len(parts) > 1:
                parts = parts[: rng.integers(1, len(parts))]  # truncated sample field
            cols.append(":".join(parts))
        if rng.random() < odd:
            cols = cols[: rng.integers(0, n_cols)]  # short line
        fixed = [chrom, pos, ".", ref, alt, ".", "PASS", "AC=1;AF=0.5" * int(rng.integers(1, 6)), fmt]
        if rng.random() < odd / 2:
            fixed = fixed[: rng.integers(1, 9)]
            cols = []
        lines.append("\t".join(fixed + cols))
    if rng.random() < 0.5:
        lines.insert(int(rng.integers(0, len(lines) + 1)), "")  # a blank line is a line
    eol = "\r\n" if crlf else "\n"
    text = eol.join(lines)
    if rng.random() < 0.5:
        text += eol
    return text


def check_against_oracle(text, chr_, regions, kept, min_gq, allow=None, mask=None, max_ploidy=3):
    from ferromic_b200 import vcf
    b = vcf.process_lines(text.encode(), chr_, regions, kept, min_gq, allow, mask, max_ploidy=max_ploidy)
    # lines the device reports as unsupported (> 7 single-base ALT alleles, genotype longer than max_ploidy)
    # are left out of the oracle's input; they must be exactly the lines with those properties
    unsupported = {l for l, m in b.errors if m.startswith("unsupported")}
    lines = ov.split_lines(text)
    for l in unsupported:
        f = lines[l].split("\t")
        assert f[4].count(",") >= 7 or any(c.split(":")[0].count("|") + c.split(":")[0].count("/") >= max_ploidy
                                           for c in f[9:])
    out, miss, stats, errors = ov.process_lines(lines, chr_, regions, kept, min_gq, allow, mask, skip=unsupported)
    assert [e for e in b.errors if e[0] not in unsupported] == errors
    assert int(b.info.n_lines) == len(ov.split_lines(text))
    assert b.n_variants == len(out)
    assert list(b.positions) == [v[0] for v in out]
    assert list(b.flags) == [v[2] for v in out]
    assert b.allele_info() == [(v[3][0], v[3][1]) for v in out]
    gt = b.genotypes()
    for i, v in enumerate(out):
        data, stride = ov.compressed(v[1])
        assert int(b.stride[i]) == stride
        got = gt[i, :, :stride].tobytes()
        assert got == data, (i, v[0])
        assert np.all(gt[i, :, stride:] == 0xFF)
    s = b.stats()
    assert s == dict(total_variants=stats.total_variants, filtered_variants=stats.filtered_variants,
                     filtered_due_to_mask=stats.filtered_due_to_mask, filtered_due_to_allow=stats.filtered_due_to_allow,
                     missing_data_variants=stats.missing_data_variants, low_gq_variants=stats.low_gq_variants,
                     mnp_variants=stats.mnp_variants, total_data_points=miss.total_data_points,
                     missing_data_points=miss.missing_data_points)
    assert set(b.positions_with_missing().tolist()) == miss.positions_with_missing
    assert set(b.filtered_positions().tolist()) == stats.filtered_positions
    return b, out


@pytest.mark.gpu
def test_device_parser_reference_cases():
    from ferromic_b200 import vcf
    text = ("chr1\t1000\t.\tA\tT\t.\tPASS\t.\tGT:GQ\t0|0:50\t0|1:60\n"
            "chr1\t1001\t.\tA\tT\t.\tPASS\t.\tGT:GQ\t0|0:20\t0|1:25\n"
            "chr1\t1002\t.\tA\tG,TT\t.\tPASS\t.\tGT:GQ\t1|2:40\t0|0:40\n")
    b, out = check_against_oracle(text, "1", REGION, [9, 10], 30)
    assert list(b.positions) == [999, 1000] and b.flags[0] == 0 and b.flags[1] == vcf.FLAG_LOW_GQ
    assert b.stats()["mnp_variants"] == 1 and b.stats()["low_gq_variants"] == 1
    assert b.allele_info()[0] == ("A", ["T"])


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_lines,n_cols,odd,crlf", [(1, 60, 3, 0.12, False), (2, 400, 7, 0.06, False),
                                                           (3, 300, 40, 0.03, True), (4, 50, 700, 0.01, False),
                                                           (5, 1500, 5, 0.0, False), (6, 8, 3000, 0.002, False),
                                                           (7, 96, 2600, 0.0, False), (8, 40, 2504, 0.0005, True)])
def test_device_parser_matches_oracle(seed, n_lines, n_cols, odd, crlf):
    rng = np.random.default_rng(seed)
    text = make_vcf(rng, n_lines, n_cols, odd=odd, crlf=crlf)
    kept = sorted(rng.choice(np.arange(9, 9 + n_cols), size=max(1, n_cols - n_cols // 4), replace=False).tolist())
    regions = [(950, 1200), (1200, 1300), (1500, 2100)]
    allow = {"1": [(1000, 1100), (1050, 1250), (1600, 1900), (-5, 3), (7, 2)], "2": [(0, 10)]}
    mask = {"1": [(1020, 1030), (1700, 1705), (-1, 5), (1990, -1)]}
    check_against_oracle(text, "chr1", regions, kept, 30, allow, mask)
    check_against_oracle(text, " 1", regions, kept, 0, None, {"2": [(0, 10 ** 9)]})
    check_against_oracle(text, "chr1", [], kept, 30, {"2": []}, None)


@pytest.mark.gpu
def test_device_parser_edges():
    from ferromic_b200 import vcf
    # empty text, one unterminated line, only blank lines
    b = vcf.process_lines(b"", "1", REGION, [9], 30)
    assert b.n_variants == 0 and int(b.info.n_lines) == 0 and b.matrix() is None
    check_against_oracle("1\t1500\t.\tC\tG\t.\t.\t.\tGT:GQ\t0|1:99", "1", REGION, [9], 30)
    check_against_oracle("\n\n\n", "1", REGION, [9], 30)
    # duplicate positions: the tie is broken by the compressed genotype bytes (process.rs:4377-4386)
    text = ("1\t1500\t.\tC\tG\t.\t.\t.\tGT:GQ\t1|1:99\t0|0:99\n"
            "1\t1500\t.\tC\tT\t.\t.\t.\tGT:GQ\t0|1:99\t1|1:99\n"
            "1\t1400\t.\tC\tT\t.\t.\t.\tGT:GQ\t0|1:99\t1|1:99\n"
            "1\t1500\t.\tC\tA\t.\t.\t.\tGT:GQ\t0:99\t1:99\n"
            "1\t1500\t.\tC\tA\t.\t.\t.\tGT:GQ\t./.:99\t.:99\n")
    b, out = check_against_oracle(text, "1", REGION, [9, 10], 30)
    assert [a[1] for a in b.allele_info()] == [["T"], ["A"], ["T"], ["G"], ["A"]]
    # a genotype longer than max_ploidy is reported, not truncated
    b = vcf.process_lines(b"1\t1500\t.\tC\tG\t.\t.\t.\tGT:GQ\t0|1|1:99\n", "1", REGION, [9], 30, max_ploidy=2)
    assert b.n_variants == 0 and b.errors[0][1].startswith("unsupported")
    with pytest.raises(Exception):
        vcf.process_lines(b"x\n", "1", REGION, [10, 9], 30)  # kept columns must increase


@pytest.mark.gpu
def test_device_parser_alignment_sweep():
    """Every alignment of the first sample field and of the line end against the 16-byte words / 4 KB tiles the
    kernels read (INFO padded byte by byte), with lines that span several tiles; text lengths that are exact
    multiples of 16 and of 4096."""
    rng = np.random.default_rng(21)
    n_cols = 1300  # ~9 KB per line: three tiles
    kept = list(range(9, 9 + n_cols))
    lines = []
    for pad in range(0, 40):
        cols = [GT_POOL[rng.integers(len(GT_POOL))] + ":" + str(int(rng.integers(10, 99))) for _ in range(n_cols)]
        if pad % 5 == 0:
            cols[int(rng.integers(n_cols))] = "./.:."
        lines.append("\t".join(["chr1", str(1000 + pad), ".", "A", "C", ".", "PASS", "I" * pad, "GT:GQ"] + cols))
    text = "\n".join(lines) + "\n"
    check_against_oracle(text, "1", REGION, kept, 30)
    one = lines[0] + "\n"
    for target in (16 * ((len(one) + 15) // 16 + 1), 4096 * ((len(one) + 4095) // 4096)):
        f = one.split("\t")
        f[7] = "I" * (target - len(one))
        padded = "\t".join(f)
        assert len(padded) == target
        check_against_oracle(padded, "1", REGION, kept, 30)
        check_against_oracle(padded[:-1], "1", REGION, kept, 30)  # unterminated: the last GQ ends the buffer


@pytest.mark.gpu
@pytest.mark.parametrize("packed", [True, False])  # packed: bit rows straight from the parser's output, no u8 matrix
@pytest.mark.parametrize("excluded", [(3, 7), ()])  # 22 samples: general to_matrix kernel; 24: whole 16-byte rows
def test_vcf_text_to_estimators_without_host_round_trip(excluded, packed):
    """Raw VCF text -> device parser -> from_variants on the device -> groups -> pi / S / Hudson, against the
    oracle estimators over the oracle-parsed variants."""
    from ferromic_b200 import _lib, vcf
    from oracle import pyoracle as orc
    from tests.synth import both_sides
    rng = np.random.default_rng(99)
    n_cols, n_lines = 24, 800
    text = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + \
        "\t".join(f"S{i}" for i in range(n_cols)) + "\n" + \
        make_vcf(rng, n_lines, n_cols, odd=0.01, chrs=("chr1",), sort=True, formats=("GT:GQ",),
                 odd_gt=[".", "./.", ".|.", "0", "1", "x", "0|."])  # biallelic, with None and haploid calls
    regions = [(900, 2300)]
    batch, names = vcf.process_vcf_text(text.encode(), "1", regions, 30, exclusion_set={f"S{i}" for i in excluded})
    assert len(names) == n_cols - len(excluded)
    kept = [9 + i for i in range(n_cols) if i not in excluded]
    body = text.split("\n", 2)[2]
    unsupported = {l for l, msg in batch.errors if msg.startswith("unsupported")}
    out, _, _, _ = ov.process_lines(ov.split_lines(body), "1", regions, kept, 30, skip=unsupported)
    assert list(batch.positions) == [v[0] for v in out] and len(out) > 300
    S = len(kept)
    for pass_only in (False, True):
        sel = [v for v in out if (v[2] == 0 or not pass_only)]
        m = batch.matrix(pass_only=pass_only, packed=packed)
        if not sel:
            assert m is None
            continue
        variants = [{"position": v[0], "genotypes": v[1]} for v in sel]
        vs = orc.variants_from_python(variants, S)
        d = orc.dense_from_variants(vs, S)
        assert (m.V, m.S, m.P, m.max_allele) == (len(sel), S, d.ploidy, d.max_allele)
        haps1, haps2 = both_sides(range(S // 2)), both_sides(range(S // 2, S))
        L = regions[0][1] - regions[0][0]
        for haps in (haps1, haps2):
            g = m.group(haps)
            s, so = g.summary(True), orc.build_summary(d, haps)
            assert np.array_equal(s["alt"], so.alt) and np.array_equal(s["called"], so.called)  # bit-exact counts
            assert s["segregating_sites"] == so.seg
            got, ref = g.pi(L, _lib.FM_PI_SPARSE), orc.pi_sparse(vs, haps, L)
            assert (np.isnan(got) and np.isnan(ref)) or abs(got - ref) <= 1e-9 * abs(ref)  # FP64: 1e-9 relative


# ----------------------------------------------------------------------------- line-sharded parse (multi-GPU)
def test_text_shard_bounds_cut_on_line_ends():
    from ferromic_b200 import vcf
    rng = np.random.default_rng(5)
    text = make_vcf(rng, 200, 6, odd=0.05).encode()
    if not text.endswith(b"\n"):
        text += b"\n"
    for world in (1, 2, 3, 8, 64, 500):
        cuts = vcf.text_shard_bounds(text, world)
        assert cuts[0] == 0 and cuts[-1] == len(text) and len(cuts) == world + 1
        assert all(a <= b for a, b in zip(cuts, cuts[1:]))
        assert all(c == 0 or c == len(text) or text[c - 1:c] == b"\n" for c in cuts)
        assert b"".join(text[a:b] for a, b in zip(cuts, cuts[1:])) == text
    assert vcf.text_shard_bounds(b"", 4) == [0, 0, 0, 0, 0]


def test_sharded_statistics_merge_like_the_whole():
    """Host logic of the line-sharded parse with the oracle as the per-shard parser: counters add, position sets
    union, per-shard variant lists concatenate to the whole (position-sorted text)."""
    from ferromic_b200 import vcf
    rng = np.random.default_rng(6)
    text = make_vcf(rng, 300, 5, odd=0.04, sort=True)
    kept, regions = [9, 10, 12, 13], [(950, 2100)]
    whole, wm, ws, _ = ov.process_lines(ov.split_lines(text), "1", regions, kept, 30)
    for world in (2, 5):
        cuts = vcf.text_shard_bounds(text.encode(), world)
        cnt, pms, pfs, outs = [], [], [], []
        for a, b in zip(cuts, cuts[1:]):
            o, m, s, _ = ov.process_lines(ov.split_lines(text.encode()[a:b].decode()), "1", regions, kept, 30)
            cnt.append([s.total_variants, s.filtered_variants, s.filtered_due_to_mask, s.filtered_due_to_allow,
                        s.missing_data_variants, s.low_gq_variants, s.mnp_variants, m.total_data_points,
                        m.missing_data_points])
            pms.append(np.array(sorted(m.positions_with_missing), dtype=np.int64))
            pfs.append(np.array(sorted(s.filtered_positions), dtype=np.int64))
            outs += o
        stats, pm, pf = vcf.merge_shard_stats(cnt, pms, pfs)
        assert stats["total_variants"] == ws.total_variants and stats["low_gq_variants"] == ws.low_gq_variants
        assert stats["missing_data_points"] == wm.missing_data_points and stats["mnp_variants"] == ws.mnp_variants
        assert set(pm.tolist()) == wm.positions_with_missing and set(pf.tolist()) == ws.filtered_positions
        assert [v[0] for v in outs] == [v[0] for v in whole]


@pytest.mark.gpu
def test_device_parser_shards_equal_the_whole():
    from ferromic_b200 import vcf
    rng = np.random.default_rng(8)
    # positions spread out: a tie of equal positions straddling a cut would be ordered by genotype bytes in the
    # whole parse but not across shards (the merge of two ranks then needs that one comparison)
    text = make_vcf(rng, 500, 30, odd=0.02, sort=True, pos_lo=1000, pos_hi=900000).encode()
    kept, regions = list(range(9, 39)), [(950, 1000000)]
    whole = vcf.process_lines(text, "1", regions, kept, 30, max_ploidy=3)
    cuts = vcf.text_shard_bounds(text, 3)
    parts = [vcf.process_lines(text[a:b], "1", regions, kept, 30, max_ploidy=3) for a, b in zip(cuts, cuts[1:])]
    stats, pm, pf = vcf.merge_shard_stats([[p.stats()[k] for k in vcf.STAT_KEYS] for p in parts],
                                          [p.positions_with_missing() for p in parts],
                                          [p.filtered_positions() for p in parts])
    assert stats == whole.stats()
    assert np.array_equal(pm, whole.positions_with_missing()) and np.array_equal(pf, whole.filtered_positions())
    assert np.array_equal(np.concatenate([p.positions for p in parts]), whole.positions)
    assert np.array_equal(np.concatenate([p.flags for p in parts]), whole.flags)
    assert np.array_equal(np.concatenate([p.genotypes() for p in parts]), whole.genotypes())
    line0 = np.cumsum([0] + [int(p.info.n_lines) for p in parts])
    assert [(l + int(line0[i]), m) for i, p in enumerate(parts) for l, m in p.errors] == whole.errors


def _gloo_worker(rank, port, q):
    import os
    import torch.distributed as dist
    from ferromic_b200 import vcf
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=2)
    try:
        rng = np.random.default_rng(6)
        text = make_vcf(rng, 300, 5, odd=0.04, sort=True).encode()
        cuts = vcf.text_shard_bounds(text, 2)
        mine = text[cuts[rank]:cuts[rank + 1]].decode()
        # no device in this container: the oracle parses the local shard, the exchange is the product's
        o, m, s, errs = ov.process_lines(ov.split_lines(mine), "1", [(950, 2100)], [9, 10, 12, 13], 30)
        cnt = [s.total_variants, s.filtered_variants, s.filtered_due_to_mask, s.filtered_due_to_allow,
               s.missing_data_variants, s.low_gq_variants, s.mnp_variants, m.total_data_points, m.missing_data_points]
        stats, pm, pf, line0 = vcf.gather_shard_stats(cnt, len(ov.split_lines(mine)), sorted(m.positions_with_missing),
                                                      sorted(s.filtered_positions))
        q.put((rank, (stats, pm.tolist(), pf.tolist(), line0, [(l + line0, msg) for l, msg in errs])))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_exchange_of_shard_statistics():
    import socket
    import torch.multiprocessing as mp
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][:3] == got[1][:3]  # every rank ends with the same totals
    rng = np.random.default_rng(6)
    text = make_vcf(rng, 300, 5, odd=0.04, sort=True)
    whole, wm, ws, werr = ov.process_lines(ov.split_lines(text), "1", [(950, 2100)], [9, 10, 12, 13], 30)
    stats, pm, pf = got[0][:3]
    assert stats["total_variants"] == ws.total_variants and stats["filtered_variants"] == ws.filtered_variants
    assert stats["total_data_points"] == wm.total_data_points and stats["missing_data_points"] == wm.missing_data_points
    assert set(pm) == wm.positions_with_missing and set(pf) == ws.filtered_positions
    assert got[0][3] == 0 and got[0][4] + got[1][4] == werr  # global line numbering of the error lines


@pytest.mark.gpu
def test_device_resident_text_entry_point():
    """fm_vcf_parse_device: text already in device memory (16-byte aligned, zero padding) gives the same batch."""
    import ctypes as C
    import torch
    from ferromic_b200 import _lib, vcf
    rng = np.random.default_rng(33)
    text = make_vcf(rng, 300, 12, odd=0.03).encode()
    kept = np.arange(9, 21, dtype=np.uint32)
    reg = np.array([[950, 2100]], dtype=np.int64)
    host = vcf.process_lines(text, "1", [(950, 2100)], kept.tolist(), 30, max_ploidy=3)
    n = len(text)
    buf = torch.zeros(((n + 15) // 16) * 16 + 32, dtype=torch.uint8, device="cuda")
    buf[:n] = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    torch.cuda.synchronize()
    assert buf.data_ptr() % 16 == 0
    h = C.c_void_p()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(_lib.lib().fm_vcf_parse_device(buf.data_ptr(), n, text[-1:], b"1", p(reg), 1, p(kept), len(kept), 30,
                                              0, None, 0, 0, None, 0, 3, C.byref(h)))
    dev = vcf.VcfBatch(h, "1", 20)
    assert dev.stats() == host.stats() and dev.errors == host.errors
    assert np.array_equal(dev.positions, host.positions) and np.array_equal(dev.flags, host.flags)
    assert np.array_equal(dev.genotypes(), host.genotypes())
