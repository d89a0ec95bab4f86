"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's VCF parse/filter stage
(SURVEY §8f rank 4): `process_variant` (process.rs:4471-4768) and the merge / sort that
`process_vcf` applies to its results (process.rs:4262-4400).  Only tests/ may import it.

Pinned against the reference's own process_variant tests (tests/test_vcf.py):
src/tests/filter_tests.rs:8-78, src/tests/mnp_test.rs:8-43, src/tests/stats_tests.rs:882-975.
Pure-Python loops: for small cases only."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Set, Tuple

FLAG_PASS, FLAG_MASK, FLAG_ALLOW, FLAG_LOW_GQ, FLAG_MISSING = 0, 1, 2, 4, 8  # process.rs:785-789
MISSING = 0xFF  # CompressedGenotypes::MISSING, process.rs:438

_U64 = 0xFFFFFFFFFFFFFFFF
# char::is_whitespace for the code points a byte-oriented parser can meet (ASCII + NEL / NBSP handled as bytes
# never match: they are multi-byte in UTF-8); str::trim removes these at both ends
_WS = "\t\n\x0b\x0c\r "


class VcfParseError(Exception):
    """VcfError::Parse(message)."""


@dataclass
class MissingDataInfo:  # process.rs:543-548
    total_data_points: int = 0
    missing_data_points: int = 0
    positions_with_missing: Set[int] = field(default_factory=set)


@dataclass
class FilteringStats:  # process.rs (FilteringStats): only the counters, not the example strings
    total_variants: int = 0
    filtered_variants: int = 0
    filtered_positions: Set[int] = field(default_factory=set)
    filtered_due_to_mask: int = 0
    filtered_due_to_allow: int = 0
    missing_data_variants: int = 0
    low_gq_variants: int = 0
    mnp_variants: int = 0


def _parse_unsigned(s: str, bits: int) -> Optional[int]:
    """<u8|u16 as FromStr>::from_str: optional '+', then ASCII digits only, no overflow."""
    if s.startswith("+"):
        s = s[1:]
    if not s or any(c < "0" or c > "9" for c in s):
        return None
    v = int(s)
    return v if v < (1 << bits) else None


def _parse_i64(s: str) -> Optional[int]:
    neg = False
    if s[:1] in ("+", "-"):
        neg = s[0] == "-"
        s = s[1:]
    if not s or any(c < "0" or c > "9" for c in s):
        return None
    v = -int(s) if neg else int(s)
    return v if -(1 << 63) <= v < (1 << 63) else None


def _normalize_chr_prefix(c: str) -> str:  # process.rs:4501-4511
    for p in ("chr", "Chr", "CHR"):
        if c.startswith(p):
            return c[len(p):]
    return c


def _rust_trim(s: str) -> str:
    return s.strip(_WS)


def position_in_zero_based_regions(pos: int, regions: Sequence[Tuple[int, int]]) -> bool:
    """process.rs:746-760 (partition_point over sorted regions)."""
    lo, hi = 0, len(regions)
    while lo < hi:  # partition_point(|r| r.end <= pos)
        mid = (lo + hi) // 2
        if regions[mid][1] <= pos:
            lo = mid + 1
        else:
            hi = mid
    return lo < len(regions) and regions[lo][0] <= pos


def _nuc(ch: str) -> str:  # process.rs:4621-4640
    return ch.upper() if ch in "ACGTacgt" and ch != "" else "N"


def process_variant(line: str, chr_: str, regions: Sequence[Tuple[int, int]], missing: MissingDataInfo,
                    kept_col_indices: Sequence[int], min_gq: int, stats: FilteringStats,
                    allow_regions: Optional[Dict[str, List[Tuple[int, int]]]] = None,
                    mask_regions: Optional[Dict[str, List[Tuple[int, int]]]] = None):
    """process.rs:4471-4768.  regions are ZeroBasedHalfOpen (start, end) pairs, sorted.
    Returns None or (position0, genotypes: list[Optional[list[int]]], flags, (pos0, ref, alts))."""
    fields = line.split("\t")
    if len(fields) < 9:
        raise VcfParseError(f"Invalid VCF line format: expected at least 9 fixed fields, found {len(fields)}")
    if kept_col_indices:
        max_idx = max(kept_col_indices)
        if len(fields) <= max_idx:
            raise VcfParseError(f"Invalid VCF line format: expected genotype field at column {max_idx + 1}, "
                                f"found {len(fields)} columns")
    vcf_chr = _normalize_chr_prefix(_rust_trim(fields[0]))
    target_chr = _normalize_chr_prefix(_rust_trim(chr_))
    if vcf_chr != target_chr:
        return None
    p1 = _parse_i64(fields[1])
    if p1 is None:
        raise VcfParseError("Invalid position")
    if p1 < 1:
        raise VcfParseError(f"Invalid 1-based pos: {p1}")
    pos0 = p1 - 1
    if not position_in_zero_based_regions(pos0, regions):
        return None
    stats.total_variants += 1
    flags = FLAG_PASS
    if allow_regions is not None:
        ar = allow_regions.get(vcf_chr)
        if ar is None or not any(s <= pos0 < e for s, e in ar):  # position_in_regions, process.rs:738-744
            flags |= FLAG_ALLOW
            stats.filtered_due_to_allow += 1
    if mask_regions is not None:
        mr = mask_regions.get(vcf_chr)
        if mr is not None:
            # ZeroBasedHalfOpen::intersect of [pos, pos+1) with (start as usize, end as usize)
            masked = any(max(pos0, s & _U64) < min(pos0 + 1, e & _U64) for s, e in mr)
            if masked:
                flags |= FLAG_MASK
                stats.filtered_due_to_mask += 1
    alt_alleles = fields[4].split(",")
    indel = False
    if len(fields[3].encode()) != 1:
        indel = True
    if not indel and any(len(a.encode()) != 1 for a in alt_alleles):
        indel = True
        if any(len(a.encode()) > 1 for a in alt_alleles):
            stats.mnp_variants += 1
    allele_info = None
    if fields[3] != "" and fields[4] != "":
        allele_info = (pos0, _nuc(fields[3][0]), [_nuc(a[0]) if a else "N" for a in alt_alleles])
    fmt = fields[8].split(":")
    if "GQ" not in fmt:
        raise VcfParseError("GQ field not found in FORMAT")
    gq_index = fmt.index("GQ")
    raw: List[Optional[List[int]]] = []
    for idx in kept_col_indices:
        gt = fields[idx]
        missing.total_data_points += 1
        alleles_str = gt.split(":")[0]
        if alleles_str in (".", "./.", ".|."):
            missing.missing_data_points += 1
            missing.positions_with_missing.add(pos0)
            raw.append(None)
            continue
        parts = alleles_str.replace("/", "|").split("|")
        vals = [_parse_unsigned(p, 8) for p in parts]
        if any(v is None for v in vals):
            missing.missing_data_points += 1
            missing.positions_with_missing.add(pos0)
            raw.append(None)
        else:
            raw.append(vals)
    low_gq = False
    for i, idx in enumerate(kept_col_indices):
        if raw[i] is None:
            continue
        sub = fields[idx].split(":")
        if gq_index >= len(sub):
            raise VcfParseError(f"GQ value missing in sample genotype field at chr{chr_}:{p1}")
        gq_str = _rust_trim(sub[gq_index])
        if gq_str in (".", ""):
            gq = 0
        else:
            gq = _parse_unsigned(gq_str, 16)
            if gq is None:
                gq = 0
        if gq < min_gq:
            low_gq = True
    has_missing = any(g is None for g in raw)
    if low_gq:
        stats.low_gq_variants += 1
        flags |= FLAG_LOW_GQ
    if has_missing:
        stats.missing_data_variants += 1
        flags |= FLAG_MISSING
    passes = flags == FLAG_PASS and not indel
    if not passes:
        stats.filtered_variants += 1
        stats.filtered_positions.add(pos0)
    if indel:
        return None
    return pos0, raw, flags, allele_info


def compressed(raw: Sequence[Optional[Sequence[int]]]) -> Tuple[bytes, int]:
    """CompressedGenotypes::new (process.rs:440-478): (data, stride)."""
    n = len(raw)
    stride = max((len(g) for g in raw if g is not None), default=0)
    if n > 0:
        stride = max(stride, 1)
    if n == 0 or stride == 0:
        return b"", stride
    flat = bytearray([MISSING]) * (n * stride)
    for s, g in enumerate(raw):
        if g is not None:
            for k, a in enumerate(g[:stride]):
                flat[s * stride + k] = a
    return bytes(flat), stride


def process_lines(lines: Sequence[str], chr_: str, regions, kept_col_indices, min_gq,
                  allow_regions=None, mask_regions=None, skip=()):
    """The data-line part of process_vcf (process.rs:4262-4400): every line through process_variant with
    line-local statistics that are merged only when the line returned Ok; a line that returns Err is
    reported and skipped; the surviving variants are sorted by (position, compressed genotype bytes).
    Returns (variants [(pos0, raw, flags, (ref, alts))], MissingDataInfo, FilteringStats, errors
    [(line_index, message)])."""
    miss, stats = MissingDataInfo(), FilteringStats()
    out, errors = [], []
    for li, line in enumerate(lines):
        if li in skip:  # lines the device path reports as unsupported (tests only)
            continue
        lm, ls = MissingDataInfo(), FilteringStats()
        try:
            r = process_variant(line, chr_, regions, lm, kept_col_indices, min_gq, ls, allow_regions, mask_regions)
        except VcfParseError as e:
            errors.append((li, str(e)))
            continue
        miss.total_data_points += lm.total_data_points
        miss.missing_data_points += lm.missing_data_points
        miss.positions_with_missing |= lm.positions_with_missing
        stats.total_variants += ls.total_variants
        stats.filtered_variants += ls.filtered_variants
        stats.filtered_positions |= ls.filtered_positions
        stats.filtered_due_to_mask += ls.filtered_due_to_mask
        stats.filtered_due_to_allow += ls.filtered_due_to_allow
        stats.missing_data_variants += ls.missing_data_variants
        stats.low_gq_variants += ls.low_gq_variants
        stats.mnp_variants += ls.mnp_variants
        if r is not None:
            pos0, raw, flags, info = r
            out.append((pos0, raw, flags, None if info is None else (info[1], info[2])))
    out.sort(key=lambda v: (v[0], compressed(v[1])[0]))  # stable, like slice::sort_by
    return out, miss, stats, errors


def split_lines(text: str) -> List[str]:
    """BufRead::read_line framing: every line keeps its terminating '\\n'; a final unterminated line is a line."""
    lines = text.split("\n")
    out = [l + "\n" for l in lines[:-1]]
    if lines[-1] != "":
        out.append(lines[-1])
    return out


# ----------------------------------------------------------------------------- compiled sibling (vcf_oracle.c)
def c_process_lines(text: bytes, chr_: str, regions, kept_col_indices, min_gq, allow_regions=None, mask_regions=None,
                    max_ploidy: int = 2, threads: int = 1):
    """oracle/vcf_oracle.c through ctypes: the same stage in plain C with pthreads (cross-check of this module and
    the CPU baseline of tools/bench_vcf.py).  Returns a dict of arrays in output order + counters + error lines."""
    import ctypes as C
    import os
    import subprocess

    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "_build", "libvcf_oracle.so")
    src = os.path.join(here, "vcf_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", here, "-s"])
    L = C.CDLL(so)
    L.orc_vcf_process_lines.restype = C.c_size_t
    key = _normalize_chr_prefix(_rust_trim(chr_))

    def mode(m):
        if m is None:
            return 0, np.zeros((0, 2), dtype=np.int64)
        if key in m:
            return 1, np.ascontiguousarray(np.asarray(m[key], dtype=np.int64).reshape(-1, 2))
        return 2, np.zeros((0, 2), dtype=np.int64)

    am, av = mode(allow_regions)
    mm, mv = mode(mask_regions)
    reg = np.ascontiguousarray(np.asarray(regions, dtype=np.int64).reshape(-1, 2))
    kept = np.ascontiguousarray(kept_col_indices, dtype=np.uint32)
    n_lines_cap = text.count(b"\n") + 1
    S, P = len(kept), max_ploidy
    counters = np.zeros(9, dtype=np.uint64)
    pos0 = np.zeros(n_lines_cap, dtype=np.int64)
    flags = np.zeros(n_lines_cap, dtype=np.uint8)
    stride = np.zeros(n_lines_cap, dtype=np.uint8)
    gt = np.zeros((n_lines_cap, S, P), dtype=np.uint8)
    err_line = np.zeros(n_lines_cap, dtype=np.uint64)
    err_code = np.zeros(n_lines_cap, dtype=np.int32)
    nv, ne = C.c_size_t(), C.c_size_t()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n_lines = L.orc_vcf_process_lines(text, C.c_size_t(len(text)), key.encode(), p(reg), C.c_size_t(len(reg)), p(kept),
                                      C.c_size_t(S), C.c_uint32(min_gq), am, p(av), C.c_size_t(len(av)), mm, p(mv),
                                      C.c_size_t(len(mv)), C.c_size_t(P), int(threads), p(counters), C.byref(nv), p(pos0),
                                      p(flags), p(stride), p(gt), C.byref(ne), p(err_line), p(err_code))
    n = nv.value
    return dict(n_lines=n_lines, positions=pos0[:n], flags=flags[:n], stride=stride[:n], gt=gt[:n],
                counters=counters, err_line=err_line[: ne.value], err_code=err_code[: ne.value])
