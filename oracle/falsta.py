"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's FALSTA track writers
(process.rs:3731-4002), one String per position like the reference.  Only tests/ may import it.

Pinned against the two reference tests that read these files back
(src/tests/stats_tests.rs:82-241 `test_missing_sites_default_to_zero_diversity`,
:1861-2034 `test_per_site_falsta_includes_hudson_components`), see tests/test_falsta.py.
Rust's `format!("{:.6}", v)` is a correctly rounded (exact-decimal-expansion, ties-to-even)
fixed formatter (core::num::flt2dec strategy::dragon::format_exact); CPython's `'%.6f' % v`
is the same function (dtoa mode 3), so it restates it."""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple


class ZeroBasedHalfOpen:
    """process.rs:190-206."""

    def __init__(self, start: int, end: int):
        self.start, self.end = start, end

    @classmethod
    def from_1based_inclusive(cls, start_inclusive: int, end_inclusive: int) -> "ZeroBasedHalfOpen":
        s = max(start_inclusive, 1)
        e = end_inclusive if end_inclusive >= s else s
        return cls(s - 1, e)

    def len(self) -> int:
        return self.end - self.start if self.end > self.start else 0

    def relative_position_1based_inclusive(self, pos: int) -> Optional[int]:  # process.rs:314-321
        p = (pos - 1) & 0xFFFFFFFFFFFFFFFF  # `(pos - 1) as usize` wraps
        if self.start <= p < self.end:
            return p - self.start + 1
        return None


def diversity_token(v: float) -> str:  # process.rs:3786-3792
    if math.isnan(v):
        return "NA"
    if v == 0.0:
        return "0"
    return "%.6f" % v


def fst_token(v: float) -> str:  # format_value, process.rs:3842-3856
    if math.isnan(v):
        return "NA"
    if math.isinf(v):
        return "Infinity" if v > 0 else "-Infinity"
    if v == 0.0:
        return "0"
    return "%.6f" % v


def diversity_falsta_text(seqname: str, region_start: int, region_end: int,
                          per_site: Sequence[Tuple[int, float, float, int, bool]]) -> str:
    """append_diversity_falsta, process.rs:3740-3806."""
    if not per_site:
        return ""
    region = ZeroBasedHalfOpen.from_1based_inclusive(region_start, region_end)
    n = region.len()
    out: List[str] = []
    for g in sorted({r[3] for r in per_site}):
        for is_filtered, which, prefix in ((False, "pi", "unfiltered_pi_"), (False, "theta", "unfiltered_theta_"),
                                           (True, "pi", "filtered_pi_"), (True, "theta", "filtered_theta_")):
            line = ["0"] * n
            any_rec = False
            for pos1, pi, th, gg, filt in per_site:
                if gg != g or bool(filt) != is_filtered:
                    continue
                rel1 = region.relative_position_1based_inclusive(pos1)
                if rel1 is not None:
                    line[rel1 - 1] = diversity_token(pi if which == "pi" else th)
                    any_rec = True
            if any_rec:
                out.append(f">{prefix}chr_{seqname}_start_{region_start}_end_{region_end}_group_{g}\n")
                out.append(",".join(line) + "\n")
    return "".join(out)


def fst_falsta_text(seqname: str, region_start: int, region_end: int, wc_sites, hudson_sites) -> str:
    """append_fst_falsta, process.rs:3809-4002.  wc_sites: (position, overall_fst, overall_numerator,
    overall_denominator, pairwise_fst, pairwise_numerator, pairwise_denominator)."""
    if not wc_sites and not hudson_sites:
        return ""
    region = ZeroBasedHalfOpen.from_1based_inclusive(region_start, region_end)
    n = region.len()
    tail = f"chr_{seqname}_start_{region_start}_end_{region_end}\n"
    out: List[str] = []

    def track(head, records, col):
        v = ["NA"] * n
        for r in records:
            rel1 = region.relative_position_1based_inclusive(int(r[0]))
            if rel1 is not None:
                v[rel1 - 1] = fst_token(r[col])
        out.append(">" + head + tail)
        out.append(",".join(v) + "\n")

    if wc_sites:
        for col, head in enumerate(("haplotype_overall_fst_summary_", "haplotype_overall_fst_numerator_",
                                    "haplotype_overall_fst_denominator_", "haplotype_0v1_pairwise_fst_summary_",
                                    "haplotype_0v1_pairwise_fst_numerator_",
                                    "haplotype_0v1_pairwise_fst_denominator_"), start=1):
            track(head, wc_sites, col)
    if hudson_sites:
        for col, head in enumerate(("hudson_pairwise_fst_hap_0v1_", "hudson_pairwise_fst_hap_0v1_numerator_",
                                    "hudson_pairwise_fst_hap_0v1_denominator_"), start=1):
            track(head, hudson_sites, col)
    return "".join(out)


def format_optional_float(v: Optional[float]) -> str:  # process.rs:3702-3713
    if v is None or math.isnan(v):
        return "NA"
    if math.isinf(v):
        return "inf" if v > 0 else "-inf"  # Rust's Display for infinities
    return "%.6f" % v


def hudson_tsv_text(rows) -> str:
    """append_hudson_tsv, process.rs:4006-4041 (csv writer, tab delimiter, no header, quotes only when needed)."""
    def pop(p):  # format_population_id, process.rs:3692-3698
        if p is None:
            return ["NA", "NA"]
        return ["NamedPopulation", p] if isinstance(p, str) else ["HaplotypeGroup", str(int(p))]

    def q(f):
        return '"' + f.replace('"', '""') + '"' if any(c in f for c in '\t"\n\r') else f

    out = []
    for chr_, rs, re_, p1, p2, dxy, pi1, pi2, pixy, fst in rows:
        f = [str(chr_), str(int(rs)), str(int(re_))] + pop(p1) + pop(p2) + \
            [format_optional_float(x) for x in (dxy, pi1, pi2, pixy, fst)]
        out.append("\t".join(q(x) for x in f) + "\n")
    return "".join(out)
