// fm_vcf.cuh -- SURVEY §8(f4): the VCF parse/filter stage (process_variant, process.rs:4471-4768)
// as device kernels over a chunk of raw VCF text.  Byte work, HBM/L2-bound; no tensor cores.
//
//   fm_k_vcf_count   per 4 KB tile: number of '\n' and '\t'            (text read #1)
//   fm_k_vcf_index   line starts + tabs-before-line from the scanned tile counts (text read #2, L2)
//   fm_k_vcf_fixed   one warp per line: the nine fixed fields -> chr / POS / region / allow / mask /
//                    REF-ALT length guard / allele info / GQ index in FORMAT   (first ~100 B of a line)
//   fm_k_vcf_samples one CTA per candidate line: the sample region is cut into 256 contiguous spans, one block scan
//                    per line gives every span its first column index, each thread parses its span's kept
//                    fields in order -> u8 genotype row, GQ / missing flags
//   fm_k_vcf_to_matrix  DenseGenotypeMatrix::from_variants (stats.rs:339-500) on the device, in output order
//
// A line is what BufRead::read_line yields: it INCLUDES its terminating '\n' (the last field carries it,
// exactly as in the reference, where "0|1\n" fails u8 parsing and GQ strings are trimmed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fm {

constexpr int VCF_TILE = 4096;  // bytes per CTA tile = 256 threads x 16 B
constexpr int VCF_MAX_ALTS = 7;
constexpr int VCF_MAX_PLOIDY = 8;

enum : uint8_t {
    VCF_CAND = 0,   // reached the genotype loop
    VCF_SKIP = 1,   // Ok(None) before any statistic: other chromosome / outside the regions
    VCF_E_FEW_FIELDS = 10,
    VCF_E_MISSING_COLUMN = 11,
    VCF_E_INVALID_POS = 12,
    VCF_E_POS_LT1 = 13,
    VCF_E_NO_GQ_FORMAT = 14,
    VCF_E_GQ_MISSING = 15,
    VCF_E_PLOIDY = 16,         // genotype longer than the caller's max_ploidy (unsupported, not a reference error)
    VCF_E_TOO_MANY_ALTS = 17,  // more than VCF_MAX_ALTS single-base ALT alleles (unsupported)
};

struct VcfLine {  // 32 bytes per text line
    int64_t pos0;             // POS - 1
    uint32_t sample_off;      // byte after the 9th tab (line end when the line has exactly nine fields)
    uint32_t missing_points;  // kept samples whose genotype is None
    uint16_t gq_index;
    uint8_t status, flags, indel /* bit0 length guard hit, bit1 counted as MNP */, stride, ref, n_alt;
    uint8_t alts[VCF_MAX_ALTS];
    uint8_t pad;
};
static_assert(sizeof(VcfLine) == 32, "VcfLine layout");

struct VcfParams {
    const uint8_t *text;
    const uint32_t *line_start;   // [n_lines + 1]
    const uint32_t *tabs_before;  // [n_lines + 1]
    uint32_t n_lines;
    uint8_t chr[64];  // target chromosome, trimmed and prefix-stripped (process.rs:4501-4514)
    uint32_t chr_len;
    const int64_t *regions;  // [n_regions][2] ZeroBasedHalfOpen (start, end), sorted
    uint32_t n_regions;
    int allow_mode, mask_mode;        // 0 = None, 1 = intervals of this chromosome, 2 = chromosome absent from the map
    const uint64_t *allow, *mask;     // merged, sorted, disjoint [s, e) (host-normalised)
    uint32_t n_allow, n_mask;
    int32_t max_idx;  // largest kept column index, -1 when no sample is kept
    uint16_t min_gq;
    uint32_t n_samples, max_ploidy;
    const int32_t *col2slot;  // [max_idx + 1] column -> kept sample slot or -1
};

__device__ __forceinline__ uint32_t vcf_eq4(uint32_t w, uint32_t pat) {
    // bit i = byte i of w equals the pattern byte: exact zero-byte test on w ^ pat (no carries cross bytes
    // because the add only sees 7-bit fields), then the four flag bits are gathered by one multiply
    const uint32_t t = w ^ pat;
    const uint32_t z = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;  // 0x80 in every zero byte
    return ((z >> 7) * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t vcf_eq16(uint4 v, uint32_t pat) {
    return vcf_eq4(v.x, pat) | (vcf_eq4(v.y, pat) << 4) | (vcf_eq4(v.z, pat) << 8) | (vcf_eq4(v.w, pat) << 12);
}
constexpr uint32_t VCF_NL = 0x0A0A0A0Au, VCF_TAB = 0x09090909u;

// block-wide exclusive scan of a packed pair of 16-bit counts (each total <= 4096); 256 threads
__device__ __forceinline__ uint32_t vcf_block_scan(uint32_t v, uint32_t *smem /*[9]*/, uint32_t &total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (uint32_t)d) inc += o;
    }
    __syncthreads();  // protects smem reuse across calls
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t s = smem[w];
        if ((uint32_t)w < warp) base += s;
        tot += s;
    }
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(256)
fm_k_vcf_count(const uint4 *__restrict__ text16, uint64_t n16, uint32_t *__restrict__ tile_nl,
               uint32_t *__restrict__ tile_tab) {
    __shared__ uint32_t s[8];
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t c = 0;
    if (i < n16) {
        const uint4 v = text16[i];  // padding bytes are zero: neither '\n' nor '\t'
        c = __popc(vcf_eq16(v, VCF_NL)) | (__popc(vcf_eq16(v, VCF_TAB)) << 16);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; ++w) t += s[w];
        tile_nl[blockIdx.x] = t & 0xFFFFu;
        tile_tab[blockIdx.x] = t >> 16;
    }
}

__global__ void __launch_bounds__(256)
fm_k_vcf_index(const uint4 *__restrict__ text16, uint64_t n16, const uint32_t *__restrict__ nl_before_tile,
               const uint32_t *__restrict__ tab_before_tile, uint32_t tab_bias, uint32_t *__restrict__ line_start,
               uint32_t *__restrict__ tabs_before) {
    // tab_before_tile comes from one scan over [newline counts, 0, tab counts, 0]: subtract the newline total
    __shared__ uint32_t s[9];
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t nlm = 0, tbm = 0;
    if (i < n16) {
        const uint4 v = text16[i];
        nlm = vcf_eq16(v, VCF_NL);
        tbm = vcf_eq16(v, VCF_TAB);
    }
    uint32_t total;
    const uint32_t ex = vcf_block_scan(__popc(nlm) | (__popc(tbm) << 16), s, total);
    uint32_t k = nl_before_tile[blockIdx.x] + (ex & 0xFFFFu);
    const uint32_t tb = tab_before_tile[blockIdx.x] - tab_bias + (ex >> 16);
    while (nlm) {
        const int b = __ffs(nlm) - 1;
        nlm &= nlm - 1;
        ++k;  // line k starts after this newline
        line_start[k] = (uint32_t)(i * 16 + b + 1);
        tabs_before[k] = tb + __popc(tbm & ((1u << b) - 1u));
    }
}

// ---------------------------------------------------------------------------------- fixed fields
__device__ __forceinline__ bool vcf_is_ws(uint8_t c) {  // ASCII members of char::is_whitespace
    return c == ' ' || (c >= 9 && c <= 13);
}
__device__ __forceinline__ uint8_t vcf_nuc(uint8_t c) {  // process.rs:4621-4640
    switch (c) {
        case 'A': case 'a': return 'A';
        case 'C': case 'c': return 'C';
        case 'G': case 'g': return 'G';
        case 'T': case 't': return 'T';
        default: return 'N';
    }
}
// partition_point(|r| r.end <= pos) then start <= pos (process.rs:746-760)
__device__ inline bool vcf_in_regions(int64_t pos, const int64_t *r, uint32_t n) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (r[2 * mid + 1] <= pos) lo = mid + 1; else hi = mid;
    }
    return lo < n && r[2 * lo] <= pos;
}
// membership in merged, sorted, disjoint unsigned intervals
__device__ inline bool vcf_in_intervals(uint64_t pos, const uint64_t *iv, uint32_t n) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (iv[2 * mid + 1] <= pos) lo = mid + 1; else hi = mid;
    }
    return lo < n && iv[2 * lo] <= pos;
}

__global__ void __launch_bounds__(256)
fm_k_vcf_fixed(VcfParams P, VcfLine *__restrict__ recs) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const uint8_t *__restrict__ tx = P.text;
    for (uint32_t line = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; line < P.n_lines; line += warps) {
        const uint32_t b = P.line_start[line], e = P.line_start[line + 1];
        const uint32_t n_tabs = P.tabs_before[line + 1] - P.tabs_before[line];
        const uint32_t n_fields = n_tabs + 1;
        VcfLine r;
        r.pos0 = (int64_t)n_fields;  // E_FEW_FIELDS / E_MISSING_COLUMN report the field count here
        r.sample_off = e;
        r.missing_points = 0;
        r.gq_index = 0;
        r.status = VCF_CAND;
        r.flags = 0;
        r.indel = 0;
        r.stride = 0;
        r.ref = 'N';
        r.n_alt = 0;
        for (int k = 0; k < VCF_MAX_ALTS; ++k) r.alts[k] = 0;
        r.pad = 0;
        if (n_fields < 9) {
            r.status = VCF_E_FEW_FIELDS;
        } else if (P.max_idx >= 0 && n_fields <= (uint32_t)P.max_idx) {
            r.status = VCF_E_MISSING_COLUMN;
        }
        if (r.status != VCF_CAND) {
            if (lane == 0) recs[line] = r;
            continue;
        }
        // the first min(9, n_tabs) tab positions, found 32 bytes at a time by the whole warp
        uint32_t tab[9];
        const uint32_t want = n_tabs < 9 ? n_tabs : 9;  // >= 8
        uint32_t found = 0;
        for (uint32_t base = b; found < want && base < e; base += 32) {
            const uint32_t p = base + lane;
            uint32_t m = __ballot_sync(0xffffffffu, p < e && tx[p] == '\t');
            while (m && found < want) {
                const uint32_t bit = __ffs(m) - 1;
                m &= m - 1;
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    if ((uint32_t)k == found) tab[k] = base + bit;
                ++found;
            }
        }
        if (lane != 0) continue;
        if (want < 9) tab[8] = e;  // exactly nine fields: FORMAT runs to the end of the line (incl. '\n')
        r.sample_off = want == 9 ? tab[8] + 1 : e;
        // ---- CHROM: trim, strip one of chr / Chr / CHR, compare (process.rs:4501-4518)
        uint32_t c0 = b, c1 = tab[0];
        while (c0 < c1 && vcf_is_ws(tx[c0])) ++c0;
        while (c1 > c0 && vcf_is_ws(tx[c1 - 1])) --c1;
        if (c1 - c0 >= 3) {
            const uint8_t x = tx[c0], y = tx[c0 + 1], z = tx[c0 + 2];
            if ((x == 'c' && y == 'h' && z == 'r') || (x == 'C' && y == 'h' && z == 'r') ||
                (x == 'C' && y == 'H' && z == 'R'))
                c0 += 3;
        }
        bool same = (c1 - c0) == P.chr_len;
        for (uint32_t k = 0; same && k < P.chr_len; ++k) same = tx[c0 + k] == P.chr[k];
        if (!same) {
            r.status = VCF_SKIP;
            recs[line] = r;
            continue;
        }
        // ---- POS: i64::from_str (optional sign, digits, overflow is an error)
        {
            uint32_t p = tab[0] + 1;
            const uint32_t pe = tab[1];
            bool neg = false, ok = true;
            if (p < pe && (tx[p] == '+' || tx[p] == '-')) {
                neg = tx[p] == '-';
                ++p;
            }
            if (p >= pe) ok = false;
            uint64_t mag = 0;
            const uint64_t lim = neg ? (1ull << 63) : (1ull << 63) - 1;
            for (; ok && p < pe; ++p) {
                const uint32_t d = (uint32_t)tx[p] - '0';
                if (d > 9) ok = false;
                else if (mag > (lim - d) / 10) ok = false;
                else mag = mag * 10 + d;
            }
            if (!ok) {
                r.status = VCF_E_INVALID_POS;
                recs[line] = r;
                continue;
            }
            const int64_t p1 = neg ? (int64_t)(0 - mag) : (int64_t)mag;
            r.pos0 = p1 - 1;  // p1 >= 1 below, or reported as the offending value - 1
            if (p1 < 1) {
                r.status = VCF_E_POS_LT1;
                recs[line] = r;
                continue;
            }
        }
        if (!vcf_in_regions(r.pos0, P.regions, P.n_regions)) {
            r.status = VCF_SKIP;
            recs[line] = r;
            continue;
        }
        // ---- allow / mask (process.rs:4549-4593)
        if (P.allow_mode == 2 || (P.allow_mode == 1 && !vcf_in_intervals((uint64_t)r.pos0, P.allow, P.n_allow)))
            r.flags |= 2;  // FLAG_ALLOW
        if (P.mask_mode == 1 && vcf_in_intervals((uint64_t)r.pos0, P.mask, P.n_mask)) r.flags |= 1;  // FLAG_MASK
        // ---- length guard + allele info (process.rs:4595-4644)
        {
            const uint32_t r0 = tab[2] + 1, r1 = tab[3];  // REF
            const uint32_t a0 = tab[3] + 1, a1 = tab[4];  // ALT
            bool indel = (r1 - r0) != 1, longer = false, not_one = false;
            uint32_t n_alt = 0, seg = a0;
            for (uint32_t p = a0; p <= a1; ++p) {
                if (p == a1 || tx[p] == ',') {
                    const uint32_t len = p - seg;
                    if (len != 1) not_one = true;
                    if (len > 1) longer = true;
                    if (n_alt < VCF_MAX_ALTS) r.alts[n_alt] = len ? vcf_nuc(tx[seg]) : 'N';
                    ++n_alt;
                    seg = p + 1;
                }
            }
            if (!indel && not_one) {
                indel = true;
                if (longer) r.indel |= 2;
            }
            if (indel) r.indel |= 1;
            r.ref = (r1 > r0) ? vcf_nuc(tx[r0]) : 'N';
            r.n_alt = (uint8_t)(n_alt < 255 ? n_alt : 255);
            if (!indel && n_alt > VCF_MAX_ALTS) r.status = VCF_E_TOO_MANY_ALTS;
        }
        // ---- FORMAT: index of the "GQ" key (process.rs:4646-4650)
        {
            const uint32_t f0 = tab[7] + 1, f1 = tab[8];
            uint32_t idx = 0, seg = f0;
            bool have = false;
            for (uint32_t p = f0; p <= f1 && !have; ++p) {
                if (p == f1 || tx[p] == ':') {
                    if (p - seg == 2 && tx[seg] == 'G' && tx[seg + 1] == 'Q') have = true;
                    else {
                        ++idx;
                        seg = p + 1;
                    }
                }
            }
            if (!have) r.status = VCF_E_NO_GQ_FORMAT;
            r.gq_index = (uint16_t)(idx < 65535 ? idx : 65535);
        }
        recs[line] = r;
    }
}

// ---------------------------------------------------------------------------------- sample fields
// One kept sample field starting at p (field end = next '\t' or line end e).  Returns the number of
// alleles (0 = genotype is None) and sets low_gq / gq_missing.
__device__ __forceinline__ uint32_t vcf_parse_sample(const uint8_t *__restrict__ tx, uint32_t p, uint32_t e,
                                                     uint32_t gq_index, uint32_t min_gq, uint32_t max_ploidy,
                                                     uint8_t *al, bool &low_gq, bool &gq_missing, bool &too_long) {
    // alleles_str = field up to the first ':' (process.rs:4659)
    uint32_t q = p;
    uint32_t n_tok = 0, val = 0, digits = 0;
    bool ok = true, plus = false;
    for (; q < e; ++q) {
        const uint8_t c = tx[q];
        if (c == ':' || c == '\t') break;
        if (c == '|' || c == '/') {
            if (digits == 0) ok = false;
            if (ok && n_tok < VCF_MAX_PLOIDY) al[n_tok] = (uint8_t)val;
            ++n_tok;
            val = 0;
            digits = 0;
            plus = false;
        } else if (c >= '0' && c <= '9') {
            val = val * 10 + (c - '0');
            if (val > 255) {
                ok = false;
                val = 256;  // stays out of range without overflowing
            }
            ++digits;
        } else if (c == '+' && digits == 0 && !plus) {
            plus = true;  // u8::from_str accepts one leading '+'
        } else {
            ok = false;  // includes ".", "./.", ".|." (None by name in the reference) and a trailing '\n'
        }
    }
    if (digits == 0) ok = false;
    if (ok && n_tok < VCF_MAX_PLOIDY) al[n_tok] = (uint8_t)val;
    ++n_tok;
    if (!ok) return 0;  // every genotype that fails u8 parsing is None (process.rs:4668-4677)
    if (n_tok > max_ploidy) {
        too_long = true;
        return 0;
    }
    // GQ: the gq_index-th ':'-separated part of the whole field, trimmed (process.rs:4694-4728)
    uint32_t part = 0, g0 = p;
    uint32_t k = p;
    for (; k < e; ++k) {
        const uint8_t d = tx[k];
        if (d == '\t') break;
        if (d == ':') {
            if (part == gq_index) break;
            ++part;
            g0 = k + 1;
        }
    }
    if (part != gq_index) {
        gq_missing = true;
        return n_tok;
    }
    uint32_t g1 = k;
    while (g0 < g1 && vcf_is_ws(tx[g0])) ++g0;
    while (g1 > g0 && vcf_is_ws(tx[g1 - 1])) --g1;
    uint32_t gq = 0;
    if (!(g1 == g0 || (g1 - g0 == 1 && tx[g0] == '.'))) {
        uint32_t s = g0;
        if (tx[s] == '+') ++s;
        bool gok = s < g1;
        for (; gok && s < g1; ++s) {
            const uint32_t d = (uint32_t)tx[s] - '0';
            if (d > 9) gok = false;
            else {
                gq = gq * 10 + d;
                if (gq > 65535) gok = false;
            }
        }
        if (!gok) gq = 0;  // unparsable GQ is treated as 0
    }
    if (gq < min_gq) low_gq = true;
    return n_tok;
}

// Fast path for the field shape that makes up almost all of a 1000-Genomes-style VCF: "a|b:QQ<tab or :>"
// with one-digit alleles and a 1-3 digit GQ as the second key.  w0 / w1 = the field's first eight bytes.
// Returns false when the field is anything else (the general parser then decides).
__device__ __forceinline__ bool vcf_fast_sample(uint32_t w0, uint32_t w1, uint32_t room, uint32_t min_gq,
                                                uint32_t &a0, uint32_t &a1, bool &low_gq) {
    const uint32_t b0 = (w0 & 0xFFu) - '0', b1 = (w0 >> 8) & 0xFFu, b2 = ((w0 >> 16) & 0xFFu) - '0', b3 = w0 >> 24;
    if (b0 > 9u || b2 > 9u || b3 != ':' || (b1 != '|' && b1 != '/')) return false;
    const uint32_t d0 = (w1 & 0xFFu) - '0', c1 = (w1 >> 8) & 0xFFu, c2 = (w1 >> 16) & 0xFFu, c3 = w1 >> 24;
    if (d0 > 9u) return false;
    uint32_t gq = d0, term = c1, used = 6;  // bytes consumed including the terminator
    if (c1 - '0' <= 9u) {
        gq = gq * 10 + (c1 - '0');
        term = c2;
        used = 7;
        if (c2 - '0' <= 9u) {
            gq = gq * 10 + (c2 - '0');
            term = c3;
            used = 8;
        }
    }
    if ((term != '\t' && term != ':') || used > room) return false;
    a0 = b0;
    a1 = b2;
    if (gq < min_gq) low_gq = true;
    return true;
}

// Unaligned 8-byte window at byte p of the text (two words), from three aligned 32-bit loads (L1 hits: the
// owning thread has just scanned these bytes).
__device__ __forceinline__ void vcf_load8(const uint8_t *__restrict__ tx, uint32_t p, uint32_t &w0, uint32_t &w1) {
    const uint32_t *a = reinterpret_cast<const uint32_t *>(tx + (p & ~3u));
    const uint32_t x0 = __ldg(a), x1 = __ldg(a + 1), x2 = __ldg(a + 2);
    const uint32_t sh = (p & 3u) * 8u;
    w0 = __funnelshift_r(x0, x1, sh);
    w1 = __funnelshift_r(x1, x2, sh);
}

// One CTA per candidate line.  The sample region [9th tab, line end) is cut into 256 contiguous spans of whole
// 16-byte words, one per thread: pass 1 counts the span's tabs, ONE block scan per line turns the counts into
// the column index of every span's first field, pass 2 walks the span's tabs in order and parses the kept
// fields (fast path for a|b:QQ, general parser otherwise).  Two barriers per line, no shared-memory staging.
__global__ void __launch_bounds__(256, 4)
fm_k_vcf_samples(VcfParams P, VcfLine *recs, const uint32_t *__restrict__ row_of_line, uint8_t *__restrict__ gt,
                 uint32_t text_cap) {
    __shared__ uint32_t s_scan[9];
    __shared__ uint32_t s_missing, s_low, s_stride, s_err;
    const uint8_t *__restrict__ tx = P.text;
    const uint32_t mp = P.max_ploidy;
    (void)text_cap;
    for (uint32_t line = blockIdx.x; line < P.n_lines; line += gridDim.x) {
        const VcfLine r = recs[line];
        if (r.status != VCF_CAND) continue;  // uniform across the CTA
        if (threadIdx.x == 0) {
            s_missing = 0;
            s_low = 0;
            s_stride = 0;
            s_err = 0;
        }
        const uint32_t e = P.line_start[line + 1];
        const bool write = !(r.indel & 1);
        const bool fast_ok = r.gq_index == 1 && mp >= 2;
        uint8_t *__restrict__ row = write ? gt + (size_t)row_of_line[line] * P.n_samples * mp : nullptr;
        uint32_t miss = 0, stride = 0, err = 0;
        bool low = false;
        if (P.max_idx >= 9) {
            // the 9th tab (at sample_off - 1) opens field 9
            const uint32_t first = r.sample_off - 1;
            const uint32_t t0 = first & ~15u;
            const uint32_t n16 = (e - t0 + 15u) >> 4;          // 16-byte words of the sample region
            const uint32_t per = (n16 + 255u) >> 8;            // words per thread
            const uint32_t q0 = threadIdx.x * per, q1 = min(n16, q0 + per);
            // ---- pass 1: tabs in my span
            uint32_t cnt = 0;
            for (uint32_t q = q0; q < q1; ++q) {
                const uint32_t my = t0 + q * 16u;
                uint32_t tbm = vcf_eq16(*reinterpret_cast<const uint4 *>(tx + my), VCF_TAB);
                if (my < first) tbm &= ~((1u << (first - my)) - 1u);  // bytes before the 9th tab
                if (e - my < 16u) tbm &= (1u << (e - my)) - 1u;       // bytes of the next line
                cnt += __popc(tbm);
            }
            uint32_t total;  // the scan's barriers also order the s_* resets above
            uint32_t f = 8u + vcf_block_scan(cnt, s_scan, total) + 1u;  // field opened by my first tab
            // ---- pass 2: my fields, in order
            for (uint32_t q = q0; q < q1 && f <= (uint32_t)P.max_idx; ++q) {
                const uint32_t my = t0 + q * 16u;
                uint32_t tbm = vcf_eq16(*reinterpret_cast<const uint4 *>(tx + my), VCF_TAB);
                if (my < first) tbm &= ~((1u << (first - my)) - 1u);
                if (e - my < 16u) tbm &= (1u << (e - my)) - 1u;
                while (tbm && f <= (uint32_t)P.max_idx) {
                    const int bit = __ffs(tbm) - 1;
                    tbm &= tbm - 1;
                    const int32_t slot = P.col2slot[f];
                    ++f;
                    if (slot < 0) continue;
                    const uint32_t p = my + bit + 1;  // first byte of the field
                    uint32_t w0, w1, a0, a1;
                    vcf_load8(tx, p, w0, w1);  // stays inside the padded buffer: p < e <= n_bytes
                    if (fast_ok && vcf_fast_sample(w0, w1, e - p, P.min_gq, a0, a1, low)) {
                        stride = stride > 2u ? stride : 2u;
                        if (write) {
                            uint8_t *dst = row + (size_t)slot * mp;
                            if (mp == 2) {
                                *reinterpret_cast<uint16_t *>(dst) = (uint16_t)(a0 | (a1 << 8));
                            } else {
                                dst[0] = (uint8_t)a0;
                                dst[1] = (uint8_t)a1;
                                for (uint32_t k = 2; k < mp; ++k) dst[k] = 0xFF;
                            }
                        }
                    } else {
                        uint8_t al[VCF_MAX_PLOIDY];
                        bool gq_missing = false, too_long = false;
                        const uint32_t n = vcf_parse_sample(tx, p, e, r.gq_index, P.min_gq, mp, al, low, gq_missing,
                                                            too_long);
                        if (too_long) err = VCF_E_PLOIDY > err ? VCF_E_PLOIDY : err;
                        else if (gq_missing) err = err ? err : VCF_E_GQ_MISSING;
                        if (n == 0) ++miss;
                        stride = n > stride ? n : stride;
                        if (write) {
                            uint8_t *dst = row + (size_t)slot * mp;
                            for (uint32_t k = 0; k < mp; ++k) dst[k] = k < n ? al[k] : (uint8_t)0xFF;
                        }
                    }
                }
            }
        } else {
            __syncthreads();  // orders the s_* resets like the scan does
        }
        if (miss) atomicAdd(&s_missing, miss);
        if (low) atomicOr(&s_low, 1u);
        if (stride) atomicMax(&s_stride, stride);
        if (err) atomicMax(&s_err, err);
        __syncthreads();
        if (threadIdx.x == 0) {
            VcfLine o = r;
            o.missing_points = s_missing;
            o.stride = (uint8_t)s_stride;
            if (s_low) o.flags |= 4;      // FLAG_LOW_GQ
            if (s_missing) o.flags |= 8;  // FLAG_MISSING
            if (s_err) o.status = (uint8_t)s_err;
            recs[line] = o;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------- from_variants on the device
// gt rows [*][S][P] with the CompressedGenotypes sentinel (0xFF in slot 0 = None, a later 0xFF ends the
// genotype) -> reference-layout u8 matrix with in-band missingness (cells >= 0x80 missing), rows in
// output order.  max_allele = largest valid allele byte.
__global__ void __launch_bounds__(256)
fm_k_vcf_to_matrix(const uint8_t *__restrict__ gt, const uint32_t *__restrict__ order, uint64_t n_rows, uint32_t S,
                   uint32_t P, uint32_t ploidy, uint8_t *__restrict__ data, uint32_t *__restrict__ max_allele) {
    const uint64_t total = n_rows * S;
    uint32_t mx = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / S;
        const uint32_t s = (uint32_t)(i - r * S);
        const uint8_t *src = gt + ((size_t)order[r] * S + s) * P;
        uint8_t *dst = data + i * ploidy;
        bool valid = true;
        for (uint32_t k = 0; k < ploidy; ++k) {
            const uint8_t a = k < P ? src[k] : (uint8_t)0xFF;
            valid = valid && a != 0xFF;
            dst[k] = valid ? a : (uint8_t)0x80;
            if (valid && a > mx) mx = a;
        }
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(max_allele, mx);
}

// Diploid batches (P == ploidy == 2) whose rows are whole 16-byte words: eight samples per thread, byte-parallel
// within each 32-bit word -- 0xFF bytes (and the second byte of a pair whose first is 0xFF) become 0x80.
__global__ void __launch_bounds__(256)
fm_k_vcf_to_matrix_p2(const uint4 *__restrict__ gt, const uint32_t *__restrict__ order, uint64_t n_rows,
                      uint32_t row_u4, uint4 *__restrict__ data, uint32_t *__restrict__ max_allele) {
    const uint64_t total = n_rows * row_u4;
    uint32_t mx = 0;  // per-byte running maximum
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / row_u4;
        const uint4 x = gt[(size_t)order[r] * row_u4 + (i - r * row_u4)];
        uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t t = ~w[k];                                                       // zero byte <=> 0xFF cell
            const uint32_t z = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;
            uint32_t m = (z >> 7) * 0xFFu;                                                  // 0xFF in every sentinel byte
            m |= (m & 0x00FF00FFu) << 8;                                                    // ... and in the byte after a pair's first
            const uint32_t v = w[k] & ~m;
            mx = __vmaxu4(mx, v);
            w[k] = v | (0x80808080u & m);
        }
        data[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    mx = max(max(mx & 0xFFu, (mx >> 8) & 0xFFu), max((mx >> 16) & 0xFFu, mx >> 24));
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(max_allele, mx);
}

// from_variants straight into PACKED rows (no u8 matrix between the parser and the bitplanes): one
// thread per output word = 32 cells c = sample * ploidy + side of row r; allele bit = the cell is called and its
// allele index is non-zero, called bit = every side up to this one holds an allele (a 0xFF sentinel ends the
// genotype, process.rs:430-478; absent sides of a short genotype are missing, stats.rs:452-460).  max_allele
// receives the largest allele index: above 1 the packed form does not apply and the caller builds the u8 matrix.
__global__ void __launch_bounds__(256)
fm_k_vcf_to_packed(const uint8_t *__restrict__ gt, const uint32_t *__restrict__ order, uint64_t n_rows, uint32_t S,
                   uint32_t P, uint32_t ploidy, uint32_t rw, uint32_t *__restrict__ abits, uint32_t *__restrict__ cbits,
                   uint32_t *__restrict__ max_allele) {
    const uint64_t total = n_rows * rw;
    const uint32_t stride = S * ploidy;
    uint32_t mx = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / rw;
        const uint32_t w = (uint32_t)(i - r * rw);
        const uint8_t *row = gt + (size_t)order[r] * S * P;
        uint32_t a = 0, c = 0;
        const uint32_t c0 = w * 32u, c1 = min(stride, c0 + 32u);
        uint32_t s = c0 / ploidy, k = c0 - s * ploidy;
        // validity of side k needs the sides before it: start at the sample's first side
        bool valid = true;
        for (uint32_t kk = 0; kk < k; ++kk) valid = valid && (kk < P ? row[(size_t)s * P + kk] : (uint8_t)0xFF) != 0xFF;
        for (uint32_t cell = c0; cell < c1; ++cell) {
            const uint8_t v = k < P ? row[(size_t)s * P + k] : (uint8_t)0xFF;
            valid = valid && v != 0xFF;
            if (valid) {
                c |= 1u << (cell - c0);
                if (v) a |= 1u << (cell - c0);
                mx = max(mx, (uint32_t)v);
            }
            if (++k == ploidy) {
                k = 0;
                ++s;
                valid = true;
            }
        }
        abits[i] = a;
        cbits[i] = c;
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(max_allele, mx);
}

__global__ void __launch_bounds__(256)
fm_k_vcf_gather_rows(const uint8_t *__restrict__ gt, const uint32_t *__restrict__ order, uint64_t n_rows,
                     uint32_t row_bytes, uint8_t *__restrict__ out) {
    const uint64_t total = n_rows * row_bytes;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / row_bytes;
        out[i] = gt[(size_t)order[r] * row_bytes + (i - r * row_bytes)];
    }
}

}  // namespace fm
