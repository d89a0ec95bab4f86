/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's VCF parse/filter stage, the
 * compiled sibling of oracle/vcf.py (which stays the line-by-line reading of process.rs): process_variant
 * (process.rs:4471-4768) over every data line with pthreads, line-local statistics merged only for
 * lines that returned Ok (process.rs:4262-4370), output sorted by (position, compressed genotype bytes)
 * (process.rs:4377-4386).  Used by tests/ (cross-check of the two restatements) and as the CPU baseline
 * of tools/bench_vcf.py.  Never linked into or called from the product. */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { CAND = 0, SKIP = 1, E_FEW = 10, E_COL = 11, E_POS = 12, E_POS1 = 13, E_NOGQ = 14, E_GQMISS = 15, E_PLOIDY = 16 };

typedef struct {
    int64_t pos0;
    uint32_t missing_points;
    uint8_t status, flags, indel, stride, ref, n_alt, alts[7];
} line_rec;

typedef struct {
    const char *text;
    const size_t *ls; /* line starts, n_lines + 1 */
    size_t n_lines;
    const char *chr;
    size_t chr_len;
    const int64_t *regions;
    size_t n_regions;
    const uint32_t *kept;
    size_t n_kept;
    uint32_t min_gq;
    int allow_mode, mask_mode;
    const int64_t *allow, *mask;
    size_t n_allow, n_mask;
    size_t P;
    line_rec *recs;
    uint8_t *gt; /* [n_lines][n_kept][P], rows of dropped lines stay unused */
    size_t t0, t1;
} job;

static int is_ws(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }
static uint8_t nuc(unsigned char c) {
    switch (c) {
        case 'A': case 'a': return 'A';
        case 'C': case 'c': return 'C';
        case 'G': case 'g': return 'G';
        case 'T': case 't': return 'T';
        default: return 'N';
    }
}
/* <uN as FromStr>: optional '+', digits only, no overflow */
static int parse_unsigned(const char *s, size_t n, uint32_t limit, uint32_t *out) {
    size_t i = 0;
    if (n && s[0] == '+') i = 1;
    if (i >= n) return 0;
    uint64_t v = 0;
    for (; i < n; ++i) {
        if (s[i] < '0' || s[i] > '9') return 0;
        v = v * 10 + (uint64_t)(s[i] - '0');
        if (v > limit) return 0;
    }
    *out = (uint32_t)v;
    return 1;
}

static void one_line(const job *J, size_t li) {
    const char *b = J->text + J->ls[li], *e = J->text + J->ls[li + 1];
    line_rec *r = &J->recs[li];
    memset(r, 0, sizeof *r);
    r->ref = 'N';
    /* fields = line.split('\t') */
    size_t n_fields = 1;
    for (const char *p = b; p < e; ++p) n_fields += *p == '\t';
    r->pos0 = (int64_t)n_fields;
    if (n_fields < 9) { r->status = E_FEW; return; }
    const uint32_t max_idx = J->n_kept ? J->kept[J->n_kept - 1] : 0;
    if (J->n_kept && n_fields <= max_idx) { r->status = E_COL; return; }
    const char *f[10];
    f[0] = b;
    {
        int k = 1;
        for (const char *p = b; p < e && k < 10; ++p)
            if (*p == '\t') f[k++] = p + 1;
        if (k < 10) f[9] = e + 1; /* exactly nine fields: FORMAT runs to the end of the line */
    }
    /* CHROM */
    const char *c0 = f[0], *c1 = f[1] - 1;
    while (c0 < c1 && is_ws((unsigned char)*c0)) ++c0;
    while (c1 > c0 && is_ws((unsigned char)c1[-1])) --c1;
    if (c1 - c0 >= 3 && (!memcmp(c0, "chr", 3) || !memcmp(c0, "Chr", 3) || !memcmp(c0, "CHR", 3))) c0 += 3;
    if ((size_t)(c1 - c0) != J->chr_len || memcmp(c0, J->chr, J->chr_len)) { r->status = SKIP; return; }
    /* POS: i64::from_str */
    {
        const char *p = f[1], *pe = f[2] - 1;
        int neg = 0, ok = 1;
        if (p < pe && (*p == '+' || *p == '-')) neg = *p++ == '-';
        if (p >= pe) ok = 0;
        uint64_t mag = 0;
        const uint64_t lim = neg ? (1ull << 63) : (1ull << 63) - 1;
        for (; ok && p < pe; ++p) {
            if (*p < '0' || *p > '9') { ok = 0; break; }
            const uint64_t d = (uint64_t)(*p - '0');
            if (mag > (lim - d) / 10) { ok = 0; break; }
            mag = mag * 10 + d;
        }
        if (!ok) { r->status = E_POS; return; }
        const int64_t p1 = neg ? (int64_t)(0 - mag) : (int64_t)mag;
        r->pos0 = (int64_t)((uint64_t)p1 - 1);
        if (p1 < 1) { r->status = E_POS1; return; }
    }
    /* regions: partition_point(|r| r.end <= pos), then start <= pos */
    {
        size_t lo = 0, hi = J->n_regions;
        while (lo < hi) {
            const size_t mid = (lo + hi) / 2;
            if (J->regions[2 * mid + 1] <= r->pos0) lo = mid + 1; else hi = mid;
        }
        if (!(lo < J->n_regions && J->regions[2 * lo] <= r->pos0)) { r->status = SKIP; return; }
    }
    if (J->allow_mode == 2) r->flags |= 2;
    if (J->allow_mode == 1) {
        int in = 0;
        for (size_t i = 0; i < J->n_allow && !in; ++i) in = r->pos0 >= J->allow[2 * i] && r->pos0 < J->allow[2 * i + 1];
        if (!in) r->flags |= 2;
    }
    if (J->mask_mode == 1) {
        const uint64_t p = (uint64_t)r->pos0;
        for (size_t i = 0; i < J->n_mask; ++i) {
            const uint64_t s = (uint64_t)J->mask[2 * i], en = (uint64_t)J->mask[2 * i + 1];
            const uint64_t a = p > s ? p : s, z = p + 1 < en ? p + 1 : en;
            if (a < z) { r->flags |= 1; break; }
        }
    }
    /* length guard, allele info */
    {
        const char *r0 = f[3], *r1 = f[4] - 1, *a0 = f[4], *a1 = f[5] - 1;
        int indel = (r1 - r0) != 1, not_one = 0, longer = 0;
        uint32_t n_alt = 0;
        const char *seg = a0;
        for (const char *p = a0; p <= a1; ++p)
            if (p == a1 || *p == ',') {
                const size_t len = (size_t)(p - seg);
                if (len != 1) not_one = 1;
                if (len > 1) longer = 1;
                if (n_alt < 7) r->alts[n_alt] = len ? nuc((unsigned char)*seg) : 'N';
                ++n_alt;
                seg = p + 1;
            }
        if (!indel && not_one) { indel = 1; if (longer) r->indel |= 2; }
        if (indel) r->indel |= 1;
        r->ref = r1 > r0 ? nuc((unsigned char)*r0) : 'N';
        r->n_alt = (uint8_t)(n_alt < 255 ? n_alt : 255);
    }
    /* GQ key in FORMAT */
    uint32_t gq_index = 0;
    {
        const char *f0 = f[8], *f1 = f[9] - 1, *seg = f[8];
        int have = 0;
        for (const char *p = f0; p <= f1 && !have; ++p)
            if (p == f1 || *p == ':') {
                if (p - seg == 2 && seg[0] == 'G' && seg[1] == 'Q') have = 1;
                else { ++gq_index; seg = p + 1; }
            }
        if (!have) { r->status = E_NOGQ; return; }
    }
    /* genotype loop, then GQ loop (an Err in the GQ loop drops the whole line) */
    uint8_t *row = J->gt + li * J->n_kept * J->P;
    const char *p = b;
    size_t col = 0, k = 0;
    int low = 0, err = 0;
    uint32_t miss = 0, stride = 0;
    while (k < J->n_kept) {
        while (col < J->kept[k]) { p = memchr(p, '\t', (size_t)(e - p)); ++p; ++col; }
        const char *fe = memchr(p, '\t', (size_t)(e - p));
        if (!fe) fe = e;
        const char *q = memchr(p, ':', (size_t)(fe - p));
        if (!q) q = fe;
        uint8_t al[16];
        uint32_t n_tok = 0;
        int ok = 1;
        const char *seg = p;
        for (const char *t = p; t <= q; ++t)
            if (t == q || *t == '|' || *t == '/') {
                uint32_t v;
                if (!parse_unsigned(seg, (size_t)(t - seg), 255, &v)) ok = 0;
                else if (n_tok < 16) al[n_tok] = (uint8_t)v;
                ++n_tok;
                seg = t + 1;
            }
        uint8_t *dst = row + k * J->P;
        if (!ok) {
            ++miss;
            memset(dst, 0xFF, J->P);
        } else if (n_tok > J->P) {
            err = E_PLOIDY;
            memset(dst, 0xFF, J->P);
        } else {
            for (size_t i = 0; i < J->P; ++i) dst[i] = i < n_tok ? al[i] : 0xFF;
            if (n_tok > stride) stride = n_tok;
            /* gq_index-th ':' part of the whole field, trimmed */
            const char *g0 = p, *g1 = NULL;
            uint32_t part = 0;
            for (const char *t = p; t <= fe; ++t)
                if (t == fe || *t == ':') {
                    if (part == gq_index) { g1 = t; break; }
                    ++part;
                    g0 = t + 1;
                }
            if (!g1) {
                if (err < E_GQMISS) err = E_GQMISS;
            } else {
                while (g0 < g1 && is_ws((unsigned char)*g0)) ++g0;
                while (g1 > g0 && is_ws((unsigned char)g1[-1])) --g1;
                uint32_t gq = 0;
                if (!(g1 == g0 || (g1 - g0 == 1 && *g0 == '.')))
                    if (!parse_unsigned(g0, (size_t)(g1 - g0), 65535, &gq)) gq = 0;
                if (gq < J->min_gq) low = 1;
            }
        }
        ++k;
    }
    r->missing_points = miss;
    r->stride = (uint8_t)stride;
    if (low) r->flags |= 4;
    if (miss) r->flags |= 8;
    if (err) r->status = (uint8_t)err;
}

static void *worker(void *arg) {
    const job *J = arg;
    for (size_t li = J->t0; li < J->t1; ++li) one_line(J, li);
    return NULL;
}

typedef struct { int64_t pos; const uint8_t *data; size_t len; size_t idx; } sort_key;
static int cmp_key(const void *a, const void *b) {
    const sort_key *x = a, *y = b;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    const size_t n = x->len < y->len ? x->len : y->len;
    const int c = memcmp(x->data, y->data, n);
    if (c) return c;
    if (x->len != y->len) return x->len < y->len ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

/* counters[9]: total_variants, filtered_variants, filtered_due_to_mask, filtered_due_to_allow,
 * missing_data_variants, low_gq_variants, mnp_variants, total_data_points, missing_data_points.
 * pos0/flags/stride: capacity n_lines; gt_out: [n_lines][n_kept][P] (first *n_variants rows filled, output order);
 * err_line/err_code: capacity n_lines.  Returns the number of lines. */
size_t orc_vcf_process_lines(const char *text, size_t n_bytes, const char *chr_norm, const int64_t *regions,
                             size_t n_regions, const uint32_t *kept, size_t n_kept, uint32_t min_gq, int allow_mode,
                             const int64_t *allow, size_t n_allow, int mask_mode, const int64_t *mask, size_t n_mask,
                             size_t P, int threads, uint64_t *counters, size_t *n_variants, int64_t *pos0,
                             uint8_t *flags, uint8_t *stride, uint8_t *gt_out, size_t *n_errors, uint64_t *err_line,
                             int32_t *err_code) {
    size_t n_lines = 0;
    for (const char *p = text; p < text + n_bytes;) {
        const char *nl = memchr(p, '\n', (size_t)(text + n_bytes - p));
        p = nl ? nl + 1 : text + n_bytes;
        ++n_lines;
    }
    size_t *ls = malloc((n_lines + 1) * sizeof *ls);
    {
        size_t i = 0;
        for (const char *p = text; p < text + n_bytes;) {
            ls[i++] = (size_t)(p - text);
            const char *nl = memchr(p, '\n', (size_t)(text + n_bytes - p));
            p = nl ? nl + 1 : text + n_bytes;
        }
        ls[n_lines] = n_bytes;
    }
    line_rec *recs = malloc((n_lines ? n_lines : 1) * sizeof *recs);
    uint8_t *gt = malloc(n_lines * n_kept * P + 1);
    if (threads < 1) threads = 1;
    job *jobs = malloc((size_t)threads * sizeof *jobs);
    pthread_t *th = malloc((size_t)threads * sizeof *th);
    for (int t = 0; t < threads; ++t) {
        job J = {text, ls, n_lines, chr_norm, strlen(chr_norm), regions, n_regions, kept, n_kept, min_gq, allow_mode,
                 mask_mode, allow, mask, n_allow, n_mask, P, recs, gt, n_lines * (size_t)t / (size_t)threads,
                 n_lines * (size_t)(t + 1) / (size_t)threads};
        jobs[t] = J;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    memset(counters, 0, 9 * sizeof *counters);
    sort_key *keys = malloc((n_lines ? n_lines : 1) * sizeof *keys);
    uint8_t *compact = malloc(n_lines * n_kept * P + 1);
    size_t nv = 0, ne = 0;
    for (size_t li = 0; li < n_lines; ++li) {
        const line_rec *r = &recs[li];
        if (r->status == SKIP) continue;
        if (r->status != CAND) { err_line[ne] = li; err_code[ne++] = r->status; continue; }
        const int indel = r->indel & 1;
        counters[0]++;
        if (r->flags & 2) counters[3]++;
        if (r->flags & 1) counters[2]++;
        if (r->indel & 2) counters[6]++;
        if (r->flags & 4) counters[5]++;
        if (r->flags & 8) counters[4]++;
        counters[7] += n_kept;
        counters[8] += r->missing_points;
        if (r->flags != 0 || indel) counters[1]++;
        if (indel) continue;
        const size_t st = r->stride ? r->stride : (n_kept ? 1 : 0);
        uint8_t *c = compact + li * n_kept * P;
        for (size_t s = 0; s < n_kept; ++s) memcpy(c + s * st, gt + (li * n_kept + s) * P, st);
        keys[nv].pos = r->pos0;
        keys[nv].data = c;
        keys[nv].len = n_kept * st;
        keys[nv].idx = li;
        ++nv;
    }
    qsort(keys, nv, sizeof *keys, cmp_key);
    for (size_t i = 0; i < nv; ++i) {
        const line_rec *r = &recs[keys[i].idx];
        pos0[i] = r->pos0;
        flags[i] = r->flags;
        stride[i] = r->stride ? r->stride : (uint8_t)(n_kept ? 1 : 0);
        memcpy(gt_out + i * n_kept * P, gt + keys[i].idx * n_kept * P, n_kept * P);
    }
    *n_variants = nv;
    *n_errors = ne;
    free(compact); free(keys); free(th); free(jobs); free(gt); free(recs); free(ls);
    return n_lines;
}
