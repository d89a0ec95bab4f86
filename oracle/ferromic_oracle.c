/*
 * ferromic_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See ferromic_oracle.h for scope, pinning and floating-point notes.
 * All citations are file:line inside the reference crate (/root/reference).
 */
#include "ferromic_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define FST_EPSILON 1e-12 /* stats.rs:26 */
#define INVALID_GROUP 0xFFFFu /* stats.rs:1080 */

void orc_free(void *p) { free(p); }

/* ---------- sparse genotype access: CompressedGenotypes::get (process.rs:479-496) ---------- */
static inline size_t gt_get(const orc_variants *vs, size_t v, size_t s, const uint8_t **gt) {
    /* returns genotype length, 0 => None */
    if (!vs || s >= vs->n_samples || vs->stride == 0) return 0;
    const uint8_t *p = vs->gt + (v * vs->n_samples + s) * vs->stride;
    if (p[0] == ORC_MISSING) return 0;
    size_t len = 0;
    while (len < vs->stride && p[len] != ORC_MISSING) len++;
    *gt = p;
    return len;
}

static inline size_t nvar(const orc_variants *vs) { return vs ? vs->n_variants : 0; }

/* ---------- QueryRegion (process.rs:559-585) ---------- */
static inline int region_contains(int64_t rs, int64_t re, int64_t pos) { return pos >= rs && pos <= re; }
static int64_t region_len(int64_t rs, int64_t re) {
    if (rs > re) return 0;
    /* ZeroBasedHalfOpen::from_0based_inclusive (process.rs:210-222) */
    int64_t as = rs < 0 ? 0 : rs;
    int64_t ae;
    if (re < as) {
        ae = as;
    } else {
        ae = (re == INT64_MAX) ? INT64_MAX : re + 1;
        if (ae < as) ae = as;
    }
    return ae > as ? ae - as : 0;
}

/* ---------- a1: from_variants (stats.rs:339-500) ---------- */
int orc_dense_from_variants(const orc_variants *vs, size_t sample_count, uint8_t **data_out,
                            uint64_t **missing_out, size_t *ploidy_out, uint8_t *max_allele_out) {
    size_t V = nvar(vs);
    if (V == 0) return 1; /* :340-342 */
    size_t max_ploidy = 0; /* :349-359 */
    for (size_t v = 0; v < V; v++)
        for (size_t s = 0; s < vs->n_samples; s++) {
            const uint8_t *g;
            size_t len = gt_get(vs, v, s, &g);
            if (len > max_ploidy) max_ploidy = len;
        }
    if (max_ploidy == 0) return 1; /* :361-364 */
    size_t stride = sample_count * max_ploidy;
    size_t total = V * stride;
    uint8_t *data = (uint8_t *)calloc(total ? total : 1, 1);
    size_t words = (total + 63) / 64;
    uint64_t *missing = (uint64_t *)calloc(words ? words : 1, sizeof(uint64_t));
    uint8_t gmax = 0;
    for (size_t v = 0; v < V; v++) {
        for (size_t s = 0; s < sample_count; s++) { /* :439-461 */
            size_t off = v * stride + s * max_ploidy;
            const uint8_t *g;
            size_t len = gt_get(vs, v, s, &g);
            if (len > 0) {
                size_t lim = len < max_ploidy ? len : max_ploidy;
                for (size_t i = 0; i < lim; i++) data[off + i] = g[i];
                for (size_t i = len; i < max_ploidy; i++) {
                    size_t idx = off + i;
                    missing[idx >> 6] |= (uint64_t)1 << (idx & 63);
                }
            } else {
                for (size_t i = 0; i < max_ploidy; i++) {
                    size_t idx = off + i;
                    missing[idx >> 6] |= (uint64_t)1 << (idx & 63);
                }
            }
        }
    }
    for (size_t i = 0; i < total; i++) /* :490 */
        if (data[i] > gmax) gmax = data[i];
    *data_out = data;
    *missing_out = missing;
    *ploidy_out = max_ploidy;
    *max_allele_out = gmax;
    return 0;
}

/* ---------- a2: memberships ---------- */
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

/* DenseMembership::build (stats.rs:1251-1284) */
size_t orc_dense_membership(const orc_dense *m, const orc_haps *h, uint64_t *offsets) {
    size_t S = m->n_samples, P = m->ploidy, n = 0;
    uint8_t *left = (uint8_t *)calloc(S ? S : 1, 1), *right = (uint8_t *)calloc(S ? S : 1, 1);
    for (size_t i = 0; i < h->n; i++) {
        uint64_t s = h->sample[i];
        if (s >= S) continue;
        if (h->side[i] == 0) {
            if (!left[s]) {
                left[s] = 1;
                offsets[n++] = s * P;
            }
        } else {
            if (P <= 1) continue;
            if (!right[s]) {
                right[s] = 1;
                offsets[n++] = s * P + 1;
            }
        }
    }
    free(left);
    free(right);
    qsort(offsets, n, sizeof(uint64_t), cmp_u64);
    return n;
}

/* HapMembership::build (stats.rs:1211-1238) */
typedef struct {
    uint8_t *left, *right;
    size_t n, total;
} hapmem;

static hapmem hapmem_build(size_t sample_count, const orc_haps *h) {
    hapmem m;
    m.n = sample_count;
    m.left = (uint8_t *)calloc(sample_count ? sample_count : 1, 1);
    m.right = (uint8_t *)calloc(sample_count ? sample_count : 1, 1);
    m.total = 0;
    for (size_t i = 0; i < h->n; i++) {
        uint64_t s = h->sample[i];
        if (s >= sample_count) continue;
        if (h->side[i] == 0) {
            if (!m.left[s]) {
                m.left[s] = 1;
                m.total++;
            }
        } else if (!m.right[s]) {
            m.right[s] = 1;
            m.total++;
        }
    }
    return m;
}
static void hapmem_free(hapmem *m) {
    free(m->left);
    free(m->right);
}
static inline int hm_left(const hapmem *m, size_t i) { return i < m->n ? m->left[i] : 0; }
static inline int hm_right(const hapmem *m, size_t i) { return i < m->n ? m->right[i] : 0; }

/* ---------- a3: innermost loops ---------- */
static inline int dense_missing(const uint64_t *bits, size_t idx) { /* stats.rs:1298-1302 */
    return (int)((bits[idx >> 6] >> (idx & 63)) & 1);
}
static inline size_t dense_sum_alt_no_missing(const uint8_t *data, size_t base, const uint64_t *off,
                                              size_t n) { /* stats.rs:1665-1674 */
    size_t sum = 0;
    const uint8_t *ptr = data + base;
    for (size_t i = 0; i < n; i++) sum += ptr[off[i]];
    return sum;
}
static inline void dense_sum_alt_with_missing(const uint8_t *data, size_t base, const uint64_t *off,
                                              size_t n, const uint64_t *bits, size_t *total_out,
                                              size_t *alt_out) { /* stats.rs:1677-1697 */
    size_t alt = 0, total = 0;
    const uint8_t *ptr = data + base;
    for (size_t i = 0; i < n; i++) {
        size_t idx = base + off[i];
        if (dense_missing(bits, idx)) continue;
        alt += ptr[off[i]];
        total += 1;
    }
    *total_out = total;
    *alt_out = alt;
}

/* ---------- a5: per-site pi variants ---------- */
static inline int dense_pi_from_counts(size_t total_called, size_t alt_count, double *out) {
    /* stats.rs:1700-1709 */
    if (total_called < 2) return 0;
    double n = (double)total_called;
    double alt = (double)alt_count;
    double ref_count = (double)(total_called - alt_count);
    double sum_sq = ref_count * ref_count + alt * alt;
    *out = n / (n - 1.0) * (1.0 - sum_sq / (n * n));
    return 1;
}
static inline int pi_from_components(size_t total_called, double sum_counts_sq, double *out) {
    /* stats.rs:2723-2733 */
    if (total_called < 2) return 0;
    double n = (double)total_called;
    double inv_n = 1.0 / n;
    double sum_p2 = sum_counts_sq * inv_n * inv_n;
    *out = n / (n - 1.0) * (1.0 - sum_p2);
    return 1;
}

/* ---------- a4: build_dense_population_summary (stats.rs:1367-1470) ---------- */
static void summary_range(const orc_dense *m, const uint64_t *off, size_t n, size_t v0, size_t v1,
                          uint32_t *alt_counts, uint32_t *called_counts, size_t *seg_out,
                          double *pi_out) {
    size_t stride = m->n_samples * m->ploidy;
    size_t seg = 0;
    double pi_total = 0.0;
    if (m->missing) {
        for (size_t v = v0; v < v1; v++) {
            size_t called, alt;
            dense_sum_alt_with_missing(m->data, v * stride, off, n, m->missing, &called, &alt);
            alt_counts[v] = (uint32_t)alt;
            called_counts[v] = (uint32_t)called;
            if (called >= 2 && alt > 0 && alt < called) seg++;
            double val;
            if (dense_pi_from_counts(called, alt, &val)) pi_total += val;
        }
    } else {
        size_t total = n;
        for (size_t v = v0; v < v1; v++) {
            size_t alt = dense_sum_alt_no_missing(m->data, v * stride, off, n);
            alt_counts[v] = (uint32_t)alt;
            called_counts[v] = (uint32_t)total;
            if (alt > 0 && alt < total) seg++;
            double val;
            if (dense_pi_from_counts(total, alt, &val)) pi_total += val;
        }
    }
    *seg_out = seg;
    *pi_out = pi_total;
}

void orc_build_summary(const orc_dense *m, const orc_haps *h, orc_summary *out) {
    uint64_t *off = (uint64_t *)malloc((h->n ? h->n : 1) * sizeof(uint64_t));
    size_t n = orc_dense_membership(m, h, off);
    out->len = m->n_variants;
    out->capacity = n;
    summary_range(m, off, n, 0, m->n_variants, out->alt, out->called, &out->seg, &out->pi_sum);
    free(off);
}

typedef struct {
    const orc_dense *m;
    const uint64_t *off;
    size_t n, v0, v1;
    uint32_t *alt, *called;
    size_t seg;
    double pi;
} summary_job;

static void *summary_worker(void *arg) {
    summary_job *j = (summary_job *)arg;
    summary_range(j->m, j->off, j->n, j->v0, j->v1, j->alt, j->called, &j->seg, &j->pi);
    return NULL;
}

void orc_build_summary_mt(const orc_dense *m, const orc_haps *h, orc_summary *out, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    uint64_t *off = (uint64_t *)malloc((h->n ? h->n : 1) * sizeof(uint64_t));
    size_t n = orc_dense_membership(m, h, off);
    out->len = m->n_variants;
    out->capacity = n;
    size_t V = m->n_variants;
    summary_job *jobs = (summary_job *)calloc((size_t)nthreads, sizeof(summary_job));
    pthread_t *tids = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    size_t chunk = (V + (size_t)nthreads - 1) / (size_t)nthreads;
    for (int t = 0; t < nthreads; t++) {
        size_t v0 = (size_t)t * chunk, v1 = v0 + chunk;
        if (v0 > V) v0 = V;
        if (v1 > V) v1 = V;
        jobs[t] = (summary_job){m, off, n, v0, v1, out->alt, out->called, 0, 0.0};
        pthread_create(&tids[t], NULL, summary_worker, &jobs[t]);
    }
    size_t seg = 0;
    double pi = 0.0;
    for (int t = 0; t < nthreads; t++) {
        pthread_join(tids[t], NULL);
        seg += jobs[t].seg;
        pi += jobs[t].pi; /* chunk order; the reference's rayon reduce is unordered */
    }
    out->seg = seg;
    out->pi_sum = pi;
    free(jobs);
    free(tids);
    free(off);
}

/* ---------- calculate_pi_from_summary_with_precomputed (stats.rs:1480-1542) ---------- */
static inline int64_t sat_sub_i64(int64_t a, int64_t b) {
    /* i64::saturating_sub */
    if (b > 0 && a < INT64_MIN + b) return INT64_MIN;
    if (b < 0 && a > INT64_MAX + b) return INT64_MAX;
    return a - b;
}

double orc_pi_from_summary(const orc_summary *s, int64_t L, int has_pre, double pre) {
    if (s->capacity <= 1) return NAN;
    if (L < 0) return 0.0;
    if (L == 0) return INFINITY;
    size_t uncallable = 0;
    for (size_t i = 0; i < s->len; i++)
        if (s->called[i] < 2) uncallable++;
    int64_t eff = sat_sub_i64(L, (int64_t)uncallable);
    if (eff == 0) return NAN;
    double sum_pi = has_pre ? pre : s->pi_sum;
    return sum_pi / (double)eff;
}

/* ---------- a6: segregating sites ---------- */
size_t orc_count_segregating_sites(const orc_variants *vs) { /* stats.rs:3808-3829 */
    size_t count = 0;
    for (size_t v = 0; v < nvar(vs); v++) {
        int have_first = 0, seg = 0;
        uint8_t first = 0;
        for (size_t s = 0; s < vs->n_samples && !seg; s++) {
            const uint8_t *g;
            size_t len = gt_get(vs, v, s, &g);
            for (size_t k = 0; k < len; k++) {
                if (!have_first) {
                    first = g[k];
                    have_first = 1;
                } else if (first != g[k]) {
                    seg = 1;
                    break;
                }
            }
        }
        count += (size_t)seg;
    }
    return count;
}

/* count_segregating_sites_for_haplotypes (stats.rs:3858-3889): raw list, no de-dup */
static size_t seg_sites_for_haplotypes(const orc_variants *vs, const orc_haps *h) {
    size_t count = 0;
    for (size_t v = 0; v < nvar(vs); v++) {
        int have_first = 0, seg = 0;
        uint8_t first = 0;
        for (size_t i = 0; i < h->n; i++) {
            const uint8_t *g;
            size_t len = gt_get(vs, v, (size_t)h->sample[i], &g);
            if (len == 0) continue;
            size_t side = h->side[i];
            if (side >= len) continue;
            uint8_t allele = g[side];
            if (!have_first) {
                first = allele;
                have_first = 1;
            } else if (first != allele) {
                seg = 1;
                break;
            }
        }
        count += (size_t)seg;
    }
    return count;
}

/* count_segregating_sites_dense (+_biallelic) (stats.rs:3891-4084) */
static size_t seg_sites_dense(const orc_dense *m, const uint64_t *off, size_t n) {
    size_t stride = m->n_samples * m->ploidy;
    size_t seg = 0;
    if (m->max_allele <= 1) { /* :4028-4084 */
        if (n < 2) return 0;
        for (size_t v = 0; v < m->n_variants; v++) {
            if (m->missing) {
                size_t called, alt;
                dense_sum_alt_with_missing(m->data, v * stride, off, n, m->missing, &called, &alt);
                if (called >= 2 && alt > 0 && alt < called) seg++;
            } else {
                size_t alt = dense_sum_alt_no_missing(m->data, v * stride, off, n);
                if (alt > 0 && alt < n) seg++;
            }
        }
        return seg;
    }
    if (n == 0) return 0; /* :3900-3902 */
    for (size_t v = 0; v < m->n_variants; v++) {
        size_t base = v * stride;
        uint8_t first = 0;
        int seen = 0, poly = 0;
        for (size_t i = 0; i < n; i++) {
            size_t idx = base + off[i];
            if (m->missing && dense_missing(m->missing, idx)) continue;
            uint8_t a = m->data[idx];
            if (seen) {
                if (a != first) {
                    poly = 1;
                    break;
                }
            } else {
                first = a;
                seen = 1;
            }
        }
        seg += (size_t)poly;
    }
    return seg;
}

size_t orc_count_segregating_sites_for_population(const orc_pop *p) { /* stats.rs:3831-3851 */
    if (p->summary) return p->summary->seg;
    if (p->dense && p->dense->ploidy == 2) {
        uint64_t *off = (uint64_t *)malloc((p->haps.n ? p->haps.n : 1) * sizeof(uint64_t));
        size_t n = orc_dense_membership(p->dense, &p->haps, off);
        size_t r = (n <= 1) ? 0 : seg_sites_dense(p->dense, off, n);
        free(off);
        return r;
    }
    return seg_sites_for_haplotypes(p->variants, &p->haps);
}

/* ---------- compute_pi_metrics_fast (stats.rs:2761-2821) ---------- */
typedef struct {
    size_t total_called;
    double sum_counts_sq;
    size_t distinct;
} pi_metrics;

static pi_metrics pi_metrics_fast(const orc_variants *vs, size_t v, const hapmem *mem) {
    uint32_t counts[256];
    uint8_t used[256];
    size_t n_used = 0, total = 0;
    memset(counts, 0, sizeof(counts));
    for (size_t s = 0; s < vs->n_samples; s++) {
        const uint8_t *g;
        size_t len = gt_get(vs, v, s, &g);
        if (len == 0) continue;
        if (hm_left(mem, s)) { /* len >= 1 here */
            uint8_t a = g[0];
            if (counts[a] == 0) used[n_used++] = a;
            counts[a]++;
            total++;
        }
        if (hm_right(mem, s) && len > 1) {
            uint8_t a = g[1];
            if (counts[a] == 0) used[n_used++] = a;
            counts[a]++;
            total++;
        }
    }
    double ssq = 0.0;
    for (size_t i = 0; i < n_used; i++) {
        double c = (double)counts[used[i]];
        ssq += c * c;
    }
    pi_metrics r = {total, ssq, n_used};
    return r;
}

/* ---------- a7: calculate_pi sparse (stats.rs:4317-4432) ---------- */
double orc_pi_sparse(const orc_variants *vs, const orc_haps *h, int64_t L) {
    if (h->n <= 1) return NAN;
    if (L < 0) return 0.0;
    if (L == 0) return INFINITY;
    size_t variant_sample_count = nvar(vs) ? vs->n_samples : 0;
    size_t hap_sample_count = 0;
    for (size_t i = 0; i < h->n; i++) {
        uint64_t s1 = h->sample[i] == UINT64_MAX ? UINT64_MAX : h->sample[i] + 1;
        if (s1 > hap_sample_count) hap_sample_count = (size_t)s1;
    }
    size_t sample_count = variant_sample_count > hap_sample_count ? variant_sample_count : hap_sample_count;
    hapmem mem = hapmem_build(sample_count, h);
    if (mem.total <= 1) {
        hapmem_free(&mem);
        return NAN;
    }
    double sum_pi = 0.0;
    size_t skipped = 0;
    for (size_t v = 0; v < nvar(vs); v++) {
        pi_metrics pm = pi_metrics_fast(vs, v, &mem);
        double val;
        if (pi_from_components(pm.total_called, pm.sum_counts_sq, &val))
            sum_pi += val;
        else if (pm.total_called < 2)
            skipped++;
    }
    hapmem_free(&mem);
    int64_t eff = sat_sub_i64(L, (int64_t)skipped);
    if (eff == 0) return NAN;
    return sum_pi / (double)eff;
}

/* ---------- dense_collect_counts (stats.rs:2823-2880) ---------- */
typedef struct {
    uint32_t counts[256];
    uint8_t used[256];
    size_t n_used;
} allele_hist;

static size_t dense_collect_counts(const orc_dense *m, const uint64_t *off, size_t n, size_t v,
                                   allele_hist *hst, double *sum_sq_out) {
    size_t stride = m->n_samples * m->ploidy;
    size_t base = v * stride;
    size_t called = 0;
    for (size_t i = 0; i < n; i++) {
        size_t idx = base + off[i];
        if (m->missing && dense_missing(m->missing, idx)) continue;
        uint8_t a = m->data[idx];
        if (hst->counts[a] == 0) hst->used[hst->n_used++] = a;
        hst->counts[a]++;
        called++;
    }
    if (!m->missing) called = n;
    double ssq = 0.0;
    for (size_t i = 0; i < hst->n_used; i++) {
        double c = (double)hst->counts[hst->used[i]];
        ssq += c * c;
    }
    *sum_sq_out = ssq;
    return called;
}
static void hist_reset(allele_hist *h) { /* stats.rs:2882-2887 */
    for (size_t i = 0; i < h->n_used; i++) h->counts[h->used[i]] = 0;
    h->n_used = 0;
}

/* calculate_pi_dense_biallelic (stats.rs:4434-4532) */
static double pi_dense_biallelic(const orc_dense *m, const uint64_t *off, size_t n, int64_t L) {
    if (n <= 1) return NAN;
    size_t stride = m->n_samples * m->ploidy;
    double sum_pi = 0.0;
    size_t skipped = 0;
    if (m->missing) {
        for (size_t v = 0; v < m->n_variants; v++) {
            size_t called, alt;
            dense_sum_alt_with_missing(m->data, v * stride, off, n, m->missing, &called, &alt);
            double val;
            if (dense_pi_from_counts(called, alt, &val))
                sum_pi += val;
            else
                skipped++;
        }
    } else {
        size_t total = n;
        double nn = (double)total;
        double scale = nn / (nn - 1.0);
        double inv_n_sq = 1.0 / (nn * nn);
        for (size_t v = 0; v < m->n_variants; v++) {
            size_t alt = dense_sum_alt_no_missing(m->data, v * stride, off, n);
            if (alt == 0 || alt == total) continue;
            double alt_f = (double)alt;
            double ref_f = (double)(total - alt);
            double sum_sq = ref_f * ref_f + alt_f * alt_f;
            sum_pi += scale * (1.0 - sum_sq * inv_n_sq);
        }
    }
    int64_t eff = sat_sub_i64(L, (int64_t)skipped);
    if (eff == 0) return NAN;
    return sum_pi / (double)eff;
}

/* calculate_pi_dense (stats.rs:4534-4597) */
static double pi_dense(const orc_dense *m, const uint64_t *off, size_t n, int64_t L) {
    if (n <= 1) return NAN;
    if (L < 0) return 0.0;
    if (L == 0) return INFINITY;
    if (m->max_allele <= 1) return pi_dense_biallelic(m, off, n, L);
    allele_hist hst;
    memset(&hst, 0, sizeof(hst));
    double sum_pi = 0.0;
    size_t skipped = 0;
    for (size_t v = 0; v < m->n_variants; v++) {
        double ssq;
        size_t called = dense_collect_counts(m, off, n, v, &hst, &ssq);
        if (called >= 2) {
            double nn = (double)called;
            double sum_p2 = ssq / (nn * nn);
            sum_pi += nn / (nn - 1.0) * (1.0 - sum_p2);
        } else {
            skipped++;
        }
        hist_reset(&hst);
    }
    int64_t eff = sat_sub_i64(L, (int64_t)skipped);
    if (eff == 0) return NAN;
    return sum_pi / (double)eff;
}

double orc_pi_for_population(const orc_pop *p) { /* stats.rs:4599-4614 */
    if (p->summary) return orc_pi_from_summary(p->summary, p->L, 0, 0.0);
    if (p->dense && p->dense->ploidy == 2) {
        uint64_t *off = (uint64_t *)malloc((p->haps.n ? p->haps.n : 1) * sizeof(uint64_t));
        size_t n = orc_dense_membership(p->dense, &p->haps, off);
        double r = pi_dense(p->dense, off, n, p->L);
        free(off);
        return r;
    }
    return orc_pi_sparse(p->variants, &p->haps, p->L);
}

/* ---------- a8 ---------- */
double orc_harmonic(size_t n) { /* stats.rs:4234-4240 */
    double sum = 0.0;
    for (size_t k = 1; k <= n; k++) sum += 1.0 / (double)k;
    return sum;
}

double orc_watterson_theta(size_t seg, size_t n, int64_t L) { /* stats.rs:4243-4307 */
    if (n <= 1) return seg == 0 ? NAN : INFINITY;
    if (L <= 0) return seg == 0 ? NAN : INFINITY;
    double hv = orc_harmonic(n - 1);
    if (hv > 0.0) return (double)seg / hv / (double)L;
    return seg == 0 ? NAN : INFINITY;
}

/* ---------- a9: calculate_per_site_diversity (stats.rs:4628-4806) ---------- */
size_t orc_per_site_diversity(const orc_variants *vs, const orc_haps *h, int64_t rs, int64_t re,
                              const int64_t *filtered, size_t n_filtered, const int64_t *mask_iv,
                              size_t n_mask, int has_mask, int64_t *pos_out, double *pi_out,
                              double *theta_out) {
    size_t sample_count = nvar(vs) ? vs->n_samples : 0; /* :4650-4653 */
    hapmem mem = hapmem_build(sample_count, h);
    size_t n_out = 0;
    if (region_len(rs, re) <= 0 || h->n < 2) { /* :4656-4681 */
        hapmem_free(&mem);
        return 0;
    }
    for (size_t v = 0; v < nvar(vs); v++) {
        int64_t pos = vs->positions[v];
        if (!region_contains(rs, re, pos)) continue;
        pi_metrics pm = pi_metrics_fast(vs, v, &mem);
        double pi_value, theta_value;
        if (pm.total_called < 2) {
            pi_value = NAN;
            theta_value = NAN;
        } else {
            if (pm.distinct > 1) {
                double denom = orc_harmonic(pm.total_called - 1);
                theta_value = denom > 0.0 ? 1.0 / denom : 0.0;
            } else {
                theta_value = 0.0;
            }
            double val;
            pi_value = pi_from_components(pm.total_called, pm.sum_counts_sq, &val) ? val : 0.0;
        }
        int masked = 0;
        if (has_mask)
            for (size_t i = 0; i < n_mask; i++)
                if (pos >= mask_iv[2 * i] && pos < mask_iv[2 * i + 1]) {
                    masked = 1;
                    break;
                }
        int is_filtered = 0;
        for (size_t i = 0; i < n_filtered; i++)
            if (filtered[i] == pos) {
                is_filtered = 1;
                break;
            }
        if (is_filtered || masked) {
            pi_value = NAN;
            theta_value = NAN;
        }
        pos_out[n_out] = pos + 1; /* :4746 */
        pi_out[n_out] = pi_value;
        theta_out[n_out] = theta_value;
        n_out++;
    }
    hapmem_free(&mem);
    return n_out;
}

/* ---------- a10/a11: Hudson per-site pieces ---------- */
static inline orc_opt some(double v) {
    orc_opt o = {v, 1};
    return o;
}
static inline orc_opt none(void) {
    orc_opt o = {0.0, 0};
    return o;
}

/* the shared (dxy, pi1, pi2) -> (fst, num, den) rule: stats.rs:1736-1757, 2984-3001, 3143-3158 */
static void fst_components(orc_opt dxy, orc_opt p1, orc_opt p2, orc_opt *fst, orc_opt *num,
                           orc_opt *den) {
    if (dxy.some && p1.some && p2.some) {
        double d = dxy.v;
        if (d > FST_EPSILON) {
            double nm = d - 0.5 * (p1.v + p2.v);
            *fst = some(nm / d);
            *num = some(nm);
            *den = some(d);
        } else {
            double pi_avg = 0.5 * (p1.v + p2.v);
            if (fabs(pi_avg) <= FST_EPSILON) {
                *fst = none();
                *num = some(0.0);
                *den = some(0.0);
            } else {
                *fst = none();
                *num = none();
                *den = none();
            }
        }
    } else {
        *fst = none();
        *num = none();
        *den = none();
    }
}

/* AlleleCountSummary via freq_summary_for_pop (stats.rs:2631-2701): counts sorted by allele,
 * sum_counts_sq accumulated incrementally as odd numbers (exact in f64). */
typedef struct {
    size_t total;
    double ssq;
    uint32_t counts[256];
} acs;

static void freq_summary(const orc_variants *vs, size_t v, const hapmem *mem, acs *out) {
    memset(out, 0, sizeof(*out));
    for (size_t s = 0; s < vs->n_samples; s++) {
        const uint8_t *g;
        size_t len = gt_get(vs, v, s, &g);
        if (len == 0) continue;
        if (hm_left(mem, s)) {
            out->ssq += (double)(2 * out->counts[g[0]] + 1);
            out->counts[g[0]]++;
            out->total++;
        }
        if (hm_right(mem, s) && len > 1) {
            out->ssq += (double)(2 * out->counts[g[1]] + 1);
            out->counts[g[1]]++;
            out->total++;
        }
    }
}

static orc_opt dxy_from_counts(const acs *c1, const acs *c2) { /* stats.rs:2907-2935 */
    if (c1->total == 0 || c2->total == 0) return none();
    double dot = 0.0;
    double inv1 = 1.0 / (double)c1->total;
    double inv2 = 1.0 / (double)c2->total;
    for (int a = 0; a < 256; a++) /* merge of sorted entries == ascending allele order */
        if (c1->counts[a] && c2->counts[a])
            dot += ((double)c1->counts[a] * inv1) * ((double)c2->counts[a] * inv2);
    double dxy = 1.0 - dot;
    dxy = fmax(dxy, 0.0);
    dxy = fmin(dxy, 1.0);
    return some(dxy);
}

static orc_hudson_site hudson_site_from_variant(const orc_variants *vs, size_t v, const hapmem *m1,
                                                const hapmem *m2) { /* stats.rs:2969-3014 */
    acs c1, c2;
    freq_summary(vs, v, m1, &c1);
    freq_summary(vs, v, m2, &c2);
    orc_hudson_site s;
    double val;
    s.pi1 = pi_from_components(c1.total, c1.ssq, &val) ? some(val) : none();
    s.pi2 = pi_from_components(c2.total, c2.ssq, &val) ? some(val) : none();
    s.d_xy = dxy_from_counts(&c1, &c2);
    fst_components(s.d_xy, s.pi1, s.pi2, &s.fst, &s.num, &s.den);
    s.position = vs->positions[v] + 1;
    s.n1 = c1.total;
    s.n2 = c2.total;
    return s;
}

static inline orc_opt dense_dxy_biallelic(size_t n1, size_t alt1, size_t n2, size_t alt2) {
    /* stats.rs:1712-1733 */
    if (n1 == 0 || n2 == 0) return none();
    double n1_f = (double)n1, n2_f = (double)n2;
    double alt1_f = (double)alt1 / n1_f;
    double alt2_f = (double)alt2 / n2_f;
    double ref1 = 1.0 - alt1_f, ref2 = 1.0 - alt2_f;
    double dot = ref1 * ref2 + alt1_f * alt2_f;
    if (dot < 0.0) dot = 0.0;
    double dxy = 1.0 - dot;
    if (dxy < 0.0)
        dxy = 0.0;
    else if (dxy > 1.0)
        dxy = 1.0;
    return some(dxy);
}

/* dot product over "used" lists exactly as stats.rs:2557-2591 / 3106-3140 */
static double hist_dot(const allele_hist *h1, size_t n1, const allele_hist *h2, size_t n2) {
    double inv1 = 1.0 / (double)n1, inv2 = 1.0 / (double)n2;
    double dot = 0.0;
    if (h1->n_used <= h2->n_used) {
        for (size_t i = 0; i < h1->n_used; i++) {
            uint8_t a = h1->used[i];
            uint32_t c1 = h1->counts[a], c2 = h2->counts[a];
            if (c1 == 0) continue;
            if (c2 != 0) dot += ((double)c1 * inv1) * ((double)c2 * inv2);
        }
    } else {
        for (size_t i = 0; i < h2->n_used; i++) {
            uint8_t a = h2->used[i];
            uint32_t c2 = h2->counts[a], c1 = h1->counts[a];
            if (c2 == 0) continue;
            if (c1 != 0) dot += ((double)c1 * inv1) * ((double)c2 * inv2);
        }
    }
    return dot;
}

/* dense_hudson_sites (stats.rs:3060-3278) */
static void dense_hudson_sites(const orc_dense *m, const orc_variants *vs, const uint64_t *off1,
                               size_t n1c, const uint64_t *off2, size_t n2c, orc_hudson_site *out) {
    size_t stride = m->n_samples * m->ploidy;
    size_t V = nvar(vs); /* iterates variants.iter().enumerate() */
    if (m->max_allele <= 1) {
        if (m->missing) { /* :3192-3217 */
            for (size_t v = 0; v < V; v++) {
                size_t n1, a1, n2, a2;
                dense_sum_alt_with_missing(m->data, v * stride, off1, n1c, m->missing, &n1, &a1);
                dense_sum_alt_with_missing(m->data, v * stride, off2, n2c, m->missing, &n2, &a2);
                orc_hudson_site s;
                double val;
                s.pi1 = dense_pi_from_counts(n1, a1, &val) ? some(val) : none();
                s.pi2 = dense_pi_from_counts(n2, a2, &val) ? some(val) : none();
                s.d_xy = dense_dxy_biallelic(n1, a1, n2, a2);
                fst_components(s.d_xy, s.pi1, s.pi2, &s.fst, &s.num, &s.den);
                s.position = vs->positions[v] + 1;
                s.n1 = n1;
                s.n2 = n2;
                out[v] = s;
            }
        } else { /* :3218-3274 */
            size_t n1t = n1c, n2t = n2c;
            double n1f = (double)n1t, n2f = (double)n2t;
            int has1 = n1t >= 2, has2 = n2t >= 2;
            double sc1 = has1 ? n1f / (n1f - 1.0) : 0.0, in1 = has1 ? 1.0 / (n1f * n1f) : 0.0;
            double sc2 = has2 ? n2f / (n2f - 1.0) : 0.0, in2 = has2 ? 1.0 / (n2f * n2f) : 0.0;
            for (size_t v = 0; v < V; v++) {
                size_t a1 = dense_sum_alt_no_missing(m->data, v * stride, off1, n1c);
                size_t a2 = dense_sum_alt_no_missing(m->data, v * stride, off2, n2c);
                orc_hudson_site s;
                if (has1) {
                    if (a1 == 0 || a1 == n1t)
                        s.pi1 = some(0.0);
                    else {
                        double af = (double)a1, rf = (double)(n1t - a1);
                        s.pi1 = some(sc1 * (1.0 - (rf * rf + af * af) * in1));
                    }
                } else
                    s.pi1 = none();
                if (has2) {
                    if (a2 == 0 || a2 == n2t)
                        s.pi2 = some(0.0);
                    else {
                        double af = (double)a2, rf = (double)(n2t - a2);
                        s.pi2 = some(sc2 * (1.0 - (rf * rf + af * af) * in2));
                    }
                } else
                    s.pi2 = none();
                s.d_xy = dense_dxy_biallelic(n1t, a1, n2t, a2);
                fst_components(s.d_xy, s.pi1, s.pi2, &s.fst, &s.num, &s.den);
                s.position = vs->positions[v] + 1;
                s.n1 = n1t;
                s.n2 = n2t;
                out[v] = s;
            }
        }
        return;
    }
    /* general: stats.rs:3072-3177 */
    allele_hist h1, h2;
    memset(&h1, 0, sizeof(h1));
    memset(&h2, 0, sizeof(h2));
    for (size_t v = 0; v < V; v++) {
        double ssq1, ssq2;
        size_t n1 = dense_collect_counts(m, off1, n1c, v, &h1, &ssq1);
        size_t n2 = dense_collect_counts(m, off2, n2c, v, &h2, &ssq2);
        orc_hudson_site s;
        if (n1 >= 2) {
            double n = (double)n1;
            s.pi1 = some(n / (n - 1.0) * (1.0 - ssq1 / (n * n)));
        } else
            s.pi1 = none();
        if (n2 >= 2) {
            double n = (double)n2;
            s.pi2 = some(n / (n - 1.0) * (1.0 - ssq2 / (n * n)));
        } else
            s.pi2 = none();
        if (n1 == 0 || n2 == 0)
            s.d_xy = none();
        else {
            double dot = hist_dot(&h1, n1, &h2, n2);
            s.d_xy = some(fmin(fmax(1.0 - dot, 0.0), 1.0));
        }
        fst_components(s.d_xy, s.pi1, s.pi2, &s.fst, &s.num, &s.den);
        s.position = vs->positions[v] + 1;
        s.n1 = n1;
        s.n2 = n2;
        out[v] = s;
        hist_reset(&h1);
        hist_reset(&h2);
    }
}

/* calculate_dxy_dense (stats.rs:2526-2611) */
static orc_opt dxy_dense(const orc_dense *m, const uint64_t *off1, size_t n1c, const uint64_t *off2,
                         size_t n2c, int64_t L) {
    if (n1c == 0 || n2c == 0) return none();
    if (L <= 0) return none();
    allele_hist h1, h2;
    memset(&h1, 0, sizeof(h1));
    memset(&h2, 0, sizeof(h2));
    double sum_dxy = 0.0;
    int64_t skipped = 0;
    for (size_t v = 0; v < m->n_variants; v++) {
        double t1, t2;
        size_t n1 = dense_collect_counts(m, off1, n1c, v, &h1, &t1);
        size_t n2 = dense_collect_counts(m, off2, n2c, v, &h2, &t2);
        if (n1 == 0 || n2 == 0) {
            skipped++;
        } else {
            double dot = hist_dot(&h1, n1, &h2, n2);
            sum_dxy += fmin(fmax(1.0 - dot, 0.0), 1.0);
        }
        hist_reset(&h1);
        hist_reset(&h2);
    }
    int64_t eff = sat_sub_i64(L, skipped);
    if (eff > 0) return some(sum_dxy / (double)eff);
    return none();
}

/* ---------- a12: aggregate_hudson_components_from_summaries (stats.rs:1554-1623) ---------- */
typedef struct {
    double num, den, pi1, pi2, dxy_all;
    size_t uncallable;
} hudson_totals;

static hudson_totals aggregate_from_summaries(const orc_summary *p1, const orc_summary *p2) {
    hudson_totals t = {0, 0, 0, 0, 0, 0};
    size_t len = p1->len < p2->len ? p1->len : p2->len;
    for (size_t i = 0; i < len; i++) {
        size_t n1 = p1->called[i], n2 = p2->called[i];
        if (n1 == 0 || n2 == 0) {
            t.uncallable++;
            continue;
        }
        size_t a1 = p1->alt[i], a2 = p2->alt[i];
        size_t r1 = n1 - a1, r2 = n2 - a2;
        double denom_pairs = (double)(n1 * n2);
        if (denom_pairs == 0.0) continue;
        double dxy = (double)(a1 * r2 + r1 * a2) / denom_pairs;
        if (dxy < 0.0)
            dxy = 0.0;
        else if (dxy > 1.0)
            dxy = 1.0;
        t.dxy_all += dxy;
        if (n1 < 2 || n2 < 2) continue;
        double d1 = (double)(n1 * (n1 - 1)), d2 = (double)(n2 * (n2 - 1));
        double pi1 = d1 > 0.0 ? 2.0 * (double)a1 * (double)r1 / d1 : 0.0;
        double pi2 = d2 > 0.0 ? 2.0 * (double)a2 * (double)r2 / d2 : 0.0;
        t.pi1 += pi1;
        t.pi2 += pi2;
        if (dxy > FST_EPSILON) {
            t.num += dxy - 0.5 * (pi1 + pi2);
            t.den += dxy;
        }
    }
    return t;
}

static void component_sums(const orc_hudson_site *s, size_t n, double *num, double *den) {
    /* stats.rs:1625-1635 */
    double a = 0.0, b = 0.0;
    for (size_t i = 0; i < n; i++)
        if (s[i].num.some && s[i].den.some) {
            a += s[i].num.v;
            b += s[i].den.v;
        }
    *num = a;
    *den = b;
}

orc_opt orc_aggregate_hudson_from_sites(const orc_hudson_site *sites, size_t n) {
    double a, b;
    component_sums(sites, n, &a, &b);
    return b > FST_EPSILON ? some(a / b) : none();
}

static int variants_compatible(const orc_variants *a, const orc_variants *b) { /* :3399-3401 */
    if (nvar(a) != nvar(b)) return 0;
    for (size_t i = 0; i < nvar(a); i++)
        if (a->positions[i] != b->positions[i]) return 0;
    return 1;
}

size_t orc_hudson_per_site(const orc_pop *p1, const orc_pop *p2, int64_t rs, int64_t re,
                           orc_hudson_site *out) { /* stats.rs:3021-3058 */
    if (!variants_compatible(p1->variants, p2->variants)) return 0;
    hapmem m1 = hapmem_build(p1->n_sample_names, &p1->haps);
    hapmem m2 = hapmem_build(p2->n_sample_names, &p2->haps);
    size_t n = 0;
    for (size_t v = 0; v < nvar(p1->variants); v++)
        if (region_contains(rs, re, p1->variants->positions[v]))
            out[n++] = hudson_site_from_variant(p1->variants, v, &m1, &m2);
    hapmem_free(&m1);
    hapmem_free(&m2);
    return n;
}

int orc_dxy_hudson(const orc_pop *p1, const orc_pop *p2, orc_opt *out) { /* stats.rs:2403-2524 */
    *out = none();
    if (p1->L <= 0) return ORC_ERR_INVALID_REGION;
    if (p1->L != p2->L) return ORC_ERR_PARSE;
    if (!variants_compatible(p1->variants, p2->variants)) return ORC_ERR_PARSE;
    if (p1->haps.n == 0 || p2->haps.n == 0) return ORC_OK;
    if (p1->summary && p2->summary) { /* dxy_from_summaries :1637-1662 */
        hudson_totals t = aggregate_from_summaries(p1->summary, p2->summary);
        int64_t eff = sat_sub_i64(p1->L, (int64_t)t.uncallable);
        if (eff > 0) *out = some(t.dxy_all / (double)eff);
        return ORC_OK;
    }
    if (p1->dense && p2->dense && p1->dense == p2->dense && p1->dense->ploidy == 2) {
        uint64_t *o1 = (uint64_t *)malloc((p1->haps.n ? p1->haps.n : 1) * sizeof(uint64_t));
        uint64_t *o2 = (uint64_t *)malloc((p2->haps.n ? p2->haps.n : 1) * sizeof(uint64_t));
        size_t n1 = orc_dense_membership(p1->dense, &p1->haps, o1);
        size_t n2 = orc_dense_membership(p2->dense, &p2->haps, o2);
        *out = dxy_dense(p1->dense, o1, n1, o2, n2, p1->L);
        free(o1);
        free(o2);
        return ORC_OK;
    }
    hapmem m1 = hapmem_build(p1->n_sample_names, &p1->haps);
    hapmem m2 = hapmem_build(p2->n_sample_names, &p2->haps);
    double sum_dxy = 0.0;
    int64_t skipped = 0;
    for (size_t v = 0; v < nvar(p1->variants); v++) {
        acs c1, c2;
        freq_summary(p1->variants, v, &m1, &c1);
        freq_summary(p1->variants, v, &m2, &c2);
        orc_opt d = dxy_from_counts(&c1, &c2);
        if (d.some)
            sum_dxy += d.v;
        else
            skipped++;
    }
    hapmem_free(&m1);
    hapmem_free(&m2);
    int64_t eff = sat_sub_i64(p1->L, skipped);
    if (eff > 0) *out = some(sum_dxy / (double)eff);
    return ORC_OK;
}

/* ---------- a13: calculate_hudson_fst_for_pair_core (stats.rs:3435-3599) ---------- */
int orc_hudson_pair(const orc_pop *p1, const orc_pop *p2, int has_region, int64_t rs, int64_t re,
                    orc_hudson_outcome *out, orc_hudson_site *sites_out, size_t *n_sites_out) {
    memset(out, 0, sizeof(*out));
    if (n_sites_out) *n_sites_out = 0;
    if (p1->L <= 0) return ORC_ERR_INVALID_REGION;
    if (p1->L != p2->L) return ORC_ERR_PARSE;
    if (!variants_compatible(p1->variants, p2->variants)) return ORC_ERR_PARSE;

    int have_summaries = p1->summary && p2->summary;
    const orc_dense *shared =
        (p1->dense && p2->dense && p1->dense == p2->dense && p1->dense->ploidy == 2) ? p1->dense : NULL;
    size_t V = nvar(p1->variants);
    orc_hudson_site *sites = sites_out;
    int own_sites = 0;
    if (!sites) {
        sites = (orc_hudson_site *)malloc((V ? V : 1) * sizeof(orc_hudson_site));
        own_sites = 1;
    }
    size_t n_sites = 0;
    double num_sum = 0.0, den_sum = 0.0;
    hudson_totals totals = {0, 0, 0, 0, 0, 0};
    int used_summaries = 0;
    if (has_region) {
        n_sites = orc_hudson_per_site(p1, p2, rs, re, sites);
        component_sums(sites, n_sites, &num_sum, &den_sum);
    } else if (have_summaries) {
        totals = aggregate_from_summaries(p1->summary, p2->summary);
        used_summaries = 1;
        num_sum = totals.num;
        den_sum = totals.den;
    } else if (shared) {
        if (V != 0) {
            uint64_t *o1 = (uint64_t *)malloc((p1->haps.n ? p1->haps.n : 1) * sizeof(uint64_t));
            uint64_t *o2 = (uint64_t *)malloc((p2->haps.n ? p2->haps.n : 1) * sizeof(uint64_t));
            size_t n1 = orc_dense_membership(shared, &p1->haps, o1);
            size_t n2 = orc_dense_membership(shared, &p2->haps, o2);
            dense_hudson_sites(shared, p1->variants, o1, n1, o2, n2, sites);
            n_sites = V;
            component_sums(sites, n_sites, &num_sum, &den_sum);
            free(o1);
            free(o2);
        }
    } else if (V != 0) {
        hapmem m1 = hapmem_build(p1->n_sample_names, &p1->haps);
        hapmem m2 = hapmem_build(p2->n_sample_names, &p2->haps);
        for (size_t v = 0; v < V; v++) sites[v] = hudson_site_from_variant(p1->variants, v, &m1, &m2);
        n_sites = V;
        component_sums(sites, n_sites, &num_sum, &den_sum);
        hapmem_free(&m1);
        hapmem_free(&m2);
    }
    orc_opt regional = den_sum > FST_EPSILON ? some(num_sum / den_sum) : none();

    double pi1_raw, pi2_raw;
    orc_opt dxy = none();
    if (used_summaries) {
        pi1_raw = orc_pi_from_summary(p1->summary, p1->L, 1, totals.pi1);
        pi2_raw = orc_pi_from_summary(p2->summary, p2->L, 1, totals.pi2);
        if (p1->haps.n != 0 && p2->haps.n != 0) {
            int64_t eff = sat_sub_i64(p1->L, (int64_t)totals.uncallable);
            if (eff > 0) dxy = some(totals.dxy_all / (double)eff);
        }
    } else {
        pi1_raw = orc_pi_for_population(p1);
        pi2_raw = orc_pi_for_population(p2);
        int rc = orc_dxy_hudson(p1, p2, &dxy);
        if (rc != ORC_OK) {
            if (own_sites) free(sites);
            return rc;
        }
    }
    out->pi1 = isfinite(pi1_raw) ? some(pi1_raw) : none();
    out->pi2 = isfinite(pi2_raw) ? some(pi2_raw) : none();
    out->d_xy = dxy;
    out->fst = regional;
    out->pi_xy_avg = (out->pi1.some && out->pi2.some) ? some(0.5 * (out->pi1.v + out->pi2.v)) : none();
    if (n_sites_out) *n_sites_out = n_sites;
    if (own_sites) free(sites);
    return ORC_OK;
}

/* ---------- a16: calculate_variance_components (stats.rs:2034-2127) ---------- */
void orc_variance_components(const uint64_t *n, const double *p, size_t rr, double global_p,
                             double *a_out, double *b_out) {
    double r = (double)rr;
    if (r < 2.0) {
        *a_out = 0.0;
        *b_out = 0.0;
        return;
    }
    size_t total = 0;
    for (size_t i = 0; i < rr; i++) total += n[i];
    double n_bar = (double)total / r;
    if ((n_bar - 1.0) < 1e-9) {
        *a_out = 0.0;
        *b_out = 0.0;
        return;
    }
    double sum_sq_diff_n = 0.0;
    for (size_t i = 0; i < rr; i++) {
        double diff = (double)n[i] - n_bar;
        sum_sq_diff_n += diff * diff;
    }
    double c_squared = (r > 0.0 && n_bar > 0.0) ? sum_sq_diff_n / (r * n_bar * n_bar) : 0.0;
    double numerator_s_squared = 0.0;
    for (size_t i = 0; i < rr; i++) {
        double diff_p = p[i] - global_p;
        numerator_s_squared += (double)n[i] * diff_p * diff_p;
    }
    double s_squared =
        ((r - 1.0) > 1e-9 && n_bar > 1e-9) ? numerator_s_squared / ((r - 1.0) * n_bar) : 0.0;
    double x_wc = global_p * (1.0 - global_p) - ((r - 1.0) / r) * s_squared;
    double a_numerator_term = s_squared - (x_wc / (n_bar - 1.0));
    double a_denominator_factor = 1.0 - (c_squared / (r - 1.0));
    *a_out = a_numerator_term / a_denominator_factor;
    *b_out = (n_bar / (n_bar - 1.0)) * x_wc;
}

/* ---------- a17: fst_estimate_from_components (stats.rs:1781-1812) ---------- */
static orc_fst_estimate classify(double a, double b, uint64_t sites) {
    /* shared threshold ladder: stats.rs:1785-1811, 2237-2270, 2297-2328 */
    orc_fst_estimate e;
    double den = a + b;
    e.sum_a = a;
    e.sum_b = b;
    e.sites = sites;
    e.value = NAN;
    if (den > FST_EPSILON) {
        e.state = 0;
        e.value = a / den;
    } else if (den < -FST_EPSILON) {
        e.state = 1;
    } else if (fabs(a) > FST_EPSILON) {
        e.state = 0;
        e.value = a / den;
    } else {
        e.state = 2;
    }
    return e;
}
orc_fst_estimate orc_fst_estimate_from_components(double a, double b) { return classify(a, b, 1); }

/* ---------- a15 + a18 ---------- */
void orc_wc_fst(const orc_variants *vs, const uint16_t *left, const uint16_t *right, size_t G,
                int64_t rs, int64_t re, size_t *n_sites_out, int64_t *site_pos, int *site_state,
                double *site_a, double *site_b, uint64_t *site_pop_sizes, uint8_t *site_has_maps,
                double *pair_a, double *pair_b, int *pair_state, orc_fst_estimate *overall,
                orc_fst_estimate *pairs, uint8_t *pair_present) {
    size_t n_pairs = G * (G > 0 ? G - 1 : 0) / 2;
    size_t *total_counts = (size_t *)calloc(G ? G : 1, sizeof(size_t));
    size_t *alt_counts = (size_t *)calloc(G ? G : 1, sizeof(size_t));
    uint64_t *st_n = (uint64_t *)calloc(G ? G : 1, sizeof(uint64_t));
    double *st_p = (double *)calloc(G ? G : 1, sizeof(double));
    double *pw_a = (double *)calloc(n_pairs ? n_pairs : 1, sizeof(double));
    double *pw_b = (double *)calloc(n_pairs ? n_pairs : 1, sizeof(double));
    uint8_t *pw_has = (uint8_t *)calloc(n_pairs ? n_pairs : 1, 1);
    /* region accumulators (calculate_overall_fst_wc, stats.rs:2145-2374), site order */
    double sum_a_total = 0.0, sum_b_total = 0.0;
    size_t n_overall = 0, n_insufficient = 0, n_sites = 0, n_sites_with_maps = 0;
    double *reg_pa = (double *)calloc(n_pairs ? n_pairs : 1, sizeof(double));
    double *reg_pb = (double *)calloc(n_pairs ? n_pairs : 1, sizeof(double));
    size_t *reg_pn = (size_t *)calloc(n_pairs ? n_pairs : 1, sizeof(size_t));

    for (size_t v = 0; v < nvar(vs); v++) {
        int64_t pos = vs->positions[v];
        if (!region_contains(rs, re, pos)) continue;
        /* ---- calculate_fst_wc_at_site_with_membership (stats.rs:1814-2032) ---- */
        uint8_t present[256];
        memset(present, 0, sizeof(present));
        for (size_t s = 0; s < vs->n_samples; s++) { /* :1826-1833 all samples, all alleles */
            const uint8_t *g;
            size_t len = gt_get(vs, v, s, &g);
            for (size_t k = 0; k < len; k++) present[g[k]] = 1;
        }
        double sum_site_a = 0.0, sum_site_b = 0.0;
        memset(pw_a, 0, n_pairs * sizeof(double));
        memset(pw_b, 0, n_pairs * sizeof(double));
        memset(pw_has, 0, n_pairs);
        int pop_sizes_populated = 0;
        uint64_t *sizes_row = site_pop_sizes ? site_pop_sizes + n_sites * G : NULL;
        if (sizes_row) memset(sizes_row, 0, G * sizeof(uint64_t));
        for (int target = 0; target < 256; target++) { /* ascending == sorted unique alleles */
            if (!present[target]) continue;
            memset(total_counts, 0, G * sizeof(size_t));
            memset(alt_counts, 0, G * sizeof(size_t));
            for (size_t s = 0; s < vs->n_samples; s++) { /* :1864-1900 */
                const uint8_t *g;
                size_t len = gt_get(vs, v, s, &g);
                if (len == 0) continue;
                uint16_t grp = left[s];
                if (grp != INVALID_GROUP) {
                    total_counts[grp]++;
                    if (g[0] == target) alt_counts[grp]++;
                }
                if (len > 1) {
                    grp = right[s];
                    if (grp != INVALID_GROUP) {
                        total_counts[grp]++;
                        if (g[1] == target) alt_counts[grp]++;
                    }
                }
            }
            size_t total_called = 0, total_target = 0, valid_groups = 0;
            for (size_t gi = 0; gi < G; gi++) { /* :1907-1922 */
                size_t total = total_counts[gi];
                if (total == 0) continue;
                st_n[valid_groups] = total;
                st_p[valid_groups] = (double)alt_counts[gi] / (double)total;
                valid_groups++;
                total_called += total;
                total_target += alt_counts[gi];
                if (!pop_sizes_populated && sizes_row) sizes_row[gi] = total;
            }
            pop_sizes_populated = 1;
            if (valid_groups < 2) continue; /* :1925-1930 */
            double global_freq = total_called > 0 ? (double)total_target / (double)total_called : 0.0;
            double ca, cb;
            orc_variance_components(st_n, st_p, valid_groups, global_freq, &ca, &cb);
            sum_site_a += ca;
            sum_site_b += cb;
            size_t pi = 0;
            for (size_t i = 0; i < G; i++)
                for (size_t j = i + 1; j < G; j++, pi++) { /* :1943-1982 */
                    size_t ta = total_counts[i], tb = total_counts[j];
                    if (ta == 0 || tb == 0) continue;
                    uint64_t pn[2] = {ta, tb};
                    double pp[2] = {(double)alt_counts[i] / (double)ta, (double)alt_counts[j] / (double)tb};
                    size_t pt = ta + tb;
                    double pg = pt > 0 ? (double)(alt_counts[i] + alt_counts[j]) / (double)pt : 0.0;
                    double xa, xb;
                    orc_variance_components(pn, pp, 2, pg, &xa, &xb);
                    pw_a[pi] += xa;
                    pw_b[pi] += xb;
                    pw_has[pi] = 1;
                }
        }
        orc_fst_estimate est;
        int has_maps;
        if (!pop_sizes_populated) { /* :1987-2001 */
            est.state = 3;
            est.value = NAN;
            est.sum_a = 0.0;
            est.sum_b = 0.0;
            est.sites = 1;
            sum_site_a = 0.0;
            sum_site_b = 0.0;
            has_maps = 0;
        } else {
            est = orc_fst_estimate_from_components(sum_site_a, sum_site_b);
            has_maps = 1;
        }
        site_pos[n_sites] = pos + 1;
        site_state[n_sites] = est.state;
        site_a[n_sites] = sum_site_a;
        site_b[n_sites] = sum_site_b;
        if (site_has_maps) site_has_maps[n_sites] = (uint8_t)has_maps;
        /* ---- region aggregation (stats.rs:2169-2204) ---- */
        if (est.state == 3)
            n_insufficient++;
        else {
            sum_a_total += sum_site_a;
            sum_b_total += sum_site_b;
            n_overall++;
        }
        if (has_maps) n_sites_with_maps++;
        for (size_t p = 0; p < n_pairs; p++) {
            int st = 3;
            double xa = 0.0, xb = 0.0;
            if (has_maps && pw_has[p]) { /* :2007-2023 */
                xa = pw_a[p];
                xb = pw_b[p];
                st = orc_fst_estimate_from_components(xa, xb).state;
                reg_pa[p] += xa;
                reg_pb[p] += xb;
                reg_pn[p]++;
            }
            if (pair_a) {
                pair_a[n_sites * n_pairs + p] = xa;
                pair_b[n_sites * n_pairs + p] = xb;
                pair_state[n_sites * n_pairs + p] = has_maps ? st : -1; /* -1: key absent */
            }
        }
        n_sites++;
    }
    *n_sites_out = n_sites;
    /* overall (stats.rs:2152-2271) */
    if (n_sites == 0 || n_overall == 0) {
        overall->state = 3;
        overall->value = NAN;
        overall->sum_a = 0.0;
        overall->sum_b = 0.0;
        overall->sites = n_sites;
    } else {
        *overall = classify(sum_a_total, sum_b_total, n_overall);
    }
    (void)n_insufficient;
    for (size_t p = 0; p < n_pairs; p++) { /* stats.rs:2285-2367 */
        if (n_sites_with_maps == 0) {
            pair_present[p] = 0; /* key never observed -> absent from the result maps */
            memset(&pairs[p], 0, sizeof(pairs[p]));
            pairs[p].state = 3;
            pairs[p].value = NAN;
            continue;
        }
        pair_present[p] = 1;
        if (reg_pn[p] > 0) {
            pairs[p] = classify(reg_pa[p], reg_pb[p], reg_pn[p]);
        } else {
            pairs[p].state = 3;
            pairs[p].value = NAN;
            pairs[p].sum_a = 0.0;
            pairs[p].sum_b = 0.0;
            pairs[p].sites = n_sites_with_maps; /* :2342-2348 */
        }
    }
    free(total_counts);
    free(alt_counts);
    free(st_n);
    free(st_p);
    free(pw_a);
    free(pw_b);
    free(pw_has);
    free(reg_pa);
    free(reg_pb);
    free(reg_pn);
}

/* ---------- a19: calculate_adjusted_sequence_length (stats.rs:3644-3736) ---------- */
typedef struct {
    int64_t s, e;
} iv;

static void hb_from_1based_inclusive(int64_t s, int64_t e, uint64_t *hs, uint64_t *he) {
    /* process.rs:193-206 */
    if (s < 1) s = 1;
    if (e < s) e = s;
    *hs = (uint64_t)(s - 1);
    *he = (uint64_t)e;
}

int64_t orc_adjusted_sequence_length(int64_t region_start, int64_t region_end, const int64_t *allow,
                                     size_t n_allow, int has_allow, const int64_t *mask,
                                     size_t n_mask, int has_mask) {
    uint64_t rs, re;
    hb_from_1based_inclusive(region_start, region_end, &rs, &re);
    size_t cap = (has_allow ? n_allow : 1) + 1;
    iv *allowed = (iv *)malloc(cap * sizeof(iv));
    size_t n_allowed = 0;
    if (has_allow) {
        for (size_t i = 0; i < n_allow; i++) {
            /* from_0based_half_open casts i64 -> usize (process.rs:346-351) */
            uint64_t as = (uint64_t)allow[2 * i], ae = (uint64_t)allow[2 * i + 1];
            uint64_t s = rs > as ? rs : as, e = re < ae ? re : ae;
            if (s < e) { /* to_1based_inclusive_tuple (process.rs:294-299) */
                allowed[n_allowed].s = (int64_t)s + 1;
                allowed[n_allowed].e = (int64_t)e;
                n_allowed++;
            }
        }
    } else {
        allowed[0].s = region_start;
        allowed[0].e = region_end;
        n_allowed = 1;
    }
    int64_t total = 0;
    /* subtract_regions (stats.rs:3739-3775) */
    for (size_t ai = 0; ai < n_allowed; ai++) {
        size_t pcap = 2 * (n_mask + 1) + 2, np = 1;
        iv *parts = (iv *)malloc(pcap * sizeof(iv));
        iv *next = (iv *)malloc(pcap * sizeof(iv));
        parts[0] = allowed[ai];
        if (has_mask) {
            for (size_t mi = 0; mi < n_mask && np > 0; mi++) {
                int64_t m_start = (int64_t)((uint64_t)mask[2 * mi]) + 1; /* converted mask */
                int64_t m_end = (int64_t)((uint64_t)mask[2 * mi + 1]);
                size_t nn = 0;
                for (size_t k = 0; k < np; k++) {
                    int64_t s = parts[k].s, e = parts[k].e;
                    if (m_end < s || m_start > e) {
                        next[nn++] = parts[k];
                        continue;
                    }
                    if (m_start > s) {
                        int64_t left_end = m_start - 1;
                        if (left_end >= s) next[nn++] = (iv){s, left_end};
                    }
                    if (m_end < e) {
                        int64_t right_start = m_end + 1;
                        if (right_start <= e) next[nn++] = (iv){right_start, e};
                    }
                }
                iv *t = parts;
                parts = next;
                next = t;
                np = nn;
            }
        }
        for (size_t k = 0; k < np; k++) { /* :3706-3713 */
            uint64_t hs, he;
            hb_from_1based_inclusive(parts[k].s, parts[k].e, &hs, &he);
            total += he > hs ? (int64_t)(he - hs) : 0;
        }
        free(parts);
        free(next);
    }
    free(allowed);
    return total;
}
