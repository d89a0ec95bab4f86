/*
 * ferromic_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the per-site population-genetics estimators of the
 * reference crate SauersML/ferromic, file src/stats.rs (+ the data model in
 * src/process.rs and the numpy ingestion in src/lib.rs).  Every function cites
 * the reference file:line it follows.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may link or call this
 * library; the product path (ferromic_b200/) never does.
 *
 * Pinning: the reference is Rust (nightly, no lockfile, no network) and cannot
 * be compiled in the authoring container, so there is no oracle/_ref build.
 * The restatement is pinned against every golden vector the reference's own
 * tests hold for this path (tests/test_oracle_golden.py transcribes them with
 * file:line).  Weir & Cockerham values are NOT pinned by any reference test
 * ("parity unpinned" for a15-a18); see DESIGN.md.
 *
 * Floating point: expressions keep the reference's operation order; the
 * library is compiled with -ffp-contract=off so no FMA contraction occurs.
 * Where the reference reduces with Rayon (unordered), the oracle sums in site
 * order.
 */
#ifndef FERROMIC_ORACLE_H
#define FERROMIC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MISSING 0xFFu /* process.rs:438 CompressedGenotypes::MISSING */

/* Sparse variants: process.rs:430-536 (Variant + CompressedGenotypes).
 * gt[(v*n_samples + s)*stride + k]; 0xFF in slot 0 => sample genotype None;
 * a later 0xFF terminates the genotype (shorter ploidy). */
typedef struct {
    size_t n_variants, n_samples, stride;
    const int64_t *positions; /* 0-based */
    const uint8_t *gt;
} orc_variants;

/* Dense matrix: stats.rs:249-331. missing may be NULL. */
typedef struct {
    const uint8_t *data;
    const uint64_t *missing;
    size_t n_variants, n_samples, ploidy;
    uint8_t max_allele;
} orc_dense;

/* Haplotype list (sample index, side 0=Left 1=Right): stats.rs:237 */
typedef struct {
    const uint64_t *sample;
    const uint8_t *side;
    size_t n;
} orc_haps;

/* DensePopulationSummary: stats.rs:1310-1344 */
typedef struct {
    uint32_t *alt;
    uint32_t *called;
    size_t len;
    size_t capacity;
    size_t seg;
    double pi_sum;
} orc_summary;

/* PopulationContext: stats.rs:230-247 */
typedef struct {
    orc_haps haps;
    const orc_variants *variants; /* may be NULL => empty slice */
    size_t n_sample_names;
    int64_t L;
    const orc_dense *dense;     /* may be NULL */
    const orc_summary *summary; /* may be NULL */
} orc_pop;

typedef struct {
    double v;
    int some;
} orc_opt;

/* SiteFstHudson: stats.rs:536-555 */
typedef struct {
    int64_t position; /* 1-based */
    orc_opt fst, d_xy, pi1, pi2, num, den;
    uint64_t n1, n2;
} orc_hudson_site;

/* HudsonFSTOutcome: stats.rs:515-532 */
typedef struct {
    orc_opt fst, d_xy, pi1, pi2, pi_xy_avg;
} orc_hudson_outcome;

/* FstEstimate: stats.rs:36-126. state: 0 Calculable, 1 ComponentsYieldIndeterminateRatio,
 * 2 NoInterPopulationVariance, 3 InsufficientDataForEstimation. */
typedef struct {
    int state;
    double value; /* only meaningful for state 0 */
    double sum_a, sum_b;
    uint64_t sites;
} orc_fst_estimate;

/* error codes mirroring VcfError (process.rs:631-640) */
#define ORC_OK 0
#define ORC_ERR_INVALID_REGION 1
#define ORC_ERR_PARSE 2

/* ---- a1: DenseGenotypeMatrix::from_variants (stats.rs:339-500) ---- */
/* Returns 0 and fills outputs (caller frees data/missing with orc_free) or 1 for None. */
int orc_dense_from_variants(const orc_variants *vs, size_t sample_count, uint8_t **data,
                            uint64_t **missing, size_t *ploidy, uint8_t *max_allele);
void orc_free(void *p);

/* ---- a2: DenseMembership::build (stats.rs:1251-1284) ---- */
/* offsets_out must hold haps.n entries; returns count. */
size_t orc_dense_membership(const orc_dense *m, const orc_haps *h, uint64_t *offsets_out);

/* ---- a4: build_dense_population_summary (stats.rs:1367-1470) ---- */
/* alt/called arrays (len n_variants) supplied by caller inside summary. */
void orc_build_summary(const orc_dense *m, const orc_haps *h, orc_summary *out);
/* Threaded variant used as the CPU baseline: same inner loop (stats.rs:1665-1697),
 * variant chunks over nthreads (mirrors rayon par_iter at stats.rs:1415-1460). */
void orc_build_summary_mt(const orc_dense *m, const orc_haps *h, orc_summary *out, int nthreads);

/* ---- a6: segregating sites ---- */
size_t orc_count_segregating_sites(const orc_variants *vs);          /* stats.rs:3808-3829 */
size_t orc_count_segregating_sites_for_population(const orc_pop *p); /* stats.rs:3831-3851 */

/* ---- a7: pi ---- */
double orc_pi_sparse(const orc_variants *vs, const orc_haps *h, int64_t L); /* stats.rs:4317-4432 */
double orc_pi_for_population(const orc_pop *p);                             /* stats.rs:4599-4614 */
double orc_pi_from_summary(const orc_summary *s, int64_t L, int has_pre, double pre); /* :1480-1542 */

/* ---- a8 ---- */
double orc_harmonic(size_t n);                                  /* stats.rs:4234-4240 */
double orc_watterson_theta(size_t seg, size_t n, int64_t L);    /* stats.rs:4243-4307 */

/* ---- a9: calculate_per_site_diversity (stats.rs:4628-4806) ---- */
/* Outputs sized >= n_variants; returns number of sites emitted. */
size_t orc_per_site_diversity(const orc_variants *vs, const orc_haps *h, int64_t region_start,
                              int64_t region_end, const int64_t *filtered, size_t n_filtered,
                              const int64_t *mask_iv, size_t n_mask, int has_mask,
                              int64_t *pos_out, double *pi_out, double *theta_out);

/* ---- a11/a13: Hudson ---- */
/* calculate_hudson_fst_for_pair_core (stats.rs:3435-3599). region may be disabled
 * (has_region=0). sites_out sized >= n_variants (may be NULL to discard). */
int orc_hudson_pair(const orc_pop *p1, const orc_pop *p2, int has_region, int64_t rs, int64_t re,
                    orc_hudson_outcome *out, orc_hudson_site *sites_out, size_t *n_sites);
/* calculate_hudson_fst_per_site (stats.rs:3021-3058): returns count (0 if incompatible). */
size_t orc_hudson_per_site(const orc_pop *p1, const orc_pop *p2, int64_t rs, int64_t re,
                           orc_hudson_site *sites_out);
/* calculate_d_xy_hudson (stats.rs:2403-2524) */
int orc_dxy_hudson(const orc_pop *p1, const orc_pop *p2, orc_opt *out);
/* aggregate_hudson_from_sites (stats.rs:3309-3316) */
orc_opt orc_aggregate_hudson_from_sites(const orc_hudson_site *sites, size_t n);

/* ---- a15-a18: Weir & Cockerham ---- */
/* calculate_fst_wc_haplotype_groups / _csv_populations body (stats.rs:706-761, 849-904)
 * after the label->index mapping (a14, done by the caller): left/right[n_samples] hold a
 * group index or 0xFFFF.  Per-site outputs are sized n_variants (only the first
 * *n_sites_out entries are written, in variant order, for variants inside the region):
 *   site_pos, site_state (overall state), site_a, site_b, site_pop_sizes [n_sites][G]
 *   pair_a/pair_b/pair_state [n_sites][n_pairs] (NULL to skip) ; pair order i<j.
 * site_has_maps[n_sites]: 1 when the per-site pair maps are populated (non-empty).
 * Region outputs: overall, pairs[n_pairs], pair_present (0 when the pair key never appears). */
void orc_wc_fst(const orc_variants *vs, const uint16_t *left, const uint16_t *right, size_t G,
                int64_t rs, int64_t re, size_t *n_sites_out, int64_t *site_pos, int *site_state,
                double *site_a, double *site_b, uint64_t *site_pop_sizes, uint8_t *site_has_maps,
                double *pair_a, double *pair_b, int *pair_state, orc_fst_estimate *overall,
                orc_fst_estimate *pairs, uint8_t *pair_present);
/* calculate_variance_components (stats.rs:2034-2127) on (n_i, p_i) */
void orc_variance_components(const uint64_t *n, const double *p, size_t r, double global_p,
                             double *a, double *b);
/* fst_estimate_from_components (stats.rs:1781-1812) */
orc_fst_estimate orc_fst_estimate_from_components(double a, double b);

/* ---- a19: calculate_adjusted_sequence_length (stats.rs:3644-3736) ---- */
int64_t orc_adjusted_sequence_length(int64_t region_start, int64_t region_end,
                                     const int64_t *allow, size_t n_allow, int has_allow,
                                     const int64_t *mask, size_t n_mask, int has_mask);

#ifdef __cplusplus
}
#endif
#endif
