// fm_wc.cuh -- K4: Weir & Cockerham variance components from per-group allele/called counts.
//
// Mirrors calculate_fst_wc_at_site_with_membership (stats.rs:1814-2032),
// calculate_variance_components (stats.rs:2034-2127) and the region aggregation of
// calculate_overall_fst_wc (stats.rs:2145-2374) for biallelic matrices.
//
// Work decomposition: one warp owns a super-batch of 8192 consecutive sites and walks it in
// site order.  For every site the 32 lanes split the G*(G-1)/2 population pairs (lane l takes
// pairs l, l+32, ...) and add the pair's (a, b) to that pair's accumulator, so every regional
// pair sum is accumulated sequentially in site order -- the same association as the reference's
// `.sum()` over sites (stats.rs:2288-2289); super-batch partials are then combined in order on
// the host.  The overall (all-population) components use half a warp per allele.
#pragma once
#include "fm_device.cuh"

namespace fm {

constexpr uint32_t kWcSitesPerSuper = 8192;  // == kSuperBatches * 32
constexpr int kWcWarpsPerCta = 4;

struct WcParams {
    const uint32_t *const *alt;  // [G + 1] device pointers, each [V]; index G = haplotypes with no group
    const uint32_t *const *cnt;  // [G + 1]
    uint32_t G, n_pairs;
    const uint16_t *pair_i, *pair_j;  // [n_pairs], i < j in label order
    uint32_t v_lo, v_hi;              // sites of the region
    uint32_t s_lo, n_super;           // global super-batch range covering [v_lo, v_hi)
    // per-site outputs (indexed v - v_lo), any may be nullptr
    int32_t *site_state;
    double *site_a, *site_b;
    uint32_t *site_sizes;        // [n_sites][G]
    double *pair_a, *pair_b;     // [n_sites][n_pairs]; NaN when the pair has no data at the site
    // per-super-batch partials
    double *part_overall;        // [n_super][2]          sum a, sum b
    uint32_t *part_counts;       // [n_super][2]          informative sites, sites with maps
    double *part_pair;           // [n_super][n_pairs][2]
    uint32_t *part_pair_n;       // [n_super][n_pairs]    informative sites per pair
};

// calculate_variance_components (stats.rs:2034-2127) for the groups with data, evaluated with a
// uniform loop over all groups; counts come from shared memory (sc: [G+1][2] = alt, called).
__device__ __forceinline__ void fm_wc_overall_allele(const uint32_t *sc, uint32_t G, bool allele_one,
                                                     double &a, double &b, bool &ok) {
    uint32_t r_groups = 0;
    uint64_t total_called = 0, total_target = 0;
    for (uint32_t g = 0; g < G; ++g) {
        const uint32_t n = sc[2 * g + 1];
        if (n == 0) continue;
        const uint32_t t = allele_one ? sc[2 * g] : n - sc[2 * g];
        ++r_groups;
        total_called += n;
        total_target += t;
    }
    a = 0.0;
    b = 0.0;
    ok = r_groups >= 2;  // stats.rs:1925-1930
    if (!ok) return;
    const double r = (double)r_groups;
    const double global_p = total_called > 0 ? (double)total_target / (double)total_called : 0.0;
    const double n_bar = (double)total_called / r;
    if ((n_bar - 1.0) < 1e-9) return;  // (0, 0)
    double sum_sq_diff_n = 0.0, numerator_s_squared = 0.0;
    for (uint32_t g = 0; g < G; ++g) {
        const uint32_t n = sc[2 * g + 1];
        if (n == 0) continue;
        const double diff = (double)n - n_bar;
        sum_sq_diff_n += diff * diff;
    }
    for (uint32_t g = 0; g < G; ++g) {
        const uint32_t n = sc[2 * g + 1];
        if (n == 0) continue;
        const uint32_t t = allele_one ? sc[2 * g] : n - sc[2 * g];
        const double freq = (double)t / (double)n;
        const double diff_p = freq - global_p;
        numerator_s_squared += (double)n * diff_p * diff_p;
    }
    const double c_squared = sum_sq_diff_n / (r * n_bar * n_bar);
    const double s_squared = ((r - 1.0) > 1e-9 && n_bar > 1e-9) ? numerator_s_squared / ((r - 1.0) * n_bar) : 0.0;
    const double x_wc = global_p * (1.0 - global_p) - ((r - 1.0) / r) * s_squared;
    const double a_numerator_term = s_squared - (x_wc / (n_bar - 1.0));
    const double a_denominator_factor = 1.0 - (c_squared / (r - 1.0));
    a = a_numerator_term / a_denominator_factor;
    b = (n_bar / (n_bar - 1.0)) * x_wc;
}

__global__ void __launch_bounds__(kWcWarpsPerCta * 32)
fm_k_wc(const WcParams P) {
    extern __shared__ __align__(16) uint8_t wc_smem[];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t G1 = P.G + 1;
    // per-warp shared memory: counts of one 32-site batch [32][G1][2] u32, pair accumulators
    // [n_pairs][2] f64 and pair counts [n_pairs] u32
    const size_t per_warp = (size_t)32 * G1 * 8 + (size_t)P.n_pairs * 16 + (size_t)P.n_pairs * 4 + 16;
    uint8_t *base = wc_smem + warp * ((per_warp + 15) & ~(size_t)15);
    double *acc = reinterpret_cast<double *>(base);
    uint32_t *cnts = reinterpret_cast<uint32_t *>(base + (size_t)P.n_pairs * 16);
    uint32_t *acc_n = cnts + (size_t)32 * G1 * 2;

    const uint32_t wpc = blockDim.x >> 5;  // the host may run fewer warps when staging is large
    const uint32_t gw = blockIdx.x * wpc + warp;
    const uint32_t GW = gridDim.x * wpc;
    const uint32_t n_sites = P.v_hi - P.v_lo;
    for (uint32_t si = gw; si < P.n_super; si += GW) {
        const uint32_t s_first = (P.s_lo + si) * kWcSitesPerSuper;
        const uint32_t lo = max(s_first, P.v_lo), hi = min(s_first + kWcSitesPerSuper, P.v_hi);
        for (uint32_t p = lane; p < P.n_pairs; p += 32) {
            acc[2 * p] = 0.0;
            acc[2 * p + 1] = 0.0;
            acc_n[p] = 0;
        }
        double sum_a = 0.0, sum_b = 0.0;  // overall, site order (kept by every lane identically)
        uint32_t n_informative = 0, n_maps = 0;
        for (uint32_t v0 = lo; v0 < hi; v0 += 32) {
            const uint32_t nb = min(32u, hi - v0);
            __syncwarp();
            // stage the batch's counts: lane = site, loop over groups (coalesced global reads)
            if (lane < nb) {
                for (uint32_t g = 0; g < G1; ++g) {
                    cnts[(lane * G1 + g) * 2] = P.alt[g][v0 + lane];
                    cnts[(lane * G1 + g) * 2 + 1] = P.cnt[g][v0 + lane];
                }
            }
            __syncwarp();
            for (uint32_t s = 0; s < nb; ++s) {
                const uint32_t *sc = cnts + (size_t)s * G1 * 2;
                const uint32_t v = v0 + s;
                // alleles present over ALL samples, members or not (stats.rs:1826-1837)
                uint64_t tot_n = 0, tot_alt = 0;
                for (uint32_t g = 0; g < G1; ++g) {
                    tot_alt += sc[2 * g];
                    tot_n += sc[2 * g + 1];
                }
                const bool has1 = tot_alt > 0, has0 = tot_n > tot_alt;
                const bool any = has0 || has1;  // pop_sizes_populated (stats.rs:1919-1923, 1987)
                // ---- overall components: lanes 0-15 evaluate allele 0, lanes 16-31 allele 1
                double oa, ob;
                bool ok;
                const bool mine_one = lane >= 16;
                fm_wc_overall_allele(sc, P.G, mine_one, oa, ob, ok);
                const bool present = mine_one ? has1 : has0;
                if (!(ok && present)) {
                    oa = 0.0;
                    ob = 0.0;
                }
                // sum over alleles in ascending order: allele 0 first (stats.rs:1859, 1939-1940)
                const double a0 = __shfl_sync(0xffffffffu, oa, 0), b0 = __shfl_sync(0xffffffffu, ob, 0);
                const double a1 = __shfl_sync(0xffffffffu, oa, 16), b1 = __shfl_sync(0xffffffffu, ob, 16);
                double site_a = 0.0, site_b = 0.0;
                site_a += a0;
                site_b += b0;
                site_a += a1;
                site_b += b1;
                int state = 3;  // InsufficientDataForEstimation: no allele at all at this site
                if (any) {
                    state = fm_fst_state(site_a, site_b);
                    sum_a += site_a;  // stats.rs:2172-2184, 2222-2229
                    sum_b += site_b;
                    ++n_informative;
                    ++n_maps;
                } else {
                    site_a = 0.0;
                    site_b = 0.0;
                }
                if (lane == 0) {
                    if (P.site_state) P.site_state[v - P.v_lo] = state;
                    if (P.site_a) P.site_a[v - P.v_lo] = site_a;
                    if (P.site_b) P.site_b[v - P.v_lo] = site_b;
                }
                if (P.site_sizes && lane < P.G) P.site_sizes[(size_t)(v - P.v_lo) * P.G + lane] = any ? sc[2 * lane + 1] : 0u;
                if (P.site_sizes)
                    for (uint32_t g = 32 + lane; g < P.G; g += 32)
                        P.site_sizes[(size_t)(v - P.v_lo) * P.G + g] = any ? sc[2 * g + 1] : 0u;
                // ---- pairwise components: lane l handles pairs l, l+32, ...
                for (uint32_t p = lane; p < P.n_pairs; p += 32) {
                    const uint32_t i = P.pair_i[p], j = P.pair_j[p];
                    const uint32_t ni = sc[2 * i + 1], nj = sc[2 * j + 1];
                    double pa = 0.0, pb = 0.0;
                    const bool has = any && ni > 0 && nj > 0;  // stats.rs:1950-1952
                    if (has) {
                        const uint32_t ai = sc[2 * i], aj = sc[2 * j];
                        double xa, xb;
                        if (has0) {
                            fm_wc_pair_components(ni, ni - ai, nj, nj - aj, xa, xb);
                            pa += xa;
                            pb += xb;
                        }
                        if (has1) {
                            fm_wc_pair_components(ni, ai, nj, aj, xa, xb);
                            pa += xa;
                            pb += xb;
                        }
                        acc[2 * p] += pa;  // site order: stats.rs:2288-2289
                        acc[2 * p + 1] += pb;
                        acc_n[p] += 1;
                    }
                    if (P.pair_a) {
                        const size_t o = (size_t)(v - P.v_lo) * P.n_pairs + p;
                        P.pair_a[o] = has ? pa : fm_nan();
                        P.pair_b[o] = has ? pb : fm_nan();
                    }
                }
            }
        }
        __syncwarp();
        for (uint32_t p = lane; p < P.n_pairs; p += 32) {
            P.part_pair[((size_t)si * P.n_pairs + p) * 2] = acc[2 * p];
            P.part_pair[((size_t)si * P.n_pairs + p) * 2 + 1] = acc[2 * p + 1];
            P.part_pair_n[(size_t)si * P.n_pairs + p] = acc_n[p];
        }
        if (lane == 0) {
            P.part_overall[2 * si] = sum_a;
            P.part_overall[2 * si + 1] = sum_b;
            P.part_counts[2 * si] = n_informative;
            P.part_counts[2 * si + 1] = n_maps;
        }
        (void)n_sites;
    }
}

}  // namespace fm
