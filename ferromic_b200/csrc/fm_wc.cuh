// fm_wc.cuh -- K4: Weir & Cockerham variance components from per-group allele/called counts.
//
// Mirrors calculate_fst_wc_at_site_with_membership (stats.rs:1814-2032),
// calculate_variance_components (stats.rs:2034-2127) and the region aggregation of
// calculate_overall_fst_wc (stats.rs:2145-2374) for biallelic matrices.
//
// Work decomposition: the host cuts every window (region) of sites at multiples of
// kWcSegSites of the site index; one CTA owns one segment and walks it in site order, 32 sites
// (one batch) at a time.  Inside the CTA
//   * all threads stage the batch's per-group counts and divide the per-group allele
//     frequencies once (shared memory), lanes of warp 0 derive the per-site totals / allele flags;
//   * "pair warps": lane l of pair warp w owns pairs (w*32 + l) + k*stride in REGISTERS and adds
//     the pair's (a, b) site after site -- every pair sum is accumulated sequentially in site
//     order inside a segment, the same association as the reference's `.sum()` over sites
//     (stats.rs:2288-2289);
//   * the "overall warp" evaluates the all-population components with one site per lane
//     (groups in order, exactly as calculate_variance_components) and then adds the 32 site
//     values in site order.
// fm_k_wc_fold adds the segment partials of a window in segment order.  Cut points depend only
// on the site index, so the result does not depend on the grid, on how windows are batched into
// calls, or on how sites are sharded over GPUs (shards are aligned to kWcSegSites).
//
// Arithmetic: every FP64 expression keeps the reference's operation order.  Work that the
// reference repeats is shared without changing a single rounding: per-group allele frequencies
// are divided once per site (not once per pair), and the terms of a pair that depend only on the
// sample sizes (n_bar, c^2, the a-denominator, n_bar/(n_bar-1)) are evaluated once for both
// alleles and reused while (n_i, n_j) do not change from site to site.  A pair that is
// monomorphic at a site contributes exactly (+0, +0) in the reference (p_i = p_j = p_bar in
// {0, 1}), so its FP64 work is skipped.
#pragma once
#include "fm_device.cuh"
#include "fm_kernels.cuh"

namespace fm {

constexpr uint32_t kWcSegSites = 1024;   // segment granularity (== 32 batches)
constexpr uint32_t kWcMaxPairWarps = 11; // + 1 overall warp = 384 threads (3 CTAs per SM at <= 56 registers)
constexpr uint32_t kWcMaxKP = 8;         // pairs per lane (template parameter KP <= this)

struct WcParams {
    const uint32_t *const *alt;  // [G + 1] device pointers, each [V]; index G = haplotypes with no group
    const uint32_t *const *cnt;  // [G + 1]
    const uint32_t *const *acount;  // multi-allelic: [G + 1] pointers to per-allele counts [V][A] (else nullptr)
    uint32_t A;                  // alleles per site slot (2 for biallelic, 2^NB for multi-allelic)
    uint32_t G, n_pairs;
    uint32_t n_pair_warps;            // pair warps per CTA (block = (n_pair_warps + 1) * 32 threads)
    const uint16_t *pair_i, *pair_j;  // [n_pairs], i < j in label order
    const uint32_t *seg_lo, *seg_hi;  // [n_seg] site ranges, walked in order by one CTA each
    uint32_t n_seg;
    uint32_t out_base;                // per-site outputs are indexed v - out_base
    // per-site outputs, any may be nullptr
    int32_t *site_state;
    double *site_a, *site_b;
    uint32_t *site_sizes;        // [n_sites][G]
    double *pair_a, *pair_b;     // [n_sites][n_pairs]; NaN when the pair has no data at the site
    // per-segment partials
    double *part_overall;        // [n_seg][2]            sum a, sum b
    uint32_t *part_counts;       // [n_seg]               sites with an estimate (!= InsufficientData)
    double *part_pair;           // [n_seg][n_pairs][2]
    uint32_t *part_pair_n;       // [n_seg][n_pairs]      informative sites per pair
};

__host__ __device__ inline size_t fm_wc_cta_smem(uint32_t G) {
    // freq [32][G] double2 | cnts [32][G+1] uint2 | info [32] u32
    return (size_t)32 * G * 16 + (size_t)32 * (G + 1) * 8 + 32 * 4 + 16;
}

template <int KP>
__global__ void __launch_bounds__(384, 3)
fm_k_wc(const WcParams P) {
    extern __shared__ __align__(16) uint8_t wc_smem[];
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint32_t G = P.G, G1 = P.G + 1, NP = P.n_pairs;
    const uint32_t NW = P.n_pair_warps;
    double2 *freq = reinterpret_cast<double2 *>(wc_smem);                  // [32][G]: (p allele 0, p allele 1)
    uint2 *cnts = reinterpret_cast<uint2 *>(freq + (size_t)32 * G);        // [32][G1]: (alt, called)
    uint32_t *info = reinterpret_cast<uint32_t *>(cnts + (size_t)32 * G1); // [32]: bit0 has0, bit1 has1
    const bool overall_warp = warp == NW;

    // pairs owned by this lane (registers)
    uint32_t pi[KP], pj[KP];
    bool pvalid[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const uint32_t p = warp * 32 + lane + (uint32_t)k * NW * 32;
        pvalid[k] = !overall_warp && p < NP;
        pi[k] = pvalid[k] ? __ldg(P.pair_i + p) : 0u;
        pj[k] = pvalid[k] ? __ldg(P.pair_j + p) : 0u;
    }

    for (uint32_t si = blockIdx.x; si < P.n_seg; si += gridDim.x) {
        const uint32_t lo = P.seg_lo[si], hi = P.seg_hi[si];
        double acc_a[KP], acc_b[KP];
        uint32_t acc_n[KP];
        // size-only terms of the lane's pairs, reused while (n_i, n_j) repeat from site to site
        uint32_t c_ni[KP], c_nj[KP];
        double c_nbar[KP], c_aden[KP], c_ratio[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            acc_a[k] = 0.0;
            acc_b[k] = 0.0;
            acc_n[k] = 0;
            c_ni[k] = 0xffffffffu;
            c_nj[k] = 0xffffffffu;
            c_nbar[k] = c_aden[k] = c_ratio[k] = 0.0;
        }
        double sum_a = 0.0, sum_b = 0.0;  // overall warp: site order, identical in every lane
        uint32_t n_informative = 0;
        for (uint32_t v0 = lo; v0 < hi; v0 += 32) {
            const uint32_t nb = min(32u, hi - v0);
            __syncthreads();  // previous batch fully consumed
            // ---- stage counts (lane = site: coalesced), one warp per group
            for (uint32_t g = warp; g < G1; g += (nt >> 5))
                if (lane < nb)
                    cnts[lane * G1 + g] = make_uint2(__ldg(P.alt[g] + v0 + lane), __ldg(P.cnt[g] + v0 + lane));
            __syncthreads();
            // ---- per-group frequencies: one division per (site, group, allele)
            for (uint32_t i = tid; i < nb * G; i += nt) {
                const uint32_t s = i / G, g = i - s * G;
                const uint2 c = cnts[s * G1 + g];
                if (c.y > 0) freq[s * G + g] = make_double2((double)(c.y - c.x) / (double)c.y, (double)c.x / (double)c.y);
            }
            // ---- alleles present over ALL samples, members or not (stats.rs:1826-1837)
            if (warp == 0 && lane < nb) {
                uint32_t t_alt = 0, t_n = 0;
                for (uint32_t g = 0; g < G1; ++g) {
                    const uint2 c = cnts[lane * G1 + g];
                    t_alt += c.x;
                    t_n += c.y;
                }
                info[lane] = (t_n > t_alt ? 1u : 0u) | (t_alt > 0 ? 2u : 0u);
            }
            __syncthreads();
            if (overall_warp) {
                // ---- overall components, one site per lane (calculate_variance_components,
                // stats.rs:2034-2127, over the groups with data: stats.rs:1907-1918)
                double site_a = 0.0, site_b = 0.0;
                int state = 3;  // InsufficientDataForEstimation: no allele at all at this site
                bool any = false;
                if (lane < nb) {
                    const uint2 *sc = cnts + lane * G1;
                    const double2 *fr = freq + lane * G;
                    const uint32_t inf = info[lane];
                    const bool has0 = inf & 1u, has1 = inf & 2u;
                    any = inf != 0;  // pop_sizes_populated (stats.rs:1919-1923, 1987)
                    uint32_t m_alt = 0, m_n = 0, m_r = 0;
                    for (uint32_t g = 0; g < G; ++g) {
                        const uint2 c = sc[g];
                        if (c.y > 0) {
                            m_alt += c.x;
                            m_n += c.y;
                            m_r += 1;
                        }
                    }
                    if (any && m_r >= 2) {  // fewer than two groups with data: no allele contributes
                        const double r = (double)m_r;
                        const double n_bar = (double)m_n / r;
                        if (!((n_bar - 1.0) < 1e-9)) {
                            const double gp0 = (double)(m_n - m_alt) / (double)m_n;
                            const double gp1 = (double)m_alt / (double)m_n;
                            double ssd = 0.0, ns0 = 0.0, ns1 = 0.0;  // sums in group order
                            for (uint32_t g = 0; g < G; ++g) {
                                const uint32_t n = sc[g].y;
                                if (n > 0) {
                                    const double nd = (double)n;
                                    const double d = nd - n_bar;
                                    ssd += d * d;
                                    const double2 f = fr[g];
                                    const double q0 = f.x - gp0, q1 = f.y - gp1;
                                    ns0 += nd * q0 * q0;
                                    ns1 += nd * q1 * q1;
                                }
                            }
                            const double c_squared = ssd / (r * n_bar * n_bar);
                            const double a_den = 1.0 - (c_squared / (r - 1.0));
                            const double nb_ratio = n_bar / (n_bar - 1.0);
                            const bool s_ok = (r - 1.0) > 1e-9 && n_bar > 1e-9;
                            // sum over alleles in ascending order: allele 0 first (stats.rs:1859, 1939-1940)
                            if (has0) {
                                const double s2 = s_ok ? ns0 / ((r - 1.0) * n_bar) : 0.0;
                                const double x = gp0 * (1.0 - gp0) - ((r - 1.0) / r) * s2;
                                site_a += (s2 - (x / (n_bar - 1.0))) / a_den;
                                site_b += nb_ratio * x;
                            }
                            if (has1) {
                                const double s2 = s_ok ? ns1 / ((r - 1.0) * n_bar) : 0.0;
                                const double x = gp1 * (1.0 - gp1) - ((r - 1.0) / r) * s2;
                                site_a += (s2 - (x / (n_bar - 1.0))) / a_den;
                                site_b += nb_ratio * x;
                            }
                        }
                    }
                    if (any) state = fm_fst_state(site_a, site_b);
                    const uint32_t o = v0 + lane - P.out_base;
                    if (P.site_state) P.site_state[o] = state;
                    if (P.site_a) P.site_a[o] = site_a;
                    if (P.site_b) P.site_b[o] = site_b;
                    if (P.site_sizes)
                        for (uint32_t g = 0; g < G; ++g) P.site_sizes[(size_t)o * G + g] = any ? sc[g].y : 0u;
                }
                // region sums in site order (stats.rs:2172-2184, 2222-2229)
                const uint32_t any_mask = __ballot_sync(0xffffffffu, any);
                for (uint32_t s = 0; s < nb; ++s) {
                    const double a_s = __shfl_sync(0xffffffffu, site_a, s);
                    const double b_s = __shfl_sync(0xffffffffu, site_b, s);
                    if ((any_mask >> s) & 1u) {
                        sum_a += a_s;
                        sum_b += b_s;
                        ++n_informative;
                    }
                }
            } else {
                // ---- pairwise components (r == 2), sites in order
                for (uint32_t s = 0; s < nb; ++s) {
                    const uint2 *sc = cnts + s * G1;
                    const double2 *fr = freq + s * G;
                    const uint32_t inf = info[s];
                    const bool has0 = inf & 1u, has1 = inf & 2u, any = inf != 0;
#pragma unroll
                    for (int k = 0; k < KP; ++k) {
                        if (!pvalid[k]) continue;
                        const uint2 ci = sc[pi[k]], cj = sc[pj[k]];
                        const uint32_t ni = ci.y, nj = cj.y;
                        double pa = 0.0, pb = 0.0;
                        const bool has = any && ni > 0 && nj > 0;  // stats.rs:1950-1952
                        if (has) {
                            const uint32_t asum = ci.x + cj.x, nsum = ni + nj;
                            if (asum != 0 && asum != nsum) {  // polymorphic in this pair
                                if (ni != c_ni[k] || nj != c_nj[k]) {
                                    c_ni[k] = ni;
                                    c_nj[k] = nj;
                                    const double n1 = (double)ni, n2 = (double)nj;
                                    const double n_bar = (double)nsum / 2.0;
                                    const double d1 = n1 - n_bar, d2 = n2 - n_bar;
                                    double ssd = 0.0;
                                    ssd += d1 * d1;
                                    ssd += d2 * d2;
                                    const double c_squared = ssd / (2.0 * n_bar * n_bar);
                                    c_nbar[k] = n_bar;
                                    c_aden[k] = 1.0 - (c_squared / 1.0);
                                    c_ratio[k] = n_bar / (n_bar - 1.0);
                                }
                                const double n_bar = c_nbar[k];
                                if (!((n_bar - 1.0) < 1e-9)) {
                                    const double n1 = (double)ni, n2 = (double)nj, nsd = (double)nsum;
                                    const double2 fi = fr[pi[k]], fj = fr[pj[k]];
#pragma unroll
                                    for (int u = 0; u < 2; ++u) {
                                        if (u == 0 ? has0 : has1) {
                                            const double gp = (double)(u == 0 ? nsum - asum : asum) / nsd;
                                            const double q1 = (u == 0 ? fi.x : fi.y) - gp;
                                            const double q2 = (u == 0 ? fj.x : fj.y) - gp;
                                            double num = 0.0;
                                            num += n1 * q1 * q1;
                                            num += n2 * q2 * q2;
                                            const double s2 = num / (1.0 * n_bar);
                                            const double x = gp * (1.0 - gp) - (1.0 / 2.0) * s2;
                                            pa += (s2 - (x / (n_bar - 1.0))) / c_aden[k];
                                            pb += c_ratio[k] * x;
                                        }
                                    }
                                }
                            }
                            acc_a[k] += pa;  // site order: stats.rs:2288-2289
                            acc_b[k] += pb;
                            acc_n[k] += 1;
                        }
                        if (P.pair_a) {
                            const uint32_t p = warp * 32 + lane + (uint32_t)k * NW * 32;
                            const size_t o = (size_t)(v0 + s - P.out_base) * NP + p;
                            P.pair_a[o] = has ? pa : fm_nan();
                            P.pair_b[o] = has ? pb : fm_nan();
                        }
                    }
                }
            }
        }
        if (overall_warp) {
            if (lane == 0) {
                P.part_overall[2 * (size_t)si] = sum_a;
                P.part_overall[2 * (size_t)si + 1] = sum_b;
                P.part_counts[si] = n_informative;
            }
        } else {
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                if (!pvalid[k]) continue;
                const uint32_t p = warp * 32 + lane + (uint32_t)k * NW * 32;
                P.part_pair[((size_t)si * NP + p) * 2] = acc_a[k];
                P.part_pair[((size_t)si * NP + p) * 2 + 1] = acc_b[k];
                P.part_pair_n[(size_t)si * NP + p] = acc_n[k];
            }
        }
    }
}

// Multi-allelic variant (max_allele 2..15): same decomposition, with the reference's loop over
// every allele present at the site (stats.rs:1849-1859, 1939-1983) instead of the two alleles of
// a biallelic site.  An allele that is absent from a pair, or fixed in it, contributes exactly
// (+0, +0) and is skipped.
__host__ __device__ inline size_t fm_wc_multi_cta_smem(uint32_t G, uint32_t A) {
    // freq [32][G][A] f64 | cn [32][G+1] u32 | ca [32][G+1][A] u32 | info [32] u32
    return (size_t)32 * G * A * 8 + (size_t)32 * (G + 1) * 4 + (size_t)32 * (G + 1) * A * 4 + 32 * 4 + 16;
}

template <int KP>
__global__ void __launch_bounds__((kWcMaxPairWarps + 1) * 32, 1)
fm_k_wc_multi(const WcParams P) {
    extern __shared__ __align__(16) uint8_t wc_smem[];
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint32_t G = P.G, G1 = P.G + 1, NP = P.n_pairs, A = P.A;
    const uint32_t NW = P.n_pair_warps;
    double *freq = reinterpret_cast<double *>(wc_smem);                          // [32][G][A]
    uint32_t *cn = reinterpret_cast<uint32_t *>(freq + (size_t)32 * G * A);      // [32][G1] called
    uint32_t *ca = cn + (size_t)32 * G1;                                         // [32][G1][A] allele counts
    uint32_t *info = ca + (size_t)32 * G1 * A;                                   // [32] mask of alleles present
    const bool overall_warp = warp == NW;
    uint32_t pi[KP], pj[KP];
    bool pvalid[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const uint32_t p = warp * 32 + lane + (uint32_t)k * NW * 32;
        pvalid[k] = !overall_warp && p < NP;
        pi[k] = pvalid[k] ? __ldg(P.pair_i + p) : 0u;
        pj[k] = pvalid[k] ? __ldg(P.pair_j + p) : 0u;
    }
    for (uint32_t si = blockIdx.x; si < P.n_seg; si += gridDim.x) {
        const uint32_t lo = P.seg_lo[si], hi = P.seg_hi[si];
        double acc_a[KP], acc_b[KP];
        uint32_t acc_n[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            acc_a[k] = 0.0;
            acc_b[k] = 0.0;
            acc_n[k] = 0;
        }
        double sum_a = 0.0, sum_b = 0.0;
        uint32_t n_informative = 0;
        for (uint32_t v0 = lo; v0 < hi; v0 += 32) {
            const uint32_t nb = min(32u, hi - v0);
            __syncthreads();
            for (uint32_t g = warp; g < G1; g += (nt >> 5))
                if (lane < nb) {
                    cn[lane * G1 + g] = __ldg(P.cnt[g] + v0 + lane);
                    const uint32_t *src = P.acount[g] + (size_t)(v0 + lane) * A;
                    for (uint32_t a = 0; a < A; ++a) ca[(lane * G1 + g) * A + a] = __ldg(src + a);
                }
            __syncthreads();
            for (uint32_t i = tid; i < nb * G * A; i += nt) {
                const uint32_t s = i / (G * A), r = i - s * G * A, g = r / A, a = r - g * A;
                const uint32_t n = cn[s * G1 + g];
                if (n > 0) freq[(s * G + g) * A + a] = (double)ca[(s * G1 + g) * A + a] / (double)n;
            }
            if (warp == 0 && lane < nb) {  // alleles present over ALL samples (stats.rs:1826-1837)
                uint32_t mask = 0;
                for (uint32_t a = 0; a < A; ++a) {
                    uint32_t t = 0;
                    for (uint32_t g = 0; g < G1; ++g) t += ca[(lane * G1 + g) * A + a];
                    if (t) mask |= 1u << a;
                }
                info[lane] = mask;
            }
            __syncthreads();
            if (overall_warp) {
                double site_a = 0.0, site_b = 0.0;
                int state = 3;
                bool any = false;
                if (lane < nb) {
                    const uint32_t *sn = cn + lane * G1;
                    const uint32_t *sa = ca + (size_t)lane * G1 * A;
                    const double *fr = freq + (size_t)lane * G * A;
                    const uint32_t present = info[lane];
                    any = present != 0;
                    uint32_t m_n = 0, m_r = 0;
                    for (uint32_t g = 0; g < G; ++g)
                        if (sn[g] > 0) {
                            m_n += sn[g];
                            m_r += 1;
                        }
                    if (any && m_r >= 2) {
                        const double r = (double)m_r;
                        const double n_bar = (double)m_n / r;
                        if (!((n_bar - 1.0) < 1e-9)) {
                            double ssd = 0.0;
                            for (uint32_t g = 0; g < G; ++g)
                                if (sn[g] > 0) {
                                    const double d = (double)sn[g] - n_bar;
                                    ssd += d * d;
                                }
                            const double c_squared = ssd / (r * n_bar * n_bar);
                            const double a_den = 1.0 - (c_squared / (r - 1.0));
                            const double nb_ratio = n_bar / (n_bar - 1.0);
                            const bool s_ok = (r - 1.0) > 1e-9 && n_bar > 1e-9;
                            for (uint32_t a = 0; a < A; ++a) {  // ascending allele order (stats.rs:1859)
                                if (!((present >> a) & 1u)) continue;
                                uint32_t m_t = 0;
                                for (uint32_t g = 0; g < G; ++g)
                                    if (sn[g] > 0) m_t += sa[g * A + a];
                                const double gp = (double)m_t / (double)m_n;
                                double ns = 0.0;
                                for (uint32_t g = 0; g < G; ++g)
                                    if (sn[g] > 0) {
                                        const double q = fr[g * A + a] - gp;
                                        ns += (double)sn[g] * q * q;
                                    }
                                const double s2 = s_ok ? ns / ((r - 1.0) * n_bar) : 0.0;
                                const double x = gp * (1.0 - gp) - ((r - 1.0) / r) * s2;
                                site_a += (s2 - (x / (n_bar - 1.0))) / a_den;
                                site_b += nb_ratio * x;
                            }
                        }
                    }
                    if (any) state = fm_fst_state(site_a, site_b);
                    const uint32_t o = v0 + lane - P.out_base;
                    if (P.site_state) P.site_state[o] = state;
                    if (P.site_a) P.site_a[o] = site_a;
                    if (P.site_b) P.site_b[o] = site_b;
                    if (P.site_sizes)
                        for (uint32_t g = 0; g < G; ++g) P.site_sizes[(size_t)o * G + g] = any ? sn[g] : 0u;
                }
                const uint32_t any_mask = __ballot_sync(0xffffffffu, any);
                for (uint32_t s = 0; s < nb; ++s) {
                    const double a_s = __shfl_sync(0xffffffffu, site_a, s);
                    const double b_s = __shfl_sync(0xffffffffu, site_b, s);
                    if ((any_mask >> s) & 1u) {
                        sum_a += a_s;
                        sum_b += b_s;
                        ++n_informative;
                    }
                }
            } else {
                for (uint32_t s = 0; s < nb; ++s) {
                    const uint32_t *sn = cn + s * G1;
                    const uint32_t *sa = ca + (size_t)s * G1 * A;
                    const double *fr = freq + (size_t)s * G * A;
                    const uint32_t present = info[s];
                    const bool any = present != 0;
#pragma unroll
                    for (int k = 0; k < KP; ++k) {
                        if (!pvalid[k]) continue;
                        const uint32_t i = pi[k], j = pj[k];
                        const uint32_t ni = sn[i], nj = sn[j];
                        double pa = 0.0, pb = 0.0;
                        const bool has = any && ni > 0 && nj > 0;
                        if (has) {
                            const uint32_t nsum = ni + nj;
                            const double n1 = (double)ni, n2 = (double)nj, nsd = (double)nsum;
                            const double n_bar = nsd / 2.0;
                            if (!((n_bar - 1.0) < 1e-9)) {
                                const double d1 = n1 - n_bar, d2 = n2 - n_bar;
                                double ssd = 0.0;
                                ssd += d1 * d1;
                                ssd += d2 * d2;
                                const double c_squared = ssd / (2.0 * n_bar * n_bar);
                                const double a_den = 1.0 - (c_squared / 1.0);
                                const double nb_ratio = n_bar / (n_bar - 1.0);
                                for (uint32_t a = 0; a < A; ++a) {
                                    if (!((present >> a) & 1u)) continue;
                                    const uint32_t asum = sa[i * A + a] + sa[j * A + a];
                                    if (asum == 0 || asum == nsum) continue;  // exactly (+0, +0)
                                    const double gp = (double)asum / nsd;
                                    const double q1 = fr[i * A + a] - gp, q2 = fr[j * A + a] - gp;
                                    double num = 0.0;
                                    num += n1 * q1 * q1;
                                    num += n2 * q2 * q2;
                                    const double s2 = num / (1.0 * n_bar);
                                    const double x = gp * (1.0 - gp) - (1.0 / 2.0) * s2;
                                    pa += (s2 - (x / (n_bar - 1.0))) / a_den;
                                    pb += nb_ratio * x;
                                }
                            }
                            acc_a[k] += pa;
                            acc_b[k] += pb;
                            acc_n[k] += 1;
                        }
                        if (P.pair_a) {
                            const uint32_t p = warp * 32 + lane + (uint32_t)k * NW * 32;
                            const size_t o = (size_t)(v0 + s - P.out_base) * NP + p;
                            P.pair_a[o] = has ? pa : fm_nan();
                            P.pair_b[o] = has ? pb : fm_nan();
                        }
                    }
                }
            }
        }
        if (overall_warp) {
            if (lane == 0) {
                P.part_overall[2 * (size_t)si] = sum_a;
                P.part_overall[2 * (size_t)si + 1] = sum_b;
                P.part_counts[si] = n_informative;
            }
        } else {
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                if (!pvalid[k]) continue;
                const uint32_t p = warp * 32 + lane + (uint32_t)k * NW * 32;
                P.part_pair[((size_t)si * NP + p) * 2] = acc_a[k];
                P.part_pair[((size_t)si * NP + p) * 2 + 1] = acc_b[k];
                P.part_pair_n[(size_t)si * NP + p] = acc_n[k];
            }
        }
    }
}

// Window totals: window w owns segments [wseg[w], wseg[w+1]); thread (w, p) adds the partials
// of pair p in segment order (p == n_pairs: the overall components).  Lanes run over p, so the
// segment rows are read coalesced.
__global__ void __launch_bounds__(128)
fm_k_wc_fold(const double *__restrict__ part_overall, const uint32_t *__restrict__ part_counts,
             const double *__restrict__ part_pair, const uint32_t *__restrict__ part_pair_n,
             const uint32_t *__restrict__ wseg, uint32_t n_windows, uint32_t n_pairs,
             double *__restrict__ out_overall /*[n_w][2]*/, uint64_t *__restrict__ out_sites /*[n_w]*/,
             double *__restrict__ out_pair /*[n_w][n_pairs][2]*/, uint64_t *__restrict__ out_pair_n) {
    const uint32_t per_w = n_pairs + 1;
    const uint64_t total = (uint64_t)n_windows * per_w;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t w = (uint32_t)(t / per_w), p = (uint32_t)(t % per_w);
        const uint32_t s0 = wseg[w], s1 = wseg[w + 1];
        double a = 0.0, b = 0.0;
        uint64_t n = 0;
        if (p == n_pairs) {
            for (uint32_t s = s0; s < s1; ++s) {
                a += part_overall[2 * (size_t)s];
                b += part_overall[2 * (size_t)s + 1];
                n += part_counts[s];
            }
            out_overall[2 * (size_t)w] = a;
            out_overall[2 * (size_t)w + 1] = b;
            out_sites[w] = n;
        } else {
#pragma unroll 4
            for (uint32_t s = s0; s < s1; ++s) {
                const size_t o = (size_t)s * n_pairs + p;
                a += part_pair[2 * o];
                b += part_pair[2 * o + 1];
                n += part_pair_n[o];
            }
            const size_t o = (size_t)w * n_pairs + p;
            out_pair[2 * o] = a;
            out_pair[2 * o + 1] = b;
            out_pair_n[o] = n;
        }
    }
}

}  // namespace fm
