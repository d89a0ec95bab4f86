// fm_wc.cuh -- K4: Weir & Cockerham variance components from per-group allele/called counts.
//
// Mirrors calculate_fst_wc_at_site_with_membership (stats.rs:1814-2032),
// calculate_variance_components (stats.rs:2034-2127) and the region aggregation of
// calculate_overall_fst_wc (stats.rs:2145-2374) for biallelic matrices.
//
// Work decomposition: the host cuts every window (region) of sites at multiples of
// kWcSegSites of the site index; one warp owns one segment and walks it in site order.  For
// every site the 32 lanes split the G*(G-1)/2 population pairs (lane l takes pairs l, l+32, ...)
// and add the pair's (a, b) to that pair's accumulator, so every pair sum is accumulated
// sequentially in site order inside a segment -- the same association as the reference's
// `.sum()` over sites (stats.rs:2288-2289); fm_k_wc_fold then adds the segment partials of a
// window in segment order.  Cut points depend only on the site index, so the result does not
// depend on the grid, on how windows are batched into calls, or on how sites are sharded over
// GPUs (shards are aligned to kWcSegSites).
//
// Arithmetic: every FP64 expression keeps the reference's operation order.  Work that the
// reference repeats is shared without changing a single rounding: per-group allele frequencies
// are divided once per site (not once per pair), and the terms of a pair that depend only on the
// sample sizes (n_bar, c^2, the a-denominator, n_bar/(n_bar-1)) are evaluated once for both
// alleles.  A pair that is monomorphic at a site contributes exactly (+0, +0) in the reference
// (p_i = p_j = p_bar in {0, 1}), so its FP64 work is skipped.
#pragma once
#include "fm_device.cuh"
#include "fm_kernels.cuh"

namespace fm {

constexpr uint32_t kWcSegSites = 1024;  // segment granularity (== 32 batches)
constexpr int kWcWarpsPerCta = 4;

struct WcParams {
    const uint32_t *const *alt;  // [G + 1] device pointers, each [V]; index G = haplotypes with no group
    const uint32_t *const *cnt;  // [G + 1]
    uint32_t G, n_pairs;
    const uint16_t *pair_i, *pair_j;  // [n_pairs], i < j in label order
    const uint32_t *seg_lo, *seg_hi;  // [n_seg] site ranges, walked in order by one warp each
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

__host__ __device__ inline size_t fm_wc_warp_smem(uint32_t G, uint32_t n_pairs) {
    // acc [n_pairs][2] f64 | freq [G][2] f64 | term [G][3] f64 | cnts [32][G+1][2] u32 | acc_n [n_pairs] u32
    size_t b = (size_t)n_pairs * 16 + (size_t)G * 16 + (size_t)G * 24 + (size_t)32 * (G + 1) * 8 + (size_t)n_pairs * 4;
    return (b + 15) & ~(size_t)15;
}

__global__ void __launch_bounds__(kWcWarpsPerCta * 32)
fm_k_wc(const WcParams P) {
    extern __shared__ __align__(16) uint8_t wc_smem[];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t G = P.G, G1 = P.G + 1, NP = P.n_pairs;
    uint8_t *base = wc_smem + warp * fm_wc_warp_smem(G, NP);
    double *acc = reinterpret_cast<double *>(base);
    double *freq = acc + (size_t)NP * 2;          // [G][2]: allele-0 and allele-1 frequency of group g
    double *term = freq + (size_t)G * 2;          // [G][3]: (n-n_bar)^2, n*(p0-pbar0)^2, n*(p1-pbar1)^2
    uint32_t *cnts = reinterpret_cast<uint32_t *>(term + (size_t)G * 3);
    uint32_t *acc_n = cnts + (size_t)32 * G1 * 2;

    const uint32_t wpc = blockDim.x >> 5;  // the host may run fewer warps when staging is large
    const uint32_t gw = blockIdx.x * wpc + warp;
    const uint32_t GW = gridDim.x * wpc;
    for (uint32_t si = gw; si < P.n_seg; si += GW) {
        const uint32_t lo = P.seg_lo[si], hi = P.seg_hi[si];
        for (uint32_t p = lane; p < NP; p += 32) {
            acc[2 * p] = 0.0;
            acc[2 * p + 1] = 0.0;
            acc_n[p] = 0;
        }
        double sum_a = 0.0, sum_b = 0.0;  // overall, site order (kept by every lane identically)
        uint32_t n_informative = 0;
        for (uint32_t v0 = lo; v0 < hi; v0 += 32) {
            const uint32_t nb = min(32u, hi - v0);
            __syncwarp();
            // stage the batch's counts: lane = site, loop over groups (coalesced global reads)
            if (lane < nb) {
                for (uint32_t g = 0; g < G1; ++g) {
                    cnts[(lane * G1 + g) * 2] = __ldg(P.alt[g] + v0 + lane);
                    cnts[(lane * G1 + g) * 2 + 1] = __ldg(P.cnt[g] + v0 + lane);
                }
            }
            __syncwarp();
            for (uint32_t s = 0; s < nb; ++s) {
                const uint32_t *sc = cnts + (size_t)s * G1 * 2;
                const uint32_t v = v0 + s;
                // ---- integer totals (exact in any order)
                // alleles present over ALL samples, members or not (stats.rs:1826-1837);
                // r, sum n, sum target over the groups with data (stats.rs:1907-1918)
                uint32_t t_alt = 0, t_n = 0, m_alt = 0, m_n = 0, m_r = 0;
                for (uint32_t g = lane; g < G1; g += 32) {
                    const uint32_t a = sc[2 * g], n = sc[2 * g + 1];
                    t_alt += a;
                    t_n += n;
                    if (g < G && n > 0) {
                        m_alt += a;
                        m_n += n;
                        m_r += 1;
                    }
                }
                t_alt = fm_warp_sum_u(t_alt);
                t_n = fm_warp_sum_u(t_n);
                m_alt = fm_warp_sum_u(m_alt);
                m_n = fm_warp_sum_u(m_n);
                m_r = fm_warp_sum_u(m_r);
                const bool has1 = t_alt > 0, has0 = t_n > t_alt;
                const bool any = has0 || has1;  // pop_sizes_populated (stats.rs:1919-1923, 1987)
                // ---- per-group frequencies, one division per group and allele
                for (uint32_t g = lane; g < G; g += 32) {
                    const uint32_t a = sc[2 * g], n = sc[2 * g + 1];
                    if (n > 0) {
                        freq[2 * g] = (double)(n - a) / (double)n;
                        freq[2 * g + 1] = (double)a / (double)n;
                    }
                }
                // ---- overall components (calculate_variance_components, stats.rs:2034-2127)
                double site_a = 0.0, site_b = 0.0;
                if (any && m_r >= 2) {  // fewer than two groups with data: no allele contributes
                    const double r = (double)m_r;
                    const double n_bar = (double)m_n / r;
                    if (!((n_bar - 1.0) < 1e-9)) {
                        const double gp0 = (double)(m_n - m_alt) / (double)m_n;
                        const double gp1 = (double)m_alt / (double)m_n;
                        __syncwarp();
                        for (uint32_t g = lane; g < G; g += 32) {
                            const uint32_t n = sc[2 * g + 1];
                            if (n > 0) {
                                const double nd = (double)n;
                                const double d = nd - n_bar;
                                term[3 * g] = d * d;
                                const double q0 = freq[2 * g] - gp0, q1 = freq[2 * g + 1] - gp1;
                                term[3 * g + 1] = nd * q0 * q0;
                                term[3 * g + 2] = nd * q1 * q1;
                            }
                        }
                        __syncwarp();
                        double ssd = 0.0, ns0 = 0.0, ns1 = 0.0;  // sums in group order
                        for (uint32_t g = 0; g < G; ++g) {
                            if (sc[2 * g + 1] > 0) {
                                ssd += term[3 * g];
                                ns0 += term[3 * g + 1];
                                ns1 += term[3 * g + 2];
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
                int state = 3;  // InsufficientDataForEstimation: no allele at all at this site
                if (any) {
                    state = fm_fst_state(site_a, site_b);
                    sum_a += site_a;  // stats.rs:2172-2184, 2222-2229
                    sum_b += site_b;
                    ++n_informative;
                }
                if (lane == 0) {
                    if (P.site_state) P.site_state[v - P.out_base] = state;
                    if (P.site_a) P.site_a[v - P.out_base] = site_a;
                    if (P.site_b) P.site_b[v - P.out_base] = site_b;
                }
                if (P.site_sizes)
                    for (uint32_t g = lane; g < G; g += 32)
                        P.site_sizes[(size_t)(v - P.out_base) * G + g] = any ? sc[2 * g + 1] : 0u;
                __syncwarp();  // freq[] written above is read by other lanes below
                // ---- pairwise components: lane l handles pairs l, l+32, ...
                for (uint32_t p = lane; p < NP; p += 32) {
                    const uint32_t i = __ldg(P.pair_i + p), j = __ldg(P.pair_j + p);
                    const uint32_t ni = sc[2 * i + 1], nj = sc[2 * j + 1];
                    double pa = 0.0, pb = 0.0;
                    const bool has = any && ni > 0 && nj > 0;  // stats.rs:1950-1952
                    if (has) {
                        const uint32_t ai = sc[2 * i], aj = sc[2 * j];
                        const uint32_t asum = ai + aj, nsum = ni + nj;
                        if (asum != 0 && asum != nsum) {  // polymorphic in this pair
                            const double n1 = (double)ni, n2 = (double)nj;
                            const double n_bar = (double)nsum / 2.0;
                            if (!((n_bar - 1.0) < 1e-9)) {
                                const double d1 = n1 - n_bar, d2 = n2 - n_bar;
                                double ssd = 0.0;
                                ssd += d1 * d1;
                                ssd += d2 * d2;
                                const double c_squared = ssd / (2.0 * n_bar * n_bar);
                                const double a_den = 1.0 - (c_squared / 1.0);
                                const double nb_ratio = n_bar / (n_bar - 1.0);
                                const double nsd = (double)nsum;
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    if (u == 0 ? has0 : has1) {
                                        const double gp = (double)(u == 0 ? nsum - asum : asum) / nsd;
                                        const double q1 = freq[2 * i + u] - gp, q2 = freq[2 * j + u] - gp;
                                        double num = 0.0;
                                        num += n1 * q1 * q1;
                                        num += n2 * q2 * q2;
                                        const double s2 = num / (1.0 * n_bar);
                                        const double x = gp * (1.0 - gp) - (1.0 / 2.0) * s2;
                                        pa += (s2 - (x / (n_bar - 1.0))) / a_den;
                                        pb += nb_ratio * x;
                                    }
                                }
                            }
                        }
                        acc[2 * p] += pa;  // site order: stats.rs:2288-2289
                        acc[2 * p + 1] += pb;
                        acc_n[p] += 1;
                    }
                    if (P.pair_a) {
                        const size_t o = (size_t)(v - P.out_base) * NP + p;
                        P.pair_a[o] = has ? pa : fm_nan();
                        P.pair_b[o] = has ? pb : fm_nan();
                    }
                }
                __syncwarp();  // freq[] / term[] are overwritten by the next site
            }
        }
        __syncwarp();
        for (uint32_t p = lane; p < NP; p += 32) {
            P.part_pair[((size_t)si * NP + p) * 2] = acc[2 * p];
            P.part_pair[((size_t)si * NP + p) * 2 + 1] = acc[2 * p + 1];
            P.part_pair_n[(size_t)si * NP + p] = acc_n[p];
        }
        if (lane == 0) {
            P.part_overall[2 * si] = sum_a;
            P.part_overall[2 * si + 1] = sum_b;
            P.part_counts[si] = n_informative;
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
