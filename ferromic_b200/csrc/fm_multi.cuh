// fm_multi.cuh -- multi-allelic sites (max_allele in 2..15): the reference's "general" code paths
//   dense_collect_counts                 stats.rs:2823-2880
//   count_segregating_sites_dense        stats.rs:3891-4026  (two distinct called alleles)
//   calculate_pi_dense (general)         stats.rs:4573-4589
//   dense_hudson_sites_general           stats.rs:3072-3177
//   calculate_dxy_dense                  stats.rs:2526-2611
//   compute_pi_metrics_fast / freq_summary / dxy_from_counts / hudson_site_from_variant (sparse)
//                                        stats.rs:2631-2821, 2907-3014
//
// Layout: a group carries NB = ceil(log2(max_allele + 1)) allele bitplanes (bit k of the allele
// index of haplotype j is bit j of plane k) plus the called plane.  The count of allele a at a
// site is popc(called & AND_k (a_k ? plane_k : ~plane_k)); counts are cached per group as
// u32 [V][2^NB] next to the called counts, and every estimator is evaluated from the cached
// counts by light kernels that share the partial-reduction shape of the biallelic path.
// Sum of squared counts is an exact integer in FP64, so its accumulation order is irrelevant; the
// D_xy dot product is accumulated in ascending allele order (the reference walks its "used" list,
// a difference of at most 1 ulp per term -- far inside the 1e-9 contract).
#pragma once
#include "fm_device.cuh"
#include "fm_kernels.cuh"

namespace fm {

// One sub-warp of LPS lanes per site; lanes stride the row's uint4 columns.
template <int NB>
__global__ void __launch_bounds__(256)
fm_k_allele_counts(const uint4 *__restrict__ planes, size_t plane_stride_u4, const uint4 *__restrict__ called,
                   uint32_t wq, uint32_t cap, uint32_t lps_log2, uint32_t v_lo, uint32_t v_hi,
                   uint32_t *__restrict__ acount, uint32_t *__restrict__ cnt) {
    constexpr uint32_t A = 1u << NB;
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lps = 1u << lps_log2, sps = 32u >> lps_log2;
    const uint32_t phase = lane & (lps - 1), slot = lane >> lps_log2;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t GW = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n_groups = (v_hi - v_lo + sps - 1) / sps;
    for (uint32_t it = gw; it < n_groups; it += GW) {
        const uint32_t v = v_lo + it * sps + slot;
        uint32_t c[A];
#pragma unroll
        for (uint32_t a = 0; a < A; ++a) c[a] = 0;
        uint32_t n_called = 0;
        if (v < v_hi) {
            for (uint32_t col = phase; col < wq; col += lps) {
                const size_t o = (size_t)v * wq + col;
                uint32_t cw[4];
                if (called) {
                    const uint4 x = called[o];
                    cw[0] = x.x; cw[1] = x.y; cw[2] = x.z; cw[3] = x.w;
                } else {  // no bitmap: every member haplotype is called; padding bits are not members
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t first = (col * 4 + j) * 32;
                        cw[j] = cap >= first + 32 ? FULL : (cap > first ? (1u << (cap - first)) - 1u : 0u);
                    }
                }
                uint32_t pw[NB][4];
#pragma unroll
                for (int k = 0; k < NB; ++k) {
                    const uint4 x = planes[(size_t)k * plane_stride_u4 + o];
                    pw[k][0] = x.x; pw[k][1] = x.y; pw[k][2] = x.z; pw[k][3] = x.w;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    n_called += __popc(cw[j]);
#pragma unroll
                    for (uint32_t a = 0; a < A; ++a) {
                        uint32_t m = cw[j];
#pragma unroll
                        for (int k = 0; k < NB; ++k) m &= ((a >> k) & 1u) ? pw[k][j] : ~pw[k][j];
                        c[a] += __popc(m);
                    }
                }
            }
        }
        for (uint32_t off = lps >> 1; off > 0; off >>= 1) {
            n_called += __shfl_xor_sync(FULL, n_called, off);
#pragma unroll
            for (uint32_t a = 0; a < A; ++a) c[a] += __shfl_xor_sync(FULL, c[a], off);
        }
        if (v < v_hi && phase == 0) {
            cnt[v] = n_called;
#pragma unroll
            for (uint32_t a = 0; a < A; ++a) acount[(size_t)v * A + a] = c[a];
        }
    }
}

struct MultiSite {
    uint32_t n;        // called haplotypes
    uint32_t distinct; // alleles with a non-zero count
    double ssq;        // sum of squared counts (exact)
};
__device__ __forceinline__ MultiSite fm_multi_site(const uint32_t *__restrict__ ac, uint32_t A, uint32_t n) {
    MultiSite s{n, 0u, 0.0};
    for (uint32_t a = 0; a < A; ++a) {
        const uint32_t c = ac[a];
        if (c) {
            const double cd = (double)c;
            s.ssq += cd * cd;
            ++s.distinct;
        }
    }
    return s;
}
// general dense per-site pi (stats.rs:3081-3094, 4576-4583): n/(n-1) * (1 - ssq / (n*n))
__device__ __forceinline__ bool fm_pi_general_dense(const MultiSite &s, double &out) {
    if (s.n < 2) return false;
    const double n = (double)s.n;
    const double sum_p2 = s.ssq / (n * n);
    out = n / (n - 1.0) * (1.0 - sum_p2);
    return true;
}
// pi_from_components (stats.rs:2723-2733)
__device__ __forceinline__ bool fm_pi_general_components(const MultiSite &s, double &out) {
    if (s.n < 2) return false;
    const double n = (double)s.n;
    const double inv_n = 1.0 / n;
    const double sum_p2 = s.ssq * inv_n * inv_n;
    out = n / (n - 1.0) * (1.0 - sum_p2);
    return true;
}
#define FM_MULTI_DENSE 0  /* general dense forms  */
#define FM_MULTI_SPARSE 1 /* sparse (pi_from_components) forms */

// Diversity from cached per-allele counts: per-batch partials (sum pi, segregating sites, sites
// with called < 2) and optional per-site pi / theta tracks (calculate_per_site_diversity,
// stats.rs:4693-4750, sparse forms).  One warp per batch, same partial shape as the biallelic path.
__global__ void __launch_bounds__(256)
fm_k_multi_div_from_counts(const uint32_t *__restrict__ acount, const uint32_t *__restrict__ cnt, uint32_t A,
                           int form, DivEpilogue e, uint32_t v_lo, uint32_t v_hi, uint32_t b_lo,
                           uint32_t n_batches) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t GW = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t bi = gw; bi < n_batches; bi += GW) {
        const uint32_t v = (b_lo + bi) * 32 + lane;
        const bool valid = v >= v_lo && v < v_hi;
        double pi_part = 0.0;
        uint32_t seg = 0, unc = 0;
        const uint32_t flags = e.site_flags ? e.site_flags[bi] : 0u;
        if (valid) {
            const MultiSite s = fm_multi_site(acount + (size_t)v * A, A, cnt[v]);
            seg = s.distinct > 1;       // stats.rs:3904-3930 (dense) / 3868-3889 (sparse)
            unc = s.n < 2;
            double val = 0.0;
            const bool has = form == FM_MULTI_DENSE ? fm_pi_general_dense(s, val) : fm_pi_general_components(s, val);
            if (has) pi_part = val;
            if (e.pi_out) {
                double pi_value, theta_value;
                if (s.n < 2 || ((flags >> lane) & 1u)) {
                    pi_value = fm_nan();
                    theta_value = fm_nan();
                } else {
                    theta_value = s.distinct > 1 ? __ldg(e.tab_theta + s.n) : 0.0;
                    double pv = 0.0;
                    pi_value = fm_pi_general_components(s, pv) ? pv : 0.0;
                }
                e.pi_out[v - v_lo] = pi_value;
                e.theta_out[v - v_lo] = theta_value;
            }
        }
        const double s_pi = fm_warp_sum(pi_part);
        const uint32_t s_seg = fm_warp_sum_u(seg), s_unc = fm_warp_sum_u(unc);
        if (lane == 0) {
            e.part_pi[bi] = s_pi;
            e.part_u[2 * bi] = s_seg;
            e.part_u[2 * bi + 1] = s_unc;
        }
    }
}

// One site of the general (multi-allelic) Hudson estimator from two groups' per-allele counts
// (stats.rs:2557-2591 / 3106-3140 / 2907-2935): accumulator contributions + the per-site values.
__device__ __forceinline__ void fm_multi_hudson_site(const uint32_t *__restrict__ c1, uint32_t n1,
                                                     const uint32_t *__restrict__ c2, uint32_t n2, uint32_t A,
                                                     int form, HudsonAcc &acc, fm_hudson_vals &o, MultiSite &s1,
                                                     MultiSite &s2) {
    s1 = fm_multi_site(c1, A, n1);
    s2 = fm_multi_site(c2, A, n2);
    double p1 = 0.0, p2 = 0.0, d = 0.0;
    const bool has1 = form == FM_MULTI_DENSE ? fm_pi_general_dense(s1, p1) : fm_pi_general_components(s1, p1);
    const bool has2 = form == FM_MULTI_DENSE ? fm_pi_general_dense(s2, p2) : fm_pi_general_components(s2, p2);
    const bool has_d = s1.n != 0 && s2.n != 0;
    if (has_d) {
        const double inv1 = 1.0 / (double)s1.n, inv2 = 1.0 / (double)s2.n;
        double dot = 0.0;
        for (uint32_t a = 0; a < A; ++a)
            if (c1[a] && c2[a]) dot += ((double)c1[a] * inv1) * ((double)c2[a] * inv2);
        d = 1.0 - dot;
        d = d > 0.0 ? d : 0.0;
        d = d < 1.0 ? d : 1.0;
    }
    fm_hudson_components(has_d, d, has1, p1, has2, p2, o);
    acc.unc1 = s1.n < 2;
    acc.unc2 = s2.n < 2;
    if (o.num == o.num && o.den == o.den) {
        acc.num = o.num;
        acc.den = o.den;
    }
    if (has_d)
        acc.dxy = d;
    else
        acc.skipped = 1;
    if (has1) acc.pi1 = p1;
    if (has2) acc.pi2 = p2;
}

// Hudson per-site values and regional partials from two groups' cached per-allele counts.
__global__ void __launch_bounds__(256)
fm_k_multi_hudson_from_counts(const uint32_t *__restrict__ ac1, const uint32_t *__restrict__ n1v,
                              const uint32_t *__restrict__ ac2, const uint32_t *__restrict__ n2v, uint32_t A,
                              int form, HudsonEpilogue e, uint32_t v_lo, uint32_t v_hi, uint32_t b_lo,
                              uint32_t n_batches) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t GW = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t bi = gw; bi < n_batches; bi += GW) {
        const uint32_t v = (b_lo + bi) * 32 + lane;
        const bool valid = v >= v_lo && v < v_hi;
        HudsonAcc acc{0.0, 0.0, 0.0, 0.0, 0.0, 0u, 0u, 0u};
        if (valid) {
            MultiSite s1, s2;
            fm_hudson_vals o;
            fm_multi_hudson_site(ac1 + (size_t)v * A, n1v[v], ac2 + (size_t)v * A, n2v[v], A, form, acc, o, s1, s2);
            if (e.fst) {
                const uint32_t i = v - v_lo;
                e.fst[i] = o.fst;
                e.dxy[i] = o.dxy;
                e.pi1[i] = o.pi1;
                e.pi2[i] = o.pi2;
                e.num[i] = o.num;
                e.den[i] = o.den;
                e.n1_out[i] = s1.n;
                e.n2_out[i] = s2.n;
            }
        }
        const double r0 = fm_warp_sum(acc.num), r1 = fm_warp_sum(acc.den), r2 = fm_warp_sum(acc.dxy),
                     r3 = fm_warp_sum(acc.pi1), r4 = fm_warp_sum(acc.pi2);
        const uint32_t u0 = fm_warp_sum_u(acc.skipped), u1 = fm_warp_sum_u(acc.unc1), u2 = fm_warp_sum_u(acc.unc2);
        if (lane == 0) {
            double *pd = e.part_d + (size_t)bi * 5;
            pd[0] = r0; pd[1] = r1; pd[2] = r2; pd[3] = r3; pd[4] = r4;
            uint32_t *pu = e.part_u + (size_t)bi * 3;
            pu[0] = u0; pu[1] = u1; pu[2] = u2;
        }
    }
}

// Window totals (one warp per window [lo, hi) of site indices) for multi-allelic groups, from the cached
// per-allele counts: the same per-site forms as the region calls above, fixed-shape butterfly per window.
__global__ void __launch_bounds__(256)
fm_k_window_div_multi(const uint32_t *__restrict__ acount, const uint32_t *__restrict__ cnt, uint32_t A, int form,
                      const uint32_t *__restrict__ win_lo, const uint32_t *__restrict__ win_hi, uint32_t n_windows,
                      uint64_t *__restrict__ seg_out, double *__restrict__ pi_out, uint64_t *__restrict__ unc_out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t GW = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w = gw; w < n_windows; w += GW) {
        double pi = 0.0;
        uint32_t seg = 0, unc = 0;
        for (uint32_t v = win_lo[w] + lane; v < win_hi[w]; v += 32) {
            const MultiSite s = fm_multi_site(acount + (size_t)v * A, A, cnt[v]);
            seg += s.distinct > 1;
            unc += s.n < 2;
            double val = 0.0;
            if (form == FM_MULTI_DENSE ? fm_pi_general_dense(s, val) : fm_pi_general_components(s, val)) pi += val;
        }
        pi = fm_warp_sum(pi);
        seg = fm_warp_sum_u(seg);
        unc = fm_warp_sum_u(unc);
        if (lane == 0) {
            seg_out[w] = seg;
            pi_out[w] = pi;
            unc_out[w] = unc;
        }
    }
}

__global__ void __launch_bounds__(256)
fm_k_window_hudson_multi(const uint32_t *__restrict__ ac1, const uint32_t *__restrict__ n1v,
                         const uint32_t *__restrict__ ac2, const uint32_t *__restrict__ n2v, uint32_t A, int form,
                         const uint32_t *__restrict__ win_lo, const uint32_t *__restrict__ win_hi, uint32_t n_windows,
                         double *__restrict__ out_d /*[n][5]*/, uint64_t *__restrict__ out_skipped) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t GW = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w = gw; w < n_windows; w += GW) {
        double s[5] = {0, 0, 0, 0, 0};
        uint32_t skipped = 0;
        for (uint32_t v = win_lo[w] + lane; v < win_hi[w]; v += 32) {
            HudsonAcc acc{0.0, 0.0, 0.0, 0.0, 0.0, 0u, 0u, 0u};
            MultiSite s1, s2;
            fm_hudson_vals o;
            fm_multi_hudson_site(ac1 + (size_t)v * A, n1v[v], ac2 + (size_t)v * A, n2v[v], A, form, acc, o, s1, s2);
            s[0] += acc.num; s[1] += acc.den; s[2] += acc.dxy; s[3] += acc.pi1; s[4] += acc.pi2;
            skipped += acc.skipped;
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) s[i] = fm_warp_sum(s[i]);
        skipped = fm_warp_sum_u(skipped);
        if (lane == 0) {
            for (int i = 0; i < 5; ++i) out_d[(size_t)w * 5 + i] = s[i];
            out_skipped[w] = skipped;
        }
    }
}

}  // namespace fm
