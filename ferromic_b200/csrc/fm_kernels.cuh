// fm_kernels.cuh -- sm_100a kernels of the per-site estimator path.
//
//   K1  fm_k_repack           u8 matrix (+ missing bitmap)  ->  per-group bitplanes
//   K2  fm_k_plane_pass<1>    one group's planes  -> alt/called counts, pi/theta tracks, S, sum pi
//   K3  fm_k_plane_pass<2>    two groups' planes  -> both counts + fused Hudson components
//   K2'/K3' light kernels evaluating estimators from cached count arrays
//   K5  fm_k_window_*         segmented (window) reductions
//   fm_k_reduce_partials      deterministic second-level reduction of per-batch partials
//
// The plane pass is a persistent kernel: every warp owns a private ring of shared-memory
// stages filled by 1-D TMA bulk copies (cp.async.bulk + mbarrier) and consumes them with
// 128-bit LDS + __popc; per-site counts are transposed with warp shuffles so that the FP64
// epilogue runs with one site per lane (32 sites = one "batch").  All reductions have a fixed
// shape keyed by the global batch index, so results do not depend on grid size or GPU count.
#pragma once
#include "fm_device.cuh"

namespace fm {

constexpr int kWarpsPerCta = 8;
constexpr int kStages = 2;
constexpr uint32_t kStageBytes = 12 * 1024;  // per warp, per stage
constexpr uint32_t kSuperBatches = 256;      // batches folded by fm_k_reduce_partials

// ------------------------------------------------------------------------------ K1 repack
// One warp builds 32-bit words with __ballot_sync: lane j handles haplotype k = 32*w + j of
// the group (offset table `off`, sorted/de-duplicated exactly like DenseMembership::build,
// stats.rs:1251-1284).  allele bit = (byte != 0) & called, called bit = !missing (or k < n
// when the matrix has no bitmap).  Words are written 32 at a time (one per lane, 128 B).
// Plane row = wq uint4 = 4*wq words; padding bits are zero.
__global__ void __launch_bounds__(256)
fm_k_repack(const uint8_t *__restrict__ data, const uint64_t *__restrict__ missing, size_t stride,
            const uint32_t *__restrict__ off, uint32_t n, uint32_t wq, uint32_t v_lo, uint32_t v_hi,
            uint32_t *__restrict__ allele, uint32_t *__restrict__ called) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t words = wq * 4;
    const uint32_t wgroups = (words + 31) / 32;  // groups of 32 words (1024 haplotypes)
    const uint64_t total = (uint64_t)(v_hi - v_lo) * wgroups;
    for (uint64_t item = warp; item < total; item += nwarps) {
        const uint32_t v = v_lo + (uint32_t)(item / wgroups);
        const uint32_t wg = (uint32_t)(item % wgroups);
        const size_t base = (size_t)v * stride;
        uint32_t my_a = 0, my_c = 0;
#pragma unroll 4
        for (uint32_t i = 0; i < 32; ++i) {
            const uint32_t w = wg * 32 + i;
            if (w >= words) break;  // warp-uniform
            const uint32_t k = w * 32 + lane;
            bool a = false, c = false;
            if (k < n) {
                const size_t idx = base + off[k];
                c = true;
                if (missing) c = !((missing[idx >> 6] >> (idx & 63)) & 1ull);
                a = c && (data[idx] != 0);
            }
            const uint32_t wa = __ballot_sync(0xffffffffu, a);
            const uint32_t wc = __ballot_sync(0xffffffffu, c);
            if (i == lane) {
                my_a = wa;
                my_c = wc;
            }
        }
        const uint32_t w = wg * 32 + lane;
        if (w < words) {
            allele[(size_t)v * words + w] = my_a;
            if (called) called[(size_t)v * words + w] = my_c;
        }
    }
}

// ------------------------------------------------------------------------------ plane pass
struct GroupPlanes {
    const uint4 *allele;
    const uint4 *called;  // nullptr => every member haplotype is called at every site
    uint32_t wq;          // uint4 per row
    uint32_t cap;         // haplotype capacity (offsets.len())
};

struct PassGeom {
    uint32_t lps;       // lanes per site: 1,2,4,...,32
    uint32_t n_chunks;  // column chunks per row (1 unless lps == 32)
    uint32_t cq;        // chunk width in uint4 (chunked mode)
    uint32_t v_lo, v_hi;
    uint32_t b_lo, n_batches;  // global batch range (batch b = sites [32b, 32b+32))
    uint32_t n_sites_total;    // V (rows available in the planes)
};

// Diversity epilogue (NG == 1): build_dense_population_summary (stats.rs:1367-1470) +
// calculate_per_site_diversity (stats.rs:4693-4750) fused.
struct DivEpilogue {
    uint32_t *alt_out, *called_out;  // [V] or nullptr
    double *pi_out, *theta_out;      // [v_hi - v_lo] or nullptr   (tracks)
    const int64_t *pos;              // [V]
    const int64_t *mask;             // merged, sorted half-open intervals [s,e) (2*n_mask) or nullptr
    const int64_t *filt;             // sorted filtered positions or nullptr
    const double *harmonic;          // H[k], k = 0..cap (forward summation, stats.rs:4234-4240)
    uint32_t n_mask, n_filt;
    int pi_form;                     // formula used for the sum-of-pi partial
    double *part_pi;                 // [n_batches]
    uint32_t *part_u;                // [n_batches][2]: segregating sites, sites with called < 2
};

// Hudson epilogue (NG == 2).
struct HudsonEpilogue {
    uint32_t *alt_out[2], *called_out[2];  // cached count arrays or nullptr
    int variant;                           // FM_HV_* (per-site form) or -1 for the summaries form
    double *fst, *dxy, *pi1, *pi2, *num, *den;  // per-site [v_hi - v_lo] or nullptr
    uint32_t *n1_out, *n2_out;
    // per batch: 0 num, 1 den, 2 dxy_sum, 3 pi1_sum, 4 pi2_sum
    double *part_d;   // [n_batches][5]
    uint32_t *part_u; // [n_batches][3]: dxy_skipped, unc1 (n1<2), unc2 (n2<2)
};

__device__ __forceinline__ bool fm_in_intervals(const int64_t *iv, uint32_t n, int64_t pos) {
    // merged + sorted half-open intervals: find last start <= pos
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (iv[2 * mid] <= pos)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo > 0 && pos < iv[2 * (lo - 1) + 1];
}
__device__ __forceinline__ bool fm_in_sorted(const int64_t *a, uint32_t n, int64_t x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < x)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo < n && a[lo] == x;
}

__device__ __forceinline__ double fm_warp_sum(double v) {
    // fixed-shape butterfly: identical association for every batch
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint32_t fm_warp_sum_u(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Per-site diversity values (shared by the fused epilogue and the light kernel).
__device__ __forceinline__ void fm_div_site(const DivEpilogue &e, uint32_t v, uint32_t v_lo,
                                            uint32_t n, uint32_t alt, double &pi_part,
                                            uint32_t &seg, uint32_t &unc) {
    seg = (n >= 2 && alt > 0 && alt < n) ? 1u : 0u;  // stats.rs:1389 / 1406
    unc = (n < 2) ? 1u : 0u;                         // stats.rs:1512-1516
    double val;
    pi_part = fm_pi_form(e.pi_form, n, alt, val) ? val : 0.0;
    if (e.alt_out) e.alt_out[v] = alt;
    if (e.called_out) e.called_out[v] = n;
    if (e.pi_out) {
        // calculate_per_site_diversity, stats.rs:4710-4743
        double pi_value, theta_value;
        if (n < 2) {
            pi_value = fm_nan();
            theta_value = fm_nan();
        } else {
            if (alt > 0 && alt < n) {  // distinct_alleles > 1
                double denom = e.harmonic[n - 1];
                theta_value = denom > 0.0 ? 1.0 / denom : 0.0;
            } else {
                theta_value = 0.0;
            }
            double p;
            pi_value = fm_pi_components(n, alt, p) ? p : 0.0;
        }
        const int64_t pos = e.pos[v];
        bool drop = false;
        if (e.filt) drop = fm_in_sorted(e.filt, e.n_filt, pos);
        if (!drop && e.mask) drop = fm_in_intervals(e.mask, e.n_mask, pos);
        if (drop) {
            pi_value = fm_nan();
            theta_value = fm_nan();
        }
        e.pi_out[v - v_lo] = pi_value;
        e.theta_out[v - v_lo] = theta_value;
    }
}

struct HudsonAcc {
    double num, den, dxy, pi1, pi2;
    uint32_t skipped, unc1, unc2;
};

// Per-site Hudson contributions. variant >= 0: per-site form (FM_HV_*), sums follow
// hudson_component_sums (stats.rs:1625-1635) + calculate_pi_dense / calculate_dxy_dense;
// variant < 0: aggregate_hudson_components_from_summaries (stats.rs:1554-1623).
__device__ __forceinline__ void fm_hudson_contrib(const HudsonEpilogue &e, uint32_t v, uint32_t v_lo,
                                                  uint32_t n1, uint32_t a1, uint32_t n2, uint32_t a2,
                                                  HudsonAcc &acc) {
    acc = HudsonAcc{0.0, 0.0, 0.0, 0.0, 0.0, 0u, 0u, 0u};
    acc.unc1 = n1 < 2;
    acc.unc2 = n2 < 2;
    if (e.variant < 0) {
        double dxy;
        if (!fm_dxy_summaries(n1, a1, n2, a2, dxy)) {
            acc.skipped = 1;
            return;
        }
        acc.dxy = dxy;
        if (n1 < 2 || n2 < 2) return;
        double p1 = fm_pi_summaries(n1, a1), p2 = fm_pi_summaries(n2, a2);
        acc.pi1 = p1;
        acc.pi2 = p2;
        if (dxy > FM_FST_EPSILON) {
            acc.num = dxy - 0.5 * (p1 + p2);
            acc.den = dxy;
        }
        return;
    }
    fm_hudson_vals o;
    fm_hudson_site(e.variant, n1, a1, n2, a2, o);
    if (o.num == o.num && o.den == o.den) {  // both Some
        acc.num = o.num;
        acc.den = o.den;
    }
    // regional Dxy via calculate_dxy_dense / sparse fold (dot form), skipped when a pop is empty
    double d;
    if (fm_dxy_dot(n1, a1, n2, a2, d))
        acc.dxy = d;
    else
        acc.skipped = 1;
    // regional pi via calculate_pi_dense(_biallelic) or calculate_pi (same form as the site pi)
    if (o.pi1 == o.pi1) acc.pi1 = o.pi1;
    if (o.pi2 == o.pi2) acc.pi2 = o.pi2;
    if (e.fst) {
        const uint32_t i = v - v_lo;
        e.fst[i] = o.fst;
        e.dxy[i] = o.dxy;
        e.pi1[i] = o.pi1;
        e.pi2[i] = o.pi2;
        e.num[i] = o.num;
        e.den[i] = o.den;
        e.n1_out[i] = n1;
        e.n2_out[i] = n2;
    }
}

template <int NG>
struct PassParams {
    GroupPlanes g[NG];
    PassGeom geom;
    DivEpilogue div;     // used when NG == 1
    HudsonEpilogue hud;  // used when NG == 2
};

template <int NG>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 1)
fm_k_plane_pass(const __grid_constant__ PassParams<NG> P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[kWarpsPerCta * kStages];

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * kWarpsPerCta + warp;
    const uint32_t GW = gridDim.x * kWarpsPerCta;
    const PassGeom &G = P.geom;

    uint8_t *my_smem = smem_raw + (size_t)warp * kStages * kStageBytes;
    uint64_t *my_bar = bars + warp * kStages;
    if (lane == 0) {
        for (int s = 0; s < kStages; ++s) fm_mbar_init(&my_bar[s], 1);
        fm_fence_mbar_init();
    }
    __syncwarp();

    const bool chunked = (G.lps == 32);
    const uint32_t lps = G.lps;
    const uint32_t sps = 32 / lps;  // sites per step (non-chunked)
    const uint32_t steps_per_batch = chunked ? 32 * G.n_chunks : lps;
    // batches handled by this warp: b = b_lo + gw, + GW, ...
    const uint32_t my_batches = (G.n_batches > gw) ? (G.n_batches - gw + GW - 1) / GW : 0;
    const uint64_t my_steps = (uint64_t)my_batches * steps_per_batch;

    // geometry of step q -> (first site, #sites, column range)
    auto step_geom = [&](uint64_t q, uint32_t &v0, uint32_t &nsites, uint32_t &c0) {
        const uint32_t bi = (uint32_t)(q / steps_per_batch);
        const uint32_t k = (uint32_t)(q % steps_per_batch);
        const uint32_t b = G.b_lo + gw + bi * GW;
        if (chunked) {
            v0 = b * 32 + k / G.n_chunks;
            c0 = (k % G.n_chunks) * G.cq;
            nsites = (v0 < G.n_sites_total) ? 1u : 0u;
        } else {
            v0 = b * 32 + k * sps;
            c0 = 0;
            nsites = (v0 < G.n_sites_total) ? min(sps, G.n_sites_total - v0) : 0u;
        }
    };
    // columns of group g present in a step starting at column c0
    auto cols_of = [&](int g, uint32_t c0) -> uint32_t {
        if (!chunked) return P.g[g].wq;
        return (c0 < P.g[g].wq) ? min(G.cq, P.g[g].wq - c0) : 0u;
    };

    auto issue = [&](uint64_t q) {  // lane 0 only
        uint32_t v0, nsites, c0;
        step_geom(q, v0, nsites, c0);
        const uint32_t stage = (uint32_t)(q % kStages);
        uint8_t *dst = my_smem + (size_t)stage * kStageBytes;
        uint32_t total = 0;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const uint32_t bytes = nsites * cols_of(g, c0) * 16u;
            total += bytes * (P.g[g].called ? 2u : 1u);
        }
        if (total == 0) {
            // nothing to load (batch tail beyond V): complete the phase with a plain arrive
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fm_smem_u32(&my_bar[stage]))
                         : "memory");
            return;
        }
        fm_mbar_expect_tx(&my_bar[stage], total);
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const uint32_t cols = cols_of(g, c0);
            const uint32_t bytes = nsites * cols * 16u;
            const uint32_t slot = (chunked ? G.cq : P.g[g].wq) * sps * 16u;  // smem bytes per plane
            if (bytes) {
                const size_t src = ((size_t)v0 * P.g[g].wq + c0);
                fm_bulk_g2s(dst, P.g[g].allele + src, bytes, &my_bar[stage]);
                if (P.g[g].called)
                    fm_bulk_g2s(dst + slot, P.g[g].called + src, bytes, &my_bar[stage]);
            }
            dst += (size_t)slot * (P.g[g].called ? 2u : 1u);
        }
    };

    if (lane == 0) {
        for (uint64_t q = 0; q < (uint64_t)kStages && q < my_steps; ++q) issue(q);
    }

    // per-lane accumulators for the site this lane is working on (alt, called) per group,
    // and the batch-transposed counts (lane i <-> site 32b+i)
    uint32_t acc_a[NG], acc_c[NG], site_a[NG], site_c[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) acc_a[g] = acc_c[g] = site_a[g] = site_c[g] = 0;

    const uint32_t slot_in_step = lane / lps;  // site slot within a step (non-chunked)
    const uint32_t phase = lane % lps;

    for (uint64_t q = 0; q < my_steps; ++q) {
        const uint32_t stage = (uint32_t)(q % kStages);
        const uint32_t parity = (uint32_t)((q / kStages) & 1);
        uint32_t v0, nsites, c0;
        step_geom(q, v0, nsites, c0);
        const uint32_t k = (uint32_t)(q % steps_per_batch);
        fm_mbar_wait(&my_bar[stage], parity);

        const uint8_t *src = my_smem + (size_t)stage * kStageBytes;
        if (!chunked) {
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const uint32_t wq = P.g[g].wq;
                const uint4 *sa = reinterpret_cast<const uint4 *>(src);
                const uint4 *sc = sa + (size_t)wq * sps;
                uint32_t a = 0, c = 0;
                if (slot_in_step < nsites) {
                    // bank-conflict-free rotation: make (row*wq + col) mod 8 distinct across the
                    // 8 lanes of every quarter-warp (see DESIGN.md "shared-memory access")
                    const uint32_t row = slot_in_step;
                    uint32_t rot = ((lane & 7u) / lps * lps + 8u * wq - (row * wq) % 8u) % 8u;
                    if (lps >= 8) rot = 0;
                    rot %= wq;
                    const uint4 *ra = sa + (size_t)row * wq;
                    const uint4 *rc = sc + (size_t)row * wq;
                    if (P.g[g].called) {
#pragma unroll 4
                        for (uint32_t u = phase; u < wq; u += lps) {
                            uint32_t col = u + rot;
                            if (col >= wq) col -= wq;
                            a += fm_popc4(ra[col]);
                            c += fm_popc4(rc[col]);
                        }
                    } else {
#pragma unroll 4
                        for (uint32_t u = phase; u < wq; u += lps) {
                            uint32_t col = u + rot;
                            if (col >= wq) col -= wq;
                            a += fm_popc4(ra[col]);
                        }
                    }
                }
                // reduce across the lps lanes of a site
                for (uint32_t o = lps >> 1; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    c += __shfl_xor_sync(0xffffffffu, c, o);
                }
                // transpose: batch lane i = k*sps + slot takes slot's totals
                const uint32_t from = (lane % sps) * lps;
                const uint32_t ta = __shfl_sync(0xffffffffu, a, from);
                const uint32_t tc = __shfl_sync(0xffffffffu, c, from);
                if (lane / sps == k) {
                    site_a[g] = ta;
                    site_c[g] = P.g[g].called ? tc : P.g[g].cap;
                }
                src += (size_t)wq * sps * 16u * (P.g[g].called ? 2u : 1u);
            }
        } else {
            const uint32_t site_in_batch = k / G.n_chunks;
            const bool last_chunk = (k % G.n_chunks) == G.n_chunks - 1;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const uint32_t cols = cols_of(g, c0);
                const uint4 *sa = reinterpret_cast<const uint4 *>(src);
                const uint4 *sc = sa + G.cq;
                if (nsites) {
                    if (P.g[g].called) {
                        for (uint32_t u = lane; u < cols; u += 32) {
                            acc_a[g] += fm_popc4(sa[u]);
                            acc_c[g] += fm_popc4(sc[u]);
                        }
                    } else {
                        for (uint32_t u = lane; u < cols; u += 32) acc_a[g] += fm_popc4(sa[u]);
                    }
                }
                if (last_chunk) {
                    const uint32_t ta = fm_warp_sum_u(acc_a[g]);
                    const uint32_t tc = fm_warp_sum_u(acc_c[g]);
                    if (lane == site_in_batch) {
                        site_a[g] = ta;
                        site_c[g] = P.g[g].called ? tc : P.g[g].cap;
                    }
                    acc_a[g] = 0;
                    acc_c[g] = 0;
                }
                src += (size_t)G.cq * 16u * (P.g[g].called ? 2u : 1u);
            }
        }
        __syncwarp();  // all lanes finished reading this stage
        if (lane == 0 && q + kStages < my_steps) issue(q + kStages);

        if (k == steps_per_batch - 1) {
            // ---- batch epilogue: lane i <-> site 32b + i
            const uint32_t bi = (uint32_t)(q / steps_per_batch);
            const uint32_t b = G.b_lo + gw + bi * GW;
            const uint32_t v = b * 32 + lane;
            const bool valid = (v >= G.v_lo) && (v < G.v_hi);
            if constexpr (NG == 1) {
                double pi_part = 0.0;
                uint32_t seg = 0, unc = 0;
                if (valid) fm_div_site(P.div, v, G.v_lo, site_c[0], site_a[0], pi_part, seg, unc);
                const double s_pi = fm_warp_sum(pi_part);
                const uint32_t s_seg = fm_warp_sum_u(seg), s_unc = fm_warp_sum_u(unc);
                if (lane == 0) {
                    const uint32_t slot = b - G.b_lo;
                    P.div.part_pi[slot] = s_pi;
                    P.div.part_u[2 * slot] = s_seg;
                    P.div.part_u[2 * slot + 1] = s_unc;
                }
            } else {
                HudsonAcc acc{0.0, 0.0, 0.0, 0.0, 0.0, 0u, 0u, 0u};
                if (valid) {
                    fm_hudson_contrib(P.hud, v, G.v_lo, site_c[0], site_a[0], site_c[1], site_a[1], acc);
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        if (P.hud.alt_out[g]) P.hud.alt_out[g][v] = site_a[g];
                        if (P.hud.called_out[g]) P.hud.called_out[g][v] = site_c[g];
                    }
                }
                const double r0 = fm_warp_sum(acc.num), r1 = fm_warp_sum(acc.den),
                             r2 = fm_warp_sum(acc.dxy), r3 = fm_warp_sum(acc.pi1),
                             r4 = fm_warp_sum(acc.pi2);
                const uint32_t u0 = fm_warp_sum_u(acc.skipped), u1 = fm_warp_sum_u(acc.unc1),
                               u2 = fm_warp_sum_u(acc.unc2);
                if (lane == 0) {
                    const uint32_t slot = b - G.b_lo;
                    double *pd = P.hud.part_d + (size_t)slot * 5;
                    pd[0] = r0; pd[1] = r1; pd[2] = r2; pd[3] = r3; pd[4] = r4;
                    uint32_t *pu = P.hud.part_u + (size_t)slot * 3;
                    pu[0] = u0; pu[1] = u1; pu[2] = u2;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------ light kernels
// Same per-site functions evaluated from cached count arrays (8 B per site and group).
// One warp per batch of 32 sites so that the per-batch partials are bit-identical to the
// fused pass.
__global__ void __launch_bounds__(256)
fm_k_div_from_counts(const uint32_t *__restrict__ alt, const uint32_t *__restrict__ called,
                     DivEpilogue e, uint32_t v_lo, uint32_t v_hi, uint32_t b_lo, uint32_t n_batches) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t GW = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t bi = gw; bi < n_batches; bi += GW) {
        const uint32_t v = (b_lo + bi) * 32 + lane;
        const bool valid = v >= v_lo && v < v_hi;
        double pi_part = 0.0;
        uint32_t seg = 0, unc = 0;
        if (valid) fm_div_site(e, v, v_lo, called[v], alt[v], pi_part, seg, unc);
        const double s_pi = fm_warp_sum(pi_part);
        const uint32_t s_seg = fm_warp_sum_u(seg), s_unc = fm_warp_sum_u(unc);
        if (lane == 0) {
            e.part_pi[bi] = s_pi;
            e.part_u[2 * bi] = s_seg;
            e.part_u[2 * bi + 1] = s_unc;
        }
    }
}

__global__ void __launch_bounds__(256)
fm_k_hudson_from_counts(const uint32_t *__restrict__ alt1, const uint32_t *__restrict__ n1,
                        const uint32_t *__restrict__ alt2, const uint32_t *__restrict__ n2,
                        HudsonEpilogue e, uint32_t v_lo, uint32_t v_hi, uint32_t b_lo,
                        uint32_t n_batches) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t GW = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t bi = gw; bi < n_batches; bi += GW) {
        const uint32_t v = (b_lo + bi) * 32 + lane;
        const bool valid = v >= v_lo && v < v_hi;
        HudsonAcc acc{0.0, 0.0, 0.0, 0.0, 0.0, 0u, 0u, 0u};
        if (valid) fm_hudson_contrib(e, v, v_lo, n1[v], alt1[v], n2[v], alt2[v], acc);
        const double r0 = fm_warp_sum(acc.num), r1 = fm_warp_sum(acc.den), r2 = fm_warp_sum(acc.dxy),
                     r3 = fm_warp_sum(acc.pi1), r4 = fm_warp_sum(acc.pi2);
        const uint32_t u0 = fm_warp_sum_u(acc.skipped), u1 = fm_warp_sum_u(acc.unc1),
                       u2 = fm_warp_sum_u(acc.unc2);
        if (lane == 0) {
            double *pd = e.part_d + (size_t)bi * 5;
            pd[0] = r0; pd[1] = r1; pd[2] = r2; pd[3] = r3; pd[4] = r4;
            uint32_t *pu = e.part_u + (size_t)bi * 3;
            pu[0] = u0; pu[1] = u1; pu[2] = u2;
        }
    }
}

// Second-level reduction: super-batch s sums batches [s*256, s*256+256) of the GLOBAL batch
// grid sequentially (fixed order), for `nd` double columns and `nu` u32 columns.
__global__ void fm_k_reduce_partials(const double *__restrict__ pd, int nd,
                                     const uint32_t *__restrict__ pu, int nu, uint32_t b_lo,
                                     uint32_t n_batches, uint32_t s_lo, uint32_t n_super,
                                     double *__restrict__ out_d, uint64_t *__restrict__ out_u) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ncol = (uint32_t)(nd + nu);
    if (t >= n_super * ncol) return;
    const uint32_t s = t / ncol, col = t % ncol;
    const uint64_t g0 = (uint64_t)(s_lo + s) * kSuperBatches, g1 = g0 + kSuperBatches;
    const uint64_t lo = g0 > b_lo ? g0 : b_lo;
    const uint64_t hi = g1 < (uint64_t)b_lo + n_batches ? g1 : (uint64_t)b_lo + n_batches;
    if (col < (uint32_t)nd) {
        double acc = 0.0;
        for (uint64_t b = lo; b < hi; ++b) acc += pd[(b - b_lo) * nd + col];
        out_d[(size_t)s * nd + col] = acc;
    } else {
        const uint32_t c = col - nd;
        uint64_t acc = 0;
        for (uint64_t b = lo; b < hi; ++b) acc += pu[(b - b_lo) * nu + c];
        out_u[(size_t)s * nu + c] = acc;
    }
}

// ------------------------------------------------------------------------------ K5 windows
// One warp per window [lo, hi) of site indices: lanes stride the window, then a fixed-shape
// butterfly.  Values are evaluated from cached counts.
__global__ void __launch_bounds__(256)
fm_k_window_div(const uint32_t *__restrict__ alt, const uint32_t *__restrict__ called,
                const uint32_t *__restrict__ win_lo, const uint32_t *__restrict__ win_hi,
                uint32_t n_windows, int pi_form, uint64_t *__restrict__ seg_out,
                double *__restrict__ pi_out, uint64_t *__restrict__ unc_out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t GW = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w = gw; w < n_windows; w += GW) {
        double pi = 0.0;
        uint32_t seg = 0, unc = 0;
        for (uint32_t v = win_lo[w] + lane; v < win_hi[w]; v += 32) {
            const uint32_t n = called[v], a = alt[v];
            seg += (n >= 2 && a > 0 && a < n);
            unc += (n < 2);
            double val;
            if (fm_pi_form(pi_form, n, a, val)) pi += val;
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
fm_k_window_hudson(const uint32_t *__restrict__ alt1, const uint32_t *__restrict__ n1,
                   const uint32_t *__restrict__ alt2, const uint32_t *__restrict__ n2,
                   const uint32_t *__restrict__ win_lo, const uint32_t *__restrict__ win_hi,
                   uint32_t n_windows, double *__restrict__ out_d /*[n][5]*/,
                   uint64_t *__restrict__ out_skipped) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t GW = (gridDim.x * blockDim.x) >> 5;
    HudsonEpilogue e{};
    e.variant = -1;
    for (uint32_t w = gw; w < n_windows; w += GW) {
        double s[5] = {0, 0, 0, 0, 0};
        uint32_t skipped = 0;
        for (uint32_t v = win_lo[w] + lane; v < win_hi[w]; v += 32) {
            HudsonAcc acc;
            fm_hudson_contrib(e, v, 0, n1[v], alt1[v], n2[v], alt2[v], acc);
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
