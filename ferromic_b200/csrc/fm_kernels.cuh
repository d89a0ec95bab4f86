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

constexpr int kWarpsPerCta = 16;
constexpr uint32_t kWarpSmemBytes = 14 * 1024;  // private staging ring of one warp (16 x 14 KB = 224 KB)
constexpr int kMaxStages = 6;
constexpr uint32_t kStepBytesTarget = 7 * 1024;  // preferred bytes per pipeline step (two stages per warp)
// Dynamic scheduler: same-address atomics serialise in L2 (~3.6 ns each on B200, which capped the
// pass at one batch per 3.6 ns), so the batch range is split over kSchedCounters counters that
// live in separate 128-byte lines; a CTA starts on its home range and steals from the others.
constexpr uint32_t kSchedCounters = 16;
constexpr uint32_t kSchedStrideWords = 32;
constexpr uint32_t kSuperBatches = 256;          // batches folded by fm_k_reduce_partials
constexpr int kBatchRing = 8;                    // claimed-batch ring per warp (> kMaxStages)

// ------------------------------------------------------------------------------ K1 repack
// One warp builds 32-bit words with __ballot_sync: lane j handles haplotype k = 32*w + j of
// the group (offset table `off`, sorted/de-duplicated exactly like DenseMembership::build,
// stats.rs:1251-1284).  allele bit = (byte != 0) & called, called bit = !missing (or k < n
// when the matrix has no bitmap).  Words are written 32 at a time (one per lane, 128 B).
// Plane row = wq uint4 = 4*wq words; padding bits are zero.  `data` starts at row v_base and
// `missing` at bitmap word word_base, so a staged chunk of rows can be repacked in place.
__global__ void __launch_bounds__(256)
fm_k_repack(const uint8_t *__restrict__ data, const uint64_t *__restrict__ missing, size_t stride,
            const uint32_t *__restrict__ off, uint32_t n, uint32_t wq, uint32_t v_base, uint64_t word_base,
            uint32_t v_lo, uint32_t v_hi, uint32_t *__restrict__ allele, uint32_t *__restrict__ called,
            uint32_t n_bits, size_t plane_stride_words, uint32_t in_band, const uint8_t *__restrict__ lut) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t words = wq * 4;
    const uint32_t wgroups = (words + 31) / 32;  // groups of 32 words (1024 haplotypes)
    const uint64_t total = (uint64_t)(v_hi - v_lo) * wgroups;
    for (uint64_t item = warp; item < total; item += nwarps) {
        const uint32_t v = v_lo + (uint32_t)(item / wgroups);
        const uint32_t wg = (uint32_t)(item % wgroups);
        const size_t base = (size_t)(v - v_base) * stride;  // offset inside `data` (row v_base first)
        const size_t bit_base = (size_t)v * stride;          // bit index in the whole-matrix bitmap
        uint32_t my_a[4] = {0, 0, 0, 0}, my_c = 0;
#pragma unroll 4
        for (uint32_t i = 0; i < 32; ++i) {
            const uint32_t w = wg * 32 + i;
            if (w >= words) break;  // warp-uniform
            const uint32_t k = w * 32 + lane;
            uint32_t byte = 0;
            bool c = false;
            if (k < n) {
                const uint32_t o = off[k];
                c = true;
                if (missing) {
                    const size_t bit = bit_base + o;
                    c = !((missing[(bit >> 6) - word_base] >> (bit & 63)) & 1ull);
                }
                byte = data[base + o];
                if (in_band) c = byte < 0x80u;  // in-band missingness: a negative int8 cell
                if (!c) byte = 0u;
                else if (lut) byte = __ldg(lut + byte);  // allele values above 15: order-preserving dense ranks
            }
            const uint32_t wc = __ballot_sync(0xffffffffu, c);
            if (n_bits == 1) {  // biallelic: any non-zero allele index is the alternate allele
                const uint32_t wa = __ballot_sync(0xffffffffu, byte != 0);
                if (i == lane) my_a[0] = wa;
            } else {            // bit b of the allele index goes to plane b
#pragma unroll
                for (uint32_t b = 0; b < 4; ++b) {  // compile-time indices keep my_a in registers
                    if (b < n_bits) {
                        const uint32_t wa = __ballot_sync(0xffffffffu, (byte >> b) & 1u);
                        if (i == lane) my_a[b] = wa;
                    }
                }
            }
            if (i == lane) my_c = wc;
        }
        const uint32_t w = wg * 32 + lane;
        if (w < words) {
#pragma unroll
            for (uint32_t b = 0; b < 4; ++b)
                if (b < n_bits) allele[b * plane_stride_words + (size_t)v * words + w] = my_a[b];
            if (called) called[(size_t)v * words + w] = my_c;
        }
    }
}

// ------------------------------------------------------------------------------ K1 v2 repack
// Row-staged, multi-group repack: one warp owns one matrix row at a time.  It pulls the row's
// u8 cells (128-bit loads) and the row's slice of the missing bitmap into its shared-memory
// slice once, then builds the bitplane words of EVERY listed group from shared memory with
// __ballot_sync -- the u8 matrix is read exactly once however many groups (a 27-group W&C
// partition, the two orientation groups of a per-site call) are repacked, and the per-lane byte
// gathers hit shared memory instead of paying an L2 round trip each (K1 v1 ran at ~0.4 TB/s).
struct RepackGroup {
    const uint32_t *off;       // sorted, de-duplicated column offsets (DenseMembership, stats.rs:1251-1284)
    uint32_t n, wq, n_bits;
    uint32_t *allele, *called; // called == nullptr when the matrix has no bitmap
    size_t plane_stride_words; // distance between allele bit planes (multi-allelic groups)
    // count-only groups (the subpopulations of a W&C partition): no bitplanes are kept, the
    // per-site alt / called counts -- all K4 ever reads -- are produced straight from the staged row
    uint32_t *alt_out, *cnt_out;
    // biallelic plane groups: the group's plane row is a bit COMPRESS of the full-row bit words by the
    // membership mask (offsets are sorted).  One plan entry (two uint4) per row word the group touches:
    // {row word, member mask, first output bit, mv0} {mv1, mv2, mv3, mv4} -- mv = the precomputed move masks
    // of the parallel-suffix compress (Hacker's Delight 7-4), so a lane extracts its word's members in
    // 5 x 4 logic ops instead of one ballot per 32 haplotypes.
    const uint4 *plan;
    uint32_t n_ent;
    // tail layout (plan path only): the plane row holds the wq FULL 16-byte words of the group; the remaining
    // n % 128 haplotypes go to tw = ceil((n % 128) / 32) u32 words per row in separate arrays (tw == 0: padded rows)
    uint32_t *tail_a, *tail_c;
    uint32_t tw;
};

__device__ __forceinline__ uint32_t fm_compress32(uint32_t x, const uint4 p0, const uint4 p1) {
    uint32_t t;
    x &= p0.y;
    t = x & p0.w; x = (x ^ t) | (t >> 1);
    t = x & p1.x; x = (x ^ t) | (t >> 2);
    t = x & p1.y; x = (x ^ t) | (t >> 4);
    t = x & p1.z; x = (x ^ t) | (t >> 8);
    t = x & p1.w; x = (x ^ t) | (t >> 16);
    return x;
}

// Count-only groups (W&C subpopulations): the row is packed ONCE into full-row allele / called
// bit words in matrix column order (4 cells per lane and step, no gather); group g then owns a
// sparse list of (row word, membership mask) entries, lane = group, and its alt / called counts
// are masked popcounts -- the cost follows the number of row words a group touches (7 for a
// contiguous 190-haplotype population), not the number of haplotypes.
struct CountTable {
    const uint32_t *ent_start;  // [n_groups + 1]
    const uint32_t *ent_word;   // [n_entries] row word index (column >> 5)
    const uint32_t *ent_mask;   // [n_entries] member columns inside that word
    uint32_t *const *alt_out;   // [n_groups] -> [V]
    uint32_t *const *cnt_out;   // [n_groups] -> [V]
    uint32_t n_groups;
};

// Packed rows (the 2-bit ingest format): row v of the cohort as rw = ceil(stride / 32) u32 words of
// allele bits (bit c & 31 of word c >> 5: cell c carries a non-zero allele and is called) and, when the matrix
// has missing data, rw words of called bits.  In this mode the kernel never sees u8 cells: the full-row bit
// words are loaded as they are (coalesced 4-byte loads) and every group is served from them.
struct PackedRows {
    const uint32_t *a;  // [rows][rw] allele bits, row v_base first; nullptr = u8 mode
    const uint32_t *c;  // [rows][rw] called bits or nullptr (every cell called)
    uint32_t rw;
};

// MODE: 0 = general (staged row, every group kind), 1 = direct rows (whole 16-byte rows read straight from global
// memory; groups are served from the full-row bit words only), 2 = packed rows.  Modes 1 and 2 compile without
// the ballot-gather paths (fewer registers, more resident warps).  BIAL: every cell is 0, 1 or (in band) missing,
// so the allele bit of a byte is its bit 0 -- three operations per four cells instead of seven.
template <int MODE, bool BIAL>
__global__ void __launch_bounds__(256, MODE == 0 ? 4 : 6)
fm_k_repack_rows(const uint8_t *__restrict__ data, size_t data_bytes, const uint64_t *__restrict__ missing,
                 size_t stride, uint32_t v_base, uint64_t word_base, uint32_t v_lo, uint32_t v_hi,
                 const RepackGroup *__restrict__ groups, uint32_t n_groups, uint32_t warp_smem_bytes,
                 uint32_t row_buf_bytes, uint32_t bit_buf_bytes, CountTable ct, uint32_t in_band,
                 uint32_t need_row_bits, uint32_t direct_rows_arg, PackedRows pk_arg, const uint8_t *__restrict__ lut) {
    constexpr uint32_t direct_rows = MODE == 1 ? 1u : 0u;
    PackedRows pk = pk_arg;
    if (MODE != 2) pk.a = nullptr;  // lets the compiler drop the packed path
    (void)direct_rows_arg;
    // direct_rows: every group is served from the full-row bit words (compress plans / count tables) and rows are
    // whole 16-byte words, so the u8 row is packed straight from global memory (coalesced 16-byte loads) and is
    // never staged: row_buf_bytes == 0, three times the resident warps per SM
    extern __shared__ __align__(16) uint8_t rp_smem[];
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *rowb = rp_smem + (size_t)warp * warp_smem_bytes;                       // staged row bytes
    uint64_t *bits = reinterpret_cast<uint64_t *>(rowb + row_buf_bytes);             // staged bitmap words
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t GW = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t v = v_lo + gw; v < v_hi; v += GW) {
        // ---- stage the row: aligned 16-byte loads over [a0, a1) covering [row0, row0 + stride)
        const size_t row0 = (size_t)(v - v_base) * stride;
        const size_t a0 = row0 & ~(size_t)15;
        const uint32_t delta = (uint32_t)(row0 - a0);
        const size_t a1 = (row0 + stride + 15) & ~(size_t)15;
        const uint32_t nq = (uint32_t)((a1 - a0) >> 4);
        __syncwarp();
        // asynchronous 16-byte copies straight into shared memory: every chunk of the row is in
        // flight at once (no register staging, no per-iteration load latency)
        for (uint32_t q = lane; !direct_rows && !pk.a && q < nq; q += 32) {
            const size_t at = a0 + ((size_t)q << 4);
            if (at + 16 <= data_bytes) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(fm_smem_u32(rowb + ((size_t)q << 4))),
                             "l"(data + at)
                             : "memory");
            } else {  // last bytes of the buffer: never read past its end
                uint32_t w[4] = {0, 0, 0, 0};
                for (uint32_t b = 0; b < 16 && at + b < data_bytes; ++b) w[b >> 2] |= (uint32_t)data[at + b] << ((b & 3) * 8);
                *reinterpret_cast<uint4 *>(rowb + ((size_t)q << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        const size_t bit0 = (size_t)v * stride;
        const uint64_t w0 = bit0 >> 6;
        if (missing) {
            const uint32_t nw = (uint32_t)(((bit0 + stride + 63) >> 6) - w0);
            for (uint32_t q = lane; q < nw; q += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(fm_smem_u32(bits + q)),
                             "l"(missing + (w0 - word_base + q))
                             : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        // 32-bit views for the inner loops: bit (brel + o) of the staged bitmap slice, byte rb[o]
        const uint32_t brel = (uint32_t)(bit0 - (w0 << 6));
        const uint32_t *bits32 = reinterpret_cast<const uint32_t *>(bits);
        const uint8_t *rb = rowb + delta;
        // ---- full-row bit words (allele != 0, called) in matrix column order; count-only groups take
        //      masked popcounts of them, biallelic plane groups compress them by their membership masks
        const uint32_t rw = (uint32_t)((stride + 31) >> 5);
        uint32_t *arow = reinterpret_cast<uint32_t *>(rowb + row_buf_bytes + bit_buf_bytes);
        uint32_t *crow = arow + rw + 1;
        uint32_t *outb = crow + rw + 1;  // 2 x out_cap words: one group's plane rows while they are assembled
        const uint32_t out_cap = (uint32_t)((stride + 127) >> 7) * 4u + 4u;  // + one padding word (16-byte granule) per row
        if (pk.a) {  // packed rows: the full-row bit words arrive ready-made
            const size_t r0 = (size_t)(v - v_base) * pk.rw;
            const uint32_t tail = (uint32_t)stride & 31u;
            for (uint32_t w = lane; w < rw; w += 32) {
                uint32_t c = pk.c ? __ldg(pk.c + r0 + w) : FULL;
                if (tail && w == rw - 1) c &= (1u << tail) - 1u;  // bits past the last cell never count
                crow[w] = c;
                arow[w] = __ldg(pk.a + r0 + w) & c;
            }
            __syncwarp();
            for (uint32_t g = lane; g < ct.n_groups; g += 32) {  // count-only groups
                uint32_t a = 0, c = 0;
                const uint32_t e1 = __ldg(ct.ent_start + g + 1);
                for (uint32_t e = __ldg(ct.ent_start + g); e < e1; ++e) {
                    const uint32_t w = __ldg(ct.ent_word + e);
                    const uint32_t cw = crow[w] & __ldg(ct.ent_mask + e);
                    c += __popc(cw);
                    a += __popc(arow[w] & cw);
                }
                ct.alt_out[g][v] = a;
                ct.cnt_out[g][v] = c;
            }
        } else if (need_row_bits) {
            const uint32_t *row32 = reinterpret_cast<const uint32_t *>(rowb + (delta & ~3u));
            const uint32_t sh8 = (delta & 3u) * 8u;
            // rows that start on a 16-byte boundary (stride % 16 == 0, e.g. 5008 or 200000 haplotypes): 16 cells per
            // lane and step (one LDS.128, four nibbles), two lanes per row word; otherwise 4 cells per lane
            const bool wide = direct_rows || (delta & 15u) == 0u;
            if (wide) {
                const uint4 *row128 = direct_rows ? reinterpret_cast<const uint4 *>(data + row0)
                                                  : reinterpret_cast<const uint4 *>(rowb + delta);
                const uint32_t nq16 = ((uint32_t)stride + 15u) >> 4;
                // bit 7 of every non-zero byte (the add only sees 7-bit fields: no carry crosses bytes), gathered by one multiply
                auto nib = [](uint32_t w) {
                    if (BIAL) return ((w & 0x01010101u) * 0x01020408u) >> 24;  // cells are 0 / 1 (missing ones are masked later)
                    return (((((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) >> 7 & 0x01010101u) * 0x01020408u) >> 24;
                };
                auto cnib = [](uint32_t w) { return (((~w >> 7) & 0x01010101u) * 0x01020408u >> 24) & 0xFu; };
                // four 16-byte loads per lane are issued before the first is consumed (direct mode reads global memory)
                for (uint32_t base = 0; base < nq16; base += 128) {
                    uint4 xq[4];
#pragma unroll
                    for (uint32_t u = 0; u < 4; ++u) {
                        const uint32_t q = base + u * 32u + lane;
                        xq[u] = make_uint4(0, 0, 0, 0);
                        if (q < nq16) xq[u] = direct_rows ? __ldg(row128 + q) : row128[q];
                    }
#pragma unroll
                    for (uint32_t u = 0; u < 4; ++u) {
                        const uint32_t q = base + u * 32u + lane;
                        if (base + u * 32u >= nq16) break;  // warp-uniform
                        uint4 x = xq[u];
                        if (q < nq16) {
                            const uint32_t left = (uint32_t)stride - q * 16u;  // cells that exist in this group
                            if (left < 16u) {
                                uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                                for (uint32_t k = 0; k < 4; ++k) {
                                    const uint32_t have = left > 4u * k ? min(4u, left - 4u * k) : 0u;
                                    xs[k] = have >= 4u ? xs[k] : (have ? xs[k] & ((1u << (8u * have)) - 1u) : 0u);
                                }
                                x = make_uint4(xs[0], xs[1], xs[2], xs[3]);
                            }
                        }
                        const uint32_t a16 = nib(x.x) | (nib(x.y) << 4) | (nib(x.z) << 8) | (nib(x.w) << 12);
                        const uint32_t ao = __shfl_xor_sync(FULL, a16, 1);
                        const uint32_t w = q >> 1;
                        if (!(lane & 1u) && w < rw) arow[w] = a16 | (ao << 16);
                        if (in_band) {  // called = cell < 0x80 (a non-negative int8); absent cells read 0 and are masked below
                            const uint32_t c16 = cnib(x.x) | (cnib(x.y) << 4) | (cnib(x.z) << 8) | (cnib(x.w) << 12);
                            const uint32_t co = __shfl_xor_sync(FULL, c16, 1);
                            if (!(lane & 1u) && w < rw) crow[w] = c16 | (co << 16);
                        }
                    }
                }
            }
            for (uint32_t i = 0; !wide && i * 128u < (uint32_t)stride; ++i) {
                const uint32_t col = (i * 32u + lane) * 4u;  // first of this lane's 4 columns
                uint32_t x = 0;
                if (col < (uint32_t)stride) {
                    const uint32_t q = i * 32u + lane;
                    x = __funnelshift_r(row32[q], row32[q + 1], sh8);
                    const uint32_t left = (uint32_t)stride - col;  // columns that exist
                    if (left < 4u) x &= (1u << (8u * left)) - 1u;
                }
                const uint32_t y = __vcmpne4(x, 0u) & 0x01010101u;
                uint32_t wv = (((y * 0x01020408u) >> 24) & 0xFu) << (4u * (lane & 7u));
                wv |= __shfl_xor_sync(FULL, wv, 1);
                wv |= __shfl_xor_sync(FULL, wv, 2);
                wv |= __shfl_xor_sync(FULL, wv, 4);
                const uint32_t w = i * 4u + (lane >> 3);
                if ((lane & 7u) == 0 && w < rw) arow[w] = wv;
                if (in_band) {  // called = cell < 0x80 (a non-negative int8)
                    const uint32_t z = (~x >> 7) & 0x01010101u;
                    uint32_t cv = (((z * 0x01020408u) >> 24) & 0xFu) << (4u * (lane & 7u));
                    cv |= __shfl_xor_sync(FULL, cv, 1);
                    cv |= __shfl_xor_sync(FULL, cv, 2);
                    cv |= __shfl_xor_sync(FULL, cv, 4);
                    if ((lane & 7u) == 0 && w < rw) crow[w] = cv;
                }
            }
            const uint32_t qb = brel >> 5, sb = brel & 31u;
            if (!in_band)
                for (uint32_t w = lane; w < rw; w += 32)
                    crow[w] = missing ? ~__funnelshift_r(bits32[w + qb], bits32[w + qb + 1], sb) : FULL;
            __syncwarp();
            for (uint32_t g = lane; g < ct.n_groups; g += 32) {  // count-only groups
                uint32_t a = 0, c = 0;
                const uint32_t e1 = __ldg(ct.ent_start + g + 1);
                for (uint32_t e = __ldg(ct.ent_start + g); e < e1; ++e) {
                    const uint32_t w = __ldg(ct.ent_word + e);
                    const uint32_t cw = crow[w] & __ldg(ct.ent_mask + e);
                    c += __popc(cw);
                    a += __popc(arow[w] & cw);
                }
                ct.alt_out[g][v] = a;  // dense_sum_alt_with_missing / _no_missing (stats.rs:1665-1697)
                ct.cnt_out[g][v] = c;
            }
        }
        // ---- every plane group's words from the staged row
        for (uint32_t gi = 0; gi < n_groups; ++gi) {
            const RepackGroup G = groups[gi];
            const uint32_t words = G.wq * 4 + G.tw;  // tw != 0 only on the plan path
            if (G.n_bits == 1 && G.plan) {
                uint32_t *oa = outb, *oc = outb + out_cap;
                for (uint32_t i = lane; i < words; i += 32) {
                    oa[i] = 0;
                    oc[i] = 0;
                }
                __syncwarp();
                for (uint32_t e = lane; e < G.n_ent; e += 32) {
                    const uint4 p0 = __ldg(G.plan + 2 * e), p1 = __ldg(G.plan + 2 * e + 1);
                    const uint32_t cw = crow[p0.x];
                    const uint32_t xa = fm_compress32(arow[p0.x] & cw, p0, p1);
                    const uint32_t sh = p0.z & 31u, wi = p0.z >> 5;
                    // straight-line: the fragment's low part and the part that spills into the next word (zero when
                    // sh == 0: the funnel shift yields the high word of (0 : x) << sh).  Nearly every fragment is
                    // non-empty and spills, so the four tests this replaces only cost issue slots and reconvergence
                    // barriers (ncu: 9 % BRA + 7 % BSSY / BSYNC of the kernel's instructions); or-ing a zero is free.
                    // The word after the row's last one is padding (repack_warp_smem).
                    atomicOr(oa + wi, xa << sh);
                    atomicOr(oa + wi + 1, __funnelshift_l(xa, 0u, sh));
                    if (G.called) {
                        const uint32_t xc = fm_compress32(cw, p0, p1);
                        atomicOr(oc + wi, xc << sh);
                        atomicOr(oc + wi + 1, __funnelshift_l(xc, 0u, sh));
                    }
                }
                __syncwarp();
                const uint32_t fw = G.wq * 4;
                for (uint32_t i = lane; i < fw; i += 32) {
                    const size_t o = (size_t)v * fw + i;
                    G.allele[o] = oa[i];
                    if (G.called) G.called[o] = oc[i];
                }
                if (lane < G.tw) {  // the row's last n % 128 haplotypes
                    const size_t o = (size_t)v * G.tw + lane;
                    G.tail_a[o] = oa[fw + lane];
                    if (G.called) G.tail_c[o] = oc[fw + lane];
                }
                __syncwarp();
                continue;
            }
            if (MODE != 0) continue;  // direct / packed launches only carry plan groups (checked on the host)
            if (G.n_bits == 1) {
                // biallelic fast path: words whose 32 haplotypes all exist run a branch-free loop
                // (one offset load, one bit test, one byte test, two ballots per word)
                const uint32_t full_words = G.n >> 5;
                for (uint32_t wb = 0; wb < words; wb += 32) {
                    uint32_t my_a = 0, my_c = 0;
                    const uint32_t lim = min(32u, words - wb);
                    const uint32_t nfull = full_words > wb ? min(lim, full_words - wb) : 0u;
                    const uint32_t *offp = G.off + (size_t)wb * 32 + lane;
#pragma unroll 8
                    for (uint32_t i = 0; i < nfull; ++i) {
                        const uint32_t o = __ldg(offp + i * 32);
                        const uint32_t cell = rb[o];
                        uint32_t cbit = 1u;
                        if (missing) {
                            const uint32_t r = brel + o;
                            cbit = ~(bits32[r >> 5] >> (r & 31u)) & 1u;
                        } else if (in_band) {
                            cbit = cell < 0x80u ? 1u : 0u;
                        }
                        const uint32_t abit = cbit & (cell != 0 ? 1u : 0u);
                        const uint32_t wc = __ballot_sync(FULL, cbit);
                        const uint32_t wa = __ballot_sync(FULL, abit);
                        my_c = (i == lane) ? wc : my_c;
                        my_a = (i == lane) ? wa : my_a;
                    }
                    for (uint32_t i = nfull; i < lim; ++i) {  // the partial word and the zero padding
                        const uint32_t k = (wb + i) * 32 + lane;
                        uint32_t cbit = 0u, abit = 0u;
                        if (k < G.n) {
                            const uint32_t o = __ldg(G.off + k);
                            const uint32_t cell = rb[o];
                            cbit = 1u;
                            if (missing) {
                                const uint32_t r = brel + o;
                                cbit = ~(bits32[r >> 5] >> (r & 31u)) & 1u;
                            } else if (in_band) {
                                cbit = cell < 0x80u ? 1u : 0u;
                            }
                            abit = cbit & (cell != 0 ? 1u : 0u);
                        }
                        const uint32_t wc = __ballot_sync(FULL, cbit);
                        const uint32_t wa = __ballot_sync(FULL, abit);
                        my_c = (i == lane) ? wc : my_c;
                        my_a = (i == lane) ? wa : my_a;
                    }
                    if (lane < lim) {
                        const size_t o = (size_t)v * words + wb + lane;
                        G.allele[o] = my_a;
                        if (G.called) G.called[o] = my_c;
                    }
                }
                continue;
            }
            for (uint32_t wb = 0; wb < words; wb += 32) {  // multi-allelic: bit b of the allele index -> plane b
                uint32_t my_a[4] = {0, 0, 0, 0}, my_c = 0;
                const uint32_t lim = min(32u, words - wb);
                for (uint32_t i = 0; i < lim; ++i) {
                    const uint32_t k = (wb + i) * 32 + lane;
                    uint32_t byte = 0;
                    bool c = false;
                    if (k < G.n) {
                        const uint32_t o = __ldg(G.off + k);
                        c = true;
                        byte = rb[o];
                        if (missing) {
                            const uint32_t r = brel + o;
                            c = !((bits32[r >> 5] >> (r & 31u)) & 1u);
                        } else if (in_band) {
                            c = byte < 0x80u;
                        }
                        if (!c) byte = 0u;
                        else if (lut) byte = __ldg(lut + byte);  // allele values above 15: dense ranks
                    }
                    const uint32_t wc = __ballot_sync(FULL, c);
#pragma unroll
                    for (uint32_t b = 0; b < 4; ++b) {  // compile-time indices keep my_a in registers
                        if (b < G.n_bits) {
                            const uint32_t wa = __ballot_sync(FULL, (byte >> b) & 1u);
                            if (i == lane) my_a[b] = wa;
                        }
                    }
                    if (i == lane) my_c = wc;
                }
                if (lane < lim) {
                    const size_t o = (size_t)v * words + wb + lane;
#pragma unroll
                    for (uint32_t b = 0; b < 4; ++b)
                        if (b < G.n_bits) G.allele[b * G.plane_stride_words + o] = my_a[b];
                    if (G.called) G.called[o] = my_c;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------ sparse missing list -> called plane
// Packed rows may arrive with a sparse list of missing cells instead of a called plane (0.15 instead of 1 bit per
// genotype over PCIe at 1 % missing; 0.09 with the one-byte gap code, ColT = uint8_t).  One warp per row rebuilds the row's called words in shared memory -- all
// ones, the listed cells cleared -- and writes them to the resident packed matrix with coalesced stores.
//   start[r - r_base] .. start[r - r_base + 1]: the row's slice of `cols` (indices relative to cols_base)
template <typename ColT>
__global__ void __launch_bounds__(256)
fm_k_expand_called(const uint64_t *__restrict__ start, const ColT *__restrict__ cols, uint64_t cols_base, uint32_t r_base,
                   uint32_t v_lo, uint32_t v_hi, uint32_t rw, uint32_t stride, uint32_t *__restrict__ cbits) {
    extern __shared__ uint32_t ex_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *crow = ex_smem + (size_t)warp * rw;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t GW = (gridDim.x * blockDim.x) >> 5;
    const uint32_t tail = stride & 31u;
    for (uint32_t v = v_lo + gw; v < v_hi; v += GW) {
        for (uint32_t w = lane; w < rw; w += 32) crow[w] = (tail && w == rw - 1) ? ((1u << tail) - 1u) : 0xffffffffu;
        __syncwarp();
        const uint64_t s0 = start[v - r_base], s1 = start[v - r_base + 1];
        if (sizeof(ColT) == 1) {
            // gap code: byte b < 255 = the missing cell b + 1 columns after the previous position, 255 = move on 255
            // columns without a cell; positions start at -1.  Warp-wide inclusive scan of the steps, 32 bytes at a time.
            uint32_t at = 0;  // position + 1 reached so far
            for (uint64_t i0 = s0; i0 < s1; i0 += 32) {
                const uint64_t i = i0 + lane;
                const uint32_t b = i < s1 ? (uint32_t)cols[i - cols_base] : 255u;
                uint32_t step = i < s1 ? (b == 255u ? 255u : b + 1u) : 0u;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, step, o);
                    if ((int)lane >= o) step += up;
                }
                const uint32_t c = at + step - 1u;  // at + step >= 1 whenever this lane holds a cell
                if (i < s1 && b != 255u && c < stride) atomicAnd(crow + (c >> 5), ~(1u << (c & 31u)));
                at += __shfl_sync(0xffffffffu, step, 31);
            }
        } else {
            for (uint64_t i = s0 + lane; i < s1; i += 32) {
                const uint32_t c = (uint32_t)cols[i - cols_base];
                if (c < stride) atomicAnd(crow + (c >> 5), ~(1u << (c & 31u)));
            }
        }
        __syncwarp();
        uint32_t *dst = cbits + (size_t)v * rw;
        for (uint32_t w = lane; w < rw; w += 32) dst[w] = crow[w];
        __syncwarp();
    }
}

// Which allele values occur in a u8 matrix (called cells only)?  256-bit presence set, for the order-preserving
// remap of matrices whose max_allele exceeds 15 (the bitplanes hold at most 4 bits per cell).
__global__ void __launch_bounds__(256)
fm_k_allele_presence(const uint8_t *__restrict__ data, const uint64_t *__restrict__ missing, uint64_t total,
                     uint32_t in_band, uint32_t *__restrict__ present /*[8]*/) {
    __shared__ uint32_t sp[8];
    if (threadIdx.x < 8) sp[threadIdx.x] = 0;
    __syncthreads();
    uint32_t mine[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = data[i];
        bool called = true;
        if (missing) called = !((missing[i >> 6] >> (i & 63)) & 1ull);
        else if (in_band) called = b < 0x80u;
        if (called) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if ((b >> 5) == (uint32_t)k) mine[k] |= 1u << (b & 31u);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (mine[k]) atomicOr(&sp[k], mine[k]);
    __syncthreads();
    if (threadIdx.x < 8 && sp[threadIdx.x]) atomicOr(&present[threadIdx.x], sp[threadIdx.x]);
}

// ------------------------------------------------------------------------------ synthetic cohorts
// Counter-based generator for benchmarks and full-size parity tests: every matrix entry is a pure
// integer function of (seed, site, column), so any slice can be re-evaluated on the CPU
// (tests/synth.py::synth_rows) without moving the matrix.  Site base frequency is U-shaped
// (x^2 mirrored), each population gets a uniform offset scaled by sigma_q, alleles are Bernoulli.
__host__ __device__ __forceinline__ uint64_t fm_splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t fm_mix(uint64_t seed, uint64_t a, uint64_t b) {
    return fm_splitmix64(fm_splitmix64(seed + a) + b);
}
__host__ __device__ __forceinline__ uint32_t fm_synth_threshold(uint64_t seed, uint64_t v, uint32_t pop,
                                                                uint32_t sigma_q) {
    const uint64_t hs = fm_mix(seed, v, 0);
    const uint32_t x = (uint32_t)(hs & 0xFFFFu);
    uint32_t y = (x * x) >> 16;
    if ((hs >> 16) & 1u) y = 65535u - y;
    const uint64_t hp = fm_mix(seed, v, 1u + pop);
    const int32_t d = ((int32_t)(hp & 0x1FFFu) - 4096) * (int32_t)sigma_q / 4096;
    int32_t t = (int32_t)y + d;
    t = t < 66 ? 66 : (t > 65470 ? 65470 : t);
    return (uint32_t)t;
}

// One thread per bitmap word = 64 consecutive entries of the linear layout.
__global__ void __launch_bounds__(256)
fm_k_synth(uint8_t *__restrict__ data, uint64_t *__restrict__ missing, uint64_t total, uint32_t stride,
           uint32_t ploidy, uint64_t first_variant, uint64_t seed, const uint16_t *__restrict__ pop_of_sample,
           uint32_t sigma_q, uint32_t miss_q) {
    const uint64_t n_words = (total + 63) / 64;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words;
         w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t e0 = w * 64;
        uint64_t v = e0 / stride;
        uint32_t col = (uint32_t)(e0 - v * stride);
        uint64_t bits = 0;
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) packed[i] = 0;
        uint32_t cur_pop = 0xFFFFFFFFu, thr = 0;
        uint64_t cur_v = ~0ull;
#pragma unroll 4
        for (uint32_t i = 0; i < 64; ++i) {
            if (e0 + i < total) {
                const uint32_t pop = pop_of_sample ? pop_of_sample[col / ploidy] : 0u;
                if (pop != cur_pop || v != cur_v) {
                    thr = fm_synth_threshold(seed, first_variant + v, pop, sigma_q);
                    cur_pop = pop;
                    cur_v = v;
                }
                const uint64_t he = fm_mix(seed, first_variant + v, 0x10000ull + col);
                const bool miss = (uint32_t)((he >> 32) & 0xFFFFu) < miss_q;
                const bool allele = (uint32_t)(he & 0xFFFFu) < thr;
                if (miss) bits |= 1ull << i;
                if (allele && !miss) packed[i >> 2] |= 1u << ((i & 3) * 8);
            }
            if (++col == stride) {
                col = 0;
                ++v;
            }
        }
        if (e0 + 64 <= total) {
            uint4 *dst = reinterpret_cast<uint4 *>(data + e0);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        } else {
            for (uint32_t i = 0; e0 + i < total; ++i) data[e0 + i] = (uint8_t)((packed[i >> 2] >> ((i & 3) * 8)) & 0xFFu);
        }
        if (missing) missing[w] = bits;
    }
}

// ------------------------------------------------------------------------------ plane pass
struct GroupPlanes {
    const uint4 *allele;
    const uint4 *called;  // nullptr => every member haplotype is called at every site
    uint32_t wq;          // uint4 per row
    uint32_t cap;         // haplotype capacity (offsets.len())
    // tail layout: the last cap % 128 haplotypes of every row live in tw u32 words per row (tail_a [V][tw], tail_c
    // likewise) instead of a padded 16-byte word; the streaming pass reads them in the batch epilogue (lane = site).
    // Cuts the padding of a 3463 + 1545 haplotype pair from 4.8 % of the plane bytes to 1 %.
    const uint32_t *tail_a, *tail_c;
    uint32_t tw;
};

// tail words of this lane's site, requested at the start of a batch and consumed in its epilogue
struct TailRegs {
    uint32_t a[3], c[3];
};
__device__ __forceinline__ void fm_tail_load(const GroupPlanes &g, uint32_t v, uint32_t n_sites_total, bool hc, TailRegs &t) {
#pragma unroll
    for (uint32_t k = 0; k < 3; ++k) {
        t.a[k] = 0;
        t.c[k] = 0;
    }
    if (g.tw && v < n_sites_total) {
#pragma unroll
        for (uint32_t k = 0; k < 3; ++k)
            if (k < g.tw) {
                t.a[k] = __ldg(g.tail_a + (size_t)v * g.tw + k);
                if (hc) t.c[k] = __ldg(g.tail_c + (size_t)v * g.tw + k);
            }
    }
}
__device__ __forceinline__ void fm_tail_add(const TailRegs &t, uint32_t &alt, uint32_t &cnt, bool hc) {
    alt += __popc(t.a[0]) + __popc(t.a[1]) + __popc(t.a[2]);
    if (hc) cnt += __popc(t.c[0]) + __popc(t.c[1]) + __popc(t.c[2]);
}

struct PassGeom {
    uint32_t lps;         // lanes per site: 1,2,4,...,32 (power of two)
    uint32_t lps_log2;
    uint32_t rounds;      // rounds of (32/lps) sites held by one pipeline step (power of two <= lps)
    uint32_t n_chunks;    // column chunks per row (1 unless lps == 32)
    uint32_t cq;          // chunk width in uint4 (chunked mode)
    uint32_t n_stages;    // pipeline depth of the per-warp ring (2..kMaxStages)
    uint32_t stage_bytes; // bytes reserved per stage (multiple of 128)
    uint32_t warp_smem_bytes;  // private staging ring of one warp (n_stages * stage_bytes <= this)
    uint32_t warps;            // warps per CTA (<= kWarpsPerCta); warps * warp_smem_bytes <= 224 KB
    uint32_t v_lo, v_hi;
    uint32_t b_lo, n_batches;  // global batch range (batch b = sites [32b, 32b+32))
    uint32_t n_sites_total;    // V (rows available in the planes)
    uint32_t *batch_counter;   // dynamic scheduler counters (zeroed before launch)
    uint32_t debug;            // tuning only: bit0 skip epilogue, bit1 skip column loop
};

// Diversity epilogue (NG == 1): build_dense_population_summary (stats.rs:1367-1470) +
// calculate_per_site_diversity (stats.rs:4693-4750) fused.
struct DivEpilogue {
    uint32_t *alt_out, *called_out;  // [V] or nullptr
    double *pi_out, *theta_out;      // [v_hi - v_lo] or nullptr   (tracks)
    const uint32_t *site_flags;      // [n_batches] bit i of word b-b_lo: site 32b+i is masked or
                                     // filtered (fm_k_site_flags); nullptr when nothing is dropped
    // per-group lookup tables indexed by the called count n = 0..cap.  They hold exactly the
    // values the reference computes per site (IEEE division is correctly rounded on host and
    // device alike): inv_n = 1/n, scale = n/(n-1), theta = 1/H_{n-1} with H by forward summation
    // (stats.rs:4234-4240, 4716-4722).
    const double *tab_inv_n, *tab_scale, *tab_theta;
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

// Mask / filtered-position lookup hoisted out of the streaming kernel: one bit per site
// (stats.rs:4731-4743: position in filtered_positions or inside any mask interval [s,e)).
// One warp per batch; the result is a 32-bit word per batch.  Positions are ascending, so the
// warp locates the batch's first position among the merged, sorted intervals with a 32-ary
// search (3 dependent loads for 2,000 intervals instead of 11) and every lane then only walks
// forward over the few intervals that start inside the batch.
__global__ void __launch_bounds__(256)
fm_k_site_flags(const int64_t *__restrict__ pos, uint32_t v_lo, uint32_t v_hi, uint32_t b_lo,
                uint32_t n_batches, const int64_t *__restrict__ mask, uint32_t n_mask,
                const int64_t *__restrict__ filt, uint32_t n_filt, uint32_t *__restrict__ flags) {
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t GW = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t bi = gw; bi < n_batches; bi += GW) {
        const uint32_t vb = (b_lo + bi) * 32;
        const uint32_t v = vb + lane;
        const bool valid = v >= v_lo && v < v_hi;
        const int64_t p = valid ? pos[v] : 0;
        bool drop = false;
        if (valid && filt) drop = fm_in_sorted(filt, n_filt, p);
        if (mask && n_mask) {
            const uint32_t first = v_lo > vb ? v_lo - vb : 0u;  // first valid lane of the batch
            const int64_t p0 = __shfl_sync(FULL, p, first & 31u);
            // last interval with start <= p0, or -1: invariant start[lo] <= p0 < start[hi]
            int lo = -1, hi = (int)n_mask;
            while (hi - lo > 1) {
                const int step = (hi - lo - 1 + 31) / 32;
                const int idx = lo + 1 + (int)lane * step;
                const bool le = idx < hi && mask[2 * (size_t)idx] <= p0;
                const int k = __popc(__ballot_sync(FULL, le));  // probes are ascending: a prefix of ones
                const int nlo = k > 0 ? lo + 1 + (k - 1) * step : lo;
                const int nhi = min(hi, lo + 1 + k * step);
                lo = nlo;
                hi = nhi;
            }
            if (valid && !drop) {
                int j = lo;
                while (j + 1 < (int)n_mask && mask[2 * (size_t)(j + 1)] <= p) ++j;
                drop = j >= 0 && p < mask[2 * (size_t)j + 1];
            }
        }
        const uint32_t w = __ballot_sync(FULL, drop);
        if (lane == 0) flags[bi] = w;
    }
}

// Per-site diversity values (shared by the fused epilogue and the light kernel).
__device__ __forceinline__ void fm_div_site(const DivEpilogue &e, uint32_t v, uint32_t v_lo,
                                            uint32_t n, uint32_t alt, bool dropped, double &pi_part,
                                            uint32_t &seg, uint32_t &unc) {
    const bool poly = (n >= 2 && alt > 0 && alt < n);
    seg = poly ? 1u : 0u;       // stats.rs:1389 / 1406
    unc = (n < 2) ? 1u : 0u;    // stats.rs:1512-1516
    double pi_comp = 0.0;       // pi_from_components (stats.rs:2723-2733) via the tables
    if (n >= 2 && (e.pi_out || e.pi_form == FM_PIFORM_COMPONENTS)) {
        const double r = (double)(n - alt), a = (double)alt;
        const double sum_counts_sq = r * r + a * a;
        const double inv_n = __ldg(e.tab_inv_n + n);
        const double sum_p2 = sum_counts_sq * inv_n * inv_n;
        pi_comp = __ldg(e.tab_scale + n) * (1.0 - sum_p2);
    }
    if (e.pi_form == FM_PIFORM_COMPONENTS) {
        pi_part = pi_comp;
    } else {
        double val;
        pi_part = fm_pi_form(e.pi_form, n, alt, val) ? val : 0.0;
    }
    if (e.alt_out) e.alt_out[v] = alt;
    if (e.called_out) e.called_out[v] = n;
    if (e.pi_out) {
        // calculate_per_site_diversity, stats.rs:4710-4743
        double pi_value, theta_value;
        if (n < 2 || dropped) {
            pi_value = fm_nan();
            theta_value = fm_nan();
        } else {
            theta_value = poly ? __ldg(e.tab_theta + n) : 0.0;  // distinct_alleles > 1
            pi_value = pi_comp;
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

// Specialisations: NG groups, LG = log2(lanes per site) (LG == 5 is the column-chunked mode for
// rows wider than a pipeline step), HC = planes carry a called bitplane.
//
// Non-chunked geometry: a pipeline step holds `rounds` x SPS consecutive sites (SPS = 32/LPS);
// in every round LPS consecutive lanes read consecutive uint4 of one row, so a quarter-warp
// always touches one contiguous 128-byte span (bank-conflict free without padding).
template <int NG, int LG, bool HC>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 1)
fm_k_plane_pass(const __grid_constant__ PassParams<NG> P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[kWarpsPerCta * kMaxStages];
    __shared__ uint32_t batch_ring[kWarpsPerCta * kBatchRing];

    constexpr uint32_t LPS = 1u << LG;       // lanes per site
    constexpr uint32_t SPS = 32u >> LG;      // sites per round
    constexpr bool CHUNKED = (LG == 5);
    constexpr uint32_t NPL = HC ? 2u : 1u;   // planes per group
    constexpr uint32_t FULL = 0xffffffffu;

    const uint32_t lane = threadIdx.x & 31;
    // broadcast through a shuffle so the compiler knows the warp index (and everything derived
    // from it: staging addresses, barrier addresses) is warp-uniform
    const uint32_t warp = __shfl_sync(FULL, threadIdx.x >> 5, 0);
    const PassGeom &G = P.geom;
    const uint32_t n_stages = G.n_stages;
    const uint32_t stage_bytes = G.stage_bytes;
    const uint32_t n_sites_total = G.n_sites_total;
    const uint32_t n_batches = G.n_batches;
    const uint32_t n_chunks = CHUNKED ? G.n_chunks : 1u;
    const uint32_t cq = G.cq;
    const uint32_t rounds = CHUNKED ? 1u : G.rounds;
    const uint32_t step_sites = CHUNKED ? 1u : SPS * rounds;
    const uint32_t spb = CHUNKED ? 32u * n_chunks : LPS / rounds;  // steps per batch

    uint32_t wq16[NG];  // row bytes per plane
#pragma unroll
    for (int g = 0; g < NG; ++g) wq16[g] = P.g[g].wq * 16u;

    const uint32_t smem_base = fm_smem_u32(smem_raw) + warp * G.warp_smem_bytes;
    const uint32_t bar_base = fm_smem_u32(bars + warp * kMaxStages);
    uint32_t *my_ring = batch_ring + warp * kBatchRing;
    if (lane == 0) {
        for (uint32_t s = 0; s < n_stages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_base + 8u * s));
        fm_fence_mbar_init();
    }
    __syncwarp();

    // ---- dynamic batch scheduler: the issue side claims batches, the consume side replays the
    // same sequence from a small per-warp ring.  Partials are keyed by the batch index, so the
    // result does not depend on which warp processed which batch.
    // The atomic is fired one batch ahead so its round trip never sits on the critical path.
    uint32_t sched_j = blockIdx.x % kSchedCounters, sched_tried = 0;
    auto range_lo = [&](uint32_t j) { return (uint32_t)(((uint64_t)n_batches * j) / kSchedCounters); };
    uint32_t prefetched = 0;
    if (lane == 0) prefetched = atomicAdd(G.batch_counter + sched_j * kSchedStrideWords, 1u);
    auto claim = [&]() -> uint32_t {
        for (;;) {
            const uint32_t idx = __shfl_sync(FULL, prefetched, 0);
            const uint32_t lo = range_lo(sched_j), hi = range_lo(sched_j + 1);
            if (idx < hi - lo) {  // fire the next claim on the same range, one batch ahead
                if (lane == 0) prefetched = atomicAdd(G.batch_counter + sched_j * kSchedStrideWords, 1u);
                return lo + idx;
            }
            if (++sched_tried == kSchedCounters) return 0xffffffffu;  // every range is drained
            sched_j = (sched_j + 1 == kSchedCounters) ? 0 : sched_j + 1;
            if (lane == 0) prefetched = atomicAdd(G.batch_counter + sched_j * kSchedStrideWords, 1u);
        }
    };

    // issue the TMA bulk copies of step k of local batch `bl` into `stage` (lane 0 only)
    auto issue = [&](uint32_t bl, uint32_t k, uint32_t stage) {
        const uint32_t b = G.b_lo + bl;
        const uint32_t bar = bar_base + 8u * stage;
        uint32_t dst = smem_base + stage * stage_bytes;
        if constexpr (CHUNKED) {
            const uint32_t v0 = b * 32 + k / n_chunks;
            const uint32_t c0 = (k % n_chunks) * cq;
            uint32_t bytes[NG], total = 0;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const uint32_t w = P.g[g].wq;
                bytes[g] = (v0 < n_sites_total && c0 < w) ? min(cq, w - c0) * 16u : 0u;
                total += bytes[g] * NPL;
            }
            if (total == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
                return;
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (bytes[g]) {
                    const size_t src = (size_t)v0 * P.g[g].wq + c0;
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                            "r"(dst), "l"(P.g[g].allele + src), "r"(bytes[g]), "r"(bar) : "memory");
                    if constexpr (HC)
                        asm volatile(
                            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                                "r"(dst + cq * 16u), "l"(P.g[g].called + src), "r"(bytes[g]), "r"(bar) : "memory");
                }
                dst += cq * 16u * NPL;
            }
        } else {
            const uint32_t v0 = b * 32 + k * step_sites;
            const uint32_t nsites = (v0 < n_sites_total) ? min(step_sites, n_sites_total - v0) : 0u;
            if (nsites == 0) {  // batch tail beyond V: complete the phase with a plain arrive
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
                return;
            }
            uint32_t total = 0;
#pragma unroll
            for (int g = 0; g < NG; ++g) total += nsites * wq16[g] * NPL;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const uint32_t bytes = nsites * wq16[g];
                const uint32_t slot = step_sites * wq16[g];
                const size_t off = (size_t)v0 * wq16[g];
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                        "r"(dst), "l"(reinterpret_cast<const uint8_t *>(P.g[g].allele) + off), "r"(bytes), "r"(bar)
                    : "memory");
                if constexpr (HC)
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                            "r"(dst + slot), "l"(reinterpret_cast<const uint8_t *>(P.g[g].called) + off), "r"(bytes),
                        "r"(bar)
                        : "memory");
                dst += slot * NPL;
            }
        }
    };

    // issue cursor (uniform across the warp)
    uint32_t iss_ring = 0, iss_k = spb, iss_stage = 0, iss_batch = 0;
    auto issue_next = [&]() {
        if (iss_k == spb) {
            if (iss_ring != 0 && iss_batch >= n_batches) return;  // scheduler drained
            iss_batch = claim();
            iss_k = 0;
            if (lane == 0) my_ring[iss_ring & (kBatchRing - 1)] = iss_batch;
            ++iss_ring;
            if (iss_batch >= n_batches) {
                iss_k = spb;
                return;
            }
        }
        if (lane == 0) issue(iss_batch, iss_k, iss_stage);
        ++iss_k;
        iss_stage = (iss_stage + 1 == n_stages) ? 0 : iss_stage + 1;
    };
    for (uint32_t s = 0; s < n_stages; ++s) issue_next();
    __syncwarp();

    // per-site totals of this lane's batch site; packed alt | called << 16 when HC (rows of the
    // non-chunked modes hold < 65536 haplotypes), plain counts in chunked mode
    uint32_t site_a[NG], site_c[NG], acc_a[NG], acc_c[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) site_a[g] = site_c[g] = acc_a[g] = acc_c[g] = 0;

    const uint32_t slot_in_round = lane >> LG;
    const uint32_t phase = lane & (LPS - 1);
    const uint32_t xfer_from = (lane & (SPS - 1)) << LG;          // source lane of the transpose
    const uint32_t xfer_round = CHUNKED ? 0u : (lane >> (5 - LG));  // round whose sites land here

    uint32_t plane_off[NG];  // byte offset of group g's allele plane inside a stage
    uint32_t nit_full[NG];   // column iterations that are in range for every lane
    uint32_t last_cols[NG];  // lanes (phases) that still have a column in the last iteration
    {
        uint32_t acc = 0;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            plane_off[g] = acc;
            acc += (CHUNKED ? cq * 16u : step_sites * wq16[g]) * NPL;
            const uint32_t w = P.g[g].wq;
            const uint32_t nit = (w + LPS - 1) / LPS;
            nit_full[g] = nit - 1;
            last_cols[g] = w - (nit - 1) * LPS;
        }
    }

    uint32_t con_ring = 0, stage = 0, parity = 0;
    for (;;) {
        const uint32_t bl = my_ring[con_ring & (kBatchRing - 1)];
        ++con_ring;
        if (bl >= n_batches) break;
        const uint32_t b = G.b_lo + bl;
        TailRegs tails[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) fm_tail_load(P.g[g], b * 32 + lane, n_sites_total, HC, tails[g]);
        for (uint32_t k = 0; k < spb; ++k) {
            {
                const uint32_t bar = bar_base + 8u * stage;
                uint32_t ok;
                do {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                        "selp.u32 %0, 1, 0, p;\n\t}"
                        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
                } while (!ok);
            }
            const uint32_t sbase = smem_base + stage * stage_bytes;
            if constexpr (!CHUNKED) {
                const uint32_t v0 = b * 32 + k * step_sites;
                auto round_body = [&](uint32_t r) {
                    const uint32_t site_local = r * SPS + slot_in_round;
                    const bool live = (v0 + site_local) < n_sites_total;
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        const uint32_t w16 = wq16[g];
                        const uint32_t ra = sbase + plane_off[g] + site_local * w16;
                        const uint32_t pb = step_sites * w16;  // distance allele -> called plane
                        uint32_t a = 0, c = 0;
                        if (live && !(G.debug & 2u)) {
                            // columns phase, phase+LPS, ...: the first nit-1 are inside the row for
                            // every lane, only the last one needs a bound check
                            uint32_t p = ra + phase * 16u;
                            const uint32_t nfull = nit_full[g];
#pragma unroll 2
                            for (uint32_t j = 0; j < nfull; ++j) {
                                const uint4 xa = fm_lds128(p);
                                a += fm_popc4(xa);
                                if constexpr (HC) {
                                    const uint4 xc = fm_lds128(p + pb);
                                    c += fm_popc4(xc);
                                }
                                p += LPS * 16u;
                            }
                            if (phase < last_cols[g]) {
                                const uint4 xa = fm_lds128(p);
                                a += fm_popc4(xa);
                                if constexpr (HC) {
                                    const uint4 xc = fm_lds128(p + pb);
                                    c += fm_popc4(xc);
                                }
                            }
                        }
                        uint32_t ac = HC ? (a | (c << 16)) : a;
#pragma unroll
                        for (uint32_t o = LPS >> 1; o > 0; o >>= 1) ac += __shfl_xor_sync(FULL, ac, o);
                        const uint32_t t = __shfl_sync(FULL, ac, xfer_from);
                        if (xfer_round == k * rounds + r) site_a[g] = t;
                    }
                };
                if (rounds == 1) {
                    round_body(0);
                } else {
                    for (uint32_t r = 0; r < rounds; r += 2) {  // rounds is a power of two
                        round_body(r);
                        round_body(r + 1);
                    }
                }
            } else {
                const uint32_t site_in_batch = k / n_chunks;
                const uint32_t ck = k - site_in_batch * n_chunks;
                const uint32_t c0 = ck * cq;
                const bool live = (b * 32 + site_in_batch) < n_sites_total;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const uint32_t w = P.g[g].wq;
                    const uint32_t cols16 = (c0 < w) ? min(cq, w - c0) * 16u : 0u;
                    const uint32_t ra = sbase + plane_off[g];
                    if (live) {
#pragma unroll 2
                        for (uint32_t u = lane * 16u; u < cols16; u += 512u) {
                            acc_a[g] += fm_popc4(fm_lds128(ra + u));
                            if constexpr (HC) acc_c[g] += fm_popc4(fm_lds128(ra + cq * 16u + u));
                        }
                    }
                    if (ck == n_chunks - 1) {
                        const uint32_t ta = fm_warp_sum_u(acc_a[g]);
                        uint32_t tc = 0;
                        if constexpr (HC) tc = fm_warp_sum_u(acc_c[g]);
                        if (lane == site_in_batch) {
                            site_a[g] = ta;
                            site_c[g] = tc;
                        }
                        acc_a[g] = 0;
                        acc_c[g] = 0;
                    }
                }
            }
            __syncwarp();   // every lane is done reading this stage
            issue_next();   // refill it with the step n_stages ahead
            stage = (stage + 1 == n_stages) ? 0 : stage + 1;
            parity ^= (stage == 0);
        }
        // ---- batch epilogue: lane i <-> site 32b + i, all 32 lanes active
        {
            uint32_t alt[NG], cnt[NG];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if constexpr (CHUNKED) {
                    alt[g] = site_a[g];
                    cnt[g] = HC ? site_c[g] : P.g[g].cap;
                } else {
                    alt[g] = HC ? (site_a[g] & 0xffffu) : site_a[g];
                    cnt[g] = HC ? (site_a[g] >> 16) : P.g[g].cap;
                }
                fm_tail_add(tails[g], alt[g], cnt[g], HC);
            }
            const uint32_t v = b * 32 + lane;
            const bool valid = (v >= G.v_lo) && (v < G.v_hi);
            if constexpr (NG == 1) {
                double pi_part = 0.0;
                uint32_t seg = 0, unc = 0;
                const uint32_t flags = P.div.site_flags ? __ldg(P.div.site_flags + bl) : 0u;
                if (valid && !(G.debug & 1u))
                    fm_div_site(P.div, v, G.v_lo, cnt[0], alt[0], (flags >> lane) & 1u, pi_part, seg, unc);
                const double s_pi = fm_warp_sum(pi_part);
                const uint32_t s_u = fm_warp_sum_u(seg | (unc << 16));
                if (lane == 0) {
                    P.div.part_pi[bl] = s_pi;
                    P.div.part_u[2 * bl] = s_u & 0xffffu;
                    P.div.part_u[2 * bl + 1] = s_u >> 16;
                }
            } else {
                HudsonAcc acc{0.0, 0.0, 0.0, 0.0, 0.0, 0u, 0u, 0u};
                if (valid) {
                    fm_hudson_contrib(P.hud, v, G.v_lo, cnt[0], alt[0], cnt[1], alt[1], acc);
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        if (P.hud.alt_out[g]) P.hud.alt_out[g][v] = alt[g];
                        if (P.hud.called_out[g]) P.hud.called_out[g][v] = cnt[g];
                    }
                }
                const double r0 = fm_warp_sum(acc.num), r1 = fm_warp_sum(acc.den),
                             r2 = fm_warp_sum(acc.dxy), r3 = fm_warp_sum(acc.pi1),
                             r4 = fm_warp_sum(acc.pi2);
                const uint32_t u01 = fm_warp_sum_u(acc.skipped | (acc.unc1 << 8) | (acc.unc2 << 16));
                if (lane == 0) {
                    double *pd = P.hud.part_d + (size_t)bl * 5;
                    pd[0] = r0; pd[1] = r1; pd[2] = r2; pd[3] = r3; pd[4] = r4;
                    uint32_t *pu = P.hud.part_u + (size_t)bl * 3;
                    pu[0] = u01 & 0xffu; pu[1] = (u01 >> 8) & 0xffu; pu[2] = u01 >> 16;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------ plane pass, several groups
// fm_k_plane_pass_seq: ONE persistent launch streams the planes of up to kMaxSeq groups one after
// the other over the same site range (the two orientation groups of a per-site diversity call).
// Same per-warp TMA rings, same dynamic scheduler and the same per-batch arithmetic as
// fm_k_plane_pass<1>; the batch index space is the concatenation of the groups' batch ranges, so
// the launch ramp-up and the end-of-kernel quantisation (a warp's last batch) are paid once, not
// once per group, and the finer batches of the smallest group fill the tail.
constexpr int kMaxSeq = 4;

struct SeqGroup {
    GroupPlanes g;
    DivEpilogue div;
    uint32_t rounds;  // rounds of (32/lps) sites per pipeline step for this group's row width
};

struct SeqParams {
    SeqGroup seg[kMaxSeq];
    uint32_t n_seg;
    PassGeom geom;  // lps, n_stages, stage_bytes (max over groups), warps, site range, scheduler counters
};

template <int LG, bool HC>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 1)
fm_k_plane_pass_seq(const __grid_constant__ SeqParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[kWarpsPerCta * kMaxStages];
    __shared__ uint32_t batch_ring[kWarpsPerCta * kBatchRing];
    static_assert(LG < 5, "column-chunked rows take the per-group kernel");
    constexpr uint32_t LPS = 1u << LG;
    constexpr uint32_t SPS = 32u >> LG;
    constexpr uint32_t NPL = HC ? 2u : 1u;
    constexpr uint32_t FULL = 0xffffffffu;

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = __shfl_sync(FULL, threadIdx.x >> 5, 0);
    const PassGeom &G = P.geom;
    const uint32_t n_stages = G.n_stages;
    const uint32_t stage_bytes = G.stage_bytes;
    const uint32_t n_sites_total = G.n_sites_total;
    const uint32_t nb_seg = G.n_batches;              // batches per group (same site range for all)
    const uint32_t n_batches = nb_seg * P.n_seg;      // concatenated batch space

    const uint32_t smem_base = fm_smem_u32(smem_raw) + warp * G.warp_smem_bytes;
    const uint32_t bar_base = fm_smem_u32(bars + warp * kMaxStages);
    uint32_t *my_ring = batch_ring + warp * kBatchRing;
    if (lane == 0) {
        for (uint32_t s = 0; s < n_stages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_base + 8u * s));
        fm_fence_mbar_init();
    }
    __syncwarp();

    uint32_t sched_j = blockIdx.x % kSchedCounters, sched_tried = 0;
    auto range_lo = [&](uint32_t j) { return (uint32_t)(((uint64_t)n_batches * j) / kSchedCounters); };
    uint32_t prefetched = 0;
    if (lane == 0) prefetched = atomicAdd(G.batch_counter + sched_j * kSchedStrideWords, 1u);
    auto claim = [&]() -> uint32_t {
        for (;;) {
            const uint32_t idx = __shfl_sync(FULL, prefetched, 0);
            const uint32_t lo = range_lo(sched_j), hi = range_lo(sched_j + 1);
            if (idx < hi - lo) {
                if (lane == 0) prefetched = atomicAdd(G.batch_counter + sched_j * kSchedStrideWords, 1u);
                return lo + idx;
            }
            if (++sched_tried == kSchedCounters) return 0xffffffffu;
            sched_j = (sched_j + 1 == kSchedCounters) ? 0 : sched_j + 1;
            if (lane == 0) prefetched = atomicAdd(G.batch_counter + sched_j * kSchedStrideWords, 1u);
        }
    };

    // ---- issue side: per-group context of the batch being issued
    uint32_t iss_ring = 0, iss_k = 0, iss_spb = 0, iss_stage = 0, iss_batch = 0;
    uint32_t iss_w16 = 0, iss_step_sites = 1, iss_seg = 0, iss_lb = 0;
    auto issue = [&](uint32_t stage) {  // lane 0 only: TMA bulk copies of step iss_k of batch iss_lb of group iss_seg
        const uint32_t bar = bar_base + 8u * stage;
        const uint32_t dst = smem_base + stage * stage_bytes;
        const uint32_t v0 = (G.b_lo + iss_lb) * 32 + iss_k * iss_step_sites;
        const uint32_t nsites = (v0 < n_sites_total) ? min(iss_step_sites, n_sites_total - v0) : 0u;
        if (nsites == 0) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
            return;
        }
        const uint32_t bytes = nsites * iss_w16;
        const uint32_t slot = iss_step_sites * iss_w16;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * NPL) : "memory");
        const size_t off = (size_t)v0 * iss_w16;
        const GroupPlanes &gp = P.seg[iss_seg].g;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(reinterpret_cast<const uint8_t *>(gp.allele) + off), "r"(bytes), "r"(bar)
                     : "memory");
        if constexpr (HC)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             dst + slot),
                         "l"(reinterpret_cast<const uint8_t *>(gp.called) + off), "r"(bytes), "r"(bar)
                         : "memory");
    };
    auto issue_next = [&]() {
        if (iss_k == iss_spb) {
            if (iss_ring != 0 && iss_batch >= n_batches) return;  // scheduler drained
            iss_batch = claim();
            iss_k = 0;
            if (lane == 0) my_ring[iss_ring & (kBatchRing - 1)] = iss_batch;
            ++iss_ring;
            if (iss_batch >= n_batches) {
                iss_spb = 0;
                return;
            }
            iss_seg = iss_batch / nb_seg;
            iss_lb = iss_batch - iss_seg * nb_seg;
            const uint32_t r = P.seg[iss_seg].rounds;
            iss_w16 = P.seg[iss_seg].g.wq * 16u;
            iss_step_sites = SPS * r;
            iss_spb = LPS / r;
        }
        if (lane == 0) issue(iss_stage);
        ++iss_k;
        iss_stage = (iss_stage + 1 == n_stages) ? 0 : iss_stage + 1;
    };
    for (uint32_t s = 0; s < n_stages; ++s) issue_next();
    __syncwarp();

    const uint32_t slot_in_round = lane >> LG;
    const uint32_t phase = lane & (LPS - 1);
    const uint32_t xfer_from = (lane & (SPS - 1)) << LG;
    const uint32_t xfer_round = lane >> (5 - LG);

    uint32_t con_ring = 0, stage = 0, parity = 0;
    for (;;) {
        const uint32_t gb = my_ring[con_ring & (kBatchRing - 1)];
        ++con_ring;
        if (gb >= n_batches) break;
        // ---- consume side: this batch's group
        const uint32_t sg = gb / nb_seg;
        const uint32_t bl = gb - sg * nb_seg;
        const SeqGroup &S = P.seg[sg];
        const uint32_t w = S.g.wq, w16 = w * 16u, rounds = S.rounds;
        const uint32_t step_sites = SPS * rounds, spb = LPS / rounds;
        const uint32_t nit = (w + LPS - 1) / LPS;
        const uint32_t nfull = nit - 1, last_cols = w - nfull * LPS;
        const uint32_t pb = step_sites * w16;  // distance allele -> called plane inside a stage
        const uint32_t b = G.b_lo + bl;
        uint32_t site_ac = 0;
        TailRegs tails;
        fm_tail_load(S.g, b * 32 + lane, n_sites_total, HC, tails);
        for (uint32_t k = 0; k < spb; ++k) {
            {
                const uint32_t bar = bar_base + 8u * stage;
                uint32_t ok;
                do {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                        "selp.u32 %0, 1, 0, p;\n\t}"
                        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
                } while (!ok);
            }
            const uint32_t sbase = smem_base + stage * stage_bytes;
            const uint32_t v0 = b * 32 + k * step_sites;
            auto round_body = [&](uint32_t r) {
                const uint32_t site_local = r * SPS + slot_in_round;
                const bool live = (v0 + site_local) < n_sites_total;
                const uint32_t ra = sbase + site_local * w16;
                uint32_t a = 0, c = 0;
                if (live) {
                    uint32_t p = ra + phase * 16u;
#pragma unroll 2
                    for (uint32_t j = 0; j < nfull; ++j) {
                        a += fm_popc4(fm_lds128(p));
                        if constexpr (HC) c += fm_popc4(fm_lds128(p + pb));
                        p += LPS * 16u;
                    }
                    if (phase < last_cols) {
                        a += fm_popc4(fm_lds128(p));
                        if constexpr (HC) c += fm_popc4(fm_lds128(p + pb));
                    }
                }
                uint32_t ac = HC ? (a | (c << 16)) : a;
#pragma unroll
                for (uint32_t o = LPS >> 1; o > 0; o >>= 1) ac += __shfl_xor_sync(FULL, ac, o);
                const uint32_t t = __shfl_sync(FULL, ac, xfer_from);
                if (xfer_round == k * rounds + r) site_ac = t;
            };
            if (rounds == 1) {
                round_body(0);
            } else {
                for (uint32_t r = 0; r < rounds; r += 2) {
                    round_body(r);
                    round_body(r + 1);
                }
            }
            __syncwarp();
            issue_next();
            stage = (stage + 1 == n_stages) ? 0 : stage + 1;
            parity ^= (stage == 0);
        }
        // ---- batch epilogue: lane i <-> site 32b + i
        {
            uint32_t alt = HC ? (site_ac & 0xffffu) : site_ac;
            uint32_t cnt = HC ? (site_ac >> 16) : S.g.cap;
            fm_tail_add(tails, alt, cnt, HC);
            const uint32_t v = b * 32 + lane;
            const bool valid = (v >= G.v_lo) && (v < G.v_hi);
            double pi_part = 0.0;
            uint32_t seg = 0, unc = 0;
            const uint32_t flags = S.div.site_flags ? __ldg(S.div.site_flags + bl) : 0u;
            if (valid) fm_div_site(S.div, v, G.v_lo, cnt, alt, (flags >> lane) & 1u, pi_part, seg, unc);
            const double s_pi = fm_warp_sum(pi_part);
            const uint32_t s_u = fm_warp_sum_u(seg | (unc << 16));
            if (lane == 0) {
                S.div.part_pi[bl] = s_pi;
                S.div.part_u[2 * bl] = s_u & 0xffffu;
                S.div.part_u[2 * bl + 1] = s_u >> 16;
            }
        }
    }
}

// ------------------------------------------------------------------------------ plane pass, segment table
// fm_k_plane_pass_tab: the same persistent per-warp-ring pass over an arbitrary LIST of plane segments read
// from a descriptor table in device memory.  Two uses:
//   * unit = one segment: many (region, group) pairs -- e.g. the per-region matrices the CLI builds one after
//     the other (process.rs:2169), each with its own population -- share ONE launch, so 64 small regions stream
//     at the rate of one large one instead of paying 64 launches + ramp-ups (cfg1 x 64);
//   * unit = two segments over the same sites (gpu == 2): a warp streams the batch of group 1, keeps its
//     counts in registers, streams the same batch of group 2 and evaluates the Hudson components of the 32
//     sites in place (stats.rs:3179-3278 / 1554-1623) -- both groups' counts and summary partials AND the
//     Hudson partials from one sweep of the planes, each group streamed with its own step geometry.
struct TabSeg {
    GroupPlanes g;
    DivEpilogue div;
    uint32_t rounds;          // rounds of (32/lps) sites per pipeline step for this segment's row width
    uint32_t v_lo, v_hi;      // site range analysed
    uint32_t b_lo;            // first batch of the range (v_lo / 32)
    uint32_t n_sites_total;   // rows available in this segment's planes
    uint32_t pad;
};

struct TabParams {
    const TabSeg *segs;            // [n_units * gpu]
    const uint32_t *unit_prefix;   // [n_units + 1]: batches of the units before u (unit u has prefix[u+1]-prefix[u])
    uint32_t n_units, gpu;         // gpu = segments (groups) per unit: 1 or 2
    const HudsonEpilogue *hud;     // [n_units] Hudson epilogues (gpu == 2) or nullptr
    PassGeom geom;                 // lps, n_stages, stage_bytes, warps, warp_smem_bytes, n_batches (total), counters
    // a single unit travels in the kernel parameters themselves (no descriptor upload: three small H2D copies
    // are ~20 us of a 250 us sharded step); use_inline selects these instead of the pointers above
    uint32_t use_inline, has_inline_hud;
    TabSeg inline_segs[2];
    HudsonEpilogue inline_hud;
    uint32_t inline_prefix[2];
};

template <int LG, bool HC>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 1)
fm_k_plane_pass_tab(const __grid_constant__ TabParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[kWarpsPerCta * kMaxStages];
    __shared__ uint32_t batch_ring[kWarpsPerCta * kBatchRing];
    static_assert(LG < 5, "column-chunked rows take the per-group kernel");
    constexpr uint32_t LPS = 1u << LG;
    constexpr uint32_t SPS = 32u >> LG;
    constexpr uint32_t NPL = HC ? 2u : 1u;
    constexpr uint32_t FULL = 0xffffffffu;

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = __shfl_sync(FULL, threadIdx.x >> 5, 0);
    const PassGeom &G = P.geom;
    const uint32_t n_stages = G.n_stages;
    const uint32_t stage_bytes = G.stage_bytes;
    const uint32_t n_batches = G.n_batches;  // whole launch
    const uint32_t gpu = P.gpu;
    const TabSeg *const segs = P.use_inline ? P.inline_segs : P.segs;
    const uint32_t *const unit_prefix = P.use_inline ? P.inline_prefix : P.unit_prefix;
    const HudsonEpilogue *const hud = P.use_inline ? (P.has_inline_hud ? &P.inline_hud : nullptr) : P.hud;

    const uint32_t smem_base = fm_smem_u32(smem_raw) + warp * G.warp_smem_bytes;
    const uint32_t bar_base = fm_smem_u32(bars + warp * kMaxStages);
    uint32_t *my_ring = batch_ring + warp * kBatchRing;
    if (lane == 0) {
        for (uint32_t s = 0; s < n_stages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_base + 8u * s));
        fm_fence_mbar_init();
    }
    __syncwarp();

    uint32_t sched_j = blockIdx.x % kSchedCounters, sched_tried = 0;
    auto range_lo = [&](uint32_t j) { return (uint32_t)(((uint64_t)n_batches * j) / kSchedCounters); };
    uint32_t prefetched = 0;
    if (lane == 0) prefetched = atomicAdd(G.batch_counter + sched_j * kSchedStrideWords, 1u);
    auto claim = [&]() -> uint32_t {
        for (;;) {
            const uint32_t idx = __shfl_sync(FULL, prefetched, 0);
            const uint32_t lo = range_lo(sched_j), hi = range_lo(sched_j + 1);
            if (idx < hi - lo) {
                if (lane == 0) prefetched = atomicAdd(G.batch_counter + sched_j * kSchedStrideWords, 1u);
                return lo + idx;
            }
            if (++sched_tried == kSchedCounters) return 0xffffffffu;
            sched_j = (sched_j + 1 == kSchedCounters) ? 0 : sched_j + 1;
            if (lane == 0) prefetched = atomicAdd(G.batch_counter + sched_j * kSchedStrideWords, 1u);
        }
    };
    // unit of a launch-wide batch index: last u with prefix[u] <= gb (warp-uniform)
    auto unit_of = [&](uint32_t gb) -> uint32_t {
        uint32_t lo = 0, hi = P.n_units;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (unit_prefix[mid] <= gb) lo = mid; else hi = mid;
        }
        return lo;
    };

    // ---- issue side: context of the (unit batch, group) whose steps are being issued
    uint32_t iss_ring = 0, iss_k = 0, iss_spb = 0, iss_stage = 0, iss_batch = 0;
    uint32_t iss_w16 = 0, iss_step_sites = 1, iss_g = 0, iss_unit = 0, iss_b = 0, iss_nst = 0;
    const uint4 *iss_allele = nullptr, *iss_called = nullptr;
    bool iss_live = false;
    auto load_iss_seg = [&]() {
        const TabSeg &S = segs[iss_unit * gpu + iss_g];
        iss_w16 = S.g.wq * 16u;
        iss_step_sites = SPS * S.rounds;
        iss_spb = LPS / S.rounds;
        iss_nst = S.n_sites_total;
        iss_allele = S.g.allele;
        iss_called = S.g.called;
    };
    auto issue = [&](uint32_t stage) {  // lane 0 only
        const uint32_t bar = bar_base + 8u * stage;
        const uint32_t dst = smem_base + stage * stage_bytes;
        const uint32_t v0 = iss_b * 32 + iss_k * iss_step_sites;
        const uint32_t nsites = (v0 < iss_nst) ? min(iss_step_sites, iss_nst - v0) : 0u;
        if (nsites == 0) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
            return;
        }
        const uint32_t bytes = nsites * iss_w16;
        const uint32_t slot = iss_step_sites * iss_w16;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * NPL) : "memory");
        const size_t off = (size_t)v0 * iss_w16;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(reinterpret_cast<const uint8_t *>(iss_allele) + off), "r"(bytes), "r"(bar)
                     : "memory");
        if constexpr (HC)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             dst + slot),
                         "l"(reinterpret_cast<const uint8_t *>(iss_called) + off), "r"(bytes), "r"(bar)
                         : "memory");
    };
    auto issue_next = [&]() {
        if (iss_k == iss_spb) {
            if (iss_live && iss_g + 1 < gpu) {  // next group of the same unit batch
                ++iss_g;
                iss_k = 0;
                load_iss_seg();
            } else {
                if (iss_ring != 0 && iss_batch >= n_batches) return;  // scheduler drained
                iss_batch = claim();
                iss_k = 0;
                if (lane == 0) my_ring[iss_ring & (kBatchRing - 1)] = iss_batch;
                ++iss_ring;
                if (iss_batch >= n_batches) {
                    iss_spb = 0;
                    iss_live = false;
                    return;
                }
                iss_live = true;
                iss_unit = unit_of(iss_batch);
                iss_g = 0;
                load_iss_seg();
                iss_b = segs[iss_unit * gpu].b_lo + (iss_batch - unit_prefix[iss_unit]);
            }
        }
        if (lane == 0) issue(iss_stage);
        ++iss_k;
        iss_stage = (iss_stage + 1 == n_stages) ? 0 : iss_stage + 1;
    };
    for (uint32_t s = 0; s < n_stages; ++s) issue_next();
    __syncwarp();

    const uint32_t slot_in_round = lane >> LG;
    const uint32_t phase = lane & (LPS - 1);
    const uint32_t xfer_from = (lane & (SPS - 1)) << LG;
    const uint32_t xfer_round = lane >> (5 - LG);

    uint32_t con_ring = 0, stage = 0, parity = 0;
    for (;;) {
        const uint32_t gb = my_ring[con_ring & (kBatchRing - 1)];
        ++con_ring;
        if (gb >= n_batches) break;
        const uint32_t unit = unit_of(gb);
        const uint32_t bl = gb - unit_prefix[unit];  // batch inside the unit's range
        uint32_t alt_g0 = 0, cnt_g0 = 0, alt_g1 = 0, cnt_g1 = 0;
        uint32_t b = 0, u_vlo = 0, u_vhi = 0;
        for (uint32_t g = 0; g < gpu; ++g) {
            const TabSeg &S = segs[unit * gpu + g];
            const uint32_t w = S.g.wq, w16 = w * 16u, rounds = S.rounds;
            const uint32_t step_sites = SPS * rounds, spb = LPS / rounds;
            const uint32_t nit = (w + LPS - 1) / LPS;
            const uint32_t nfull = nit - 1, last_cols = w - nfull * LPS;
            const uint32_t pb = step_sites * w16;
            const uint32_t nst = S.n_sites_total;
            b = S.b_lo + bl;
            u_vlo = S.v_lo;
            u_vhi = S.v_hi;
            uint32_t site_ac = 0;
            TailRegs tails;
            fm_tail_load(S.g, b * 32 + lane, nst, HC, tails);
            for (uint32_t k = 0; k < spb; ++k) {
                {
                    const uint32_t bar = bar_base + 8u * stage;
                    uint32_t ok;
                    do {
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t"
                            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                            "selp.u32 %0, 1, 0, p;\n\t}"
                            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
                    } while (!ok);
                }
                const uint32_t sbase = smem_base + stage * stage_bytes;
                const uint32_t v0 = b * 32 + k * step_sites;
                auto round_body = [&](uint32_t r) {
                    const uint32_t site_local = r * SPS + slot_in_round;
                    const bool live = (v0 + site_local) < nst;
                    const uint32_t ra = sbase + site_local * w16;
                    uint32_t a = 0, c = 0;
                    if (live) {
                        uint32_t p = ra + phase * 16u;
#pragma unroll 2
                        for (uint32_t j = 0; j < nfull; ++j) {
                            a += fm_popc4(fm_lds128(p));
                            if constexpr (HC) c += fm_popc4(fm_lds128(p + pb));
                            p += LPS * 16u;
                        }
                        if (phase < last_cols) {
                            a += fm_popc4(fm_lds128(p));
                            if constexpr (HC) c += fm_popc4(fm_lds128(p + pb));
                        }
                    }
                    uint32_t ac = HC ? (a | (c << 16)) : a;
#pragma unroll
                    for (uint32_t o = LPS >> 1; o > 0; o >>= 1) ac += __shfl_xor_sync(FULL, ac, o);
                    const uint32_t t = __shfl_sync(FULL, ac, xfer_from);
                    if (xfer_round == k * rounds + r) site_ac = t;
                };
                if (rounds == 1) {
                    round_body(0);
                } else {
                    for (uint32_t r = 0; r < rounds; r += 2) {
                        round_body(r);
                        round_body(r + 1);
                    }
                }
                __syncwarp();
                issue_next();
                stage = (stage + 1 == n_stages) ? 0 : stage + 1;
                parity ^= (stage == 0);
            }
            // ---- group epilogue: lane i <-> site 32b + i (counts, summary partials, optional tracks)
            uint32_t alt = HC ? (site_ac & 0xffffu) : site_ac;
            uint32_t cnt = HC ? (site_ac >> 16) : S.g.cap;
            fm_tail_add(tails, alt, cnt, HC);
            if (g == 0) {
                alt_g0 = alt;
                cnt_g0 = cnt;
            } else {
                alt_g1 = alt;
                cnt_g1 = cnt;
            }
            const uint32_t v = b * 32 + lane;
            const bool valid = (v >= S.v_lo) && (v < S.v_hi);
            if (S.div.part_pi) {
                double pi_part = 0.0;
                uint32_t seg = 0, unc = 0;
                const uint32_t flags = S.div.site_flags ? __ldg(S.div.site_flags + bl) : 0u;
                if (valid) fm_div_site(S.div, v, S.v_lo, cnt, alt, (flags >> lane) & 1u, pi_part, seg, unc);
                const double s_pi = fm_warp_sum(pi_part);
                const uint32_t s_u = fm_warp_sum_u(seg | (unc << 16));
                if (lane == 0) {
                    S.div.part_pi[bl] = s_pi;
                    S.div.part_u[2 * bl] = s_u & 0xffffu;
                    S.div.part_u[2 * bl + 1] = s_u >> 16;
                }
            }
        }
        if (gpu == 2 && hud) {  // ---- Hudson components of the batch from both groups' counts
            const HudsonEpilogue &H = hud[unit];
            const uint32_t v = b * 32 + lane;
            HudsonAcc acc{0.0, 0.0, 0.0, 0.0, 0.0, 0u, 0u, 0u};
            if (v >= u_vlo && v < u_vhi) fm_hudson_contrib(H, v, u_vlo, cnt_g0, alt_g0, cnt_g1, alt_g1, acc);
            const double r0 = fm_warp_sum(acc.num), r1 = fm_warp_sum(acc.den), r2 = fm_warp_sum(acc.dxy),
                         r3 = fm_warp_sum(acc.pi1), r4 = fm_warp_sum(acc.pi2);
            const uint32_t u01 = fm_warp_sum_u(acc.skipped | (acc.unc1 << 8) | (acc.unc2 << 16));
            if (lane == 0) {
                double *pd = H.part_d + (size_t)bl * 5;
                pd[0] = r0; pd[1] = r1; pd[2] = r2; pd[3] = r3; pd[4] = r4;
                uint32_t *pu = H.part_u + (size_t)bl * 3;
                pu[0] = u01 & 0xffu; pu[1] = (u01 >> 8) & 0xffu; pu[2] = u01 >> 16;
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
        const uint32_t flags = e.site_flags ? e.site_flags[bi] : 0u;
        if (valid) fm_div_site(e, v, v_lo, called[v], alt[v], (flags >> lane) & 1u, pi_part, seg, unc);
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

// Second-level reduction: super-batch s folds batches [s*256, s*256+256) of the GLOBAL batch
// grid with a fixed shape -- lane l sums its 8 consecutive batches in order, then a butterfly
// over lanes -- for `nd` double columns and `nu` u32 columns.  One warp per super-batch.
__global__ void __launch_bounds__(128)
fm_k_reduce_partials(const double *__restrict__ pd, int nd, const uint32_t *__restrict__ pu, int nu,
                     uint32_t b_lo, uint32_t n_batches, uint32_t s_lo, uint32_t n_super,
                     double *__restrict__ out_d, uint64_t *__restrict__ out_u) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= n_super) return;
    constexpr uint32_t per_lane = kSuperBatches / 32;
    const uint64_t g0 = (uint64_t)(s_lo + s) * kSuperBatches + (uint64_t)lane * per_lane;
    const uint64_t b_hi = (uint64_t)b_lo + n_batches;
    for (int col = 0; col < nd; ++col) {
        double acc = 0.0;
#pragma unroll
        for (uint32_t i = 0; i < per_lane; ++i) {
            const uint64_t b = g0 + i;
            if (b >= b_lo && b < b_hi) acc += pd[(b - b_lo) * nd + col];
        }
        acc = fm_warp_sum(acc);
        if (lane == 0) out_d[(size_t)s * nd + col] = acc;
    }
    for (int col = 0; col < nu; ++col) {
        uint32_t acc = 0;
#pragma unroll
        for (uint32_t i = 0; i < per_lane; ++i) {
            const uint64_t b = g0 + i;
            if (b >= b_lo && b < b_hi) acc += pu[(b - b_lo) * nu + col];
        }
        acc = fm_warp_sum_u(acc);
        if (lane == 0) out_u[(size_t)s * nu + col] = acc;
    }
}

// The same fold for many units at once (fm_k_plane_pass_tab launches): entry e = (first partial of the unit in
// the shared partial arrays, the unit's first global batch, its batch count, the global super-batch index); one
// warp per entry, the association of fm_k_reduce_partials -- a unit reduced here gives the bits it would give alone.
__global__ void __launch_bounds__(128)
fm_k_reduce_partials_tab(const double *__restrict__ pd, int nd, const uint32_t *__restrict__ pu, int nu,
                         const uint4 *__restrict__ ents, uint32_t n_ent, double *__restrict__ out_d,
                         uint64_t *__restrict__ out_u) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= n_ent) return;
    const uint4 E = ents[e];
    constexpr uint32_t per_lane = kSuperBatches / 32;
    const uint64_t g0 = (uint64_t)E.w * kSuperBatches + (uint64_t)lane * per_lane;
    const uint64_t b_lo = E.y, b_hi = (uint64_t)E.y + E.z;
    for (int col = 0; col < nd; ++col) {
        double acc = 0.0;
#pragma unroll
        for (uint32_t i = 0; i < per_lane; ++i) {
            const uint64_t b = g0 + i;
            if (b >= b_lo && b < b_hi) acc += pd[((size_t)E.x + (b - b_lo)) * nd + col];
        }
        acc = fm_warp_sum(acc);
        if (lane == 0) out_d[(size_t)e * nd + col] = acc;
    }
    for (int col = 0; col < nu; ++col) {
        uint32_t acc = 0;
#pragma unroll
        for (uint32_t i = 0; i < per_lane; ++i) {
            const uint64_t b = g0 + i;
            if (b >= b_lo && b < b_hi) acc += pu[((size_t)E.x + (b - b_lo)) * nu + col];
        }
        acc = fm_warp_sum_u(acc);
        if (lane == 0) out_u[(size_t)e * nu + col] = acc;
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
