// fm_host_pack.cpp -- host side of the packed (2 bits per genotype) ingest format, SURVEY 8 f1.
//
// fm_pack_rows turns rows of the reference's dense matrix -- u8 allele indices plus either the
// packed missing bitmap of DenseGenotypeMatrix (stats.rs:1298-1302), in-band missingness (the
// negative cells of the int8 array Population.from_numpy receives, lib.rs:1168-1199) or no
// missing data at all -- into full-row bit words: per row rw = ceil(stride / 32) u32 words of
// allele bits (cell is called and non-zero) and rw words of called bits, bit c & 31 of word c >> 5
// for cell c.  That is what fm_ingest_rows_packed / fm_matrix_create_packed take: 0.25 B per
// genotype over PCIe instead of 1.125 B.  A caller that parses text (process.rs:2602-2660) can
// write these words directly and never build the u8 matrix; fm_pack_rows is for callers that
// already hold one (lib.rs:1135-1227).
//
// Plain host code: several threads over row ranges, AVX2 compare + movemask when the CPU has it
// (32 cells per instruction), a 64-bit SWAR gather otherwise.  No CUDA calls.
#include "../../include/ferromic_gpu.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define FM_X86 1
#else
#define FM_X86 0
#endif

namespace {

// 8 cells -> 8 bits (bit i = cell i satisfies the predicate).  t has bit 7 of every byte set where
// the predicate holds; the multiply gathers those bits into the top byte.
inline uint32_t gather8(uint64_t t) { return (uint32_t)((((t >> 7) & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56); }
inline uint64_t nonzero_hi(uint64_t x) {  // bit 7 of every non-zero byte
    return (((x & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full) | x) & 0x8080808080808080ull;
}

// 32 consecutive bits of the LSB-first bitmap starting at absolute bit index `bit`
inline uint32_t bitmap_u32(const uint64_t *bm, uint64_t bit, uint64_t n_words_total) {
    const uint64_t w = bit >> 6;
    const uint32_t sh = (uint32_t)(bit & 63);
    uint64_t lo = bm[w] >> sh;
    if (sh > 32 && w + 1 < n_words_total) lo |= bm[w + 1] << (64 - sh);
    return (uint32_t)lo;
}

void pack_range_generic(const uint8_t *rows, const uint64_t *missing, uint64_t n_bitmap_words, int mode,
                        size_t first_row, size_t r_lo, size_t r_hi, size_t stride, uint32_t *abits,
                        uint32_t *cbits) {
    const size_t rw = (stride + 31) / 32;
    for (size_t r = r_lo; r < r_hi; ++r) {
        const uint8_t *src = rows + r * stride;
        uint32_t *a = abits + r * rw;
        uint32_t *c = cbits ? cbits + r * rw : nullptr;
        const uint64_t bit0 = (uint64_t)(first_row + r) * stride;
        for (size_t w = 0; w < rw; ++w) {
            const size_t c0 = w * 32;
            const size_t n = std::min<size_t>(32, stride - c0);
            uint32_t nz = 0, neg = 0;
            size_t i = 0;
            for (; i + 8 <= n; i += 8) {
                uint64_t x;
                std::memcpy(&x, src + c0 + i, 8);
                nz |= gather8(nonzero_hi(x)) << i;
                neg |= gather8(x & 0x8080808080808080ull) << i;
            }
            for (; i < n; ++i) {
                const uint8_t b = src[c0 + i];
                nz |= (uint32_t)(b != 0) << i;
                neg |= (uint32_t)(b >> 7) << i;
            }
            const uint32_t valid = n == 32 ? 0xffffffffu : ((1u << n) - 1u);
            uint32_t called = valid;
            if (mode == FM_MISSING_BITMAP)
                called = ~bitmap_u32(missing, bit0 + c0, n_bitmap_words) & valid;
            else if (mode == FM_MISSING_IN_BAND)
                called = ~neg & valid;
            a[w] = nz & called;
            if (c) c[w] = called;
        }
    }
}

// 64 consecutive bits of the bitmap starting at absolute bit index `bit`, branch-free: two unaligned 8-byte
// reads at the bit's byte.  `safe_bytes` = bytes of the bitmap that may be read; the caller keeps the last
// words of the matrix on the careful path.
inline uint64_t bitmap_u64_fast(const uint8_t *bm8, uint64_t bit) {
    const uint64_t byte = bit >> 3;
    const uint32_t s = (uint32_t)(bit & 7);
    uint64_t lo, hi;
    std::memcpy(&lo, bm8 + byte, 8);
    std::memcpy(&hi, bm8 + byte + 8, 8);
    return (lo >> s) | ((hi << 1) << (63 - s));
}

#if FM_X86
// AVX-512BW: one vptestmb turns 64 cells into 64 allele bits, vpmovb2m gives the in-band sign bits.
__attribute__((target("avx512f,avx512bw"))) void pack_range_avx512(const uint8_t *rows, const uint64_t *missing,
                                                                   uint64_t n_bitmap_words, int mode, size_t first_row,
                                                                   size_t r_lo, size_t r_hi, size_t stride,
                                                                   uint32_t *abits, uint32_t *cbits) {
    const size_t rw = (stride + 31) / 32;
    const size_t n64 = stride / 64;
    const uint8_t *bm8 = reinterpret_cast<const uint8_t *>(missing);
    const uint64_t bm_bytes = n_bitmap_words * 8;
    for (size_t r = r_lo; r < r_hi; ++r) {
        const uint8_t *src = rows + r * stride;
        uint32_t *a = abits + r * rw;
        uint32_t *c = cbits ? cbits + r * rw : nullptr;
        const uint64_t bit0 = (uint64_t)(first_row + r) * stride;
        // the fast bitmap read touches 16 bytes from the bit's byte: stay clear of the buffer's end
        const bool bm_fast = mode != FM_MISSING_BITMAP || ((bit0 + stride) >> 3) + 16 <= bm_bytes;
        size_t k = 0;
        if (bm_fast) {
            for (; k < n64; ++k) {
                const __m512i x = _mm512_loadu_si512(src + k * 64);
                const uint64_t nz = _mm512_test_epi8_mask(x, x);
                uint64_t called = ~0ull;
                if (mode == FM_MISSING_BITMAP)
                    called = ~bitmap_u64_fast(bm8, bit0 + k * 64);
                else if (mode == FM_MISSING_IN_BAND)
                    called = ~(uint64_t)_mm512_movepi8_mask(x);
                const uint64_t av = nz & called;
                a[2 * k] = (uint32_t)av;
                a[2 * k + 1] = (uint32_t)(av >> 32);
                if (c) {
                    c[2 * k] = (uint32_t)called;
                    c[2 * k + 1] = (uint32_t)(called >> 32);
                }
            }
        }
        // what is left of the row (everything, near the end of the bitmap): 32 cells at a time, scalar tail
        for (size_t w = 2 * k; w < rw; ++w) {
            const size_t c0 = w * 32, n = std::min<size_t>(32, stride - c0);
            uint32_t nz = 0, neg = 0;
            for (size_t i = 0; i < n; ++i) {
                nz |= (uint32_t)(src[c0 + i] != 0) << i;
                neg |= (uint32_t)(src[c0 + i] >> 7) << i;
            }
            const uint32_t valid = n == 32 ? 0xffffffffu : ((1u << n) - 1u);
            uint32_t called = valid;
            if (mode == FM_MISSING_BITMAP)
                called = ~bitmap_u32(missing, bit0 + c0, n_bitmap_words) & valid;
            else if (mode == FM_MISSING_IN_BAND)
                called = ~neg & valid;
            a[w] = nz & called;
            if (c) c[w] = called;
        }
    }
}

__attribute__((target("avx2"))) void pack_range_avx2(const uint8_t *rows, const uint64_t *missing,
                                                     uint64_t n_bitmap_words, int mode, size_t first_row, size_t r_lo,
                                                     size_t r_hi, size_t stride, uint32_t *abits, uint32_t *cbits) {
    const size_t rw = (stride + 31) / 32;
    const size_t full = stride / 32;
    const __m256i zero = _mm256_setzero_si256();
    for (size_t r = r_lo; r < r_hi; ++r) {
        const uint8_t *src = rows + r * stride;
        uint32_t *a = abits + r * rw;
        uint32_t *c = cbits ? cbits + r * rw : nullptr;
        const uint64_t bit0 = (uint64_t)(first_row + r) * stride;
        for (size_t w = 0; w < full; ++w) {
            const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + w * 32));
            const uint32_t nz = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, zero));
            uint32_t called = 0xffffffffu;
            if (mode == FM_MISSING_BITMAP)
                called = ~bitmap_u32(missing, bit0 + w * 32, n_bitmap_words);
            else if (mode == FM_MISSING_IN_BAND)
                called = ~(uint32_t)_mm256_movemask_epi8(x);  // sign bit = negative int8 = missing
            a[w] = nz & called;
            if (c) c[w] = called;
        }
    }
    if (full < rw)  // the partial last word of every row
        for (size_t r = r_lo; r < r_hi; ++r) {
            const size_t c0 = full * 32, n = stride - c0;
            const uint8_t *src = rows + r * stride + c0;
            uint32_t nz = 0, neg = 0;
            for (size_t i = 0; i < n; ++i) {
                nz |= (uint32_t)(src[i] != 0) << i;
                neg |= (uint32_t)(src[i] >> 7) << i;
            }
            const uint32_t valid = (1u << n) - 1u;
            uint32_t called = valid;
            if (mode == FM_MISSING_BITMAP)
                called = ~bitmap_u32(missing, (uint64_t)(first_row + r) * stride + c0, n_bitmap_words) & valid;
            else if (mode == FM_MISSING_IN_BAND)
                called = ~neg & valid;
            abits[r * rw + full] = nz & called;
            if (cbits) cbits[r * rw + full] = called;
        }
}
#endif

}  // namespace

// Returns nullptr on success or a static message for FM_ERR_INVALID_ARG (the C-ABI wrapper fm_pack_rows lives in
// fm_gpu.cu so that fm_last_error() covers it).
const char *fm_host_pack_rows(const uint8_t *rows, const uint64_t *missing_whole_or_null, int missing_mode,
                              size_t first_row, size_t n_rows, size_t n_total_rows, size_t stride,
                              uint32_t *allele_bits, uint32_t *called_bits_or_null, int n_threads) {
    if (missing_mode != FM_MISSING_NONE && missing_mode != FM_MISSING_BITMAP && missing_mode != FM_MISSING_IN_BAND)
        return "fm_pack_rows: bad missing_mode";
    if (n_rows == 0 || stride == 0) return nullptr;
    if (!rows || !allele_bits) return "fm_pack_rows: rows / allele_bits is NULL";
    if (missing_mode == FM_MISSING_BITMAP && !missing_whole_or_null)
        return "fm_pack_rows: FM_MISSING_BITMAP needs the whole matrix's bitmap";
    if (missing_mode != FM_MISSING_NONE && !called_bits_or_null)
        return "fm_pack_rows: called_bits is required when the matrix has missing data";
    if (first_row > n_total_rows || n_rows > n_total_rows - first_row) return "fm_pack_rows: row range outside the matrix";
    const uint64_t n_bitmap_words = ((uint64_t)n_total_rows * stride + 63) / 64;
    unsigned T = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
    T = (unsigned)std::min<size_t>(std::min<unsigned>(T, 64), std::max<size_t>(1, n_rows * stride / (1u << 20)));
#if FM_X86
    const bool avx2 = __builtin_cpu_supports("avx2");
    const bool avx512 = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f") && !getenv("FM_PACK_NO_AVX512");
#else
    const bool avx2 = false;
#endif
    auto work = [&](size_t lo, size_t hi) {
#if FM_X86
        if (avx512) {
            pack_range_avx512(rows, missing_whole_or_null, n_bitmap_words, missing_mode, first_row, lo, hi, stride,
                              allele_bits, called_bits_or_null);
            return;
        }
        if (avx2) {
            pack_range_avx2(rows, missing_whole_or_null, n_bitmap_words, missing_mode, first_row, lo, hi, stride,
                            allele_bits, called_bits_or_null);
            return;
        }
#endif
        pack_range_generic(rows, missing_whole_or_null, n_bitmap_words, missing_mode, first_row, lo, hi, stride,
                           allele_bits, called_bits_or_null);
    };
    if (T <= 1) {
        work(0, n_rows);
        return nullptr;
    }
    std::vector<std::thread> pool;
    const size_t per = (n_rows + T - 1) / T;
    try {
        for (unsigned t = 1; t < T; ++t) {
            const size_t lo = std::min(n_rows, per * t), hi = std::min(n_rows, per * (t + 1));
            if (hi > lo) pool.emplace_back(work, lo, hi);
        }
    } catch (...) {  // thread creation failed: finish everything on this thread
        for (auto &th : pool) th.join();
        work(0, n_rows);
        return nullptr;
    }
    work(0, std::min(n_rows, per));
    for (auto &th : pool) th.join();
    return nullptr;
}

// Sparse-missing variant (include/ferromic_gpu.h "packed rows, sparse missing list"): the allele bits as above, and
// instead of a called plane the columns of the missing cells of every row, in ascending order, as a CSR list
// (row_start[r] .. row_start[r + 1], relative to this call).  With 1 % missing cells that is ~0.15 bits per
// genotype instead of 1.  Works on top of the dense packer: every thread packs a block of rows into a small
// scratch called plane, walks its zero bits, and the per-thread lists are concatenated in row order.
const char *fm_host_pack_rows_sparse(const uint8_t *rows, const uint64_t *missing_whole_or_null, int missing_mode,
                                     size_t first_row, size_t n_rows, size_t n_total_rows, size_t stride,
                                     uint32_t *allele_bits, uint64_t *row_start, void *missing_cols, size_t capacity,
                                     int col_bytes, int n_threads, size_t *needed) {
    if (needed) *needed = 0;
    if (col_bytes != 1 && col_bytes != 2 && col_bytes != 4) return "fm_pack_rows_sparse: col_bytes must be 1 (gap code), 2 or 4";
    if (col_bytes == 2 && stride > 65536) return "fm_pack_rows_sparse: 16-bit columns need a row stride <= 65536";
    if (!row_start) return "fm_pack_rows_sparse: row_start is NULL";
    row_start[0] = 0;
    if (n_rows == 0 || stride == 0) {
        for (size_t r = 0; r < n_rows; ++r) row_start[r + 1] = 0;
        return nullptr;
    }
    if (missing_mode == FM_MISSING_NONE) {  // nothing is missing: plain allele bits, empty lists
        const char *e = fm_host_pack_rows(rows, nullptr, FM_MISSING_NONE, first_row, n_rows, n_total_rows, stride,
                                          allele_bits, nullptr, n_threads);
        for (size_t r = 0; r < n_rows; ++r) row_start[r + 1] = 0;
        return e;
    }
    if (!rows || !allele_bits) return "fm_pack_rows_sparse: rows / allele_bits is NULL";
    if (missing_mode == FM_MISSING_BITMAP && !missing_whole_or_null)
        return "fm_pack_rows_sparse: FM_MISSING_BITMAP needs the whole matrix's bitmap";
    if (first_row > n_total_rows || n_rows > n_total_rows - first_row) return "fm_pack_rows_sparse: row range outside the matrix";
    const size_t rw = (stride + 31) / 32;
    unsigned T = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
    T = (unsigned)std::min<size_t>(std::min<unsigned>(T, 64), std::max<size_t>(1, n_rows * stride / (1u << 20)));
    const size_t per = (n_rows + T - 1) / T;
    std::vector<std::vector<uint32_t>> lists(T);
    std::vector<std::vector<uint8_t>> gaps(T);  // col_bytes == 1: the rows' gap codes (FM gap code, ferromic_gpu.h)
    std::vector<std::vector<uint32_t>> counts(T);
    auto work = [&](unsigned t) {
        const size_t lo = std::min(n_rows, per * t), hi = std::min(n_rows, per * (t + 1));
        std::vector<uint32_t> &L = lists[t];
        std::vector<uint8_t> &Gp = gaps[t];
        std::vector<uint32_t> &Cn = counts[t];
        Cn.resize(hi - lo);
        const size_t block = std::max<size_t>(1, (64u << 10) / (rw * 4));  // scratch called plane: ~64 KB, stays in L2
        std::vector<uint32_t> cb(block * rw);
        for (size_t b0 = lo; b0 < hi; b0 += block) {
            const size_t b1 = std::min(hi, b0 + block);
            // dense packer on this block (single-threaded inside: we are already on a worker thread)
            fm_host_pack_rows(rows + b0 * stride, missing_whole_or_null, missing_mode, first_row + b0, b1 - b0,
                              n_total_rows, stride, allele_bits + b0 * rw, cb.data(), 1);
            for (size_t r = b0; r < b1; ++r) {
                const uint32_t *c = cb.data() + (r - b0) * rw;
                uint32_t n = 0;
                int64_t prev = -1;  // gap code: last emitted (or skipped-to) column
                for (size_t w = 0; w < rw; ++w) {
                    const size_t cells = std::min<size_t>(32, stride - w * 32);
                    uint32_t miss = ~c[w] & (cells == 32 ? 0xffffffffu : ((1u << cells) - 1u));
                    while (miss) {
                        const uint32_t bit = (uint32_t)__builtin_ctz(miss);
                        miss &= miss - 1;
                        const uint32_t col = (uint32_t)(w * 32 + bit);
                        if (col_bytes == 1) {
                            int64_t gap = (int64_t)col - prev;  // >= 1
                            while (gap > 255) {                 // 255: skip 255 columns, no cell
                                Gp.push_back(255);
                                gap -= 255;
                                ++n;
                            }
                            Gp.push_back((uint8_t)(gap - 1));   // b < 255: the cell b + 1 columns further on
                            prev = col;
                            ++n;
                        } else {
                            L.push_back(col);
                            ++n;
                        }
                    }
                }
                Cn[r - lo] = n;
            }
        }
    };
    {
        std::vector<std::thread> pool;
        try {
            for (unsigned t = 1; t < T; ++t) pool.emplace_back(work, t);
        } catch (...) {
            for (auto &th : pool) th.join();
            return "fm_pack_rows_sparse: could not start worker threads";
        }
        work(0);
        for (auto &th : pool) th.join();
    }
    // row starts (relative to this call) and the concatenation of the per-thread lists
    std::vector<uint64_t> base(T + 1, 0);
    size_t r = 0;
    uint64_t acc = 0;
    for (unsigned t = 0; t < T; ++t) {
        base[t] = acc;
        for (uint32_t n : counts[t]) {
            acc += n;
            row_start[++r] = acc;
        }
    }
    base[T] = acc;
    if (needed) *needed = (size_t)acc;
    if (acc > capacity || (acc && !missing_cols)) return "fm_pack_rows_sparse: missing_cols capacity too small (see needed)";
    auto copy = [&](unsigned t) {
        const std::vector<uint32_t> &L = lists[t];
        if (col_bytes == 1) {
            if (!gaps[t].empty()) std::memcpy(static_cast<uint8_t *>(missing_cols) + base[t], gaps[t].data(), gaps[t].size());
        } else if (col_bytes == 4) {
            std::memcpy(static_cast<uint32_t *>(missing_cols) + base[t], L.data(), L.size() * 4);
        } else {
            uint16_t *dst = static_cast<uint16_t *>(missing_cols) + base[t];
            for (size_t i = 0; i < L.size(); ++i) dst[i] = (uint16_t)L[i];
        }
    };
    {
        std::vector<std::thread> pool;
        try {
            for (unsigned t = 1; t < T; ++t) pool.emplace_back(copy, t);
        } catch (...) {
            for (auto &th : pool) th.join();
            for (unsigned t = 0; t < T; ++t) copy(t);
            return nullptr;
        }
        copy(0);
        for (auto &th : pool) th.join();
    }
    return nullptr;
}

// test hook: the portable SWAR path, whatever the CPU supports
const char *fm_host_pack_rows_generic(const uint8_t *rows, const uint64_t *missing, int mode, size_t first_row,
                                      size_t n_rows, size_t n_total_rows, size_t stride, uint32_t *abits,
                                      uint32_t *cbits) {
    if (n_rows == 0 || stride == 0) return nullptr;
    pack_range_generic(rows, missing, ((uint64_t)n_total_rows * stride + 63) / 64, mode, first_row, 0, n_rows, stride,
                       abits, cbits);
    return nullptr;
}
