// fm_comm.cuh -- the path's only exchange step, done over NVLink peer memory instead of a
// library collective: every rank owns a small "mailbox" in its HBM that all peers map
// (cudaIpc between processes, direct pointers inside one process).  One 128-thread kernel per
// exchange
//   1. folds this rank's per-super-batch partials into its region totals (fixed order),
//   2. stores the totals into slot [rank] of EVERY peer's mailbox with plain P2P stores,
//      fences at system scope and publishes the step number in the peer's flag word,
//   3. waits (acquire loads on its own mailbox) until every peer's flag reached the step,
//   4. adds the slots in rank order -> merged totals, identical bits on every rank.
// This is an all-gather + rank-ordered sum fused into the final reduction kernel of the step:
// no host round trip, no extra launch, ~2-3 us on NVSwitch, and the step stays fully
// asynchronous on its stream.  The mailbox is double-buffered by step parity: a peer can run at
// most one exchange ahead because it needs this rank's flag to finish the next one.
#pragma once
#include "fm_device.cuh"

namespace fm {

constexpr uint32_t kCommMaxRanks = 16;
constexpr uint32_t kCommMaxValues = 2048;  // 8-byte words per rank per exchange
constexpr uint32_t kCommStageWords = 1024;  // staging of super-batch partial rows for the fused fold

struct CommMailbox {                       // lives in device memory of its owner
    unsigned long long flags[kCommMaxRanks * 16];  // flags[r*16]: last step rank r has fully written (own line);
                                                   // flags[r*16 + 8]: rank r is closing (fm_k_comm_goodbye)
    unsigned long long data[2][kCommMaxRanks][kCommMaxValues];
};

struct CommFold {                // optional step 1: totals[c] = sum over super-batches, in order
    const double *sd;            // [n_super][nd]
    const unsigned long long *su;  // [n_super][nu]
    uint32_t n_super, nd, nu;
};

struct CommParams {
    CommMailbox *peers[kCommMaxRanks];  // this rank's view of every rank's mailbox (peers[rank] = own)
    uint32_t rank, world;
    unsigned long long step;            // 1, 2, 3, ... identical on all ranks
    uint32_t n_words;                   // 8-byte words contributed per rank
    uint32_t n_double;                  // first n_double words are doubles (summed as FP64), the rest u64
    const unsigned long long *local;    // [n_words] when no fold is requested
    CommFold fold[4];                   // up to 4 folded column groups, concatenated: doubles first? no --
                                        // each fold contributes nd doubles then nu integers (see host)
    uint32_t n_fold;
    unsigned long long *gathered;       // [world][n_words] or nullptr
    unsigned long long *merged;         // [n_words] rank-ordered sums or nullptr
    uint32_t *status;                   // 0 ok, 1 timeout waiting for a peer
    unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long fm_ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fm_st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long fm_ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fm_st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long fm_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// The fold of fm_k_comm_exchange evaluated on the host: lane-strided row sums + xor butterfly, bit for bit.
inline double fm_comm_fold_host(const double *rows, size_t n_super, size_t stride, size_t col) {
    double lane[32];
    const size_t R = (n_super + 31) / 32;
    for (size_t l = 0; l < 32; ++l) {
        const size_t r0 = l * R < n_super ? l * R : n_super, r1 = r0 + R < n_super ? r0 + R : n_super;
        double acc = 0.0;
        for (size_t r = r0; r < r1; ++r) acc += rows[r * stride + col];
        lane[l] = acc;
    }
    for (int o = 16; o > 0; o >>= 1) {
        double next[32];
        for (int l = 0; l < 32; ++l) next[l] = lane[l] + lane[l ^ o];
        for (int l = 0; l < 32; ++l) lane[l] = next[l];
    }
    return lane[0];
}

// is word i of the message a double?  (folds: each contributes nd doubles followed by nu integers)
__device__ __forceinline__ bool fm_comm_is_double(const CommParams &P, uint32_t i) {
    if (P.n_fold == 0) return i < P.n_double;
    uint32_t base = 0;
    for (uint32_t f = 0; f < P.n_fold; ++f) {
        const uint32_t w = P.fold[f].nd + P.fold[f].nu;
        if (i < base + w) return (i - base) < P.fold[f].nd;
        base += w;
    }
    return false;
}

__global__ void __launch_bounds__(128)
fm_k_comm_exchange(const CommParams P) {
    __shared__ unsigned long long vals[kCommMaxValues];
    __shared__ uint32_t timed_out;
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t buf = (uint32_t)(P.step & 1ull);
    if (tid == 0) timed_out = 0;
    // ---- 1. this rank's contribution
    if (P.n_fold == 0) {
        for (uint32_t i = tid; i < P.n_words; i += nt) vals[i] = P.local[i];
    } else {
        // warp f folds column group f.  Fixed shape (fm_comm_fold_host below is the same arithmetic on the host):
        // lane l adds rows [l*R, (l+1)*R), R = ceil(n_super / 32), in row order for every column -- independent
        // loads, so their latencies overlap -- then the 32 lane sums are combined by the xor butterfly of
        // fm_warp_sum.  A few microseconds for hundreds of super-batches; the sequential walk took 0.14 us per row.
        const uint32_t warp = tid >> 5, lane = tid & 31;
        if (warp < P.n_fold) {
            uint32_t base = 0;
            for (uint32_t f = 0; f < warp; ++f) base += P.fold[f].nd + P.fold[f].nu;
            const CommFold &F = P.fold[warp];
            const uint32_t R = (F.n_super + 31) / 32;
            const uint32_t r0 = min(F.n_super, lane * R), r1 = min(F.n_super, r0 + R);
            for (uint32_t col = 0; col < F.nd; ++col) {
                double acc = 0.0;
                for (uint32_t r = r0; r < r1; ++r) acc += F.sd[(size_t)r * F.nd + col];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) vals[base + col] = (unsigned long long)__double_as_longlong(acc);
            }
            for (uint32_t col = 0; col < F.nu; ++col) {
                unsigned long long acc = 0;
                for (uint32_t r = r0; r < r1; ++r) acc += F.su[(size_t)r * F.nu + col];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) vals[base + F.nd + col] = acc;
            }
        }
    }
    __syncthreads();
    // ---- 2. push into every mailbox (own included), then publish the step
    for (uint32_t p = 0; p < P.world; ++p) {
        unsigned long long *dst = P.peers[p]->data[buf][P.rank];
        for (uint32_t i = tid; i < P.n_words; i += nt) dst[i] = vals[i];
    }
    // slot stores -> CTA barrier -> release store of the flag by the publishing thread: the
    // barrier orders every thread's stores before the publisher's release (fence cumulativity, the
    // pattern of a grid-wide barrier), so one system-scope release per flag is enough
    __syncthreads();
    if (tid < P.world) fm_st_release_sys(&P.peers[tid]->flags[P.rank * 16], P.step);
    // ---- 3. wait for every rank's contribution to land in OUR mailbox
    CommMailbox *mine = P.peers[P.rank];
    if (tid < P.world) {
        const unsigned long long t0 = fm_globaltimer();
        while (fm_ld_relaxed_sys(&mine->flags[tid * 16]) < P.step) {  // cheap polling, one fence after
            if (P.timeout_ns && fm_globaltimer() - t0 > P.timeout_ns) {  // 0: wait for ever, like a collective
                timed_out = 1;
                break;
            }
            __nanosleep(32);
        }
        (void)fm_ld_acquire_sys(&mine->flags[tid * 16]);  // acquire: slot loads below are ordered after the flag
    }
    __syncthreads();
    // the status also travels right behind the merged words, so that one copy fetches both
    if (timed_out) {
        if (tid == 0) {
            *P.status = 1;
            if (P.merged) P.merged[P.n_words] = 1ull;
        }
        return;
    }
    if (tid == 0) {
        *P.status = 0;
        if (P.merged) P.merged[P.n_words] = 0ull;
    }
    // ---- 4. gather / rank-ordered sum
    for (uint32_t i = tid; i < P.n_words; i += nt) {
        const bool is_d = fm_comm_is_double(P, i);
        double accd = 0.0;
        unsigned long long accu = 0;
        for (uint32_t r = 0; r < P.world; ++r) {
            const unsigned long long w = __ldcg(&mine->data[buf][r][i]);  // peer-written: bypass L1
            if (P.gathered) P.gathered[(size_t)r * P.n_words + i] = w;
            if (is_d)
                accd += __longlong_as_double((long long)w);
            else
                accu += w;
        }
        if (P.merged) P.merged[i] = is_d ? (unsigned long long)__double_as_longlong(accd) : accu;
    }
}

// Closing handshake of fm_comm_destroy: every rank tells every peer "I will never write into your mailbox
// again" and waits for the same word from all of them before its own mailbox is freed -- without it a slower
// peer could still be issuing P2P stores into memory that has gone back to the driver.  status: 0 ok,
// 1 a peer did not close within the timeout (the mailbox is then leaked on purpose, never freed under a writer).
__global__ void __launch_bounds__(32)
fm_k_comm_goodbye(CommMailbox *const *peers_unused, const CommParams P) {
    (void)peers_unused;
    const uint32_t tid = threadIdx.x;
    if (tid < P.world) fm_st_release_sys(&P.peers[tid]->flags[P.rank * 16 + 8], 1ull);
    CommMailbox *mine = P.peers[P.rank];
    bool late = false;
    if (tid < P.world) {
        const unsigned long long t0 = fm_globaltimer();
        while (fm_ld_relaxed_sys(&mine->flags[tid * 16 + 8]) == 0ull) {
            if (P.timeout_ns && fm_globaltimer() - t0 > P.timeout_ns) {
                late = true;
                break;
            }
            __nanosleep(256);
        }
    }
    if (__any_sync(0xffffffffu, late) && tid == 0) *P.status = 1;
}

}  // namespace fm
