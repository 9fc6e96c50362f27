// Contig-in-truth offset: leftmost exact occurrence of every contig in its segment's truth.
#pragma once
#include "bs_place.cuh"

namespace bs {

// ------------------------------------------------------------------------------------------
// contig-in-truth offset (upstream lib/BreakageScorer.cpp:273-274): leftmost exact occurrence
// of the whole contig in its segment's truth, assigned only if at least one read was placed.
//
// k_startpos_build: one thread per contig inserts the contig's SEED (first 32 bases) into the
// seed table of its group (up to SP_GROUP contigs of one segment) in global memory: the slot is
// claimed by compare-and-swap with linear probing, contigs with equal seeds are chained.
// k_startpos_index: the truth streams past the group's table, copied into shared memory; the
// thread that finds a seed hit verifies the whole contig on packed words itself; the leftmost
// verified position wins through atomicMin.  O(L_truth + sum L_c) per segment instead of
// O(C * L_truth).  Contigs without a packable seed (shorter than 32, or a non-ACGT byte in it)
// are left to k_startpos, which also turns the atomicMin scratch into the final values.
// The table is built and consumed in different kernels, and the scan uses no warp-level
// primitive: an earlier version (table built with a compare-and-swap loop in shared memory by
// the scanning block itself, warp-cooperative verification behind ballots and shuffles) missed
// about 1 % of the hits of a 1000-contig segment on the GPU, depending on how it was compiled.
// ------------------------------------------------------------------------------------------

struct StartposArgs {
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const uint8_t *ctg_chars;
    const int32_t *ctg_seg;
    const int64_t *seg_contig_start;  // [S+1]
    const int64_t *tr_off;
    const int64_t *tr_woff;
    const uint64_t *tr_words;
    const uint32_t *tr_mask;
    const uint8_t *tr_chars;
    const int32_t *total;
    int64_t n_contigs;
    int32_t n_seg;
    int32_t splits;       // blocks per segment (each scans a slice of the truth)
    uint32_t *best;       // [C] leftmost verified position, POS_INF on entry (k_startpos_index writes, k_startpos reads)
    int32_t *startpos;    // optional: truth.find(contig) if a read was placed, else 0 (upstream semantics)
    int32_t *exact;       // optional: truth.find(contig) for EVERY contig (-1: not a substring); feeds the edit distance
    int32_t search_all;   // search also for contigs without a placed read (needed for `exact`)
    // seed tables, one per group of contigs (startpos_group): all-ones keys and zero heads on entry
    unsigned long long *tab_key;  // [groups][SP_SLOTS + 1] seed owning the slot
    int32_t *tab_head;            // [groups][SP_SLOTS + 1] group-local contig index + 1 of the first contig with that seed
    int32_t *tab_next;            // [C] next contig of the group with the same seed
};

constexpr int SP_SLOTS = 2048;  // seed table slots per group; slot SP_SLOTS is reserved for the all-ones seed (= the empty marker)
constexpr int SP_GROUP = 1024;  // contigs per group
constexpr int SP_BITMAP_WORDS = 65536 / 32;  // k_startpos_index<true>: one bit per 8-base prefix of a seed
constexpr int SP_RUN = 8;           // consecutive truth positions per thread and scan step of k_startpos_index (divides 32)
constexpr int SP_VERIFY_BATCH = 4;  // contig words compared per verification step of k_startpos_index

// table of the k-th group of a segment whose contigs start at c0: distinct for every (segment, k)
// without a prefix sum, and below n_contigs / SP_GROUP + n_seg + 1
BS_HD int64_t startpos_group(int64_t c0, int64_t seg, int64_t k) { return c0 / SP_GROUP + seg + k; }
BS_HD int64_t startpos_group_count(int64_t n_contigs, int64_t n_seg) { return n_contigs / SP_GROUP + n_seg + 1; }

// a contig the seed index can look for: 32 valid leading bases, fits in the truth, had a read placed
__device__ __forceinline__ bool startpos_indexable(const StartposArgs &a, int64_t c, int64_t LT) {
    const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
    return L >= 32 && L <= LT && (a.search_all || a.total[c] != 0) && a.ctg_mask[a.ctg_woff[c]] == 0;
}

__global__ void k_startpos_build(StartposArgs a) {
    const unsigned long long EMPTY = ~0ull;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < a.n_contigs; c += stride) {
        const int64_t seg = a.ctg_seg[c];
        if (!startpos_indexable(a, c, a.tr_off[seg + 1] - a.tr_off[seg])) continue;
        const int64_t c0 = a.seg_contig_start[seg];
        const int64_t tab = startpos_group(c0, seg, (c - c0) / SP_GROUP) * (SP_SLOTS + 1);
        const unsigned long long key = a.ctg_words[a.ctg_woff[c]];
        uint32_t h = SP_SLOTS;
        if (key != EMPTY) {
            h = seed_hash(key) & (SP_SLOTS - 1);
            for (int probe = 0; probe < SP_SLOTS; probe++) {  // at most SP_GROUP of the SP_SLOTS slots are ever taken
                const unsigned long long prev = atomicCAS(&a.tab_key[tab + h], EMPTY, key);
                if (prev == EMPTY || prev == key) break;
                h = (h + 1) & (SP_SLOTS - 1);
            }
        }
        a.tab_next[c] = atomicExch(&a.tab_head[tab + h], (int32_t)((c - c0) % SP_GROUP) + 1);
    }
}

// group-local index + 1 of the first contig whose seed is `seed` (0: none) in a group's table in shared memory
__device__ __forceinline__ int startpos_probe(const unsigned long long *s_key, const int32_t *s_head, unsigned long long seed) {
    const unsigned long long EMPTY = ~0ull;
    if (seed == EMPTY) return s_head[SP_SLOTS];
    uint32_t h = seed_hash(seed) & (SP_SLOTS - 1);
#pragma unroll 1
    for (int probe = 0; probe < SP_SLOTS; probe++) {
        const unsigned long long k = s_key[h];
        if (k == EMPTY) break;
        if (k == seed) return s_head[h];
        h = (h + 1) & (SP_SLOTS - 1);
    }
    return 0;
}

// Seed hit at truth position q for the chain of contigs starting at group-local index ci_head - 1: the calling
// thread compares every contig of the chain with the truth on packed words (thread-local, no warp primitive)
// and keeps the leftmost verified position.
__device__ __forceinline__ void startpos_verify(const StartposArgs &a, int ci_head, const int32_t *s_next, int64_t g0, int64_t q, int64_t LT,
                                            const uint64_t *tw, const uint32_t *tm, const uint8_t *tc) {
    const int64_t idx = q >> 5;
    const uint32_t o = (uint32_t)(q & 31);
    for (int ci = ci_head; ci != 0; ci = s_next[ci - 1]) {
        const int64_t c = g0 + ci - 1;
        const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
        if (q + L > LT) continue;
        if ((uint32_t)q >= *(volatile uint32_t *)&a.best[c]) continue;  // a position further left is already known
        const uint64_t *cw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *cm = a.ctg_mask + a.ctg_woff[c];
        const int64_t nw = (L + 31) >> 5;
        bool ok = true, any_invalid = false;
        uint64_t t_lo = tw[idx];
        uint32_t m_lo = tm[idx];
        // SP_VERIFY_BATCH words per step without an early exit in between: the loads of a step are
        // independent and in flight together (one thread walks a 50 kb contig in ~400 round trips
        // to L2 instead of ~1600; a batch of 4 keeps the kernel at 64 registers, four blocks per SM)
        for (int64_t j0 = 0; ok && j0 < nw; j0 += SP_VERIFY_BATCH) {
            uint64_t t_hi[SP_VERIFY_BATCH], cwj[SP_VERIFY_BATCH];
            uint32_t m_hi[SP_VERIFY_BATCH], cmj[SP_VERIFY_BATCH];
#pragma unroll
            for (int u = 0; u < SP_VERIFY_BATCH; u++) {
                const bool in = j0 + u < nw;  // (the two pad words of a sequence keep idx + j + 1 inside for j < nw)
                t_hi[u] = in ? tw[idx + j0 + u + 1] : 0;
                m_hi[u] = in ? tm[idx + j0 + u + 1] : 0;
                cwj[u] = in ? cw[j0 + u] : 0;
                cmj[u] = in ? cm[j0 + u] : 0;
            }
            uint64_t diff = 0;
            uint32_t mdiff = 0, minv = 0;
#pragma unroll
            for (int u = 0; u < SP_VERIFY_BATCH; u++) {
                if (j0 + u < nw) {
                    const int64_t left = L - 32 * (j0 + u);
                    const int rem = left < 32 ? (int)left : 32;
                    // byte equality == equal 2-bit codes AND equal validity AND equal raw bytes where invalid
                    diff |= (window64(t_lo, t_hi[u], o) ^ cwj[u]) & keep_bases(rem);
                    mdiff |= (window32(m_lo, m_hi[u], o) ^ cmj[u]) & keep_bits(rem);
                    minv |= cmj[u] & keep_bits(rem);
                    t_lo = t_hi[u];
                    m_lo = m_hi[u];
                }
            }
            ok = diff == 0 && mdiff == 0;
            any_invalid |= minv != 0;
        }
        if (ok && any_invalid) {
            const uint8_t *cc = a.ctg_chars + a.ctg_off[c];
            for (int64_t i = 0; ok && i < L; i++) ok = tc[q + i] == cc[i];
        }
        if (ok) atomicMin(&a.best[c], (uint32_t)q);
    }
}

template <bool BITMAP>
__global__ void __launch_bounds__(256, 4) k_startpos_index(StartposArgs a) {
    __shared__ unsigned long long s_key[SP_SLOTS + 1];
    __shared__ int32_t s_head[SP_SLOTS + 1];
    __shared__ int32_t s_next[SP_GROUP];
    __shared__ uint32_t s_bits[BITMAP ? SP_BITMAP_WORDS : 1];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int seg = blockIdx.x / a.splits, part = blockIdx.x % a.splits;
    const int64_t LT = a.tr_off[seg + 1] - a.tr_off[seg];
    const uint64_t *tw = a.tr_words + a.tr_woff[seg];
    const uint32_t *tm = a.tr_mask + a.tr_woff[seg];
    const uint8_t *tc = a.tr_chars + a.tr_off[seg];
    const int64_t nq = LT - 31;  // positions that can hold a 32-base seed
    if (nq <= 0) return;
    int64_t span = (nq + a.splits - 1) / a.splits;
    span = (span + 31) / 32 * 32;
    const int64_t q_begin = (int64_t)part * span;
    const int64_t q_end = q_begin + span < nq ? q_begin + span : nq;
    if (q_begin >= q_end) return;
    const int64_t c0 = a.seg_contig_start[seg], c1 = a.seg_contig_start[seg + 1];
    for (int64_t g0 = c0; g0 < c1; g0 += SP_GROUP) {
        const int gn = (int)(c1 - g0 < SP_GROUP ? c1 - g0 : SP_GROUP);
        const int64_t tab = startpos_group(c0, seg, (g0 - c0) / SP_GROUP) * (SP_SLOTS + 1);
        __syncthreads();
        int any = 0;
        for (int i = tid; i <= SP_SLOTS; i += nthr) {
            const int32_t hd = a.tab_head[tab + i];
            s_head[i] = hd;
            s_key[i] = a.tab_key[tab + i];
            any |= hd;
        }
        for (int i = tid; i < gn; i += nthr) s_next[i] = a.tab_next[g0 + i];  // (only the entries of chained contigs are ever read)
        if (!__syncthreads_or(any)) continue;  // no contig of this group is in the index
        if constexpr (BITMAP) {
            // one bit per value of a seed's first eight bases: a position whose bit is clear cannot be a seed hit
            for (int i = tid; i < SP_BITMAP_WORDS; i += nthr) s_bits[i] = 0;
            __syncthreads();
            for (int i = tid; i <= SP_SLOTS; i += nthr) {
                if (s_head[i] == 0) continue;
                const uint32_t k16 = i == SP_SLOTS ? 0xffffu : (uint32_t)(s_key[i] >> 48);
                atomicOr(&s_bits[k16 >> 5], 1u << (k16 & 31));
            }
            __syncthreads();
        }
        // SP_RUN consecutive positions per thread: they share one pair of truth words and one pair of mask words
        // (SP_RUN divides 32 and slices start at multiples of 32), so a position costs two funnel shifts, a hash
        // and a shared-memory probe -- no global load of its own
        for (int64_t qb = q_begin; qb < q_end; qb += (int64_t)nthr * SP_RUN) {
            const int64_t q0 = qb + (int64_t)tid * SP_RUN;
            if (q0 >= q_end) continue;
            const int64_t idx = q0 >> 5;
            const uint32_t o0 = (uint32_t)(q0 & 31);
            const uint32_t m_a = __ldg(&tm[idx]), m_b = __ldg(&tm[idx + 1]);
            const uint64_t w_a = __ldg(&tw[idx]), w_b = __ldg(&tw[idx + 1]);
            uint32_t hits = 0;  // positions of the run whose seed is in the table
            if constexpr (BITMAP) {
                // the first eight bases of the run's eight windows all lie in the top 32 bits of the window at the run's
                // start: a shift, a shared-memory word and a bit test per position; the table is probed for bitmap hits only
                const uint32_t r_hi = (uint32_t)(window64(w_a, w_b, o0) >> 32);
                uint32_t maybe = 0;
#pragma unroll
                for (int u = 0; u < SP_RUN; u++) {
                    const uint32_t k16 = (r_hi >> (16 - 2 * u)) & 0xffffu;
                    maybe |= ((s_bits[k16 >> 5] >> (k16 & 31)) & 1u) << u;
                }
                while (maybe) {
                    const int u = __ffs((int)maybe) - 1;
                    maybe &= maybe - 1;
                    const uint32_t o = o0 + (uint32_t)u;
                    if (window32(m_a, m_b, o) == 0 && startpos_probe(s_key, s_head, window64(w_a, w_b, o)) != 0) hits |= 1u << u;
                }
            } else {
#pragma unroll
                for (int u = 0; u < SP_RUN; u++) {
                    const uint32_t o = o0 + (uint32_t)u;
                    const unsigned long long seed = window64(w_a, w_b, o);
                    bool hit = startpos_probe(s_key, s_head, seed) != 0;
                    hit = hit && window32(m_a, m_b, o) == 0;  // (no byte outside ACGT in the window)
                    hits |= (uint32_t)hit << u;
                }
            }
            // seed hits (rare): this thread verifies the whole contig; one copy of that code, outside the unrolled run
            while (hits) {
                const int u = __ffs((int)hits) - 1;
                hits &= hits - 1;
                const int64_t q = q0 + u;
                if (q >= q_end) break;
                const int cand = startpos_probe(s_key, s_head, window64(w_a, w_b, o0 + (uint32_t)u));
                startpos_verify(a, cand, s_next, g0, q, LT, tw, tm, tc);
            }
        }
    }
}

// final values, and the scan for the contigs the index cannot take
__global__ void k_startpos(StartposArgs a) {
    __shared__ int s_best;
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int64_t c = blockIdx.x; c < a.n_contigs; c += gridDim.x) {
        const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
        const int64_t seg = a.ctg_seg[c];
        const int64_t LT = a.tr_off[seg + 1] - a.tr_off[seg];
        const bool placed = a.total[c] != 0;
        // upstream assigns truth.find(contig) only inside the read-hit branch, else the value stays 0
        auto publish = [&](int found) {
            if (a.startpos) a.startpos[c] = placed ? found : 0;
            if (a.exact) a.exact[c] = found;
        };
        if (!placed && !a.search_all) {
            if (tid == 0) publish(0);
            continue;
        }
        if (L == 0 || L > LT) {  // "" is found at 0; longer than the truth: npos
            if (tid == 0) publish(L == 0 ? 0 : -1);
            continue;
        }
        if (a.best && startpos_indexable(a, c, LT)) {
            if (tid == 0) { const uint32_t b = a.best[c]; publish(b == POS_INF ? -1 : (int32_t)b); }
            continue;
        }
        const uint64_t *cw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *cm = a.ctg_mask + a.ctg_woff[c];
        const uint8_t *cc = a.ctg_chars + a.ctg_off[c];
        const uint64_t *tw = a.tr_words + a.tr_woff[seg];
        const uint32_t *tm = a.tr_mask + a.tr_woff[seg];
        const uint8_t *tc = a.tr_chars + a.tr_off[seg];
        const int64_t nw = (L + 31) >> 5;
        const uint64_t c0 = cw[0];
        const uint32_t m0 = cm[0];
        const int first = L < 32 ? (int)L : 32;
        __syncthreads();
        if (tid == 0) s_best = 0x7fffffff;
        __syncthreads();
        const int64_t nq = LT - L + 1;
        for (int64_t qb = 0; qb < nq; qb += nthr) {
            const int64_t q = qb + tid;
            if (q < nq) {
                const int64_t idx = q >> 5;
                const uint32_t o = (uint32_t)(q & 31);
                // byte equality == equal 2-bit codes AND equal validity AND equal raw bytes where invalid
                bool ok = ((window64(tw[idx], tw[idx + 1], o) ^ c0) & keep_bases(first)) == 0 &&
                          ((window32(tm[idx], tm[idx + 1], o) ^ m0) & keep_bits(first)) == 0;
                bool any_invalid = (m0 & keep_bits(first)) != 0;
                for (int64_t j = 1; ok && j < nw; j++) {
                    const int rem = (L - 32 * j) < 32 ? (int)(L - 32 * j) : 32;
                    const uint32_t mj = cm[j];
                    ok = ((window64(tw[idx + j], tw[idx + j + 1], o) ^ cw[j]) & keep_bases(rem)) == 0 &&
                         ((window32(tm[idx + j], tm[idx + j + 1], o) ^ mj) & keep_bits(rem)) == 0;
                    any_invalid |= (mj & keep_bits(rem)) != 0;
                }
                if (ok && any_invalid) {
                    for (int64_t i = 0; ok && i < L; i++) ok = tc[q + i] == cc[i];
                }
                if (ok) atomicMin(&s_best, (int)q);
            }
            if (__syncthreads_or(s_best != 0x7fffffff)) break;  // positions are visited in ascending blocks
        }
        __syncthreads();
        if (tid == 0) publish(s_best == 0x7fffffff ? -1 : s_best);
    }
}

}  // namespace bs
