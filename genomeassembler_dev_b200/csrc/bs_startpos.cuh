// Contig-in-truth offset: leftmost exact occurrence of every contig in its segment's truth.
#pragma once
#include "bs_place.cuh"

namespace bs {

// ------------------------------------------------------------------------------------------
// contig-in-truth offset (upstream lib/BreakageScorer.cpp:273-274): leftmost exact occurrence
// of the whole contig in its segment's truth, assigned only if at least one read was placed.
//
// k_startpos_index: the truth streams past a small shared-memory table of the segment's contig
// SEEDS (first 32 bases); a seed hit is verified by the whole warp on packed words; the leftmost
// verified position wins through atomicMin.  O(L_truth + sum L_c) per segment instead of
// O(C * L_truth).  Contigs without a packable seed (shorter than 32, or a non-ACGT byte in it)
// are left to k_startpos, which also turns the atomicMin scratch into the final values.
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
};

constexpr int SP_SLOTS = 2048;  // seed table slots per pass
constexpr int SP_GROUP = 1024;  // contigs per pass

// a contig the seed index can look for: 32 valid leading bases, fits in the truth, had a read placed
__device__ __forceinline__ bool startpos_indexable(const StartposArgs &a, int64_t c, int64_t LT) {
    const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
    return L >= 32 && L <= LT && (a.search_all || a.total[c] != 0) && a.ctg_mask[a.ctg_woff[c]] == 0;
}

__global__ void __launch_bounds__(256) k_startpos_index(StartposArgs a) {
    __shared__ unsigned long long s_key[SP_SLOTS + 1];  // slot SP_SLOTS is reserved for the all-ones seed (= the empty marker)
    __shared__ int32_t s_head[SP_SLOTS + 1];  // local contig index + 1 of the first contig with that seed
    __shared__ int32_t s_next[SP_GROUP];  // next contig with the same seed
    const unsigned long long EMPTY = ~0ull;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
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
        __syncthreads();
        for (int i = tid; i <= SP_SLOTS; i += nthr) { s_head[i] = 0; s_key[i] = EMPTY; }
        __syncthreads();
        // parallel build: claim the seed's slot by compare-and-swap (linear probing), then push the
        // contig on the slot's chain
        for (int i = tid; i < gn; i += nthr) {
            const int64_t c = g0 + i;
            if (!startpos_indexable(a, c, LT)) continue;
            const unsigned long long key = a.ctg_words[a.ctg_woff[c]];
            uint32_t h = SP_SLOTS;
            if (key != EMPTY) {
                h = seed_hash(key) & (SP_SLOTS - 1);
                for (;;) {
                    const unsigned long long prev = atomicCAS(&s_key[h], EMPTY, key);
                    if (prev == EMPTY || prev == key) break;
                    h = (h + 1) & (SP_SLOTS - 1);
                }
            }
            s_next[i] = atomicExch(&s_head[h], i + 1);
        }
        __syncthreads();
        for (int64_t qb = q_begin; qb < q_end; qb += nthr) {
            const int64_t q = qb + tid;
            int cand = 0;
            if (q < q_end) {
                const int64_t idx = q >> 5;
                const uint32_t o = (uint32_t)(q & 31);
                if (window32(__ldg(&tm[idx]), __ldg(&tm[idx + 1]), o) == 0) {
                    const unsigned long long seed = window64(__ldg(&tw[idx]), __ldg(&tw[idx + 1]), o);
                    if (seed == EMPTY) cand = s_head[SP_SLOTS];
                    else {
                        uint32_t h = seed_hash(seed) & (SP_SLOTS - 1);
                        for (;;) {
                            const unsigned long long k = s_key[h];
                            if (k == EMPTY) break;
                            if (k == seed) { cand = s_head[h]; break; }
                            h = (h + 1) & (SP_SLOTS - 1);
                        }
                    }
                }
            }
            // seed hits of the warp, one after the other, each verified by all 32 lanes
            unsigned hits = __ballot_sync(FULL_MASK, cand != 0);
            while (hits) {
                const int src = __ffs((int)hits) - 1;
                hits &= hits - 1;
                const int64_t qc = __shfl_sync(FULL_MASK, q, src);
                int ci = __shfl_sync(FULL_MASK, cand, src);
                for (; ci != 0; ci = s_next[ci - 1]) {
                    const int64_t c = g0 + ci - 1;
                    const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
                    if (qc + L > LT) continue;
                    // a position further left may already be known; one lane reads so that the whole
                    // warp takes the same branch (the value can change under us)
                    uint32_t known = 0;
                    if (lane == 0) known = *(volatile uint32_t *)&a.best[c];
                    known = __shfl_sync(FULL_MASK, known, 0);
                    if ((uint32_t)qc >= known) continue;
                    const uint64_t *cw = a.ctg_words + a.ctg_woff[c];
                    const uint32_t *cm = a.ctg_mask + a.ctg_woff[c];
                    const int64_t idx = qc >> 5;
                    const uint32_t o = (uint32_t)(qc & 31);
                    const int64_t nw = (L + 31) >> 5;
                    bool ok = true, any_invalid = false;
                    for (int64_t j0 = 0; ok && j0 < nw; j0 += 32) {
                        const int64_t j = j0 + lane;
                        bool okl = true;
                        if (j < nw) {
                            const int rem = (L - 32 * j) < 32 ? (int)(L - 32 * j) : 32;
                            const uint32_t mj = cm[j];
                            // byte equality == equal 2-bit codes AND equal validity AND equal raw bytes where invalid
                            okl = ((window64(tw[idx + j], tw[idx + j + 1], o) ^ cw[j]) & keep_bases(rem)) == 0 &&
                                  ((window32(tm[idx + j], tm[idx + j + 1], o) ^ mj) & keep_bits(rem)) == 0;
                            any_invalid |= (mj & keep_bits(rem)) != 0;
                        }
                        ok = __ballot_sync(FULL_MASK, !okl) == 0;
                    }
                    if (ok && __ballot_sync(FULL_MASK, any_invalid) != 0) {
                        const uint8_t *cc = a.ctg_chars + a.ctg_off[c];
                        for (int64_t i0 = 0; ok && i0 < L; i0 += 32) {
                            const int64_t i = i0 + lane;
                            const bool okl = i < L ? tc[qc + i] == cc[i] : true;
                            ok = __ballot_sync(FULL_MASK, !okl) == 0;
                        }
                    }
                    if (ok && lane == 0) atomicMin(&a.best[c], (uint32_t)qc);
                }
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
