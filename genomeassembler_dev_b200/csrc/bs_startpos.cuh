// Contig-in-truth offset: leftmost exact occurrence of every contig in its segment's truth.
#pragma once
#include "bs_place.cuh"

namespace bs {

// ------------------------------------------------------------------------------------------
// contig-in-truth offset (upstream lib/BreakageScorer.cpp:273-274): leftmost exact occurrence
// of the whole contig in its segment's truth, assigned only if at least one read was placed.
//
// k_startpos_build : one thread per contig inserts the contig's SEED (first 32 bases) into the seed
//                    table of its group in global memory (slot claimed by compare-and-swap with linear
//                    probing, contigs with equal seeds chained) and sets the bit of the seed's PREFIX
//                    in the group's bitmap.
// k_startpos_scan  : the truth streams past the bitmap (shared memory: a shift, a word and a bit test
//                    per position); the seed table is probed only for bitmap hits; every (contig,
//                    position) with an equal seed goes to a candidate queue in global memory.
// k_startpos_verify: one WARP per queued candidate compares the whole contig with the truth on packed
//                    words, 64 words per step; the leftmost verified position wins through atomicMin.
// k_startpos       : final values, and the scan for the contigs the index cannot take (shorter than
//                    32 bases, or a non-ACGT byte in the seed).
// O(L_truth + sum L_c) per segment instead of O(C * L_truth).
//
// Two table geometries.  Segments with few contigs (cfg-2: 5-60): groups of SP_GROUP contigs, table
// (2048 slots) and bitmap (8-base prefixes, 8 KB) copied into shared memory by every scanning block.
// BIG (a segment with more than SP_GROUP contigs -- cfg-4's 10^4 scaffolds, cfg-5's 10^5 contigs):
// ONE group per segment, table in global memory (L2 resident), bitmap over 10-base prefixes (128 KB of
// shared memory), so the truth is streamed once instead of once per 1024 contigs.
//
// History: the first indexed version verified a seed hit inside the scanning thread; ONE thread then
// walked a 50 kb contig in ~400 dependent L2 round trips and the whole kernel waited for the longest
// such chain (0.46 ms per 1000 segments whatever the scan cost: profiles/r02a).  Queue + warp-per-
// candidate verification removes that chain from the scan.
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
    uint32_t *best;       // [C] leftmost verified position, POS_INF on entry (k_startpos_verify writes, k_startpos reads)
    int32_t *startpos;    // optional: truth.find(contig) if a read was placed, else 0 (upstream semantics)
    int32_t *exact;       // optional: truth.find(contig) for EVERY contig (-1: not a substring); feeds the edit distance
    int32_t search_all;   // search also for contigs without a placed read (needed for `exact`)
    // seed tables: all-ones keys, zero heads and zero bitmaps on entry
    const int64_t *seg_tab_off;   // [S] first table slot of the segment's groups
    const int32_t *seg_tab_mask;  // [S] BIG: slots - 1 of the segment's one table (a power of two - 1); else unused
    unsigned long long *tab_key;  // seed owning the slot; a table has slots + 1 entries: the last one is reserved for the all-ones seed (= the empty marker)
    int32_t *tab_head;            // group-local contig index + 1 of the first contig with that seed
    int32_t *tab_next;            // [C] next contig of the group with the same seed
    uint32_t *bitmap;             // per group: one bit per seed prefix
    int32_t big;                  // table geometry (see above)
    // candidate queue
    uint2 *queue;                 // (contig, truth position)
    int32_t *q_count;             // zero on entry
    int32_t q_cap;
};

constexpr int SP_SLOTS = 2048;  // seed table slots per group (small geometry)
constexpr int SP_GROUP = 1024;  // contigs per group (small geometry)
constexpr int SP_PFX_SMALL = 16, SP_PFX_BIG = 20;  // bits of a seed's prefix indexed by the bitmap (8 / 10 bases)
constexpr int SP_BITMAP_WORDS = (1 << SP_PFX_SMALL) / 32, SP_BITMAP_WORDS_BIG = (1 << SP_PFX_BIG) / 32;
constexpr int SP_RUN = 8;           // consecutive truth positions per thread and scan step (divides 32)
constexpr int SP_VERIFY_WORDS = 2;  // contig words per lane and verification step

BS_HD int64_t startpos_groups_of(int64_t n_contigs_of_segment) { return (n_contigs_of_segment + SP_GROUP - 1) / SP_GROUP; }

// a contig the seed index can look for: 32 valid leading bases, fits in the truth, had a read placed
__device__ __forceinline__ bool startpos_indexable(const StartposArgs &a, int64_t c, int64_t LT) {
    const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
    return L >= 32 && L <= LT && (a.search_all || a.total[c] != 0) && a.ctg_mask[a.ctg_woff[c]] == 0;
}

// geometry of the table that holds contig number `local` (0-based inside its segment)
struct StartposTable {
    int64_t tab;      // first slot
    uint32_t mask;    // slots - 1; slot mask + 1 is the reserved one
    int64_t bm;       // first bitmap word
    int64_t g0;       // segment-local index of the group's first contig
};
__device__ __forceinline__ StartposTable startpos_table(const StartposArgs &a, int64_t seg, int64_t local) {
    StartposTable t;
    if (a.big) {
        t.tab = a.seg_tab_off[seg];
        t.mask = (uint32_t)a.seg_tab_mask[seg];
        t.bm = seg * (int64_t)SP_BITMAP_WORDS_BIG;
        t.g0 = 0;
    } else {
        const int64_t g = local / SP_GROUP;
        t.tab = a.seg_tab_off[seg] + g * (SP_SLOTS + 1);
        t.mask = SP_SLOTS - 1;
        t.bm = (t.tab / (SP_SLOTS + 1)) * SP_BITMAP_WORDS;
        t.g0 = g * SP_GROUP;
    }
    return t;
}

__global__ void k_startpos_build(StartposArgs a) {
    const unsigned long long EMPTY = ~0ull;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < a.n_contigs; c += stride) {
        const int64_t seg = a.ctg_seg[c];
        if (!startpos_indexable(a, c, a.tr_off[seg + 1] - a.tr_off[seg])) continue;
        const int64_t local = c - a.seg_contig_start[seg];
        const StartposTable t = startpos_table(a, seg, local);
        const unsigned long long key = a.ctg_words[a.ctg_woff[c]];
        uint32_t h = t.mask + 1;
        if (key != EMPTY) {
            h = seed_hash(key) & t.mask;
            for (uint32_t probe = 0; probe <= t.mask; probe++) {  // at most half of the slots are ever taken
                const unsigned long long prev = atomicCAS(&a.tab_key[t.tab + h], EMPTY, key);
                if (prev == EMPTY || prev == key) break;
                h = (h + 1) & t.mask;
            }
        }
        a.tab_next[c] = atomicExch(&a.tab_head[t.tab + h], (int32_t)(local - t.g0) + 1);
        const uint32_t pfx = (uint32_t)(key >> (64 - (a.big ? SP_PFX_BIG : SP_PFX_SMALL)));
        atomicOr(&a.bitmap[t.bm + (pfx >> 5)], 1u << (pfx & 31));
    }
}

// group-local index + 1 of the first contig whose seed is `seed` (0: none)
__device__ __forceinline__ int startpos_probe(const unsigned long long *key, const int32_t *head, uint32_t mask, unsigned long long seed) {
    const unsigned long long EMPTY = ~0ull;
    if (seed == EMPTY) return head[mask + 1];
    uint32_t h = seed_hash(seed) & mask;
#pragma unroll 1
    for (uint32_t probe = 0; probe <= mask; probe++) {
        const unsigned long long k = key[h];
        if (k == EMPTY) break;
        if (k == seed) return head[h];
        h = (h + 1) & mask;
    }
    return 0;
}

// does contig c occur at truth position q?  One thread, packed words (the scan's fallback when the queue is full).
// (plain pointers, not the argument struct: a reference to kernel parameters in a non-inlined call would make every
// thread copy the struct to its stack)
__device__ BS_NOINLINE bool startpos_verify_one(const uint64_t *cw, const uint32_t *cm, const uint8_t *cc, int64_t L, int64_t q,
                                                const uint64_t *tw, const uint32_t *tm, const uint8_t *tc) {
    const int64_t idx = q >> 5, nw = (L + 31) >> 5;
    const uint32_t o = (uint32_t)(q & 31);
    bool any_invalid = false;
    for (int64_t j = 0; j < nw; j++) {
        const int64_t left = L - 32 * j;
        const int rem = left < 32 ? (int)left : 32;
        const uint32_t mj = cm[j];
        // byte equality == equal 2-bit codes AND equal validity AND equal raw bytes where invalid
        if ((window64(tw[idx + j], tw[idx + j + 1], o) ^ cw[j]) & keep_bases(rem)) return false;
        if ((window32(tm[idx + j], tm[idx + j + 1], o) ^ mj) & keep_bits(rem)) return false;
        any_invalid |= (mj & keep_bits(rem)) != 0;
    }
    if (any_invalid) {
        for (int64_t i = 0; i < L; i++)
            if (tc[q + i] != cc[i]) return false;
    }
    return true;
}

// seed hit at truth position q: every contig of the chain that fits goes to the candidate queue
__device__ __forceinline__ void startpos_enqueue(const StartposArgs &a, int ci_head, int64_t cbase, int64_t q, int64_t LT,
                                                 const uint64_t *tw, const uint32_t *tm, const uint8_t *tc) {
    for (int ci = ci_head; ci != 0; ci = a.tab_next[cbase + ci - 1]) {
        const int64_t c = cbase + ci - 1;
        const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
        if (q + L > LT) continue;
        const int slot = atomicAdd(a.q_count, 1);
        if (slot < a.q_cap) a.queue[slot] = make_uint2((uint32_t)c, (uint32_t)q);
        else if (startpos_verify_one(a.ctg_words + a.ctg_woff[c], a.ctg_mask + a.ctg_woff[c], a.ctg_chars + a.ctg_off[c], L, q, tw, tm, tc))
            atomicMin(&a.best[c], (uint32_t)q);  // queue full: verify here
    }
}

template <bool BIG>
__global__ void __launch_bounds__(BIG ? 1024 : 256, BIG ? 1 : 4) k_startpos_scan(StartposArgs a) {
    constexpr int PFX = BIG ? SP_PFX_BIG : SP_PFX_SMALL;
    constexpr int BM_WORDS = BIG ? SP_BITMAP_WORDS_BIG : SP_BITMAP_WORDS;
    __shared__ unsigned long long s_key[BIG ? 1 : SP_SLOTS + 1];
    __shared__ int32_t s_head[BIG ? 1 : SP_SLOTS + 1];
    __shared__ uint32_t s_bits_small[BIG ? 1 : SP_BITMAP_WORDS];
    uint32_t *s_bits = BIG ? (uint32_t *)bs_dyn_smem() : s_bits_small;
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
    const int64_t group = BIG ? (c1 - c0 > 0 ? c1 - c0 : 1) : SP_GROUP;
    for (int64_t l0 = 0; l0 < c1 - c0; l0 += group) {
        const StartposTable t = startpos_table(a, seg, l0);
        const unsigned long long *key = BIG ? a.tab_key + t.tab : s_key;
        const int32_t *head = BIG ? a.tab_head + t.tab : s_head;
        __syncthreads();
        int any = 0;
        for (int i = tid; i < BM_WORDS; i += nthr) {
            const uint32_t w = a.bitmap[t.bm + i];
            s_bits[i] = w;
            any |= (int)(w != 0);
        }
        if constexpr (!BIG) {
            for (int i = tid; i <= SP_SLOTS; i += nthr) {
                s_head[i] = a.tab_head[t.tab + i];
                s_key[i] = a.tab_key[t.tab + i];
            }
        }
        if (!__syncthreads_or(any)) continue;  // no contig of this group is in the index
        // SP_RUN consecutive positions per thread: they share one pair of truth words and one pair of mask words
        // (SP_RUN divides 32 and slices start at multiples of 32).  The prefixes of the run's windows all lie in the
        // leading bits of the window at the run's start: a shift, a shared-memory word and a bit test per position.
        for (int64_t qb = q_begin; qb < q_end; qb += (int64_t)nthr * SP_RUN) {
            const int64_t q0 = qb + (int64_t)tid * SP_RUN;
            if (q0 >= q_end) continue;
            const int64_t idx = q0 >> 5;
            const uint32_t o0 = (uint32_t)(q0 & 31);
            const uint32_t m_a = __ldg(&tm[idx]), m_b = __ldg(&tm[idx + 1]);
            const uint64_t w_a = __ldg(&tw[idx]), w_b = __ldg(&tw[idx + 1]);
            const uint64_t r = window64(w_a, w_b, o0);
            uint32_t maybe = 0;
#pragma unroll
            for (int u = 0; u < SP_RUN; u++) {
                const uint32_t pfx = (uint32_t)(r >> (64 - PFX - 2 * u)) & ((1u << PFX) - 1u);
                maybe |= ((s_bits[pfx >> 5] >> (pfx & 31)) & 1u) << u;
            }
            while (maybe) {  // (rare) prefix present: probe the table, queue the chain
                const int u = __ffs((int)maybe) - 1;
                maybe &= maybe - 1;
                const int64_t q = q0 + u;
                if (q >= q_end) break;
                const uint32_t o = o0 + (uint32_t)u;
                if (window32(m_a, m_b, o) != 0) continue;  // a byte outside ACGT in the window
                const int cand = startpos_probe(key, head, t.mask, window64(w_a, w_b, o));
                if (cand != 0) startpos_enqueue(a, cand, c0 + l0, q, LT, tw, tm, tc);
            }
        }
    }
}

// one warp per queued candidate: the whole contig against the truth at that position
__global__ void __launch_bounds__(256) k_startpos_verify(StartposArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int n = *a.q_count < a.q_cap ? *a.q_count : a.q_cap;
    for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < n; e += warps) {  // (warp-uniform)
        const uint2 cd = a.queue[e];
        const int64_t c = cd.x, q = cd.y;
        const int64_t seg = a.ctg_seg[c];
        const uint64_t *tw = a.tr_words + a.tr_woff[seg];
        const uint32_t *tm = a.tr_mask + a.tr_woff[seg];
        const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
        const uint64_t *cw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *cm = a.ctg_mask + a.ctg_woff[c];
        const int64_t idx = q >> 5, nw = (L + 31) >> 5;
        const uint32_t o = (uint32_t)(q & 31);
        // a position further left may already be known; one lane reads so that the whole warp takes the same branch
        uint32_t known = 0;
        if (lane == 0) known = *(volatile uint32_t *)&a.best[c];
        known = __shfl_sync(FULL_MASK, known, 0);
        if ((uint32_t)q >= known) continue;
        bool ok = true, any_invalid = false;
        for (int64_t j0 = 0; ok && j0 < nw; j0 += 32 * SP_VERIFY_WORDS) {  // (ok is warp-uniform: every lane leaves together)
            uint64_t diff = 0;
            uint32_t mdiff = 0, minv = 0;
#pragma unroll
            for (int u = 0; u < SP_VERIFY_WORDS; u++) {
                const int64_t j = j0 + 32 * u + lane;
                if (j < nw) {
                    const int64_t left = L - 32 * j;
                    const int rem = left < 32 ? (int)left : 32;
                    const uint32_t mj = cm[j];
                    // byte equality == equal 2-bit codes AND equal validity AND equal raw bytes where invalid
                    diff |= (window64(tw[idx + j], tw[idx + j + 1], o) ^ cw[j]) & keep_bases(rem);
                    mdiff |= (window32(tm[idx + j], tm[idx + j + 1], o) ^ mj) & keep_bits(rem);
                    minv |= mj & keep_bits(rem);
                }
            }
            ok = __all_sync(FULL_MASK, diff == 0 && mdiff == 0);
            any_invalid |= minv != 0;
        }
        if (ok && __any_sync(FULL_MASK, any_invalid)) {
            const uint8_t *cc = a.ctg_chars + a.ctg_off[c];
            const uint8_t *tc = a.tr_chars + a.tr_off[seg];
            for (int64_t i0 = 0; ok && i0 < L; i0 += 32) {
                const int64_t i = i0 + lane;
                ok = __all_sync(FULL_MASK, i < L ? tc[q + i] == cc[i] : true);
            }
        }
        if (ok && lane == 0) atomicMin(&a.best[c], (uint32_t)q);
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
