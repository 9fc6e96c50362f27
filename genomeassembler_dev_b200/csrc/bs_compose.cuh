// Compositional scoring of scaffold sets: the candidates of upstream assemble_contigs
// (lib/BreakageScorer.cpp:105-171) are chains of the same few base contigs, so the reads are placed once per BASE
// contig and every scaffold is scored from its parts.
//
// A scaffold is text = base[b_0] + base[b_1][ov_1..] + ...; every part is a whole substring of it, part i starting at
// s_i = dst_i - ov_i (dst_i: where the bases that part i adds begin = the end of the scaffold before it).  For a read r
// and an occurrence [p, p + len) in the scaffold let i be the last part with s_i <= p:
//   * the occurrence ends inside part i  -> it is an occurrence in base[b_i] at p - s_i: the leftmost such p over the
//     parts is min_i (s_i + leftmost_{b_i}(r)), from the ONE placement of the reads in the base contigs;
//   * it runs past the end of part i      -> p lies in [dst_{i+1} - (len - 1), s_{i+1}): a junction window of at most
//     max_read_len - 1 positions, probed against the read index like k_place_index probes every position.
// The minimum over both is std::string::find's answer (lib/BreakageScorer.cpp:241) -- a superset of occurrences is
// examined, every one of them verified, and duplicates do not change a minimum.
#pragma once
#include "bs_place.cuh"
#include "bs_score.cuh"
#include "bs_ks.cuh"

namespace bs {

static_assert(POS_INF == POS_INF_ROW, "score_finish reads the rows of k_place_compose");

struct ScaffoldParts {
    const int64_t *part_start;  // [n_scaffolds + 1]
    const int32_t *part_base;   // [P] base contig of a part
    const int32_t *part_ov;     // [P] overlap with the scaffold so far
    const int32_t *part_dst;    // [P] scaffold position of the first base the part adds
};

// ---- scaffold text from the base contigs' text (one block per scaffold at a time) ----
struct ComposeTextArgs {
    ScaffoldParts sp;
    const uint8_t *base_chars;
    const int64_t *base_off;
    const int64_t *ctg_off;  // [n_scaffolds + 1] offsets of the scaffold texts
    uint8_t *ctg_chars;
    int64_t n_scaffolds;
};

__global__ void k_compose_text(ComposeTextArgs a) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int64_t c = blockIdx.x; c < a.n_scaffolds; c += gridDim.x) {
        uint8_t *out = a.ctg_chars + a.ctg_off[c];
        for (int64_t i = a.sp.part_start[c]; i < a.sp.part_start[c + 1]; i++) {
            const int32_t b = a.sp.part_base[i], ov = a.sp.part_ov[i];
            const uint8_t *src = a.base_chars + a.base_off[b] + ov;
            const int64_t n = a.base_off[b + 1] - a.base_off[b] - ov;
            uint8_t *dst = out + a.sp.part_dst[i];
            // bytes up to the first 8-byte boundary of the destination, then aligned 8-byte stores assembled from
            // byte loads (the base contigs are a few tens of kB: they sit in L1 / L2), then the tail
            int64_t head = (8 - (int64_t)((uintptr_t)dst & 7)) & 7;
            if (head > n) head = n;
            for (int64_t k = tid; k < head; k += nthr) dst[k] = src[k];
            const int64_t nw = (n - head) >> 3;
            for (int64_t k = tid; k < nw; k += nthr) {
                const uint8_t *s = src + head + 8 * k;
                uint64_t v = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) v |= (uint64_t)s[j] << (8 * j);
                *reinterpret_cast<uint64_t *>(dst + head + 8 * k) = v;
            }
            for (int64_t k = head + 8 * nw + tid; k < n; k += nthr) dst[k] = src[k];
        }
    }
}

// ---- ... or, when no kernel will look at the text (base contigs of ACGT only: a read or truth with another byte cannot
// match them, so the byte-comparison paths are never taken), the packed words and masks straight from the base contigs'
// packed words: a thread per output word, funnel shifts over the parts the word spans ----
struct ComposeWordsArgs {
    ScaffoldParts sp;
    const int64_t *base_woff;
    const uint64_t *base_words;
    const uint32_t *base_mask;
    const int64_t *ctg_off;   // [n_scaffolds + 1] text offsets (lengths)
    const int64_t *ctg_woff;  // [n_scaffolds + 1] word offsets
    uint64_t *ctg_words;
    uint32_t *ctg_mask;
    int64_t n_scaffolds;
};

__global__ void k_compose_words(ComposeWordsArgs a) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int64_t c = blockIdx.x; c < a.n_scaffolds; c += gridDim.x) {
        const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
        const int64_t nwords = a.ctg_woff[c + 1] - a.ctg_woff[c];
        uint64_t *ow = a.ctg_words + a.ctg_woff[c];
        uint32_t *om = a.ctg_mask + a.ctg_woff[c];
        const int64_t ps = a.sp.part_start[c], pe = a.sp.part_start[c + 1];
        for (int64_t w = tid; w < nwords; w += nthr) {
            int64_t pos = 32 * w;
            uint64_t word = 0;
            uint32_t mask = 0;
            if (pos >= L) mask = ~0u;  // past the end and the two pad words: code 0, mask set
            else {
                int64_t lo = ps, hi = pe - 1;  // last part whose new bases start at or before pos
                while (lo < hi) {
                    const int64_t mid = (lo + hi + 1) >> 1;
                    if ((int64_t)__ldg(&a.sp.part_dst[mid]) <= pos) lo = mid; else hi = mid - 1;
                }
                int64_t i = lo;
                int filled = 0;
                while (filled < 32) {
                    if (pos >= L) { mask |= keep_bits(32 - filled) >> filled; break; }
                    const int64_t dst = __ldg(&a.sp.part_dst[i]);
                    const int64_t end = i + 1 < pe ? (int64_t)__ldg(&a.sp.part_dst[i + 1]) : L;
                    int n = 32 - filled;
                    if (end - pos < n) n = (int)(end - pos);
                    if (n > 0) {
                        const int64_t q = (int64_t)__ldg(&a.sp.part_ov[i]) + (pos - dst);
                        const int64_t bo = __ldg(&a.base_woff[__ldg(&a.sp.part_base[i])]) + (q >> 5);
                        const uint32_t o = (uint32_t)(q & 31);
                        word |= (window64(__ldg(&a.base_words[bo]), __ldg(&a.base_words[bo + 1]), o) & keep_bases(n)) >> (2 * filled);
                        mask |= (window32(__ldg(&a.base_mask[bo]), __ldg(&a.base_mask[bo + 1]), o) & keep_bits(n)) >> filled;
                        filled += n;
                        pos += n;
                    }
                    if (pos >= end) i++;
                }
            }
            ow[w] = word;
            om[w] = mask;
        }
    }
}

// ---- what a base contig contributes to every scaffold it is part of, computed once ----
struct BaseSideArgs {
    const int32_t *base_pos;   // [n_base][n_reads] leftmost positions out of k_place_index, -1: none
    int64_t n_base, n_reads;
    uint2 *hits;               // [n_base][n_reads] out: (read, position) of the reads placed in a base contig, ascending read ids
    int32_t *hit_di;           // [n_base][n_reads] out: dense index of the break 8-mer of that placement if it lies inside the
                               // base contig, base[q - 4 .. q + 4) (-1: a base outside ACGT), else -2 (the scaffold decides)
    int32_t *cnt;              // [n_base] out: how many
    const int64_t *base_off;   // [n_base + 1]
    const int64_t *base_woff;
    const uint64_t *base_words;
    const uint32_t *base_mask;
    int32_t kmer;
    // for the KS statistic of the rolling-window probabilities (k_ks_compose), all NULL / 0 when it is not wanted:
    const WinEntry *win;       // [4^kmer]
    int32_t rank_zero;
    uint16_t *base_rank;       // [sum L_b] out: rank of the window that starts at a base contig position (rank_zero: a base outside ACGT)
};

// one block per base contig: ordered compaction of its row of leftmost positions (warp ballots per round of reads), so
// that the list -- and with it which thread of k_place_compose adds which read's probability -- is the same in every run
__global__ void k_base_side(BaseSideArgs a) {
    __shared__ int s_wcnt[32];
    __shared__ int64_t s_base;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const int64_t b = blockIdx.x, N = a.n_reads;
    const int32_t *row = a.base_pos + b * N;
    const int64_t Lb = a.base_off[b + 1] - a.base_off[b];
    const uint64_t *bw = a.base_words + a.base_woff[b];
    const uint32_t *bm = a.base_mask + a.base_woff[b];
    if (tid == 0) s_base = 0;
    __syncthreads();
    // blockDim reads per round: a ballot per warp gives the lane's place among its warp's placed reads, the warps'
    // counts (shared memory) the warp's place in the round (the loop bounds are block-uniform: every lane votes)
    for (int64_t r0 = 0; r0 < N; r0 += nthr) {
        const int64_t r = r0 + tid;
        const int32_t q = r < N ? row[r] : -1;
        const unsigned m = __ballot_sync(FULL_MASK, q >= 0);
        if (lane == 0) s_wcnt[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < nwarp; w++) {
            if (w < warp) before += s_wcnt[w];
            total += s_wcnt[w];
        }
        if (q >= 0) {
            const int64_t o = b * N + s_base + before + __popc(m & ((1u << lane) - 1u));
            a.hits[o] = make_uint2((uint32_t)r, (uint32_t)q);
            int32_t di = -2;
            if (a.kmer == 8 && q >= 4 && (int64_t)q + 4 <= Lb) di = dense_index_at(bw, bm, (int64_t)q - 4, 8);
            a.hit_di[o] = di;
        }
        __syncthreads();
        if (tid == 0) s_base += total;
    }
    __syncthreads();
    if (tid == 0) a.cnt[b] = (int32_t)s_base;
    if (a.base_rank) {
        const int kshift = 64 - 2 * a.kmer;
        const uint32_t kbits = keep_bits(a.kmer);
        uint16_t *out = a.base_rank + a.base_off[b];
        for (int64_t q = tid; q + a.kmer <= Lb; q += nthr) {
            const int64_t wi = q >> 5;
            const uint32_t o = (uint32_t)(q & 31);
            int32_t rk = a.rank_zero;
            if (!(window32(bm[wi], bm[wi + 1], o) & kbits)) rk = a.win[(uint32_t)(window64(bw[wi], bw[wi + 1], o) >> kshift)].rank;
            out[q] = (uint16_t)rk;
        }
    }
}

// ---- placement (and scoring) of a scaffold from its parts ----
struct PlaceComposeArgs {
    PlaceIxArgs p;            // scaffolds as the contig set (order, counter, words, mask, text), reads, read index; w / total / pos
    ScaffoldParts sp;
    const int32_t *part_jid;  // [P] list of the reads that cross the junction in front of a part (an entry of the extended base
                              // set: the text around the junction, placed once), or -1: first part / probed here
    const uint2 *base_hits;   // [n_base][n_reads] (read, leftmost position) of the reads placed in a base contig (k_base_side)
    const int32_t *base_hit_di;  // [n_base][n_reads] their break 8-mers where the base contig alone decides them
    const int32_t *base_cnt;  // [n_base] how many
    int64_t n_reads;          // reads of the (one) segment
    int32_t max_read_len;
    int32_t hash_slots;       // SCORE: slots of the break k-mer hash table in dynamic shared memory (a power of two)
    uint32_t *rows;           // [gridDim][row_stride] per-block row of leftmost positions, or NULL: the row follows the hash table in dynamic shared memory
    int64_t row_stride;
    ScoreArgs sc;             // SCORE instantiation: the break k-mers are scored here (no position weights are written)
};

#ifndef BS_COMPOSE_THREADS
#define BS_COMPOSE_THREADS 512
#endif
constexpr int COMPOSE_THREADS = BS_COMPOSE_THREADS;
#ifdef BS_CPU_EMUL
constexpr int COMPOSE_PART_CHUNK = 4;    // (emulation: small, so that the tests reach the chunked form)
#else
constexpr int COMPOSE_PART_CHUNK = 64;   // parts of a scaffold staged in shared memory at a time
#endif
constexpr uint32_t ROW_SCORED = 0x80000000u;  // flag on a row entry: the read's break has been added by the part that owns it

// ROWS_SMEM: the row of leftmost positions follows the hash table in dynamic shared memory (its own instantiation, so that
// the row's loads, stores and atomicMin are shared-memory instructions and not generic ones), else it is a.rows' (global)
template <bool SCORE, bool ROWS_SMEM>
__global__ void __launch_bounds__(COMPOSE_THREADS, 2) k_place_compose(PlaceComposeArgs a) {
    __shared__ int s_item, s_placed;
    // the lists a scaffold draws its placements from, two slots per staged part: [2i] the reads crossing the junction in
    // front of part i (if that junction has a list), [2i + 1] the reads placed in part i's base contig
    constexpr int NSRC = 2 * COMPOSE_PART_CHUNK;
    __shared__ int32_t s_pstart[NSRC], s_pbase[NSRC], s_pcnt[NSRC];  // scaffold position of the list's position 0, list, entries
    __shared__ int32_t s_qmax[NSRC], s_own[NSRC];  // entries at or beyond qmax are not placements here; an entry owns its read below own
    __shared__ int32_t s_pstep[NSRC + 1];          // warp steps (32 list entries each) before a source of the staged chunk
    __shared__ int32_t s_qstep[COMPOSE_PART_CHUNK + 1];  // ... counting the parts' lists alone (the scoring walk)
    __shared__ ScoreSharedCore s_score;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    uint32_t *hash = (uint32_t *)bs_dyn_smem();
    const int hash_slots = SCORE ? a.hash_slots : 0;
    uint32_t *row;
    if constexpr (ROWS_SMEM) row = (uint32_t *)bs_dyn_smem() + hash_slots;
    else row = a.rows + (int64_t)blockIdx.x * a.row_stride;
    if constexpr (SCORE) score_shared_init(a.sc, s_score, hash, hash_slots);
    const int64_t N = a.n_reads;
    const int span = a.max_read_len - 1;  // start positions before a junction from which a read can cross it
    const int ngroup = (span + 31) / 32;  // ... in groups of 32: one warp step each
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_item = atomicAdd(a.p.work_counter, 1);
            s_placed = 0;
            if constexpr (SCORE) { s_score.novf = 0; s_score.maxc = 0; s_score.nz = 0; s_score.fallback = 0; }
        }
        __syncthreads();
        const int item = s_item;
        if (item >= a.p.n_items) break;
        const int c = a.p.order[item];
        const int s = a.p.ctg_seg[c];
        const int64_t coff = a.p.ctg_off[c];
        const int64_t L = a.p.ctg_off[c + 1] - coff;
        const uint64_t *gw = a.p.ctg_words + a.p.ctg_woff[c];
        const uint32_t *gm = a.p.ctg_mask + a.p.ctg_woff[c];
        const uint8_t *cc = a.p.ctg_chars ? a.p.ctg_chars + coff : nullptr;  // (NULL: no text was composed -- nothing but ACGT anywhere in it)
        const int64_t ps = a.sp.part_start[c], pe = a.sp.part_start[c + 1];
        const int64_t r0 = a.p.ix.seg_read_start[s];

        // the first chunk of parts goes into shared memory together with the row's reset, so that A can follow B without a barrier
        // (the lists of the staged parts are walked as ONE sequence of warp steps, dealt out to the warps round-robin: no
        // part leaves a block-wide loop with a mostly idle tail)
        auto stage_parts = [&](int64_t pc, int np) {
            for (int i = tid; i < np; i += nthr) {
                const int64_t g = pc + i;
                const int32_t b = a.sp.part_base[g], dst = a.sp.part_dst[g], ov = a.sp.part_ov[g];
                const int32_t jid = g > ps ? a.part_jid[g] : -1;
                s_pstart[2 * i] = dst - span;  // the junction's text starts span bases before the end of the scaffold so far
                s_pbase[2 * i] = jid >= 0 ? jid : 0;
                s_pcnt[2 * i] = jid >= 0 ? a.base_cnt[jid] : 0;
                s_qmax[2 * i] = span > ov ? span - ov : 0;  // ... and a crossing read starts before part i does (never, if the overlap is that long)
                s_own[2 * i] = 0x7fffffff;
                s_pstart[2 * i + 1] = dst - ov;
                s_pbase[2 * i + 1] = b;
                s_pcnt[2 * i + 1] = a.base_cnt[b];
                s_qmax[2 * i + 1] = 0x7fffffff;
                s_own[2 * i + 1] = g + 1 < pe ? a.sp.part_dst[g + 1] - a.sp.part_ov[g + 1] : 0x7fffffff;  // the next part's start
            }
        };
        auto stage_steps = [&](int np) {  // (after a barrier behind stage_parts; followed by one)
            if (tid == 0) {
                int acc = 0, accp = 0;
                for (int i = 0; i < 2 * np; i++) { s_pstep[i] = acc; acc += (s_pcnt[i] + 31) >> 5; }
                s_pstep[2 * np] = acc;
                for (int i = 0; i < np; i++) { s_qstep[i] = accp; accp += (s_pcnt[2 * i + 1] + 31) >> 5; }
                s_qstep[np] = accp;
            }
        };
        const int np0 = (int)(pe - ps < COMPOSE_PART_CHUNK ? pe - ps : COMPOSE_PART_CHUNK);
        stage_parts(ps, np0);
        for (int64_t r = tid; r < N; r += nthr) row[r] = POS_INF;
        __syncthreads();
        stage_steps(np0);
        __syncthreads();

        // ---- B: occurrences that cross a junction (first: long dependent chains, few of them; a warp step = 32
        // consecutive start positions of one junction's window; no barrier between this and A, both only lower row entries) ----
        const int64_t nj = pe - ps - 1;
        if (nj > 0 && span > 0 && N > 0) {
            const int S = a.p.ix.seed_len[s];
            const uint32_t *head = a.p.ix.head + a.p.ix.tab_off[s];
            const uint32_t hmask = (uint32_t)a.p.ix.tab_mask[s];
            const uint64_t keepS = keep_bases(S);
            const uint32_t keepSm = keep_bits(S);
            const int nstep = (int)nj * ngroup;
            for (int it = warp; it < nstep; it += nwarp) {
                const int jj = it / ngroup, off = (it - jj * ngroup) * 32 + lane;
                const int64_t j = ps + 1 + jj;
                if (__ldg(&a.part_jid[j]) >= 0) continue;  // (warp-uniform) this junction's crossing reads come from a list
                const int64_t e = __ldg(&a.sp.part_dst[j]);        // end of the scaffold before part j
                const int64_t sj = e - __ldg(&a.sp.part_ov[j]);    // where part j starts as a whole
                const int64_t p = e - span + off;
                if (off >= span || p < 0 || p >= sj || p + S > L) continue;
                const int64_t wi = p >> 5;
                const uint32_t o = (uint32_t)(p & 31);
                if (window32(__ldg(&gm[wi]), __ldg(&gm[wi + 1]), o) & keepSm) continue;  // a non-ACGT base in the seed window
                const uint64_t seed = window64(__ldg(&gw[wi]), __ldg(&gw[wi + 1]), o) & keepS;
                const uint32_t tag = seed_tag(seed);
                for (uint32_t q = head[seed_hash(seed) & hmask]; q != 0;) {
                    const int64_t n = (int64_t)q - 1;
                    const uint2 en = a.p.ix.next[n];
                    q = en.x;
                    if (en.y != tag) continue;
                    const int len = read_length(a.p.reads, n);
                    if (p + len > L || p + len <= e) continue;  // (an occurrence that ends before the junction is a part's own)
                    if (!verify_at(a.p, gw, gm, cc, p, n, len, __ldg(&a.p.reads.words[n * a.p.reads.W]))) continue;
                    atomicMin(&row[n - r0], (uint32_t)p);
                }
            }
            // reads outside the index (a byte outside ACGT in the seed): text comparison in every junction window
            int64_t k = 0;
            for (uint32_t q = a.p.ix.odd_head[s]; q != 0;) {
                const int64_t n = (int64_t)q - 1;
                q = a.p.ix.next[n].x;
                const int len = read_length(a.p.reads, n);
                if (len <= 1) continue;  // ("" is found at 0 through the first part; one byte cannot cross a junction)
                const uint8_t *rc = a.p.reads.chars + read_begin(a.p.reads, n);
                for (int64_t jj = 0; jj < nj; jj++, k++) {
                    if (k % nthr != tid) continue;
                    const int64_t j = ps + 1 + jj;
                    if (a.part_jid[j] >= 0) continue;
                    const int64_t e = a.sp.part_dst[j], sj = e - a.sp.part_ov[j];
                    int64_t lo = e - (len - 1);
                    if (lo < 0) lo = 0;
                    int64_t hi = sj - 1 + len;  // one past the last byte a match starting before sj can cover
                    if (hi > L) hi = L;
                    if (hi - lo < len) continue;
                    int64_t f = -1;
                    if (cc) f = find_bytes(cc + lo, hi - lo, rc, len);
                    else  // (no text: the scaffold is ACGT only; the read's bytes against the bases decoded from the words)
                        for (int64_t t = 0; t + len <= hi - lo && f < 0; t++)
                            if (match_text_packed(gw, gm, lo + t, rc, len)) f = t;
                    if (f >= 0) atomicMin(&row[n - r0], (uint32_t)(lo + f));
                }
            }
        }

        // ---- A: leftmost occurrence inside a part, over the parts: every part scatters the (read, position) list of its
        // base contig into the row (atomicMin).  The parts go through shared memory a chunk at a time ----
        for (int64_t pc = ps; pc < pe; pc += COMPOSE_PART_CHUNK) {
            const int np = (int)(pe - pc < COMPOSE_PART_CHUNK ? pe - pc : COMPOSE_PART_CHUNK);
            if (pc != ps) {
                __syncthreads();
                stage_parts(pc, np);
                __syncthreads();
                stage_steps(np);
                __syncthreads();
            }
            const int nstep = s_pstep[2 * np];
            for (int it = warp, i = 0; it < nstep; it += nwarp) {
                while (it >= s_pstep[i + 1]) i++;
                const int k = (it - s_pstep[i]) * 32 + lane;
                if (k < s_pcnt[i]) {
                    const uint2 e = __ldg(&a.base_hits[(int64_t)s_pbase[i] * N + k]);
                    if (e.y < (uint32_t)s_qmax[i]) atomicMin(&row[e.x], (uint32_t)(s_pstart[i] + (int32_t)e.y));
                }
            }
        }
        __syncthreads();

        // ---- C: leftmost positions -> reads placed, positions out, position weights or scores ----
        int placed = 0;
        for (int64_t r = tid; r < N; r += nthr) {
            const uint32_t p = row[r];
            if (p == POS_INF) continue;
            placed++;
            if (a.p.pos) a.p.pos[a.p.pos_off[c] + r] = (int32_t)p;
            if constexpr (!SCORE) atomicAdd(&a.p.w[coff + c + p], 1);
        }
        if (placed) atomicAdd(&s_placed, placed);
        __syncthreads();
        const int32_t total = s_placed;
        if (tid == 0) a.p.total[c] = total;
        if constexpr (SCORE) {
            ScoreState st;
            IncrTally tl{0, 0, 0};
            score_begin_total(a.sc, total, st, hash, hash_slots);
            if (total != 0) {
                // (1) the reads a part placed: the parts' lists once more.  A list entry is THE placement of its read iff it
                // gives the row's position and its part is the last one that starts at or before that position (two parts
                // can only give the same position inside their overlap, where the later part's list holds the read as
                // well): every read is added by exactly one thread, always the same one, whatever the timing.  The break
                // 8-mer of a placement inside its base contig came with the list; the others are read off the scaffold.
                // Which thread adds which read depends on the scaffold's parts alone -- not on the other scaffolds of the
                // call, nor on which junctions have lists (those reads are added in (2), by read id): a scaffold's sums are
                // the same bits in any subset of the set, on any number of GPUs.
                for (int64_t pc = ps; pc < pe; pc += COMPOSE_PART_CHUNK) {
                    const int np = (int)(pe - pc < COMPOSE_PART_CHUNK ? pe - pc : COMPOSE_PART_CHUNK);
                    if (pe - ps > COMPOSE_PART_CHUNK) {  // (else the only chunk is still staged)
                        __syncthreads();
                        stage_parts(pc, np);
                        __syncthreads();
                        stage_steps(np);
                        __syncthreads();
                    }
                    const int nstep = s_qstep[np];
                    for (int it = warp, ip = 0; it < nstep; it += nwarp) {
                        while (it >= s_qstep[ip + 1]) ip++;
                        const int i = 2 * ip + 1;  // the part's slot
                        const int k = (it - s_qstep[ip]) * 32 + lane;
                        if (k < s_pcnt[i]) {
                            const int64_t hb = (int64_t)s_pbase[i] * N;
                            const uint2 e = __ldg(&a.base_hits[hb + k]);
                            const uint32_t p = (uint32_t)(s_pstart[i] + (int32_t)e.y);
                            if (p >= (uint32_t)s_own[i] || row[e.x] != p) continue;
                            row[e.x] = p | ROW_SCORED;
                            int32_t di = a.sc.kmer == 8 ? __ldg(&a.base_hit_di[hb + k]) : -2;
                            if (di == -2) {
                                const BreakWindow bw = break_window((int64_t)p, a.sc.kmer, L);
                                di = dense_index_at(gw, gm, bw.start, bw.len);
                            }
                            TabEntry te;
                            te.prob = 0.0;
                            te.row = -1;
                            if (di >= 0) te = a.sc.tab[di];
                            score_add<true>(a.sc, s_score, st, c, di, te.prob, te.row, 1, &tl);
                        }
                    }
                }
                __syncthreads();
                // (2) the reads only a junction window placed, and the flags off again
                for (int64_t r = tid; r < N; r += nthr) {
                    const uint32_t v = row[r];
                    if (v == POS_INF) continue;
                    if (v & ROW_SCORED) { row[r] = v & ~ROW_SCORED; continue; }
                    const BreakWindow bw = break_window((int64_t)v, a.sc.kmer, L);
                    const int di = dense_index_at(gw, gm, bw.start, bw.len);
                    TabEntry te;
                    te.prob = 0.0;
                    te.row = -1;
                    if (di >= 0) te = a.sc.tab[di];
                    score_add<true>(a.sc, s_score, st, c, di, te.prob, te.row, 1, &tl);
                }
            }
            score_finish<true>(a.sc, s_score, st, c, L, gw, gm, nullptr, row, N, &tl);
        }
    }
}

// ---- KS statistic of a scaffold's rolling-window probabilities from its parts ----
// The windows of a scaffold are the windows of its parts (their ranks were computed once per base contig: a coalesced
// 2-byte load instead of a 16-byte table gather per window) plus the few windows that start before a part and end in it.
// A running "first window not yet counted" walks over the parts, so every window start 0 .. L - kmer is counted once
// whatever the overlaps and however short a part is.  Histogram layout and sweep are k_prob_dist_ks's.
struct KsComposeArgs {
    ProbDistArgs pd;  // scaffolds as the contig set, table, truth side, histogram geometry, ks out (prob_dist must be NULL)
    ScaffoldParts sp;
    const uint16_t *base_rank;
    const int64_t *base_off;
};

template <bool PACKED>
__global__ void __launch_bounds__(768, 2) k_ks_compose(KsComposeArgs b) {
    const ProbDistArgs &a = b.pd;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    __shared__ int64_t s_wsum[32];
    __shared__ int64_t s_wmax[32];
    __shared__ int s_item;
    __shared__ int32_t s_rlo[COMPOSE_PART_CHUNK], s_rn[COMPOSE_PART_CHUNK], s_glo[COMPOSE_PART_CHUNK + 1], s_gn[COMPOSE_PART_CHUNK + 1];
    __shared__ int64_t s_src[COMPOSE_PART_CHUNK];
    uint32_t *s_hist = (uint32_t *)bs_dyn_smem();
    const int kshift = 64 - 2 * a.kmer;
    const uint32_t kbits = keep_bits(a.kmer);
    for (int i = tid; i < a.hist_words; i += nthr) s_hist[i] = 0;
    auto count = [&](int32_t rk) {
        const int lw = hist_logical_word<PACKED>(rk);
        atomicAdd(&s_hist[lw + (lw >> 5)], PACKED ? 1u << (16 * (rk & 1)) : 1u);
    };
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(a.work_counter, 1);
        __syncthreads();
        if (s_item >= a.n_contigs) break;
        const int64_t c = a.order[s_item];
        const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
        const uint64_t *gw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *gm = a.ctg_mask + a.ctg_woff[c];
        int64_t nwin = L - a.kmer + 1;
        if (nwin < 0) nwin = 0;
        const int64_t seg = a.ctg_seg[c];
        const LeLt *yx = a.yx + seg * a.R_x;
        const int64_t n_y = a.R_y > 0 ? a.ycum[seg * a.R_y + a.R_y - 1] : 0;
        auto count_generic = [&](int64_t p) {  // a window read off the scaffold's own words
            const int64_t wi = p >> 5;
            const uint32_t o = (uint32_t)(p & 31);
            int32_t rk = a.rank_zero;
            if (!(window32(__ldg(&gm[wi]), __ldg(&gm[wi + 1]), o) & kbits))
                rk = a.win[(uint32_t)(window64(__ldg(&gw[wi]), __ldg(&gw[wi + 1]), o) >> kshift)].rank;
            count(rk);
        };
        const int64_t ps = b.sp.part_start[c], pe = b.sp.part_start[c + 1];
        for (int64_t pc = ps; pc < pe; pc += COMPOSE_PART_CHUNK) {
            const int np = (int)(pe - pc < COMPOSE_PART_CHUNK ? pe - pc : COMPOSE_PART_CHUNK);
            __syncthreads();
            // window ranges of the chunk's parts, a thread per part: starts and ends of the parts ascend (checked at the entry
            // point), so the first window not yet counted when part g comes up is fixed by part g - 1 alone
            for (int i = tid; i < np; i += nthr) {
                const int64_t g = pc + i;
                const int32_t bb = b.sp.part_base[g];
                const int64_t si = (int64_t)b.sp.part_dst[g] - b.sp.part_ov[g], b0 = b.base_off[bb], ei = si + (b.base_off[bb + 1] - b0);
                int64_t cov = 0;
                if (g > ps) {
                    const int32_t bq = b.sp.part_base[g - 1];
                    const int64_t sq = (int64_t)b.sp.part_dst[g - 1] - b.sp.part_ov[g - 1], eq = sq + (b.base_off[bq + 1] - b.base_off[bq]);
                    cov = eq - a.kmer + 1 > sq ? eq - a.kmer + 1 : sq;
                    if (cov < 0) cov = 0;
                    if (cov > nwin) cov = nwin;
                }
                const int64_t gend = si < nwin ? si : nwin;  // windows that start before the part: off the scaffold
                s_glo[i] = (int32_t)cov;
                s_gn[i] = (int32_t)(gend > cov ? gend - cov : 0);
                if (si > cov) cov = si;
                int64_t last = ei - a.kmer;  // last window that lies inside the part
                if (last > nwin - 1) last = nwin - 1;
                s_rlo[i] = (int32_t)cov;
                s_rn[i] = (int32_t)(last >= cov ? last - cov + 1 : 0);
                s_src[i] = b0 + (cov - si);
                if (g + 1 == pe) {  // after the last part of the scaffold: what is left (nothing, for a well-formed set)
                    if (last + 1 > cov) cov = last + 1;
                    s_glo[np] = (int32_t)cov;
                    s_gn[np] = (int32_t)(nwin > cov ? nwin - cov : 0);
                }
            }
            if (tid == 0 && pc + np < pe) s_gn[np] = 0;
            __syncthreads();
            for (int i = 0; i < np; i++) {
                const uint16_t *src = b.base_rank + s_src[i];
                const int n = s_rn[i];
                for (int k = tid; k < n; k += nthr) count((int32_t)__ldg(&src[k]));
                for (int k = tid; k < s_gn[i]; k += nthr) count_generic((int64_t)s_glo[i] + k);
            }
            for (int k = tid; k < s_gn[np]; k += nthr) count_generic((int64_t)s_glo[np] + k);
        }
        __syncthreads();
        // ---- D = sup |F_x - F_y| at every x value that is present, exact 64-bit numerators (k_prob_dist_ks's).  A scaffold
        // fills most of the histogram, so the sweep is dense: a thread owns a run of consecutive words (an odd number of
        // them: its neighbours' runs start in other banks), sums it, the block scans the sums, the thread walks its run
        // again with the running count in hand and leaves it zeroed.  No bitmap of touched words is kept. ----
        const bool defined = nwin > 0 && n_y > 0;
        int64_t best = 0;
        const int nwords = hist_logical_words(a.R_x, PACKED);
        const int per = ((nwords + nthr - 1) / nthr) | 1;
        const int w_lo = tid * per < nwords ? tid * per : nwords, w_hi = w_lo + per < nwords ? w_lo + per : nwords;
        uint32_t cnt_t = 0;
        for (int w = w_lo; w < w_hi; w++) {
            const uint32_t v = s_hist[w + (w >> 5)];
            cnt_t += PACKED ? (v & 0xffffu) + (v >> 16) : v;
        }
        uint32_t incl = cnt_t;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL_MASK, incl, dd);
            if (lane >= dd) incl += o;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t run = incl - cnt_t;
        for (int w = 0; w < warp; w++) run += (uint32_t)s_wsum[w];
        for (int w = w_lo; w < w_hi; w++) {
            const int pw = w + (w >> 5);
            const uint32_t v = s_hist[pw];
            if (v == 0u) continue;
            s_hist[pw] = 0;
#pragma unroll
            for (int h = 0; h < (PACKED ? 2 : 1); h++) {
                const uint32_t cnt = PACKED ? (v >> (16 * h)) & 0xffffu : v;
                if (cnt == 0 || !defined) continue;
                const int i = PACKED ? 2 * w + h : w;
                const int64_t d = ks_numerator(run, run + cnt, yx[i], (uint32_t)n_y, (uint32_t)nwin);
                run += cnt;
                if (d > best) best = d;
            }
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            const int64_t o = __shfl_xor_sync(FULL_MASK, best, m);
            if (o > best) best = o;
        }
        if (lane == 0) s_wmax[warp] = best;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < nwarp; w++) if (s_wmax[w] > best) best = s_wmax[w];
            a.ks[c] = defined ? (double)best / ((double)nwin * (double)n_y) : __longlong_as_double(0x7ff8000000000000ll);
        }
    }
}

}  // namespace bs
