// Placement kernels: leftmost exact occurrence of every read in every contig of its segment.
#pragma once
#include "bs_seq.cuh"

namespace bs {

// ------------------------------------------------------------------------------------------
// placement: leftmost exact occurrence of every read of a segment in one contig
// (upstream lib/BreakageScorer.cpp:235-243, std::string::find)
// ------------------------------------------------------------------------------------------

struct PlaceItem {
    int32_t contig;
    int32_t seg;
    int64_t read_begin;  // global read indices [read_begin, read_end)
    int64_t read_end;
};

struct PlaceArgs {
    const PlaceItem *items;
    // contigs
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const uint8_t *ctg_chars;
    // reads
    ReadSet reads;
    const int32_t *seg_seed_len;    // [S] bases of the seed (1..32) = min read length of the segment, capped
    const int64_t *seg_read_start;  // [S+1]
    // outputs
    int32_t *w;      // position weights: contig c, position p at ctg_off[c] + c + p
    int32_t *total;  // [C] number of reads placed (kmer_breaks)
    int32_t *pos;    // optional: contig c, read n at pos_off[c] + (n - seg_read_start[seg])
    const int64_t *pos_off;
    int32_t tile_len;   // positions per shared-memory tile (multiple of 32, <= 65504)
    int32_t hash_size;  // power of two
    int32_t found_words;  // shared-memory words reserved for the "already placed" bitmask
    int32_t scan_mode;  // 1: exhaustive scan instead of the seed index
};

BS_HD size_t place_smem_bytes(int tile_len, int hash_size, int W, int found_words) {
    const size_t tile_words = (size_t)tile_len / 32 + W + 2;
    return tile_words * 8 + (size_t)hash_size * 4 + tile_words * 4 + (size_t)found_words * 4 + (size_t)tile_len * 2 + 16;
}

// does read n (all-ACGT, length len) occur at tile-relative position pr?
__device__ __forceinline__ bool match_packed(const uint64_t *s_seq, const uint32_t *s_mask, int pr,
                                             const ReadSet &rd, int64_t n, int len, uint64_t r0) {
    const int idx = pr >> 5;
    const uint32_t o = (uint32_t)(pr & 31);
    const int nw = (len + 31) >> 5;
    for (int j = 0; j < nw; j++) {
        const int rem = len - 32 * j;
        const uint64_t rw = (j == 0) ? r0 : __ldg(&rd.words[n * rd.W + j]);
        const uint64_t cw = window64(s_seq[idx + j], s_seq[idx + j + 1], o);
        if ((cw ^ rw) & keep_bases(rem)) return false;
        if (window32(s_mask[idx + j], s_mask[idx + j + 1], o) & keep_bits(rem)) return false;
    }
    return true;
}

__global__ void k_place(PlaceArgs a) {
    const PlaceItem it = a.items[blockIdx.x];
    const int c = it.contig;
    const int64_t coff = a.ctg_off[c];
    const int64_t L = a.ctg_off[c + 1] - coff;
    const int64_t nW = a.ctg_woff[c + 1] - a.ctg_woff[c];
    const uint64_t *gw = a.ctg_words + a.ctg_woff[c];
    const uint32_t *gm = a.ctg_mask + a.ctg_woff[c];
    const int S = a.seg_seed_len[it.seg];
    const int Lt = a.tile_len;
    const int H = a.hash_size;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int tile_words = Lt / 32 + a.reads.W + 2;

    unsigned char *sm = bs_dyn_smem();
    uint64_t *s_seq = (uint64_t *)sm;           sm += (size_t)tile_words * 8;
    uint32_t *s_head = (uint32_t *)sm;          sm += (size_t)H * 4;
    uint32_t *s_mask = (uint32_t *)sm;          sm += (size_t)tile_words * 4;
    uint32_t *s_found = (uint32_t *)sm;         sm += (size_t)a.found_words * 4;
    uint16_t *s_next = (uint16_t *)sm;

    const int64_t nchunk = it.read_end - it.read_begin;
    const bool multi_tile = L > Lt;
    if (multi_tile) {
        for (int i = tid; i < a.found_words; i += nthr) s_found[i] = 0;
    }
    int placed = 0;  // meaningful on lane 0 of each warp

    for (int64_t t0 = 0; t0 == 0 || t0 < L; t0 += Lt) {
        // ---- stage the tile: packed bases + validity mask (with read-length overhang) ----
        __syncthreads();
        for (int i = tid; i < tile_words; i += nthr) {
            const int64_t gi = (t0 >> 5) + i;
            uint64_t wv = 0;
            uint32_t mv = ~0u;
            if (gi < nW) { wv = gw[gi]; mv = gm[gi]; }
            s_seq[i] = wv;
            s_mask[i] = mv;
        }
        for (int i = tid; i < H; i += nthr) s_head[i] = 0;
        __syncthreads();
        // ---- seed index: hash of the S-base window at every position -> chain of positions ----
        if (!a.scan_mode) {
            for (int pr = tid; pr < Lt; pr += nthr) {
                if (t0 + pr + S > L) break;
                const int idx = pr >> 5;
                const uint32_t o = (uint32_t)(pr & 31);
                if (window32(s_mask[idx], s_mask[idx + 1], o) & keep_bits(S)) continue;  // window holds a non-ACGT base
                const uint64_t seed = window64(s_seq[idx], s_seq[idx + 1], o) & keep_bases(S);
                const uint32_t h = seed_hash(seed) & (uint32_t)(H - 1);
                const uint32_t old = atomicExch(&s_head[h], (uint32_t)pr + 1u);
                s_next[pr] = (uint16_t)old;
            }
        }
        __syncthreads();
        // ---- stream the reads of the segment past the tile ----
        for (int64_t base = 0; base < nchunk; base += nthr) {
            const int64_t k = base + tid;
            bool hit = false;
            if (k < nchunk && !(multi_tile && ((s_found[k >> 5] >> (k & 31)) & 1u))) {
                const int64_t n = it.read_begin + k;
                const int len = read_length(a.reads, n);
                int64_t best = -1;
                if (len == 0) {
                    if (t0 == 0) best = 0;  // std::string::find("") == 0
                } else if (a.reads.flags[n] & 1) {
                    // read with bytes outside ACGT: exact byte comparison against the contig text
                    const uint8_t *rc = a.reads.chars + read_begin(a.reads, n);
                    const uint8_t *cc = a.ctg_chars + coff;
                    int64_t pend = t0 + Lt;
                    if (pend > L - len + 1) pend = L - len + 1;
                    for (int64_t p = t0; p < pend; p++) {
                        int i = 0;
                        while (i < len && cc[p + i] == rc[i]) i++;
                        if (i == len) { best = p; break; }
                    }
                } else {
                    const uint64_t r0 = __ldg(&a.reads.words[n * a.reads.W]);
                    if (a.scan_mode) {
                        for (int pr = 0; pr < Lt; pr++) {
                            if (t0 + pr + len > L) break;
                            if (match_packed(s_seq, s_mask, pr, a.reads, n, len, r0)) { best = t0 + pr; break; }
                        }
                    } else {
                        const uint64_t seed = r0 & keep_bases(S);
                        const uint32_t h = seed_hash(seed) & (uint32_t)(H - 1);
                        int bestpr = 0x7fffffff;
                        for (uint32_t q = s_head[h]; q != 0; q = s_next[q - 1]) {
                            const int pr = (int)q - 1;
                            if (pr >= bestpr || t0 + pr + len > L) continue;
                            if (match_packed(s_seq, s_mask, pr, a.reads, n, len, r0)) bestpr = pr;
                        }
                        if (bestpr != 0x7fffffff) best = t0 + bestpr;
                    }
                }
                if (best >= 0) {
                    hit = true;
                    atomicAdd(&a.w[coff + c + best], 1);
                    if (a.pos) a.pos[a.pos_off[c] + (n - a.seg_read_start[it.seg])] = (int32_t)best;
                    if (multi_tile) atomicOr(&s_found[k >> 5], 1u << (k & 31));
                }
            }
            const unsigned m = __ballot_sync(FULL_MASK, hit);
            if ((tid & 31) == 0) placed += __popc(m);
        }
    }
    if ((tid & 31) == 0 && placed) atomicAdd(&a.total[c], placed);
}

// ------------------------------------------------------------------------------------------
// placement through the read index (default).  The contig streams past the segment's read
// index: every contig position looks its seed up, candidates are verified on the packed words
// (XOR of whole words under the validity mask), and the LEFTMOST position per (read, contig)
// wins through an atomicMin on a per-block scratch row -- std::string::find semantics of
// upstream lib/BreakageScorer.cpp:241.  Work per contig is O(L_c + hits) instead of O(U * L_c).
// ------------------------------------------------------------------------------------------

constexpr uint32_t POS_INF = 0x7f7f7f7fu;  // memset-able "no position yet"

struct PlaceIxArgs {
    const int32_t *order;  // [n_items] contig ids, longest first
    int32_t n_items;
    int32_t *work_counter;  // zero on entry
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const uint8_t *ctg_chars;
    const int32_t *ctg_seg;
    ReadSet reads;
    ReadIndex ix;
    uint32_t *best;       // [gridDim][best_stride], all POS_INF on entry and on exit
    int64_t best_stride;  // >= reads of the largest segment
    int32_t *w;           // position weights: contig c, position p at ctg_off[c] + c + p
    int32_t *total;       // [C] reads placed (kmer_breaks)
    int32_t *pos;         // optional
    const int64_t *pos_off;
    int32_t hit_cap;      // shared-memory list of reads placed in the current contig ...
    uint32_t *hits_ovf;   // [gridDim][best_stride] ... and its continuation in global memory (never full: distinct reads)
    // HASHED variant (segments whose read count makes a dense row per resident block too large, cfg-5):
    // `best` is a per-block open-addressed table instead, best_stride = its capacity (a power of two).
    uint32_t *keys;       // [gridDim][best_stride] read id + 1, all 0 on entry and on a successful exit
    int32_t *overflow;    // set when a contig places more than best_stride/2 distinct reads: the host
                          // repeats the launch with a larger table
};

// slot of a read in the hashed scratch (Knuth's multiplicative hash of the read's local id)
__device__ __forceinline__ uint32_t scratch_slot(uint32_t key, uint32_t mask) { return ((key * 2654435761u) >> 7) & mask; }

// leftmost byte-exact occurrence of read n in contig text, or -1 (reads that cannot be packed)
__device__ __forceinline__ int64_t find_bytes(const uint8_t *cc, int64_t L, const uint8_t *rc, int len) {
    if (len == 0) return 0;  // std::string::find("") == 0
    for (int64_t p = 0; p + len <= L; p++) {
        int i = 0;
        while (i < len && cc[p + i] == rc[i]) i++;
        if (i == len) return p;
    }
    return -1;
}

// A contig of ACGT only that is kept as packed words alone (bs_compose.cuh): the bytes of a read against the bases decoded
// from the words.  For reads whose flag is set -- the flag is conservative: it also marks a read that sits next to a byte
// outside ACGT in the read buffer, and such a read can match.
__device__ __forceinline__ bool match_text_packed(const uint64_t *gw, const uint32_t *gm, int64_t p, const uint8_t *rc, int len) {
    for (int i = 0; i < len; i++) {
        const int64_t q = p + i;
        const int o = (int)(q & 31);
        if ((__ldg(&gm[q >> 5]) >> (31 - o)) & 1u) return false;  // (a masked position: past the end)
        const uint32_t code = (uint32_t)(__ldg(&gw[q >> 5]) >> (62 - 2 * o)) & 3u;
        if (rc[i] != (uint8_t)("ACGT"[code])) return false;
    }
    return true;
}

// does read n (length len, packed words rw) equal the contig at position p?
__device__ __forceinline__ bool verify_at(const PlaceIxArgs &a, const uint64_t *gw, const uint32_t *gm, const uint8_t *cc,
                                          int64_t p, int64_t n, int len, uint64_t w0) {
    if (a.reads.flags[n] & 1) {  // bytes outside ACGT somewhere in (or next to) the read: compare the text
        const uint8_t *rc = a.reads.chars + read_begin(a.reads, n);
        if (cc == nullptr) return match_text_packed(gw, gm, p, rc, len);  // (bs_compose.cuh: a contig of ACGT only, kept without its text)
        for (int i = 0; i < len; i++)
            if (cc[p + i] != rc[i]) return false;
        return true;
    }
    const int64_t idx = p >> 5;
    const uint32_t o = (uint32_t)(p & 31);
    const int nw = (len + 31) >> 5;
    const uint64_t *rw = a.reads.words + n * a.reads.W;
    uint64_t c_lo = __ldg(&gw[idx]);
    uint32_t m_lo = __ldg(&gm[idx]);
    for (int j = 0; j < nw; j++) {
        const int rem = len - 32 * j;
        const uint64_t c_hi = __ldg(&gw[idx + j + 1]);
        const uint32_t m_hi = __ldg(&gm[idx + j + 1]);
        const uint64_t r = (j == 0) ? w0 : __ldg(&rw[j]);
        if ((window64(c_lo, c_hi, o) ^ r) & keep_bases(rem)) return false;
        if (window32(m_lo, m_hi, o) & keep_bits(rem)) return false;
        c_lo = c_hi;
        m_lo = m_hi;
    }
    return true;
}

#ifndef BS_PLACE_IX_THREADS
#define BS_PLACE_IX_THREADS 256
#endif
#ifndef BS_PLACE_IX_BLOCKS
#define BS_PLACE_IX_BLOCKS 6
#endif
#ifndef BS_PLACE_HIT_CAP
#define BS_PLACE_HIT_CAP 1024
#endif
constexpr int PLACE_IX_THREADS = BS_PLACE_IX_THREADS, PLACE_IX_BLOCKS = BS_PLACE_IX_BLOCKS;  // resident blocks per SM the kernel is built for
#ifndef BS_PLACE_CAND_CAP
#define BS_PLACE_CAND_CAP 256
#endif
constexpr int PLACE_CAND_CAP = BS_PLACE_CAND_CAP;  // seed hits of one WARP iteration (32 * PLACE_POS positions) awaiting verification
#ifndef BS_PLACE_POS
#define BS_PLACE_POS 4
#endif
constexpr int PLACE_POS = BS_PLACE_POS;  // consecutive contig positions per thread and iteration (divides 32): bucket-head gathers in flight

BS_HD size_t place_index_smem_bytes(int hit_cap, int nthr) { return (size_t)hit_cap * 4 + (size_t)(nthr / 32) * PLACE_CAND_CAP * 8; }

template <bool HASHED>
__global__ void __launch_bounds__(PLACE_IX_THREADS, PLACE_IX_BLOCKS) k_place_index(PlaceIxArgs a) {
    uint32_t *s_hits = (uint32_t *)bs_dyn_smem();
    __shared__ int s_item, s_nhit, s_placed, s_ncand[32];
    __shared__ int s_full;  // HASHED: the current contig does not fit the table
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    // every warp queues and verifies its own candidates: no block barrier inside the position loop
    uint2 *s_cand = (uint2 *)(s_hits + a.hit_cap) + warp * PLACE_CAND_CAP;  // (read id, contig position) with an equal seed tag
    uint32_t *best = a.best + (int64_t)blockIdx.x * a.best_stride;
    uint32_t *hits_ovf = a.hits_ovf + (int64_t)blockIdx.x * a.best_stride;
    uint32_t *keys = HASHED ? a.keys + (int64_t)blockIdx.x * a.best_stride : nullptr;
    const uint32_t smask = (uint32_t)a.best_stride - 1u;
    const int fill_max = (int)(a.best_stride >> 1);
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_item = atomicAdd(a.work_counter, 1); s_nhit = 0; s_placed = 0;
            if (HASHED) {
                s_full = 0;
                if (*(volatile int32_t *)a.overflow) s_item = a.n_items;  // another block gave up: the launch is repeated
            }
        }
        if (lane == 0) s_ncand[warp] = 0;
        __syncthreads();
        const int item = s_item;
        if (item >= a.n_items) break;
        const int c = a.order[item];
        const int s = a.ctg_seg[c];
        const int64_t coff = a.ctg_off[c];
        const int64_t L = a.ctg_off[c + 1] - coff;
        const uint64_t *gw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *gm = a.ctg_mask + a.ctg_woff[c];
        const uint8_t *cc = a.ctg_chars + coff;
        const int64_t r0 = a.ix.seg_read_start[s];
        const int64_t n_seg_reads = a.ix.seg_read_start[s + 1] - r0;
        const int S = a.ix.seed_len[s];
        const uint32_t *head = a.ix.head + a.ix.tab_off[s];
        const uint32_t hmask = (uint32_t)a.ix.tab_mask[s];
        const uint64_t keepS = keep_bases(S);
        const uint32_t keepSm = keep_bits(S);

        // the list of reads placed in this contig (dense: local read ids; hashed: table slots)
        auto remember = [&](uint32_t v) {
            const int slot = atomicAdd(&s_nhit, 1);
            if (slot < a.hit_cap) s_hits[slot] = v;
            else hits_ovf[slot - a.hit_cap] = v;
        };
        // read nl (local id) occurs at position p: keep the leftmost position
        auto record = [&](uint32_t nl, uint32_t p) {
            if constexpr (!HASHED) {
                const uint32_t old = atomicMin(&best[nl], p);
                if (old == POS_INF) remember(nl);  // first time this read is seen in this contig
            } else {
                // at most fill_max (+ one per thread in flight) of the 2 * fill_max slots are ever taken, so
                // the probe ends; all table traffic is atomic (L2), nothing stale out of the L1
                if (*(volatile int *)&s_nhit >= fill_max) { s_full = 1; return; }
                const uint32_t key = nl + 1u;
                for (uint32_t slot = scratch_slot(key, smask);; slot = (slot + 1u) & smask) {
                    const uint32_t old = atomicCAS(&keys[slot], 0u, key);
                    if (old == 0u) remember(slot);
                    if (old == 0u || old == key) { atomicMin(&best[slot], p); break; }
                }
            }
        };
        // a read with an equal seed at position p: verify the rest
        auto verify_and_record = [&](int64_t n, int64_t p) {
            const int len = read_length(a.reads, n);
            if (p + len > L) return;
            const uint64_t w0 = __ldg(&a.reads.words[n * a.reads.W]);
            if (!verify_at(a, gw, gm, cc, p, n, len, w0)) return;
            record((uint32_t)(n - r0), (uint32_t)p);
        };

        if (n_seg_reads > 0) {
            // ---- every contig position against the index.  Phase 1 (sparse, cheap): four positions
            // per thread in flight, bucket head -> chain entry; equal seed tags are queued per warp.
            // Phase 2 (dense): the warp's lanes verify one queued candidate each on the packed words. ----
            for (int64_t p0 = 0; p0 + S <= L; p0 += PLACE_POS * (int64_t)nthr) {
                // a thread takes PLACE_POS CONSECUTIVE positions (never straddling a word boundary)
                const int64_t pb = p0 + PLACE_POS * (int64_t)tid;
                uint32_t q4[PLACE_POS];
                uint64_t seed4[PLACE_POS];
#pragma unroll
                for (int u = 0; u < PLACE_POS; u++) { q4[u] = 0; seed4[u] = 0; }
                if (pb + S <= L) {
                    const int64_t idx = pb >> 5;
                    const uint32_t o = (uint32_t)(pb & 31);
                    const uint64_t cw0 = __ldg(&gw[idx]), cw1 = __ldg(&gw[idx + 1]);
                    const uint32_t cm0 = __ldg(&gm[idx]), cm1 = __ldg(&gm[idx + 1]);
#pragma unroll
                    for (int u = 0; u < PLACE_POS; u++) {
                        if (pb + u + S <= L && !(window32(cm0, cm1, o + u) & keepSm)) {  // else: a non-ACGT base in the seed window
                            seed4[u] = window64(cw0, cw1, o + u) & keepS;
                            q4[u] = head[seed_hash(seed4[u]) & hmask];
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < PLACE_POS; u++) {
                    const int64_t p = pb + u;
                    uint32_t q = q4[u];
                    const uint32_t tag = seed_tag(seed4[u]);
                    while (q != 0) {  // the chain entry carries a 32-bit tag of the read's seed: the packed
                        const int64_t n = (int64_t)q - 1;  // reads themselves are only touched in phase 2
                        const uint2 e = a.ix.next[n];
                        q = e.x;
                        if (e.y != tag) continue;
                        const int slot = atomicAdd(&s_ncand[warp], 1);
                        if (slot < PLACE_CAND_CAP) s_cand[slot] = make_uint2((uint32_t)n, (uint32_t)p);
                        else verify_and_record(n, p);  // queue full (long chains): verify in place
                    }
                }
                __syncwarp();
                const int nc = s_ncand[warp] < PLACE_CAND_CAP ? s_ncand[warp] : PLACE_CAND_CAP;
                for (int i = lane; i < nc; i += 32) {
                    const uint2 cd = s_cand[i];
                    verify_and_record((int64_t)cd.x, (int64_t)cd.y);
                }
                __syncwarp();
                if (lane == 0) s_ncand[warp] = 0;
                __syncwarp();
            }
            // ---- reads outside the index: text comparison, one read per thread ----
            int i = 0;
            for (uint32_t q = a.ix.odd_head[s]; q != 0; i++) {
                const int64_t n = (int64_t)q - 1;
                q = a.ix.next[n].x;
                if (i % nthr != tid) continue;
                const int64_t p = find_bytes(cc, L, a.reads.chars + read_begin(a.reads, n), read_length(a.reads, n));
                if (p < 0) continue;
                if constexpr (!HASHED) {
                    best[n - r0] = (uint32_t)p;  // each such read is visited by exactly one thread
                    remember((uint32_t)(n - r0));
                } else {
                    record((uint32_t)(n - r0), (uint32_t)p);
                }
            }
        }
        __threadfence_block();
        __syncthreads();
        if (HASHED && s_full) {  // (block-uniform) leave the table as it is: the host clears it and repeats the launch
            if (tid == 0) atomicExch(a.overflow, 1);
            break;
        }
        // ---- leftmost positions -> position weights ----
        const int nh = s_nhit;
        int placed = 0;
        for (int h = tid; h < nh; h += nthr) {
            uint32_t nl = h < a.hit_cap ? s_hits[h] : hits_ovf[h - a.hit_cap];
            const uint32_t p = atomicExch(&best[nl], POS_INF);
            if (HASHED) nl = atomicExch(&keys[nl], 0u) - 1u;  // slot -> the read that owns it
            atomicAdd(&a.w[coff + c + p], 1);
            if (a.pos) a.pos[a.pos_off[c] + nl] = (int32_t)p;
            placed++;
        }
        if (placed) atomicAdd(&s_placed, placed);
        __syncthreads();
        if (tid == 0) a.total[c] = s_placed;
    }
}

}  // namespace bs
