// CUDA kernels of the breakage scorer (sm_100a).  Integer compare-and-count work: no tensor
// cores; the rules that matter are coalesced streaming, shared-memory staging of the contig
// and grids sized to the 148 SMs.
//
// The same source compiles under tests/emul/cuda_emul.h (BS_CPU_EMUL) so that the device
// algorithm can be checked against the oracle on a machine without a GPU; that build is test
// infrastructure, not a fallback.
#pragma once
#include "bs_common.h"

#ifndef BS_CPU_EMUL
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned char *bs_dyn_smem() {
    extern __shared__ __align__(16) unsigned char bs_smem_raw[];
    return bs_smem_raw;
}
#define BS_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

namespace bs {

constexpr unsigned FULL_MASK = 0xffffffffu;

// ------------------------------------------------------------------------------------------
// descriptors (plain structs passed by value)
// ------------------------------------------------------------------------------------------

// a set of strings packed AoS: string i owns words [woff[i], woff[i+1]) = ceil(L_i/32) + 2
// (two pad words: sequence 0, mask all-ones) and the same range of mask words
struct SeqSet {
    const uint8_t *chars;  // ASCII
    const int64_t *off;    // [n+1] char offsets
    const int64_t *woff;   // [n+1] word offsets
    uint64_t *words;
    uint32_t *mask;
    int64_t n;
    int64_t total_words;
};

// reads packed SoA: word j of read i at words[j * n + i]
struct ReadSet {
    const uint8_t *chars;
    const int64_t *off;  // NULL => uniform_len, dense
    int32_t uniform_len;
    int32_t W;  // words per read
    int64_t n;
    uint64_t *words;
    uint8_t *flags;  // bit0: read holds a byte outside ACGT
};

BS_HD int64_t read_begin(const ReadSet &r, int64_t i) { return r.off ? r.off[i] : i * (int64_t)r.uniform_len; }
BS_HD int32_t read_length(const ReadSet &r, int64_t i) { return r.off ? (int32_t)(r.off[i + 1] - r.off[i]) : r.uniform_len; }

// ------------------------------------------------------------------------------------------
// 2-bit packing
// ------------------------------------------------------------------------------------------

// one thread per packed word of a SeqSet
__global__ void k_pack_seqs(SeqSet s) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < s.total_words; wi += stride) {
        // string owning word wi: largest i with woff[i] <= wi
        int64_t lo = 0, hi = s.n - 1;
        while (lo < hi) {
            int64_t mid = (lo + hi + 1) >> 1;
            if (s.woff[mid] <= wi) lo = mid; else hi = mid - 1;
        }
        const int64_t c0 = s.off[lo];
        const int64_t L = s.off[lo + 1] - c0;
        const int64_t b0 = (wi - s.woff[lo]) * 32;
        uint64_t w = 0;
        uint32_t m = 0;
        for (int b = 0; b < 32; b++) {
            const int64_t p = b0 + b;
            uint32_t code = 0, bad = 1;
            if (p < L) {
                const uint32_t ch = s.chars[c0 + p];
                if (base_valid(ch)) { code = base_code(ch); bad = 0; }
            }
            w = (w << 2) | code;
            m = (m << 1) | bad;
        }
        s.words[wi] = w;
        s.mask[wi] = m;
    }
}

// one thread per read: all W words + the "has non-ACGT byte" flag
__global__ void k_pack_reads(ReadSet r) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < r.n; i += stride) {
        const int64_t c0 = read_begin(r, i);
        const int32_t len = read_length(r, i);
        uint32_t any_bad = 0;
        for (int j = 0; j < r.W; j++) {
            uint64_t w = 0;
            for (int b = 0; b < 32; b++) {
                const int p = j * 32 + b;
                uint32_t code = 0;
                if (p < len) {
                    const uint32_t ch = r.chars[c0 + p];
                    if (base_valid(ch)) code = base_code(ch); else any_bad = 1;
                }
                w = (w << 2) | code;
            }
            r.words[(int64_t)j * r.n + i] = w;
        }
        r.flags[i] = (uint8_t)any_bad;
    }
}

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
        const uint64_t rw = (j == 0) ? r0 : __ldg(&rd.words[(int64_t)j * rd.n + n]);
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
                    const uint64_t r0 = __ldg(&a.reads.words[n]);
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
// block-level helpers
// ------------------------------------------------------------------------------------------

// sum of v over the block, fixed tree => bit-reproducible; result valid on every thread
__device__ __forceinline__ double block_sum(double v, double *s_red) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    __syncthreads();
    s_red[tid] = v;
    __syncthreads();
    for (int s = nthr >> 1; s > 0; s >>= 1) {
        if (tid < s) s_red[tid] += s_red[tid + s];
        __syncthreads();
    }
    return s_red[0];
}
__device__ __forceinline__ double block_max(double v, double *s_red) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    __syncthreads();
    s_red[tid] = v;
    __syncthreads();
    for (int s = nthr >> 1; s > 0; s >>= 1) {
        if (tid < s) { const double o = s_red[tid + s]; if (o > s_red[tid]) s_red[tid] = o; }
        __syncthreads();
    }
    return s_red[0];
}
// exclusive prefix of v over the block (thread order); s_scan holds blockDim ints
__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *s_scan, int64_t *total) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    __syncthreads();
    s_scan[tid] = v;
    __syncthreads();
    for (int d = 1; d < nthr; d <<= 1) {
        int64_t add = (tid >= d) ? s_scan[tid - d] : 0;
        __syncthreads();
        s_scan[tid] += add;
        __syncthreads();
    }
    if (total) *total = s_scan[nthr - 1];
    return s_scan[tid] - v;
}

// ------------------------------------------------------------------------------------------
// break k-mers -> probability-weighted sums (upstream lib/BreakageScorer.cpp:244-303)
// ------------------------------------------------------------------------------------------

struct ScoreArgs {
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const int32_t *w;
    const int32_t *total;
    const double *tab_prob;  // [DENSE_SIZE] probability of a dense k-mer index (0.0 if absent)
    const int32_t *tab_row;  // [DENSE_SIZE] table row of a dense k-mer index (-1 if absent)
    int32_t kmer;
    int32_t T;
    int64_t n_contigs;
    int32_t *sequence_len;
    double *bp_score;
    double *norm_by_break_freqs;
    double *norm_by_len;
    int32_t *kmer_breaks;
    int32_t *hist;  // optional [C][T+1]
};

// one block per contig; sums in position order with a fixed tree (bit-reproducible, the same
// for every GPU count)
__global__ void k_break_score(ScoreArgs a) {
    double *s_red = (double *)bs_dyn_smem();
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int64_t c = blockIdx.x; c < a.n_contigs; c += gridDim.x) {
        const int64_t coff = a.ctg_off[c];
        const int64_t L = a.ctg_off[c + 1] - coff;
        const uint64_t *gw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *gm = a.ctg_mask + a.ctg_woff[c];
        const int32_t *w = a.w + coff + c;
        const int32_t total = a.total[c];
        double s1 = 0.0, s2 = 0.0;
        if (total != 0) {
            const int64_t np = L > 0 ? L : 1;
            for (int64_t p = tid; p < np; p += nthr) {
                const int32_t wv = w[p];
                if (wv == 0) continue;
                const BreakWindow bw = break_window(p, a.kmer, L);
                const int di = dense_index_at(gw, gm, bw.start, bw.len);
                const int32_t row = di >= 0 ? a.tab_row[di] : -1;
                if (row >= 0) {
                    const double pr = a.tab_prob[di];
                    s1 += pr * (double)wv;
                    s2 += pr * ((double)wv / (double)total);
                }
                if (a.hist) atomicAdd(&a.hist[c * (int64_t)(a.T + 1) + (row >= 0 ? row : a.T)], wv);
            }
        }
        s1 = block_sum(s1, s_red);
        s2 = block_sum(s2, s_red);
        if (tid == 0) {
            if (a.sequence_len) a.sequence_len[c] = (int32_t)L;
            if (a.bp_score) a.bp_score[c] = s1;
            if (a.norm_by_break_freqs) a.norm_by_break_freqs[c] = s2;
            if (a.norm_by_len) a.norm_by_len[c] = s1 / (double)(int32_t)L;
            if (a.kmer_breaks) a.kmer_breaks[c] = total;
        }
    }
}

// ------------------------------------------------------------------------------------------
// truth-side distribution for the KS statistics (kmer_from_seq, upstream
// lib/GenerateReads.R:243-259): per segment, counts of truth windows per distinct table value
// ------------------------------------------------------------------------------------------

struct SpectrumArgs {
    const int64_t *tr_off;
    const int64_t *tr_woff;
    const uint64_t *tr_words;
    const uint32_t *tr_mask;
    const int32_t *rank_y;  // [4^kmer] rank of the window's truth-table value, -1 = not in table (NA)
    int32_t *ycnt;          // [S][R_y]
    int32_t R_y;
    int32_t kmer;
    int32_t blocks_per_seg;
};

__global__ void k_truth_spectrum(SpectrumArgs a) {
    const int64_t s = blockIdx.x / a.blocks_per_seg;
    const int b = blockIdx.x % a.blocks_per_seg;
    const int64_t L = a.tr_off[s + 1] - a.tr_off[s];
    const uint64_t *gw = a.tr_words + a.tr_woff[s];
    const uint32_t *gm = a.tr_mask + a.tr_woff[s];
    const int64_t nwin = L - a.kmer + 1;
    const int doff = dense_offset(a.kmer);
    for (int64_t p = (int64_t)b * blockDim.x + threadIdx.x; p < nwin; p += (int64_t)a.blocks_per_seg * blockDim.x) {
        const int di = dense_index_at(gw, gm, p, a.kmer);
        if (di < 0) continue;
        const int32_t rk = a.rank_y[di - doff];
        if (rk >= 0) atomicAdd(&a.ycnt[s * a.R_y + rk], 1);
    }
}

// in-place inclusive prefix sum of every row of a [rows][R] int32 matrix; one block per row
__global__ void k_row_cumsum(int32_t *m, int32_t R) {
    int64_t *s_scan = (int64_t *)bs_dyn_smem();
    int32_t *row = m + (int64_t)blockIdx.x * R;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int per = (R + nthr - 1) / nthr;
    const int lo = tid * per, hi = (lo + per < R) ? lo + per : R;
    int64_t sum = 0;
    for (int i = lo; i < hi; i++) sum += row[i];
    int64_t run = block_exclusive_scan(sum, s_scan, nullptr);
    for (int i = lo; i < hi; i++) { run += row[i]; row[i] = (int32_t)run; }
}

// ------------------------------------------------------------------------------------------
// rolling-window probabilities + KS statistic against the truth distribution
// (upstream lib/BreakageScorer.cpp:200-215 and lib/DeNovoAssembler.R:416-424)
// ------------------------------------------------------------------------------------------

struct ProbDistArgs {
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const int32_t *ctg_seg;  // [C] segment of a contig
    const double *tab_prob;
    const int32_t *tab_row;
    const int32_t *rank_x;   // [4^kmer] rank of a window's scoring-table value among xv (0.0 included)
    const int32_t *le_idx;   // [R_x] index into the y cumulative counts of the last y value <= xv[k], -1 none
    const int32_t *lt_idx;   // [R_x] ... of the last y value < xv[k], -1 none
    const int32_t *ycum;     // [S][R_y] inclusive cumulative counts of the truth distribution
    int32_t R_x, R_y;
    int32_t rank_zero;       // rank of the value 0.0 (windows not in the table)
    int32_t kmer;
    int64_t n_contigs;
    double *prob_dist;       // optional
    const int64_t *pd_off;
    double *ks;              // optional [C]
    int32_t *rank_scratch;   // [gridDim][R_x] global rank histogram when R_x does not fit shared memory, else NULL
};

BS_HD size_t probdist_smem_bytes(int R_x, int nthr, bool hist_in_smem) {
    return (hist_in_smem ? (size_t)R_x * 4 : 0) + (size_t)nthr * 8 + 16;
}

__global__ void k_prob_dist_ks(ProbDistArgs a) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    unsigned char *sm = bs_dyn_smem();
    double *s_red = (double *)sm;
    int64_t *s_scan = (int64_t *)sm;  // same storage, used at different times
    int32_t *s_hist = a.rank_scratch ? a.rank_scratch + (int64_t)blockIdx.x * a.R_x : (int32_t *)(sm + (size_t)nthr * 8);
    const bool want_ks = a.ks != nullptr;
    const int doff = (a.kmer >= 1 && a.kmer <= MAXK) ? dense_offset(a.kmer) : 0;
    for (int64_t c = blockIdx.x; c < a.n_contigs; c += gridDim.x) {
        const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
        const uint64_t *gw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *gm = a.ctg_mask + a.ctg_woff[c];
        int64_t nwin = L - a.kmer + 1;
        if (nwin < 0) nwin = 0;
        if (want_ks) {
            __syncthreads();
            for (int i = tid; i < a.R_x; i += nthr) s_hist[i] = 0;
            __syncthreads();
        }
        double *pd = a.prob_dist ? a.prob_dist + a.pd_off[c] : nullptr;
        for (int64_t p = tid; p < nwin; p += nthr) {
            const int di = dense_index_at(gw, gm, p, a.kmer);
            double val = 0.0;
            int32_t rk = a.rank_zero;
            if (di >= 0 && a.tab_row[di] >= 0) {
                val = a.tab_prob[di];
                if (want_ks) rk = a.rank_x[di - doff];
            }
            if (pd) pd[p] = val;
            if (want_ks) atomicAdd(&s_hist[rk], 1);
        }
        if (!want_ks) continue;
        __syncthreads();
        // D = sup |F_x - F_y| over the pooled distinct values, evaluated at every x value that
        // is present: just below it (F_x of the previous step vs #{y < v}) and at it.
        const int64_t seg = a.ctg_seg[c];
        const int32_t *ycum = a.ycum + seg * a.R_y;
        const int64_t n_y = a.R_y > 0 ? ycum[a.R_y - 1] : 0;
        const int per = (a.R_x + nthr - 1) / nthr;
        const int lo = tid * per, hi = (lo + per < a.R_x) ? lo + per : a.R_x;
        int64_t sum = 0;
        for (int i = lo; i < hi; i++) sum += s_hist[i];
        int64_t run = block_exclusive_scan(sum, s_scan, nullptr);
        double d = 0.0;
        if (nwin > 0 && n_y > 0) {
            const double inx = (double)nwin, iny = (double)n_y;
            for (int i = lo; i < hi; i++) {
                const int32_t cnt = s_hist[i];
                if (cnt == 0) continue;
                const int32_t li = a.lt_idx[i], ei = a.le_idx[i];
                const double lt = li >= 0 ? (double)ycum[li] : 0.0;
                const double le = ei >= 0 ? (double)ycum[ei] : 0.0;
                double d1 = (double)run / inx - lt / iny;
                run += cnt;
                double d2 = (double)run / inx - le / iny;
                if (d1 < 0) d1 = -d1;
                if (d2 < 0) d2 = -d2;
                if (d1 > d) d = d1;
                if (d2 > d) d = d2;
            }
        }
        d = block_max(d, s_red);
        if (tid == 0) a.ks[c] = (nwin > 0 && n_y > 0) ? d : __longlong_as_double(0x7ff8000000000000ll);
    }
}

// ------------------------------------------------------------------------------------------
// KS statistic of the normalised break histogram (variant lib/DeNovoAssembler.cpp:395-420):
// x = count_row / total for every table row, y = truth distribution
// ------------------------------------------------------------------------------------------

struct PathFreqArgs {
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const int32_t *ctg_seg;
    const int32_t *w;
    const int32_t *total;
    const int32_t *tab_row;
    const double *yv;      // [R_y] sorted distinct truth-table values
    const int32_t *ycum;   // [S][R_y]
    int32_t *scratch;      // [gridDim][T+1], all zero on entry and on exit
    int32_t *ovf_cnt;      // [gridDim][OVF_CAP] counts >= CC_DENSE (unsorted), and
    int32_t *status;       // set to 1 if a block ran out of overflow space
    int32_t R_y;
    int32_t kmer;
    int32_t T;
    int64_t n_contigs;
    double *ks;            // [C]
};
constexpr int CC_DENSE = 4096;  // counts below this are tallied in a dense shared-memory array
constexpr int OVF_CAP = 4096;   // per-block capacity for larger counts

// number of y values <= v (le) and < v (lt), from the cumulative counts
__device__ __forceinline__ void y_counts_at(const double *yv, const int32_t *ycum, int R_y, double v,
                                            double *le, double *lt) {
    // first index with yv > v
    int lo = 0, hi = R_y;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (yv[mid] <= v) lo = mid + 1; else hi = mid; }
    const int ule = lo;  // #distinct <= v
    lo = 0; hi = R_y;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (yv[mid] < v) lo = mid + 1; else hi = mid; }
    const int ult = lo;  // #distinct < v
    *le = ule > 0 ? (double)ycum[ule - 1] : 0.0;
    *lt = ult > 0 ? (double)ycum[ult - 1] : 0.0;
}

__global__ void k_ks_path_freq(PathFreqArgs a) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    unsigned char *sm = bs_dyn_smem();
    double *s_red = (double *)sm;
    int64_t *s_scan = (int64_t *)sm;
    int32_t *s_cc = (int32_t *)(sm + (size_t)nthr * 8);  // [CC_DENSE] rows having count j
    __shared__ int s_novf;
    int32_t *scratch = a.scratch + (int64_t)blockIdx.x * (a.T + 1);
    int32_t *ovf = a.ovf_cnt + (int64_t)blockIdx.x * OVF_CAP;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    for (int64_t c = blockIdx.x; c < a.n_contigs; c += gridDim.x) {
        const int64_t coff = a.ctg_off[c];
        const int64_t L = a.ctg_off[c + 1] - coff;
        const int32_t total = a.total[c];
        if (total == 0) {  // 0/0 for every row: R drops the NaNs and ks.test stops on empty x
            if (tid == 0) a.ks[c] = qnan;
            continue;
        }
        const uint64_t *gw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *gm = a.ctg_mask + a.ctg_woff[c];
        const int32_t *w = a.w + coff + c;
        const int64_t np = L > 0 ? L : 1;
        __syncthreads();
        for (int i = tid; i < CC_DENSE; i += nthr) s_cc[i] = 0;
        if (tid == 0) s_novf = 0;
        __syncthreads();
        // pass 1: per-row counts in the block's scratch histogram
        for (int64_t p = tid; p < np; p += nthr) {
            const int32_t wv = w[p];
            if (wv == 0) continue;
            const BreakWindow bw = break_window(p, a.kmer, L);
            const int di = dense_index_at(gw, gm, bw.start, bw.len);
            const int32_t row = di >= 0 ? a.tab_row[di] : -1;
            if (row >= 0) atomicAdd(&scratch[row], wv);
        }
        __syncthreads();
        // pass 2: whoever swaps a row's count out first owns it; tally rows per count value
        for (int64_t p = tid; p < np; p += nthr) {
            const int32_t wv = w[p];
            if (wv == 0) continue;
            const BreakWindow bw = break_window(p, a.kmer, L);
            const int di = dense_index_at(gw, gm, bw.start, bw.len);
            const int32_t row = di >= 0 ? a.tab_row[di] : -1;
            if (row < 0) continue;
            const int32_t cnt = atomicExch(&scratch[row], 0);
            if (cnt == 0) continue;
            if (cnt < CC_DENSE) atomicAdd(&s_cc[cnt], 1);
            else {
                const int slot = atomicAdd(&s_novf, 1);
                if (slot < OVF_CAP) ovf[slot] = cnt; else *a.status = 1;
            }
        }
        __syncthreads();
        // rows with a non-zero count, to get the number of zero rows
        const int per = CC_DENSE / nthr > 0 ? (CC_DENSE + nthr - 1) / nthr : 1;
        const int lo = tid * per, hi = (lo + per < CC_DENSE) ? lo + per : CC_DENSE;
        int64_t sum = 0;
        for (int j = (lo > 0 ? lo : 1); j < hi; j++) sum += s_cc[j];
        int64_t nz_dense = 0;
        int64_t run = block_exclusive_scan(sum, s_scan, &nz_dense);
        const int novf = s_novf < OVF_CAP ? s_novf : OVF_CAP;
        const int64_t n_zero = (int64_t)a.T - nz_dense - novf;
        const int32_t *ycum = a.ycum + (int64_t)a.ctg_seg[c] * a.R_y;
        const int64_t n_y = a.R_y > 0 ? ycum[a.R_y - 1] : 0;
        double d = 0.0;
        if (n_y > 0 && a.T > 0) {
            const double inx = (double)a.T, iny = (double)n_y;
            // x value 0 (rows never broken)
            if (tid == 0 && n_zero > 0) {
                double le, lt;
                y_counts_at(a.yv, ycum, a.R_y, 0.0, &le, &lt);
                double d1 = lt / iny, d2 = (double)n_zero / inx - le / iny;
                if (d1 < 0) d1 = -d1;
                if (d2 < 0) d2 = -d2;
                d = d1 > d2 ? d1 : d2;
            }
            run += n_zero;
            for (int j = (lo > 0 ? lo : 1); j < hi; j++) {
                const int32_t cnt = s_cc[j];
                if (cnt == 0) continue;
                const double v = (double)j / (double)total;
                double le, lt;
                y_counts_at(a.yv, ycum, a.R_y, v, &le, &lt);
                double d1 = (double)run / inx - lt / iny;
                run += cnt;
                double d2 = (double)run / inx - le / iny;
                if (d1 < 0) d1 = -d1;
                if (d2 < 0) d2 = -d2;
                if (d1 > d) d = d1;
                if (d2 > d) d = d2;
            }
            // counts >= CC_DENSE: few; thread 0 walks them in ascending order
            if (tid == 0 && novf > 0) {
                int64_t below = n_zero + nz_dense;
                int32_t last = CC_DENSE - 1;
                for (int done = 0; done < novf;) {
                    int32_t cur = 0x7fffffff;
                    int mult = 0;
                    for (int i = 0; i < novf; i++) {
                        const int32_t v = ovf[i];
                        if (v > last && v < cur) { cur = v; mult = 1; }
                        else if (v == cur) mult++;
                    }
                    const double v = (double)cur / (double)total;
                    double le, lt;
                    y_counts_at(a.yv, ycum, a.R_y, v, &le, &lt);
                    double d1 = (double)below / inx - lt / iny;
                    below += mult;
                    double d2 = (double)below / inx - le / iny;
                    if (d1 < 0) d1 = -d1;
                    if (d2 < 0) d2 = -d2;
                    if (d1 > d) d = d1;
                    if (d2 > d) d = d2;
                    last = cur;
                    done += mult;
                }
            }
        }
        d = block_max(d, s_red);
        if (tid == 0) a.ks[c] = (n_y > 0 && a.T > 0) ? d : qnan;
    }
}

// ------------------------------------------------------------------------------------------
// contig-in-truth offset (upstream lib/BreakageScorer.cpp:273-274): leftmost exact occurrence
// of the whole contig in its segment's truth, assigned only if at least one read was placed
// ------------------------------------------------------------------------------------------

struct StartposArgs {
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const uint8_t *ctg_chars;
    const int32_t *ctg_seg;
    const int64_t *tr_off;
    const int64_t *tr_woff;
    const uint64_t *tr_words;
    const uint32_t *tr_mask;
    const uint8_t *tr_chars;
    const int32_t *total;
    int64_t n_contigs;
    int32_t *startpos;
};

__global__ void k_startpos(StartposArgs a) {
    __shared__ int s_best;
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int64_t c = blockIdx.x; c < a.n_contigs; c += gridDim.x) {
        const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
        const int64_t seg = a.ctg_seg[c];
        const int64_t LT = a.tr_off[seg + 1] - a.tr_off[seg];
        if (a.total[c] == 0 || L == 0 || L > LT) {
            // no read placed: upstream never assigns (stays 0); "" is found at 0; too long: npos
            if (tid == 0) a.startpos[c] = (a.total[c] != 0 && L > LT) ? -1 : 0;
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
        if (tid == 0) a.startpos[c] = s_best == 0x7fffffff ? -1 : s_best;
    }
}

}  // namespace bs
