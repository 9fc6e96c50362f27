// Truth-side distribution and the rolling-window probabilities with their KS statistic.
#pragma once
#include "bs_score.cuh"

namespace bs {

// ------------------------------------------------------------------------------------------
// truth-side distribution for the KS statistics (kmer_from_seq, upstream
// lib/GenerateReads.R:243-259): per segment, counts of truth windows per distinct table value
// ------------------------------------------------------------------------------------------

// per x-value rank: index into the truth cumulative counts of the last y value <= / < that x value
// (table level), or the cumulative counts themselves (segment level, yx below)
struct alignas(8) LeLt {  // one 8-byte gather
    int32_t le, lt;
};

// KS numerator at one x value: |F_x - F_y| * nwin * n_y just below the value (n_lt windows, y_lt truth
// windows smaller) and at it.  Every factor is below 2^31: 32 x 32 -> 64-bit products.
BS_HD int64_t ks_numerator(uint32_t n_lt, uint32_t n_le, LeLt y, uint32_t n_y, uint32_t nwin) {
    int64_t d1 = (int64_t)((uint64_t)n_lt * n_y) - (int64_t)((uint64_t)(uint32_t)y.lt * nwin);
    int64_t d2 = (int64_t)((uint64_t)n_le * n_y) - (int64_t)((uint64_t)(uint32_t)y.le * nwin);
    if (d1 < 0) d1 = -d1;
    if (d2 < 0) d2 = -d2;
    return d1 > d2 ? d1 : d2;
}

struct SpectrumArgs {
    const int64_t *tr_off;
    const int64_t *tr_woff;
    const uint64_t *tr_words;
    const uint32_t *tr_mask;
    const int32_t *rank_y;  // [4^kmer] rank of the window's truth-table value, -1 = not in table (NA)
    int32_t *ycnt;          // [S][R_y]
    const LeLt *lelt;       // [R_x] table-level indices (NULL: no yx output)
    LeLt *yx;               // [S][R_x] #{truth windows <= x value}, #{truth windows < x value} per x-value rank
    int32_t R_x;
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

// The same for truths of fewer than 65 536 windows, a block per segment at a time: counts in shared
// memory (two 16-bit counters per word), cumulated in place, written out once, coalesced.
// TAB: the block also keeps the rank table in shared memory (16-bit ranks, 128 KB for 8-mers) and
// walks over segments: a window's rank is then a shared-memory lookup instead of a divergent
// global gather (32 L1 wavefronts per warp), which is what bounds the plain variant.
constexpr int SPECTRUM_TAB_THREADS = 1024;
constexpr uint32_t SPECTRUM_NO_RANK = 0xffffu;

template <bool TAB>
__global__ void __launch_bounds__(TAB ? SPECTRUM_TAB_THREADS : 512, TAB ? 1 : 3) k_truth_spectrum_smem(SpectrumArgs a, int n_seg) {
    uint32_t *s_h = (uint32_t *)bs_dyn_smem();
    __shared__ uint32_t s_wsum[32];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int kshift = 64 - 2 * a.kmer;
    const uint32_t kbits = keep_bits(a.kmer);
    const int nword = (a.R_y + 1) >> 1;
    uint16_t *s_tab = reinterpret_cast<uint16_t *>(s_h + nword);  // TAB: [4^kmer]
    if (TAB) {
        const int n_codes = 1 << (2 * a.kmer);
        for (int i = tid; i < n_codes; i += nthr) {
            const int32_t r = a.rank_y[i];
            s_tab[i] = (uint16_t)(r < 0 ? SPECTRUM_NO_RANK : (uint32_t)r);
        }
    }
    for (int64_t s = blockIdx.x; s < n_seg; s += gridDim.x) {
        const int64_t L = a.tr_off[s + 1] - a.tr_off[s];
        const uint64_t *gw = a.tr_words + a.tr_woff[s];
        const uint32_t *gm = a.tr_mask + a.tr_woff[s];
        const int64_t nwin = L - a.kmer + 1;
        __syncthreads();  // the previous segment's counts have been written out (and the table is loaded)
        for (int i = tid; i < nword; i += nthr) s_h[i] = 0;
        __syncthreads();
        // a thread takes eight CONSECUTIVE windows (never straddling a 32-base word boundary): the
        // truth words are loaded once per eight windows (the next group's while this one is counted)
        constexpr int PER = 8;
        const int64_t step = PER * (int64_t)nthr;
        int64_t pb = PER * (int64_t)tid;
        uint64_t w0 = 0, w1 = 0;
        uint32_t m0 = 0, m1 = 0;
        if (pb < nwin) { w0 = __ldg(&gw[pb >> 5]); w1 = __ldg(&gw[(pb >> 5) + 1]); m0 = __ldg(&gm[pb >> 5]); m1 = __ldg(&gm[(pb >> 5) + 1]); }
        for (; pb < nwin; pb += step) {
            const uint64_t c0 = w0, c1 = w1;
            const uint32_t d0 = m0, d1 = m1;
            const int64_t nb = pb + step;
            if (nb < nwin) { w0 = __ldg(&gw[nb >> 5]); w1 = __ldg(&gw[(nb >> 5) + 1]); m0 = __ldg(&gm[nb >> 5]); m1 = __ldg(&gm[(nb >> 5) + 1]); }
            const uint32_t o = (uint32_t)(pb & 31);
            int32_t rk[PER];
#pragma unroll
            for (int u = 0; u < PER; u++) {
                rk[u] = -1;
                if (pb + u < nwin && !(window32(d0, d1, o + u) & kbits)) {
                    const uint64_t code = window64(c0, c1, o + u) >> kshift;
                    if (TAB) {
                        const uint32_t r = s_tab[code];
                        rk[u] = r == SPECTRUM_NO_RANK ? -1 : (int32_t)r;
                    } else {
                        rk[u] = __ldg(&a.rank_y[code]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < PER; u++)
                if (rk[u] >= 0) atomicAdd(&s_h[rk[u] >> 1], 1u << (16 * (rk[u] & 1)));
        }
        __syncthreads();
        // inclusive prefix over ranks: every thread owns an even number of consecutive ranks
        int per = (a.R_y + nthr - 1) / nthr;
        per += per & 1;
        const int lo = tid * per < a.R_y ? tid * per : a.R_y;
        const int hi = lo + per < a.R_y ? lo + per : a.R_y;
        uint32_t sum = 0;
        for (int i = lo >> 1; i < (hi + 1) >> 1; i++) { const uint32_t w = s_h[i]; sum += (w & 0xffffu) + (w >> 16); }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
        for (int w = 0; w < warp; w++) run += s_wsum[w];
        for (int i = lo >> 1; i < (hi + 1) >> 1; i++) {
            const uint32_t w = s_h[i];
            const uint32_t c0 = run + (w & 0xffffu), c1 = c0 + (w >> 16);
            s_h[i] = c0 | (c1 << 16);  // totals stay below 65 536
            run = c1;
        }
        __syncthreads();
        int32_t *out = a.ycnt + s * a.R_y;
        for (int i = tid; i < a.R_y; i += nthr) out[i] = (int32_t)((s_h[i >> 1] >> (16 * (i & 1))) & 0xffffu);
        if (a.lelt) {  // the same counts looked up per x-value rank: the KS sweep then needs one gather, not two
            LeLt *yx = a.yx + s * a.R_x;
            for (int i = tid; i < a.R_x; i += nthr) {
                const LeLt q = a.lelt[i];
                LeLt o;
                o.le = q.le >= 0 ? (int32_t)((s_h[q.le >> 1] >> (16 * (q.le & 1))) & 0xffffu) : 0;
                o.lt = q.lt >= 0 ? (int32_t)((s_h[q.lt >> 1] >> (16 * (q.lt & 1))) & 0xffffu) : 0;
                yx[i] = o;
            }
        }
    }
}

// yx from the cumulative counts in global memory (after k_truth_spectrum + k_row_cumsum)
__global__ void k_yx_gather(SpectrumArgs a, int64_t n_seg) {
    const int64_t total = n_seg * a.R_x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = t / a.R_x;
        const int i = (int)(t - s * a.R_x);
        const LeLt q = a.lelt[i];
        const int32_t *ycum = a.ycnt + s * a.R_y;
        LeLt o;
        o.le = q.le >= 0 ? ycum[q.le] : 0;
        o.lt = q.lt >= 0 ? ycum[q.lt] : 0;
        a.yx[t] = o;
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

// one entry per k-mer code of the window length: what a window contributes to path_prob_dist
// (0.0 when the k-mer is not a table row) and the rank of that value among the distinct x values
struct alignas(16) WinEntry {
    double prob;
    int32_t rank;
    int32_t row;  // table row of the k-mer, -1 when it is not a row (what TabEntry holds for the same k-mer)
};

struct ProbDistArgs {
    const int32_t *order;     // [C] contig ids, longest first
    int32_t *work_counter;    // zero on entry
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const int32_t *ctg_seg;  // [C] segment of a contig
    const WinEntry *win;     // [4^kmer], NULL when kmer is outside 1..8 (every window is "not in the table")
    const LeLt *yx;          // [S][R_x] truth windows <= / < every x value
    const int32_t *ycum;     // [S][R_y] inclusive cumulative counts of the truth distribution
    int32_t R_x, R_y;
    int32_t rank_zero;       // rank of the value 0.0 (windows not in the table)
    int32_t kmer;
    int64_t n_contigs;
    double *prob_dist;       // optional
    const int64_t *pd_off;
    double *ks;              // optional [C]
    uint32_t *rank_scratch;  // [gridDim][hist_words + n_ranges] global rank histogram when it does not fit shared memory, else NULL
    int32_t hist_words;      // 32-bit words of one rank histogram (hist_phys_words)
    int32_t n_ranges;        // hist_ranges: ranges of 32 histogram words
    // FUSE instantiations also do k_break_score's work (kmer == 8 only): the break k-mer of contig position p is the
    // rolling window p - 4 (upstream lib/BreakageScorer.cpp:244-267 with kmer/2 == 4 and the fixed expanded size 8), whose
    // table entry this kernel has just gathered -- so the position weights are read beside the windows and the second
    // scan of the contig, its table gathers and its launch disappear.  Positions 0..3 and L-3..L-1 (the start rules and
    // the clamped end) take the generic path.
    ScoreArgs sc;
    int64_t fuse_min_len;    // contigs of at least this many bases are scored on the way (FUSE_MIN_LEN unless tuned)
};

// Which contigs are scored on the way: those of at least FUSE_MIN_LEN bases (16 kb).  Measured (profiles/r02o): one scan instead of
// two pays for long contigs (cfg-4's 32 kb scaffolds: 7.3 -> 5.5 ms) and costs for short ones (cfg-2's set: 1.44 -> 1.63 ms;
// the KS-A kernels lose occupancy to the scoring state and the per-contig epilogues add up).  The rule looks at the contig
// alone, so a contig's sums do not depend on what else is in the call, the chunk or the shard.
#ifdef BS_CPU_EMUL
constexpr int64_t FUSE_MIN_LEN = 2560;  // (emulation: the tests' contigs are short)
#else
constexpr int64_t FUSE_MIN_LEN = 16384;
#endif

// the part of the fused scoring that sits in the window loop: window pw (valid: all bases ACGT) with table entry e
template <bool FUSE>
__device__ __forceinline__ void fused_score_window(const ProbDistArgs &a, ScoreShared *sh, ScoreState &st, int64_t c, const int32_t *wsc,
                                                   int64_t pw, bool valid, uint32_t code, const WinEntry &e) {
    if constexpr (FUSE) {
        if (st.total != 0) {
            const int32_t wv = wsc[pw + 4];
            if (wv != 0 && valid) score_add(a.sc, *sh, st, c, dense_offset(8) + (int)code, e.prob, e.row, wv);
        }
    }
}
// ... and after it: the irregular positions, then sums / outputs / KS of the break histogram (barriers inside)
template <bool FUSE>
__device__ __forceinline__ void fused_score_tail(const ProbDistArgs &a, ScoreShared *sh, ScoreState &st, int64_t c, int64_t L, const uint64_t *gw,
                                                 const uint32_t *gm, const int32_t *wsc) {
    if constexpr (FUSE) {
        const int tid = threadIdx.x;
        const int64_t np = L > 0 ? L : 1;
        if (st.total != 0 && tid < 7) {
            const int64_t p = tid < 4 ? tid : L - 3 + (tid - 4);
            const bool mine = tid < 4 ? p < np : (p >= 4 && p < np);
            if (mine) {
                const int32_t wv = wsc[p];
                if (wv != 0) score_add_position(a.sc, *sh, st, c, gw, gm, p, L, wv);
            }
        }
        score_finish(a.sc, *sh, st, c, L, gw, gm, wsc);
    }
}


// Rank histogram layout.  PACKED: two 16-bit counters per 32-bit word (every contig of the launch
// has < 65536 windows), else one counter per word.  Words are grouped into RANGES of 32 (one
// thread sweeps one range) with one pad word after every range, which keeps a warp's
// simultaneous sweeps on different banks.  A bitmap word per range marks the non-empty words, so
// a sweep touches only those: its cost follows the contig's windows, not the number of ranks.
template <bool PACKED>
__device__ __forceinline__ int hist_logical_word(int r) { return PACKED ? r >> 1 : r; }
BS_HD int hist_logical_words(int R_x, bool packed) { return packed ? (R_x + 1) / 2 : R_x; }
BS_HD int hist_phys_words(int R_x, bool packed) {
    const int w = hist_logical_words(R_x, packed);
    return w + (w >> 5) + 2;
}
BS_HD int hist_ranges(int R_x, bool packed) { return (hist_logical_words(R_x, packed) + 31) / 32; }

template <bool PACKED, bool IN_SMEM, bool FUSE>
__global__ void __launch_bounds__(768, 2) k_prob_dist_ks(ProbDistArgs a) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    __shared__ int64_t s_wsum[32];
    __shared__ int64_t s_wmax[32];
    __shared__ int s_item;
    ScoreShared *sh = nullptr;
    if constexpr (FUSE) {
        __shared__ ScoreShared s_score;
        sh = &s_score;
        score_shared_init(a.sc, *sh);
    }
    // dynamic shared memory (or the global scratch row): histogram words, then the range bitmaps
    uint32_t *s_hist = IN_SMEM ? (uint32_t *)bs_dyn_smem() : a.rank_scratch + (int64_t)blockIdx.x * (a.hist_words + a.n_ranges);
    uint32_t *s_bm = s_hist + a.hist_words;
    const bool want_ks = a.ks != nullptr;
    const int kshift = 64 - 2 * a.kmer;
    const uint32_t kbits = keep_bits(a.kmer);
    if (want_ks) {
        for (int i = tid; i < a.hist_words + a.n_ranges; i += nthr) s_hist[i] = 0;
    }
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_item = atomicAdd(a.work_counter, 1);
            if constexpr (FUSE) { sh->novf = 0; sh->maxc = 0; sh->nz = 0; }
        }
        __syncthreads();
        if (s_item >= a.n_contigs) break;
        const int64_t c = a.order[s_item];
        const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
        const uint64_t *gw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *gm = a.ctg_mask + a.ctg_woff[c];
        int64_t nwin = L - a.kmer + 1;
        if (nwin < 0) nwin = 0;
        double *pd = a.prob_dist ? a.prob_dist + a.pd_off[c] : nullptr;
        ScoreState st;
        const int32_t *wsc = nullptr;
        const bool fz = FUSE && L >= a.fuse_min_len;  // (block-uniform) a property of the contig alone: see FUSE_MIN_LEN
        if constexpr (FUSE) {
            score_begin(a.sc, *sh, c, st);
            if (!fz) st.total = 0;  // scored by k_break_score: nothing is accumulated here
            wsc = a.sc.w + a.ctg_off[c] + c;
        }
        // truth side of this contig's segment, loaded early so that the latency hides behind the windows
        const int64_t seg = a.ctg_seg[c];
        const LeLt *yx = want_ks ? a.yx + seg * a.R_x : nullptr;
        const int64_t n_y = (want_ks && a.R_y > 0) ? a.ycum[seg * a.R_y + a.R_y - 1] : 0;
        // ---- windows: table value out, rank histogram in.  A thread takes four CONSECUTIVE
        // positions (they never straddle a 32-base word boundary), so the contig words are
        // loaded once per four windows; the four table gathers are in flight together ----
        for (int64_t p0 = 0; p0 < nwin; p0 += 4 * (int64_t)nthr) {
            // a WARP takes 128 consecutive positions, lane l the positions 32 u + l of them (u = 0..3): the contig
            // words of a step are the same for every lane (one broadcast load), the four table gathers of a thread
            // are in flight together and the 8-byte results of a step go out as one coalesced 256-byte store
            const int64_t base = p0 + 128 * (int64_t)warp;
            double val[4];
            int32_t rk[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { val[u] = 0.0; rk[u] = a.rank_zero; }
            if (base < nwin && a.win) {
                const int64_t wi = base >> 5;
                uint64_t w0 = __ldg(&gw[wi]);
                uint32_t m0 = __ldg(&gm[wi]);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (base + 32 * u < nwin) {  // (warp-uniform: the next word exists, strings carry two pad words)
                        const uint64_t w1 = __ldg(&gw[wi + u + 1]);
                        const uint32_t m1 = __ldg(&gm[wi + u + 1]);
                        if (base + 32 * u + lane < nwin) {
                            const bool valid = !(window32(m0, m1, (uint32_t)lane) & kbits);
                            const uint32_t code = (uint32_t)(window64(w0, w1, (uint32_t)lane) >> kshift);
                            WinEntry e;
                            e.prob = 0.0; e.rank = a.rank_zero; e.row = -1;
                            if (valid) e = a.win[code];  // one 16-byte gather
                            val[u] = e.prob;
                            rk[u] = e.rank;
                            fused_score_window<FUSE>(a, sh, st, c, wsc, base + 32 * u + lane, valid, code, e);
                        }
                        w0 = w1;
                        m0 = m1;
                    }
                }
            }
            const int64_t pb = base + lane;  // this thread's positions: pb + 32 u
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (pb + 32 * u < nwin) {
                    if (pd) pd[pb + 32 * u] = val[u];
                    if (want_ks) {
                        const int lw = hist_logical_word<PACKED>(rk[u]);
                        atomicAdd(&s_hist[lw + (lw >> 5)], PACKED ? 1u << (16 * (rk[u] & 1)) : 1u);
                        atomicOr(&s_bm[lw >> 5], 1u << (lw & 31));
                    }
                }
            }
        }
        if (fz) fused_score_tail<FUSE>(a, sh, st, c, L, gw, gm, wsc);
        if (!want_ks) continue;
        __syncthreads();
        // ---- D = sup |F_x - F_y| over the pooled distinct values, evaluated at every x value that
        // is present: just below it (F_x of the previous step vs #{y < v}) and at it ----
        const bool defined = nwin > 0 && n_y > 0;
        // |F_x - F_y| = |run * n_y - ycount * nwin| / (nwin * n_y): the numerator is maximised in exact
        // 64-bit integers (both factors are below 2^31), one division at the end
        int64_t best = 0;
        uint32_t carry = 0;  // windows in the ranges of earlier rounds (same value on every thread)
        for (int r0 = 0; r0 < a.n_ranges; r0 += nthr) {  // one range of 32 words per thread and round
            const int r = r0 + tid;
            const uint32_t m = r < a.n_ranges ? s_bm[r] : 0u;
            if (m) s_bm[r] = 0;
            const int w0 = 33 * r;  // physical index of the range's first word
            uint32_t cnt_r = 0;
            for (uint32_t mm = m; mm; mm &= mm - 1) {
                const uint32_t w = s_hist[w0 + __ffs((int)mm) - 1];
                cnt_r += PACKED ? (w & 0xffffu) + (w >> 16) : w;
            }
            // exclusive prefix of the range counts over the block (+ what earlier rounds held)
            uint32_t incl = cnt_r;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL_MASK, incl, dd);
                if (lane >= dd) incl += o;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            uint32_t run = carry + incl - cnt_r;
            for (int w = 0; w < nwarp; w++) {
                if (w < warp) run += s_wsum[w];
                carry += s_wsum[w];
            }
            for (uint32_t mm = m; mm; mm &= mm - 1) {
                const int k = __ffs((int)mm) - 1;
                const uint32_t w = s_hist[w0 + k];
                s_hist[w0 + k] = 0;  // leave the histogram zeroed for the next contig
#pragma unroll
                for (int h = 0; h < (PACKED ? 2 : 1); h++) {
                    const uint32_t cnt = PACKED ? (w >> (16 * h)) & 0xffffu : w;
                    if (cnt == 0 || !defined) continue;
                    const int i = PACKED ? 64 * r + 2 * k + h : 32 * r + k;
                    const int64_t d = ks_numerator(run, run + cnt, yx[i], (uint32_t)n_y, (uint32_t)nwin);
                    run += cnt;
                    if (d > best) best = d;
                }
            }
            __syncthreads();  // s_wsum is rewritten by the next round
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

// ------------------------------------------------------------------------------------------
// The same for SHORT contigs (at most KS_SMALL_MAX windows: three in four of a velvet-style set).
// Instead of a rank histogram as wide as the table (66 KB, swept per contig) the block sorts the
// contig's window ranks by a counting pass over ranges of 64 ranks: per-range counts -> offsets
// (which are already the number of windows in smaller ranges) -> scatter; every window then counts
// the handful of windows that share its range to get its own position.  20 KB of shared memory for
// the real table, so eight blocks of 256 threads are resident per SM and short contigs stop paying
// for the width of the table.
// ------------------------------------------------------------------------------------------

#ifndef BS_KS_SMALL_MAX
#define BS_KS_SMALL_MAX 2048
#endif
#ifndef BS_KS_SMALL_THREADS
#define BS_KS_SMALL_THREADS 256
#endif
constexpr int KS_SMALL_MAX = BS_KS_SMALL_MAX;  // windows per contig the short-contig kernel takes
constexpr int KS_SMALL_THREADS = BS_KS_SMALL_THREADS;
static_assert(FUSE_MIN_LEN > KS_SMALL_MAX + MAXK, "a contig scored on the way must be one the long-contig kernel takes");
constexpr int KS_SMALL_DIRECT = 24;  // windows sharing a range of 64 ranks up to which every window counts its own neighbours

BS_HD int ks_small_ranges(int R_x) { return (R_x + 63) >> 6; }
BS_HD size_t ks_small_smem_bytes(int R_x) { return (size_t)KS_SMALL_MAX * 8 + (size_t)(2 * ks_small_ranges(R_x) + 4) * 4; }

__global__ void __launch_bounds__(KS_SMALL_THREADS, 6) k_prob_dist_ks_small(ProbDistArgs a) {
    constexpr bool FUSE = false;  // short contigs are scored by k_break_score (see FUSE_MIN_LEN)
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    __shared__ int64_t s_wmax[32];
    __shared__ uint32_t s_wsum[32];
    __shared__ int s_item;
    ScoreShared *sh = nullptr;
    if constexpr (FUSE) {
        __shared__ ScoreShared s_score;
        sh = &s_score;
        score_shared_init(a.sc, *sh);
    }
    uint32_t *s_rank = (uint32_t *)bs_dyn_smem();   // [KS_SMALL_MAX] rank of the window at p
    uint32_t *s_keys = s_rank + KS_SMALL_MAX;       // [KS_SMALL_MAX] the same ranks grouped by range
    const int n_rng = ks_small_ranges(a.R_x);
    uint32_t *s_off = s_keys + KS_SMALL_MAX;        // [n_rng + 1] first slot of a range = windows in smaller ranges
    uint32_t *s_cur = s_off + n_rng + 2;            // [n_rng] counts, then scatter cursors
    const int kshift = 64 - 2 * a.kmer;
    const uint32_t kbits = keep_bits(a.kmer);
    const int per = (n_rng + nthr - 1) / nthr;      // ranges per thread in the prefix sum
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_item = atomicAdd(a.work_counter, 1);
            if constexpr (FUSE) { sh->novf = 0; sh->maxc = 0; sh->nz = 0; }
        }
        for (int i = tid; i < n_rng; i += nthr) s_cur[i] = 0;
        __syncthreads();
        if (s_item >= a.n_contigs) break;
        const int64_t c = a.order[s_item];
        const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
        const uint64_t *gw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *gm = a.ctg_mask + a.ctg_woff[c];
        int nwin = (int)(L - a.kmer + 1);
        if (nwin < 0) nwin = 0;
        double *pd = a.prob_dist ? a.prob_dist + a.pd_off[c] : nullptr;
        ScoreState st;
        const int32_t *wsc = nullptr;
        if constexpr (FUSE) {
            score_begin(a.sc, *sh, c, st);
            wsc = a.sc.w + a.ctg_off[c] + c;
        }
        const int64_t seg = a.ctg_seg[c];
        const LeLt *yx = a.yx + seg * a.R_x;
        const int64_t n_y = a.R_y > 0 ? a.ycum[seg * a.R_y + a.R_y - 1] : 0;
        // ---- windows: table value out, rank kept, per-range counts ----
        for (int p0 = 0; p0 < nwin; p0 += 4 * nthr) {
            // (same mapping as k_prob_dist_ks: a warp takes 128 consecutive positions, lane l the positions 32 u + l)
            const int base = p0 + 128 * warp;
            double val[4];
            int32_t rk[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { val[u] = 0.0; rk[u] = a.rank_zero; }
            if (base < nwin && a.win) {
                const int wi = base >> 5;
                uint64_t w0 = __ldg(&gw[wi]);
                uint32_t m0 = __ldg(&gm[wi]);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (base + 32 * u < nwin) {
                        const uint64_t w1 = __ldg(&gw[wi + u + 1]);
                        const uint32_t m1 = __ldg(&gm[wi + u + 1]);
                        if (base + 32 * u + lane < nwin) {
                            const bool valid = !(window32(m0, m1, (uint32_t)lane) & kbits);
                            const uint32_t code = (uint32_t)(window64(w0, w1, (uint32_t)lane) >> kshift);
                            WinEntry e;
                            e.prob = 0.0; e.rank = a.rank_zero; e.row = -1;
                            if (valid) e = a.win[code];
                            val[u] = e.prob;
                            rk[u] = e.rank;
                            fused_score_window<FUSE>(a, sh, st, c, wsc, (int64_t)base + 32 * u + lane, valid, code, e);
                        }
                        w0 = w1;
                        m0 = m1;
                    }
                }
            }
            const int pb = base + lane;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (pb + 32 * u < nwin) {
                    if (pd) pd[pb + 32 * u] = val[u];
                    s_rank[pb + 32 * u] = (uint32_t)rk[u];
                    atomicAdd(&s_cur[rk[u] >> 6], 1u);
                }
            }
        }
        fused_score_tail<FUSE>(a, sh, st, c, L, gw, gm, wsc);
        __syncthreads();
        // ---- exclusive prefix of the range counts ----
        const int lo_r = tid * per < n_rng ? tid * per : n_rng, hi_r = lo_r + per < n_rng ? lo_r + per : n_rng;
        uint32_t sum = 0;
        for (int i = lo_r; i < hi_r; i++) sum += s_cur[i];
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t run0 = incl - sum;
        for (int w = 0; w < warp; w++) run0 += s_wsum[w];
        for (int i = lo_r; i < hi_r; i++) {
            const uint32_t cnt = s_cur[i];
            s_off[i] = run0;
            s_cur[i] = run0;  // scatter cursor
            run0 += cnt;
        }
        if (tid == 0) s_off[n_rng] = (uint32_t)nwin;
        __syncthreads();
        // ---- scatter the ranks into their ranges ----
        for (int p = tid; p < nwin; p += nthr) {
            const uint32_t r = s_rank[p];
            s_keys[atomicAdd(&s_cur[r >> 6], 1u)] = r;
        }
        __syncthreads();
        // ---- statistic just below and at every x value that is present ----
        int64_t best = 0;
        const bool defined = nwin > 0 && n_y > 0;
        if (defined) {
            const uint32_t ny = (uint32_t)n_y, nw = (uint32_t)nwin;
            // a thread per window: windows of smaller / not larger rank = the range's first slot + those
            // among the few windows sharing its range
            for (int i = tid; i < nwin; i += nthr) {
                const uint32_t key = s_keys[i];
                const uint32_t g0 = s_off[key >> 6], g1 = s_off[(key >> 6) + 1];
                if (g1 - g0 > (uint32_t)KS_SMALL_DIRECT) continue;
                uint32_t n_lt = g0, n_le = g0;
                for (uint32_t j = g0; j < g1; j++) {
                    const uint32_t k = s_keys[j];
                    n_lt += k < key;
                    n_le += k <= key;
                }
                const int64_t d = ks_numerator(n_lt, n_le, yx[key], ny, nw);
                if (d > best) best = d;
            }
            // crowded ranges (low-complexity contigs): a thread per range walks its distinct ranks in
            // order through a 64-bit presence mask, so the work stays linear in the windows
            for (int r = tid; r < n_rng; r += nthr) {
                const uint32_t g0 = s_off[r], g1 = s_off[r + 1];
                if (g1 - g0 <= (uint32_t)KS_SMALL_DIRECT) continue;
                uint64_t present = 0;
                for (uint32_t i = g0; i < g1; i++) present |= 1ull << (s_keys[i] & 63u);
                uint32_t run = g0;  // windows with a smaller rank
                while (present) {
                    const int bit = __ffsll((long long)present) - 1;
                    present &= present - 1;
                    const uint32_t rank = ((uint32_t)r << 6) + (uint32_t)bit;
                    uint32_t cnt = 0;
                    for (uint32_t i = g0; i < g1; i++) cnt += s_keys[i] == rank;
                    const int64_t d = ks_numerator(run, run + cnt, yx[rank], ny, nw);
                    run += cnt;
                    if (d > best) best = d;
                }
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
