// Break k-mers -> probability-weighted sums, optional histogram, KS of the normalised break histogram.
#pragma once
#include "bs_seq.cuh"

namespace bs {

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
// break k-mers -> probability-weighted sums (upstream lib/BreakageScorer.cpp:244-303), the
// optional dense histogram, and the KS statistic of the normalised break histogram (variant
// lib/DeNovoAssembler.cpp:395-420: x = count_row / total for every table row, y = truth
// distribution) -- one pass family over the position weights of a contig
// ------------------------------------------------------------------------------------------

// scoring table over the dense k-mer index space (all lengths 1..8), one 16-byte gather per break
struct alignas(16) TabEntry {
    double prob;
    int32_t row;
    int32_t pad;
};

#ifndef BS_SCORE_THREADS
#define BS_SCORE_THREADS 256
#endif
#ifndef BS_SCORE_BLOCKS
#define BS_SCORE_BLOCKS 6
#endif
#ifndef BS_SCORE_CC_DENSE
#define BS_SCORE_CC_DENSE 1024
#endif
#ifndef BS_SCORE_HASH_SLOTS
#define BS_SCORE_HASH_SLOTS 4096
#endif
constexpr int SCORE_THREADS = BS_SCORE_THREADS, SCORE_BLOCKS = BS_SCORE_BLOCKS;  // block shape k_break_score is built for
constexpr int CC_DENSE = BS_SCORE_CC_DENSE;  // counts below this are tallied in a dense shared-memory array
constexpr int OVF_CAP = 4096;   // per-block capacity for larger counts
// KS-B needs, per contig, the number of table rows with every count value.  Contigs that placed at most
// HASH_LIMIT reads (so at most that many distinct break k-mers) count their rows in a shared-memory hash table:
// one 32-bit slot = (dense k-mer index + 1) << 15 | count.  Larger contigs use the block's row of the global
// scratch (T+1 counters): rare, so those rows stay out of DRAM (every contig going through them cost 315 MB of
// write-backs per 1000 segments, profiles/r02a).
#ifdef BS_CPU_EMUL
constexpr int HASH_SLOTS = 128;  // (emulation: small, so that the tests run both forms)
#else
constexpr int HASH_SLOTS = BS_SCORE_HASH_SLOTS;
#endif
constexpr int HASH_LIMIT = HASH_SLOTS * 2 / 3;
static_assert(HASH_LIMIT < (1 << 15) && DENSE_SIZE + 1 < (1 << 17), "slot layout: 17-bit key, 15-bit count");

struct ScoreArgs {
    const int32_t *order;   // [C] contig ids, longest first
    int32_t *work_counter;  // zero on entry
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const int32_t *ctg_seg;
    const int32_t *w;
    const int32_t *total;
    const TabEntry *tab;     // [DENSE_SIZE] probability and table row of a dense k-mer index (row -1 if absent)
    int32_t kmer;
    int32_t T;
    int64_t n_contigs;
    int32_t *sequence_len;
    double *bp_score;
    double *norm_by_break_freqs;
    double *norm_by_len;
    int32_t *kmer_breaks;
    int32_t *hist;  // optional [C][T+1]
    // KS of the normalised break histogram (all NULL / 0 when not wanted)
    double *ks_b;          // [C]
    const double *yv;      // [R_y] sorted distinct truth-table values
    const int32_t *ycum;   // [S][R_y] inclusive cumulative counts of the truth distribution
    int32_t R_y;
    int32_t zero_le, zero_lt;  // indices into ycum of the last y value <= 0.0 / < 0.0 (-1: none)
    double y_max;          // largest truth-table value
    int32_t *scratch;      // [gridDim][T+1], all zero on entry and on exit
    int32_t *ovf_cnt;      // [gridDim][OVF_CAP] counts >= CC_DENSE (unsorted)
    int32_t *status;       // set to 1 if a block ran out of overflow space
};

// number of y values <= v (le) and < v (lt), from the cumulative counts
__device__ __forceinline__ void y_counts_at(const double *yv, const int32_t *ycum, int R_y, double y_max, double v,
                                            double *le, double *lt) {
    if (R_y <= 0) { *le = 0.0; *lt = 0.0; return; }
    if (v > y_max) { *le = *lt = (double)ycum[R_y - 1]; return; }  // the usual case: count/total >> any table value
    int lo = 0, hi = R_y;  // first index with yv > v
    while (lo < hi) { int mid = (lo + hi) >> 1; if (yv[mid] <= v) lo = mid + 1; else hi = mid; }
    const int ule = lo;  // #distinct <= v
    lo = 0; hi = R_y;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (yv[mid] < v) lo = mid + 1; else hi = mid; }
    const int ult = lo;  // #distinct < v
    *le = ule > 0 ? (double)ycum[ule - 1] : 0.0;
    *lt = ult > 0 ? (double)ycum[ult - 1] : 0.0;
}

// sum over the block in a fixed order (lane tree, then warps in order): bit-reproducible, the
// same for every GPU count; valid on thread 0
__device__ __forceinline__ double block_sum_fixed(double v, double *s_w) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(FULL_MASK, v, m);
    __syncthreads();
    if (lane == 0) s_w[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) for (int i = 0; i < nwarp; i++) t += s_w[i];
    return t;
}

// ---- the per-contig scoring state, shared by k_break_score and by the KS-A kernels when they score on the way
// (bs_ks.cuh: with kmer == 8 the break k-mer of position p IS the rolling window p - 4, already gathered there) ----

struct ScoreSharedCore {
    double w[32];
    int32_t cc[CC_DENSE];        // rows having count j
    int novf, maxc, nz;          // reset by thread 0 where the block fetches its next contig
    int fallback;                // INCR (below): a count reached CC_DENSE, the rows are tallied by the sweep after all
};
// INCR instantiations of score_add / score_finish (bs_compose.cuh: every call adds ONE read) keep the number of rows per
// count value up to date while the hash table fills -- a new key is a row of count 1, an increment moves a row from its
// old count to the next -- so that no sweep over the table is needed afterwards, only its reset.  A thread's share of the
// rows of count 1 and of all rows stays in registers until score_finish.
struct IncrTally {
    int nz, c1, mx;
};
struct ScoreShared : ScoreSharedCore {
    uint32_t hash[HASH_SLOTS];   // (dense index + 1) << 15 | count, 0 = empty; all empty between contigs
};

struct ScoreState {
    int32_t total;
    bool in_smem;     // (block-uniform) distinct break k-mers <= placed reads <= two thirds of the hash table (and < 2^15)
    uint32_t *hash;   // the block's shared-memory hash table: ScoreShared::hash, or a larger one in dynamic shared memory (bs_compose.cuh)
    int hs;           // slots of the table this contig uses: a power of two >= 2 * total (pass 2 sweeps what a small contig
    uint32_t hmask;   // can have touched, not all HASH_SLOTS), all HASH_SLOTS for the largest contigs that still fit
    double s1, s2;
    double inv_total;  // 1.0 / total: what (double)wv / (double)total is for wv == 1, the usual weight (one division per contig, not per break)
    int32_t *scratch, *ovf;
};

__device__ __forceinline__ void score_shared_init(const ScoreArgs &a, ScoreSharedCore &sh, uint32_t *hash, int slots) {  // once per kernel, before the first barrier
    if (a.ks_b == nullptr) return;
    for (int i = threadIdx.x; i < CC_DENSE; i += blockDim.x) sh.cc[i] = 0;
    for (int i = threadIdx.x; i < slots; i += blockDim.x) hash[i] = 0u;
}
__device__ __forceinline__ void score_shared_init(const ScoreArgs &a, ScoreShared &sh) { score_shared_init(a, sh, sh.hash, HASH_SLOTS); }

// (total: reads placed in the contig; given by the caller where it has just been counted, bs_compose.cuh)
// hash / slots: the shared-memory table (slots a power of two)
__device__ __forceinline__ void score_begin_total(const ScoreArgs &a, int32_t total, ScoreState &st, uint32_t *hash, int slots) {
    const bool want_ks = a.ks_b != nullptr;
    st.total = total;
    st.hash = hash;
    st.in_smem = st.total <= slots * 2 / 3 && st.total < (1 << 15);
    st.hs = 64 < slots ? 64 : slots;
    while (st.hs < 2 * st.total && st.hs < slots) st.hs <<= 1;
    st.hmask = (uint32_t)st.hs - 1u;
    st.s1 = st.s2 = 0.0;
    st.inv_total = 1.0 / (double)total;
    st.scratch = want_ks ? a.scratch + (int64_t)blockIdx.x * (a.T + 1) : nullptr;
    st.ovf = want_ks ? a.ovf_cnt + (int64_t)blockIdx.x * OVF_CAP : nullptr;
}
__device__ __forceinline__ void score_begin(const ScoreArgs &a, ScoreShared &sh, int64_t c, ScoreState &st) { score_begin_total(a, a.total[c], st, sh.hash, HASH_SLOTS); }

// wv reads broke the k-mer with dense index di (table entry te) in contig c
template <bool INCR = false>
__device__ __forceinline__ void score_add(const ScoreArgs &a, ScoreSharedCore &sh, ScoreState &st, int64_t c, int di, double prob, int32_t row,
                                          int32_t wv, IncrTally *t = nullptr) {
    if (row >= 0) {
        st.s1 += prob * (double)wv;
        st.s2 += prob * (wv == 1 ? st.inv_total : (double)wv / (double)st.total);
        if (a.ks_b != nullptr) {
            if (st.in_smem) {
                // distinct dense indices are distinct table rows: count per k-mer.  At most HASH_LIMIT
                // of the HASH_SLOTS slots are ever taken, so the probe ends.
                const uint32_t key = (uint32_t)di + 1u;
                for (uint32_t h = (key * 2654435761u) >> 16;; h++) {
                    h &= st.hmask;
                    uint32_t cur = st.hash[h];
                    if (cur == 0u) {
                        cur = atomicCAS(&st.hash[h], 0u, (key << 15) | (uint32_t)wv);
                        if (cur == 0u) {
                            if constexpr (INCR) { t->nz++; t->c1++; if (t->mx < 1) t->mx = 1; }  // (wv == 1)
                            break;
                        }
                    }
                    if ((cur >> 15) == key) {
                        const uint32_t old = atomicAdd(&st.hash[h], (uint32_t)wv);
                        if constexpr (INCR) {  // the row moves from count `was` to was + 1
                            const int was = (int)(old & 0x7fffu);
                            if (was + 1 >= CC_DENSE) sh.fallback = 1;
                            else {
                                if (was == 1) t->c1--; else atomicAdd(&sh.cc[was], -1);
                                atomicAdd(&sh.cc[was + 1], 1);
                                if (t->mx < was + 1) t->mx = was + 1;
                            }
                        }
                        break;
                    }
                }
            } else {
                atomicAdd(&st.scratch[row], wv);
            }
        }
    }
    if (a.hist) atomicAdd(&a.hist[c * (int64_t)(a.T + 1) + (row >= 0 ? row : a.T)], wv);
}

// the generic form: position p of a contig of length L (any kmer, any position)
__device__ __forceinline__ void score_add_position(const ScoreArgs &a, ScoreSharedCore &sh, ScoreState &st, int64_t c, const uint64_t *gw,
                                                   const uint32_t *gm, int64_t p, int64_t L, int32_t wv) {
    const BreakWindow bw = break_window(p, a.kmer, L);
    const int di = dense_index_at(gw, gm, bw.start, bw.len);
    TabEntry te;
    te.prob = 0.0;
    te.row = -1;
    if (di >= 0) te = a.tab[di];
    score_add(a, sh, st, c, di, te.prob, te.row, wv);
}

// sums, outputs and the KS of the break histogram; every thread of the block calls it (barriers inside).
// The broken positions of a contig that counted its rows in the global scratch are found again through the position
// weights w, or -- rows != NULL, the compositional path, which has no weights -- through the block's row of leftmost
// positions per read (POS_INF_ROW: read not placed).
constexpr uint32_t POS_INF_ROW = 0x7f7f7f7fu;
template <bool INCR = false>
__device__ __forceinline__ void score_finish(const ScoreArgs &a, ScoreSharedCore &sh, ScoreState &st, int64_t c, int64_t L, const uint64_t *gw,
                                             const uint32_t *gm, const int32_t *w, const uint32_t *rows = nullptr, int64_t n_rows = 0,
                                             IncrTally *t = nullptr) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const bool want_ks = a.ks_b != nullptr;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    const int32_t total = st.total;
    const int64_t np = L > 0 ? L : 1;
    const double s1 = block_sum_fixed(st.s1, sh.w);
    const double s2 = block_sum_fixed(st.s2, sh.w);
    if (tid == 0) {
        if (a.sequence_len) a.sequence_len[c] = (int32_t)L;
        if (a.bp_score) a.bp_score[c] = s1;
        if (a.norm_by_break_freqs) a.norm_by_break_freqs[c] = s2;
        if (a.norm_by_len) a.norm_by_len[c] = s1 / (double)(int32_t)L;
        if (a.kmer_breaks) a.kmer_breaks[c] = total;
    }
    if (!want_ks) return;
    if (total == 0) {  // 0/0 for every row: R drops the NaNs and ks.test stops on empty x
        if (tid == 0) a.ks_b[c] = qnan;
        return;
    }
    __threadfence_block();
    __syncthreads();
    // pass 2: tally rows per count value
    // (a thread keeps its own tallies of the rows it saw, of those with count 1 -- nearly all -- and of the largest
    // count; they reach the shared counters once, behind the sweep: no same-address atomic per row)
    int t_nz = 0, t_c1 = 0, t_max = 0;
    auto tally = [&](int32_t cnt) {
        if (cnt == 0) return;
        t_nz++;
        if (cnt == 1) { t_c1++; if (t_max < 1) t_max = 1; }
        else if (cnt < CC_DENSE) { atomicAdd(&sh.cc[cnt], 1); if (t_max < cnt) t_max = cnt; }
        else {
            const int slot = atomicAdd(&sh.novf, 1);
            if (slot < OVF_CAP) st.ovf[slot] = cnt; else *a.status = 1;
        }
    };
    bool swept = false;
    if constexpr (INCR) {
        if (st.in_smem) {
            if (!sh.fallback) {  // (block-uniform: read behind the barrier) the rows are tallied already: reset the table
                for (int i = tid; i < st.hs; i += nthr) st.hash[i] = 0u;
                t_nz = t->nz; t_c1 = t->c1; t_max = t->mx;
                swept = true;
            } else {  // a count reached CC_DENSE: forget the running tallies, sweep as usual
                for (int i = tid; i < CC_DENSE; i += nthr) sh.cc[i] = 0;
                __syncthreads();
                if (tid == 0) { sh.nz = 0; sh.maxc = 0; }
                __syncthreads();
            }
        }
    }
    if (swept) {
    } else if (st.in_smem) {  // the usual case: the hash table's slots, emptied on the way
        for (int i = tid; i < st.hs; i += nthr) {
            const uint32_t v = st.hash[i];
            if (v != 0u) { st.hash[i] = 0u; tally((int32_t)(v & 0x7fffu)); }
        }
    } else if (rows != nullptr) {
        for (int64_t r = tid; r < n_rows; r += nthr) {
            const uint32_t p = rows[r];
            if (p == POS_INF_ROW) continue;
            const BreakWindow bw = break_window((int64_t)p, a.kmer, L);
            const int di = dense_index_at(gw, gm, bw.start, bw.len);
            const int32_t trow = di >= 0 ? a.tab[di].row : -1;
            tally(trow >= 0 ? atomicExch(&st.scratch[trow], 0) : 0);
        }
    } else
    // whoever swaps a row's count out of the scratch first owns it
    for (int64_t p0 = 0; p0 < np; p0 += 4 * (int64_t)nthr) {
        int32_t wv[4], row[4], cnt[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t p = p0 + (int64_t)u * nthr + tid;
            wv[u] = p < np ? w[p] : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            row[u] = -1;
            if (wv[u] != 0) {
                const BreakWindow bw = break_window(p0 + (int64_t)u * nthr + tid, a.kmer, L);
                const int di = dense_index_at(gw, gm, bw.start, bw.len);
                if (di >= 0) row[u] = a.tab[di].row;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) cnt[u] = row[u] >= 0 ? atomicExch(&st.scratch[row[u]], 0) : 0;
#pragma unroll
        for (int u = 0; u < 4; u++) tally(cnt[u]);
    }
    if (t_nz) atomicAdd(&sh.nz, t_nz);
    if (t_c1) atomicAdd(&sh.cc[1], t_c1);  // (INCR: a thread's share can be negative)
    if (t_max) atomicMax(&sh.maxc, t_max);
    __syncthreads();
    // the distinct x values are 0 and count/total for the few distinct counts: thread 0 walks them
    if (tid == 0) {
        const int32_t *ycum = a.ycum + (int64_t)a.ctg_seg[c] * a.R_y;
        const int64_t n_y = a.R_y > 0 ? ycum[a.R_y - 1] : 0;
        const int novf = sh.novf < OVF_CAP ? sh.novf : OVF_CAP;
        const int32_t *ovf = st.ovf;
        double d = 0.0;
        if (n_y > 0 && a.T > 0) {
            const double inx = (double)a.T, iny = (double)n_y;
            int64_t run = (int64_t)a.T - sh.nz;  // rows never broken: x value 0
            if (run > 0) {
                const double le = a.zero_le >= 0 ? (double)ycum[a.zero_le] : 0.0;
                const double lt = a.zero_lt >= 0 ? (double)ycum[a.zero_lt] : 0.0;
                double d1 = lt / iny, d2 = (double)run / inx - le / iny;
                if (d1 < 0) d1 = -d1;
                if (d2 < 0) d2 = -d2;
                d = d1 > d2 ? d1 : d2;
            }
            for (int j = 1; j <= sh.maxc; j++) {
                const int32_t cnt = sh.cc[j];
                if (cnt == 0) continue;
                sh.cc[j] = 0;
                double le, lt;
                y_counts_at(a.yv, ycum, a.R_y, a.y_max, (double)j / (double)total, &le, &lt);
                double d1 = (double)run / inx - lt / iny;
                run += cnt;
                double d2 = (double)run / inx - le / iny;
                if (d1 < 0) d1 = -d1;
                if (d2 < 0) d2 = -d2;
                if (d1 > d) d = d1;
                if (d2 > d) d = d2;
            }
            // counts >= CC_DENSE: few; walked in ascending order
            int32_t last = CC_DENSE - 1;
            for (int done = 0; done < novf;) {
                int32_t cur = 0x7fffffff;
                int mult = 0;
                for (int i = 0; i < novf; i++) {
                    const int32_t v = ovf[i];
                    if (v > last && v < cur) { cur = v; mult = 1; }
                    else if (v == cur) mult++;
                }
                double le, lt;
                y_counts_at(a.yv, ycum, a.R_y, a.y_max, (double)cur / (double)total, &le, &lt);
                double d1 = (double)run / inx - lt / iny;
                run += mult;
                double d2 = (double)run / inx - le / iny;
                if (d1 < 0) d1 = -d1;
                if (d2 < 0) d2 = -d2;
                if (d1 > d) d = d1;
                if (d2 > d) d = d2;
                last = cur;
                done += mult;
            }
        } else {
            for (int j = 1; j <= sh.maxc; j++) sh.cc[j] = 0;
        }
        a.ks_b[c] = (n_y > 0 && a.T > 0) ? d : qnan;
    }
}

__global__ void __launch_bounds__(SCORE_THREADS, SCORE_BLOCKS) k_break_score(ScoreArgs a) {
    __shared__ ScoreShared sh;
    __shared__ int s_item;
    const int tid = threadIdx.x, nthr = blockDim.x;
    score_shared_init(a, sh);
    for (;;) {
        __syncthreads();
        if (tid == 0) { s_item = atomicAdd(a.work_counter, 1); sh.novf = 0; sh.maxc = 0; sh.nz = 0; }
        __syncthreads();
        if (s_item >= a.n_contigs) break;
        const int64_t c = a.order[s_item];
        const int64_t coff = a.ctg_off[c];
        const int64_t L = a.ctg_off[c + 1] - coff;
        const uint64_t *gw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *gm = a.ctg_mask + a.ctg_woff[c];
        const int32_t *w = a.w + coff + c;
        const int64_t np = L > 0 ? L : 1;
        ScoreState st;
        score_begin(a, sh, c, st);
        // pass 1: weighted sums in position order (+ histogram, + per-row counts for the KS); four
        // positions per thread in flight so that the table gathers overlap
        if (st.total != 0) {
            for (int64_t p0 = 0; p0 < np; p0 += 4 * (int64_t)nthr) {
                int32_t wv[4], di[4];
                TabEntry te[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int64_t p = p0 + (int64_t)u * nthr + tid;
                    wv[u] = p < np ? w[p] : 0;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    te[u].prob = 0.0;
                    te[u].row = -1;
                    di[u] = -1;
                    if (wv[u] != 0) {
                        const BreakWindow bw = break_window(p0 + (int64_t)u * nthr + tid, a.kmer, L);
                        di[u] = dense_index_at(gw, gm, bw.start, bw.len);
                        if (di[u] >= 0) te[u] = a.tab[di[u]];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (wv[u] != 0) score_add(a, sh, st, c, di[u], te[u].prob, te[u].row, wv[u]);
            }
        }
        score_finish(a, sh, st, c, L, gw, gm, w);
    }
}

}  // namespace bs
