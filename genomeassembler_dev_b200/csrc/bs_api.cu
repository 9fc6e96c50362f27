// C-ABI of the breakage scorer (include/breakscore.h): context, resident tables, work buffers
// and the orchestration of the kernels in bs_kernels.cuh.  Host C++ only talks to the CUDA
// runtime; nothing here depends on torch, Python or R.
//
// A scoring call is cut into CHUNKS of whole segments (independent experiments).  Chunks run
// through a three-stream pipeline -- H2D of chunk k+1, kernels of chunk k and D2H of chunk k-1
// overlap -- over two workspaces, so an end-to-end call from host buffers is bound by the PCIe
// copy of the ASCII reads, not by the sum of copy and compute.
#ifdef BS_CPU_EMUL
#include "cuda_emul.h"  // tests/emul: CPU emulation used by the CPU-only test-suite, never shipped
#endif
#include "bs_kernels.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <new>
#include <numeric>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/breakscore.h"

// NVTX ranges around the pipeline stages (header-only NVTX3: a no-op unless a tool such as nsys / ncu is attached)
#ifndef BS_CPU_EMUL
#include <nvtx3/nvToolsExt.h>
#define BS_NVTX_PUSH(name) nvtxRangePushA(name)
#define BS_NVTX_POP() nvtxRangePop()
#else
#define BS_NVTX_PUSH(name) ((void)0)
#define BS_NVTX_POP() ((void)0)
#endif

namespace {

char g_create_error[512] = "";

#ifdef BS_CPU_EMUL
constexpr int kPlaceThreads = 64;
constexpr int kScoreThreads = 64;
constexpr int kBreakScoreThreads = 64;
constexpr int kStartposThreads = 64;
constexpr int kStartposBigThreads = 64;
constexpr int kLevThreads = 64;
constexpr int kSimThreads = 64;
constexpr int kKsThreads = 64;
constexpr int kSpectrumThreads = 64;
constexpr int kSpectrumTabThreads = 64;
constexpr int kPackThreads = 64;
constexpr int kPlaceIxThreads = 64;
constexpr int kComposeThreads = 64;
constexpr int kHitCap = 16;           // (emulation: small, so that the tests reach the global continuation of the hit list)
#else
constexpr int kPlaceThreads = 256;
constexpr int kScoreThreads = 256;
constexpr int kBreakScoreThreads = bs::SCORE_THREADS;
constexpr int kStartposThreads = 256;
constexpr int kStartposBigThreads = 1024;  // one block per SM holds the 128 KB prefix bitmap
constexpr int kLevThreads = 128;
constexpr int kSimThreads = 256;
constexpr int kKsThreads = 768;     // one sweep round covers the 515 ranges of the real table's rank histogram
constexpr int kSpectrumThreads = 512;
constexpr int kSpectrumTabThreads = bs::SPECTRUM_TAB_THREADS;
constexpr int kPackThreads = 256;
constexpr int kPlaceIxThreads = bs::PLACE_IX_THREADS;
constexpr int kComposeThreads = bs::COMPOSE_THREADS;
constexpr int kHitCap = BS_PLACE_HIT_CAP;  // reads placed per contig kept in shared memory (k_place_index); more go to global memory
#endif
constexpr int kMaxTile = 8192;        // contig positions per shared-memory tile (tile placement modes)
constexpr int64_t kMaxChunk = 32768;  // reads per placement work item (tile placement modes)
constexpr int64_t kMinChunk = 2048;
constexpr int kWorkspaces = 3;  // chunks in flight: one copying in, one computing, one copying out

enum Stage { ST_H2D, ST_PACK, ST_PLACE, ST_SCORE, ST_SPECTRUM, ST_PROBDIST, ST_PATHFREQ, ST_STARTPOS, ST_D2H, ST_LEV, ST_COUNT };

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct KsCache {
    int kmer = -1;
    uint64_t table_version = 0;
    int R_x = 0, R_y = 0, rank_zero = 0;
    DevBuf win, rank_y, lelt, yv;
    double y_max = 0.0;
    int zero_le = -1, zero_lt = -1;  // truth cumulative-count indices of the last y value <= 0 / < 0
};

// device + pinned staging memory of one in-flight chunk
struct Workspace {
    void *h_meta = nullptr;
    size_t h_meta_cap = 0;
    DevBuf meta, read_chars, read_off, ctg_chars, tr_chars;
    DevBuf rwords, rflags, cwords, cmask, twords, tmask;
    DevBuf w, total, ycnt, yx, yx2, head, next, odd_head, spbest, exact, sp_key, sp_head, sp_next, sp_bitmap, sp_queue;
    DevBuf out_i32, out_f64, pd, pd2, hist, pos;
    DevBuf base_chars, base_words, base_mask, base_pos, base_w, base_total, base_hits, base_cnt, base_di, base_rank;  // compositional scoring: the base contigs of a scaffold set
    cudaEvent_t ev_h2d = nullptr, ev_compute = nullptr, ev_d2h = nullptr;
    bool in_flight = false;
};

struct TimedSpan {
    int stage;
    cudaEvent_t a, b;
};

// a run of whole segments processed together
struct Chunk {
    int64_t s0, s1;  // segments
    int64_t r0, r1;  // reads
    int64_t c0, c1;  // contigs
};

}  // namespace

struct bs_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t own_stream = nullptr;   // compute stream unless the caller sets one
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t copy_stream = nullptr;  // H2D
    cudaStream_t out_stream = nullptr;   // D2H
    char err[512] = "";
    int64_t launches = 0;
    int64_t chunk_bytes_host = (int64_t)96 << 20;    // ASCII bytes per chunk when inputs come from the host
    int64_t chunk_bytes_dev = (int64_t)4096 << 20;   // ... when they are already on the device

    // table
    bool has_table = false;
    uint64_t table_version = 0;
    int64_t T = 0;
    std::vector<double> prob_dense;    // [DENSE_SIZE]
    std::vector<double> tprob_dense;   // [DENSE_SIZE] truth-side values (follows prob_dense unless set)
    std::vector<int32_t> row_dense;    // [DENSE_SIZE]
    bool has_truth_table = false;
    DevBuf d_tab;  // bs::TabEntry[DENSE_SIZE]
    KsCache ks;
    // optional second scoring table over the same rows (the R driver's "random" pass, lib/DeNovoAssembler.R:325-333):
    // scored in the same call from the same placement (BS_WANT_SECOND_TABLE)
    bool has_table2 = false;
    std::vector<double> prob2_dense;
    DevBuf d_tab2;
    KsCache ks2;

    Workspace ws[kWorkspaces];
    unsigned ws_cursor = 0;  // workspaces rotate across calls too: an asynchronous (device-result) call may still own one
    // scratch shared by all chunks (kernels of different chunks never overlap: one compute stream)
    DevBuf d_best, d_hits_ovf, d_keys, d_scratch, d_ovf, d_status, d_rank_scratch, d_counters, d_hbuf;
    DevBuf sim_meta, sim_chars, sim_words, sim_mask, sim_cdf, sim_starts, sim_kept, sim_reads;  // bs_simulate_reads
    size_t best_elems = 0;
    bool best_dirty = true;

    bool status_pending = false;  // a device-result call wanted KS-B: d_status has not been read since
    // pinned staging of this context's share of the contigs in bs_score_multi (grow-only; a pageable std::vector
    // went up at a fifth of the PCIe rate)
    char *h_multi = nullptr;
    size_t h_multi_cap = 0;
    // bs_score_multi with a large read set: the reads cross PCIe ONCE (context 0) and reach the other GPUs by peer
    // copies over NVLink; contigs and truth of this context's share go up beside them
    DevBuf m_reads, m_ctgs, m_truth;
    cudaEvent_t ev_share = nullptr;

    // interrupt poll of the calling thread (bs_ctx_set_poll)
    bs_poll_fn poll = nullptr;
    void *poll_user = nullptr;

    // timing
    bool timing = false;
    std::vector<TimedSpan> spans;
    size_t spans_used = 0;
};

namespace {

int fail(bs_ctx *ctx, int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx ? ctx->err : g_create_error, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define BS_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, BS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),  \
                        __FILE__, __LINE__);                                                       \
    } while (0)
#define BS_TRY(expr)                 \
    do {                             \
        int rc_ = (expr);            \
        if (rc_ != BS_OK) return rc_; \
    } while (0)

int ensure(bs_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap && b.p) return BS_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(ctx, BS_ERR_ALLOC, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return BS_OK;
}

void release(DevBuf &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

// host-side assembly of the small per-chunk metadata arrays into one pinned block, one H2D
struct MetaBuilder {
    std::vector<unsigned char> bytes;
    template <class T>
    size_t add(const T *src, size_t n) {
        size_t off = (bytes.size() + 15) & ~(size_t)15;
        bytes.resize(off + std::max<size_t>(n, 1) * sizeof(T));
        if (n) std::memcpy(bytes.data() + off, src, n * sizeof(T));
        return off;
    }
};

struct StageTimer {
    bs_ctx *ctx;
    cudaStream_t st;
    TimedSpan *span = nullptr;
    StageTimer(bs_ctx *c, Stage s, cudaStream_t stream) : ctx(c), st(stream) {
        static const char *const names[ST_COUNT] = {"bs:h2d", "bs:pack", "bs:place", "bs:score", "bs:truth_spectrum", "bs:prob_dist_ks",
                                                    "bs:ks_path_freq", "bs:startpos", "bs:d2h", "bs:lev"};
        BS_NVTX_PUSH(names[(int)s]);
        if (!ctx->timing) return;
        if (ctx->spans_used == ctx->spans.size()) {
            TimedSpan t{(int)s, nullptr, nullptr};
            if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess) return;
            ctx->spans.push_back(t);
        }
        span = &ctx->spans[ctx->spans_used++];
        span->stage = (int)s;
        cudaEventRecord(span->a, st);
    }
    ~StageTimer() {
        if (span) cudaEventRecord(span->b, st);
        BS_NVTX_POP();
    }
};

int grid_for(int64_t work_items, int threads, int cap) {
    int64_t g = (work_items + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

// resident blocks per SM of a persistent kernel (grid = SMs x this): what the hardware can hold
template <class K>
int blocks_per_sm(K kernel, int threads, size_t smem) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) n = 1;
    return n;
}

int sync_all(bs_ctx *ctx) {
    BS_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    BS_CUDA(cudaStreamSynchronize(ctx->stream));
    BS_CUDA(cudaStreamSynchronize(ctx->out_stream));
    return BS_OK;
}

int upload_table(bs_ctx *ctx, const std::vector<double> &prob_dense, DevBuf &d_tab) {
    std::vector<bs::TabEntry> tab(bs::DENSE_SIZE);
    for (int i = 0; i < bs::DENSE_SIZE; i++) {
        tab[i].prob = prob_dense[i];
        tab[i].row = ctx->row_dense[i];
        tab[i].pad = 0;
    }
    BS_TRY(ensure(ctx, d_tab, tab.size() * sizeof(bs::TabEntry)));
    BS_CUDA(cudaMemcpyAsync(d_tab.p, tab.data(), tab.size() * sizeof(bs::TabEntry), cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaStreamSynchronize(ctx->stream));
    return BS_OK;
}

// sorted distinct values + rank maps for the KS statistics of window length kmer
int prepare_ks(bs_ctx *ctx, int kmer, KsCache &k, const std::vector<double> &xprob_dense) {
    if (k.kmer == kmer && k.table_version == ctx->table_version) return BS_OK;
    const bool in_range = kmer >= 1 && kmer <= bs::MAXK;
    const int ncode = in_range ? 1 << (2 * kmer) : 1;
    const int doff = in_range ? bs::dense_offset(kmer) : 0;
    std::vector<double> xv{0.0}, yv;
    if (in_range) {
        for (int c = 0; c < ncode; c++)
            if (ctx->row_dense[doff + c] >= 0) {
                xv.push_back(xprob_dense[doff + c]);
                yv.push_back(ctx->tprob_dense[doff + c]);
            }
    }
    std::sort(xv.begin(), xv.end());
    xv.erase(std::unique(xv.begin(), xv.end()), xv.end());
    std::sort(yv.begin(), yv.end());
    yv.erase(std::unique(yv.begin(), yv.end()), yv.end());
    std::vector<int32_t> rank_x(ncode, 0), rank_y(ncode, -1);
    const int rank_zero = (int)(std::lower_bound(xv.begin(), xv.end(), 0.0) - xv.begin());
    if (in_range) {
        for (int c = 0; c < ncode; c++) {
            if (ctx->row_dense[doff + c] < 0) { rank_x[c] = rank_zero; continue; }
            rank_x[c] = (int32_t)(std::lower_bound(xv.begin(), xv.end(), xprob_dense[doff + c]) - xv.begin());
            rank_y[c] = (int32_t)(std::lower_bound(yv.begin(), yv.end(), ctx->tprob_dense[doff + c]) - yv.begin());
        }
    }
    std::vector<bs::LeLt> lelt(xv.size());
    for (size_t i = 0; i < xv.size(); i++) {
        lelt[i].le = (int32_t)(std::upper_bound(yv.begin(), yv.end(), xv[i]) - yv.begin()) - 1;
        lelt[i].lt = (int32_t)(std::lower_bound(yv.begin(), yv.end(), xv[i]) - yv.begin()) - 1;
    }
    std::vector<bs::WinEntry> win((size_t)ncode);
    for (int c = 0; c < ncode; c++) {
        const bool in_table = in_range && ctx->row_dense[doff + c] >= 0;
        win[c].prob = in_table ? xprob_dense[doff + c] : 0.0;
        win[c].rank = rank_x[c];
        win[c].row = in_table ? ctx->row_dense[doff + c] : -1;
    }
    k.y_max = yv.empty() ? 0.0 : yv.back();
    k.zero_le = (int)(std::upper_bound(yv.begin(), yv.end(), 0.0) - yv.begin()) - 1;
    k.zero_lt = (int)(std::lower_bound(yv.begin(), yv.end(), 0.0) - yv.begin()) - 1;
    if (yv.empty()) yv.push_back(0.0);  // keep the device array non-empty; R_y stays 0
    BS_TRY(ensure(ctx, k.win, win.size() * sizeof(bs::WinEntry)));
    BS_TRY(ensure(ctx, k.rank_y, rank_y.size() * 4));
    BS_TRY(ensure(ctx, k.lelt, lelt.size() * sizeof(bs::LeLt)));
    BS_TRY(ensure(ctx, k.yv, yv.size() * 8));
    BS_CUDA(cudaMemcpyAsync(k.win.p, win.data(), win.size() * sizeof(bs::WinEntry), cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaMemcpyAsync(k.rank_y.p, rank_y.data(), rank_y.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaMemcpyAsync(k.lelt.p, lelt.data(), lelt.size() * sizeof(bs::LeLt), cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaMemcpyAsync(k.yv.p, yv.data(), yv.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaStreamSynchronize(ctx->stream));
    k.kmer = kmer;
    k.table_version = ctx->table_version;
    k.R_x = (int)xv.size();
    int ry = 0;  // number of distinct truth-side values actually present
    for (int c = 0; c < ncode; c++) ry = std::max(ry, rank_y[c] + 1);
    k.R_y = ry;
    k.rank_zero = rank_zero;
    return BS_OK;
}

int prepare_ks(bs_ctx *ctx, int kmer) { return prepare_ks(ctx, kmer, ctx->ks, ctx->prob_dense); }

int check_offsets(bs_ctx *ctx, const char *what, const int64_t *off, int64_t n) {
    if (!off) return fail(ctx, BS_ERR_INVALID, "%s offsets are NULL", what);
    if (off[0] < 0) return fail(ctx, BS_ERR_INVALID, "%s offsets start below 0", what);
    for (int64_t i = 0; i < n; i++)
        if (off[i + 1] < off[i]) return fail(ctx, BS_ERR_INVALID, "%s offsets are not monotone at %lld", what, (long long)i);
    return BS_OK;
}

// everything of a call that does not change from chunk to chunk
struct CallEnv {
    const bs_batch *b;
    bs_result *res;
    int kmer;
    uint32_t flags;
    bool dev_chars, dev_res, want_ks, want_pd, want_pos, want_hist, want_sp, want_lev, second;
    bool w_out, w_in;  // two-phase scoring: stop after the placement / start from given position weights
    const char *read_chars;  // base of read 0 (shifted when offsets turned out to be uniform)
    const int64_t *roff;     // read offsets or NULL (every read has rlen bytes, dense)
    int32_t rlen;
    int64_t T;
    // compositional scoring (bs_score_scaffolds): the contigs of the call are scaffolds given as parts of base contigs
    const bs_scaffold_set *comp = nullptr;
    const int32_t *part_dst = nullptr;  // [parts] scaffold position of the first base a part adds
    bool comp_score = false;            // the break k-mers are scored inside the compositional placement kernel
};

int64_t read_byte_begin(const CallEnv &e, int64_t n) { return e.roff ? e.roff[n] : n * (int64_t)e.rlen; }


// One chunk on its way through the pipeline: chunk-local metadata, device views and result
// destinations (prepare), then one method per stage of the compute stream, then the results.
struct ChunkRun {
    bs_ctx *ctx;
    Workspace &ws;
    const CallEnv &e;
    const Chunk &ch;
    const bs_batch *b;
    bs_result *res;
    int64_t S, N, C, T;
    int kmer;
    cudaStream_t st;
    int grid_cap = 0;

    // sizes found while building the metadata
    int W = 1;
    bool tile_mode = false, ks_a = false, ks_b = false;
    int64_t max_ctg = 0, max_tr = 0, max_seg_reads = 0, head_total = 0, w_elems = 0, pd_elems = 0, pos_elems = 0;
    int tile_len = 0, hash_size = 0, found_words = 0;  // tile placement modes
    size_t place_smem = 0, n_items = 0;

    // device views of the chunk
    const int64_t *d_ctg_off = nullptr, *d_ctg_woff = nullptr, *d_tr_off = nullptr, *d_tr_woff = nullptr, *d_seg_rs = nullptr,
                  *d_seg_cs = nullptr, *d_tab_off = nullptr, *d_pd_off = nullptr, *d_pos_off = nullptr, *d_roff = nullptr;
    const int32_t *d_ctg_seg = nullptr, *d_seed = nullptr, *d_tab_mask = nullptr, *d_order = nullptr;
    const int64_t *d_sp_tab_off = nullptr;    // contig-in-truth seed tables: first slot per segment
    const int32_t *d_sp_tab_mask = nullptr;   // ... and slots - 1 (one table per segment geometry)
    bool sp_big = false;
    int64_t sp_slots = 0, sp_bitmap_words = 0;
    const int32_t *d_order_small = nullptr, *d_order_large = nullptr;  // work order split at KS_SMALL_MAX windows
    const int32_t *d_order_score = nullptr;                            // contigs below FUSE_MIN_LEN
    int64_t n_small = 0, n_large = 0, n_score = 0;
    int64_t fuse_min_len = bs::FUSE_MIN_LEN;  // (BS_FUSE_MIN_LEN: tuning runs)
    const bs::PlaceItem *d_items = nullptr;
    const uint8_t *d_rchars = nullptr, *d_cchars = nullptr, *d_tchars = nullptr;
    bs::SeqSet cs, ts;
    bs::ReadSet rs;
    bs::ReadIndex ix;
    // compositional scoring: the base contigs as their own small contig set, the parts of the scaffolds
    bs::SeqSet bset;
    bs::ScaffoldParts sparts;
    const int64_t *d_base_off = nullptr, *d_base_woff = nullptr, *d_base_pos_off = nullptr;
    const int32_t *d_base_seg = nullptr, *d_base_order = nullptr;
    int64_t n_base = 0, base_bytes = 0, base_w_elems = 0, max_read_len = 0;
    // ... extended by the texts around the distinct junctions (previous base contig's tail + next one's head): placed like base
    // contigs, once, so that a scaffold takes its junction-crossing reads from a list as well (reads of one length only)
    int64_t n_ext = 0, ext_bytes = 0;
    std::vector<char> junc_chars;
    const int32_t *d_part_jid = nullptr;
    bool comp_text = true;  // the scaffold texts are composed on the device (else only their packed words: base contigs of ACGT only)
    int32_t *w_ptr = nullptr, *total_ptr = nullptr;  // position weights and reads placed per contig (workspace, or the caller's with BS_WEIGHTS_*)

    // result destinations on the device (user arrays with BS_DEVICE_RESULT, else the workspace)
    int32_t *o_len = nullptr, *o_breaks = nullptr, *o_startpos = nullptr, *o_lev = nullptr, *o_hist = nullptr, *o_pos = nullptr;
    double *o_score = nullptr, *o_norm = nullptr, *o_bylen = nullptr, *o_ksa = nullptr, *o_ksb = nullptr, *o_pd = nullptr;
    // ... for the second table (same order)
    double *o_score2 = nullptr, *o_norm2 = nullptr, *o_bylen2 = nullptr, *o_ksa2 = nullptr, *o_ksb2 = nullptr, *o_pd2 = nullptr;

    ChunkRun(bs_ctx *c, Workspace &w, const CallEnv &env, const Chunk &chunk)
        : ctx(c), ws(w), e(env), ch(chunk), b(env.b), res(env.res), S(chunk.s1 - chunk.s0), N(chunk.r1 - chunk.r0),
          C(chunk.c1 - chunk.c0), T(env.T), kmer(env.kmer), st(c->stream), grid_cap(c->sm_count * 32) {}

    int prepare();    // metadata, buffers, H2D, result destinations, memsets
    int pack();       // 2-bit packing of contigs, truths and reads (+ read index)
    int place();      // leftmost placement of every read in every contig of its segment
    int place_composed();  // ... of a scaffold set: reads placed in the base contigs, scaffolds from their parts
    bool ks_compose() const;  // KS-A of a scaffold set from the window ranks of its base contigs (k_ks_compose)
    int spectrum();   // truth-side distribution of the KS statistics
    bool fused(int which) const;  // scores computed inside the KS-A kernels (kmer == 8, no dense histogram, those kernels run)
    int score_args(int which, int64_t rows, bs::ScoreArgs &sa);  // k_break_score's arguments (rows: blocks that may use a scratch row)
    int score(int which);      // weighted sums, histogram, KS-B (which: 0 scoring table, 1 second table)
    int prob_dist(int which);  // path_prob_dist and KS-A
    int second_table();        // both again for the second table, from the same placement
    int startpos();   // contig-in-truth offset
    int lev();        // infix edit distance
    int results();    // D2H (or the tail of the device-result path)
};

int ChunkRun::prepare() {
    // the workspace's previous chunk must have left it (its results are on their way or home)
    if (ws.in_flight) {
        BS_CUDA(cudaEventSynchronize(ws.ev_d2h));
        ws.in_flight = false;
    }

    // ---------------- chunk-local metadata ----------------
    const bool comp = e.comp != nullptr;  // the contigs are scaffolds given as parts: their text is composed on the device
    const int64_t ctg_b0 = b->contig_off[ch.c0], ctg_bytes = b->contig_off[ch.c1] - ctg_b0;
    const int64_t tr_b0 = b->truth_off[ch.s0], tr_bytes = b->truth_off[ch.s1] - tr_b0;
    const int64_t rd_b0 = read_byte_begin(e, ch.r0), read_bytes = read_byte_begin(e, ch.r1) - rd_b0;

    std::vector<int64_t> ctg_off(C + 1), ctg_woff(C + 1), tr_off(S + 1), tr_woff(S + 1), seg_rs(S + 1), seg_cs(S + 1), tab_off(std::max<int64_t>(S, 1));
    std::vector<int32_t> ctg_seg(std::max<int64_t>(C, 1)), seed_len(std::max<int64_t>(S, 1)), tab_mask(std::max<int64_t>(S, 1));
    ctg_woff[0] = 0;
    for (int64_t c = 0; c <= C; c++) ctg_off[c] = b->contig_off[ch.c0 + c] - ctg_b0;
    for (int64_t c = 0; c < C; c++) {
        const int64_t L = ctg_off[c + 1] - ctg_off[c];
        max_ctg = std::max(max_ctg, L);
        ctg_woff[c + 1] = ctg_woff[c] + (L + 31) / 32 + 2;
    }
    tr_woff[0] = 0;
    int64_t max_read = 0;
    for (int64_t s = 0; s <= S; s++) {
        tr_off[s] = b->truth_off[ch.s0 + s] - tr_b0;
        seg_rs[s] = b->seg_read_start[ch.s0 + s] - ch.r0;
        seg_cs[s] = b->seg_contig_start[ch.s0 + s] - ch.c0;
    }
    for (int64_t s = 0; s < S; s++) {
        const int64_t L = tr_off[s + 1] - tr_off[s];
        max_tr = std::max(max_tr, L);
        tr_woff[s + 1] = tr_woff[s] + (L + 31) / 32 + 2;
        for (int64_t c = b->seg_contig_start[ch.s0 + s]; c < b->seg_contig_start[ch.s0 + s + 1]; c++) ctg_seg[c - ch.c0] = (int32_t)s;
        const int64_t ns = seg_rs[s + 1] - seg_rs[s];
        max_seg_reads = std::max(max_seg_reads, ns);
        int64_t mn = 32;
        if (e.roff) {
            for (int64_t n = ch.r0 + seg_rs[s]; n < ch.r0 + seg_rs[s + 1]; n++) {
                const int64_t l = e.roff[n + 1] - e.roff[n];
                max_read = std::max(max_read, l);
                if (l > 0 && l < mn) mn = l;
            }
        } else {
            max_read = e.rlen;
            mn = std::min<int64_t>(32, std::max<int64_t>(1, e.rlen));
        }
        seed_len[s] = (int32_t)mn;
        int64_t hs = 64;
        while (hs < 2 * ns) hs <<= 1;
        tab_off[s] = head_total;
        tab_mask[s] = (int32_t)(hs - 1);
        head_total += hs;
    }
    // contig-in-truth seed tables (bs_startpos.cuh): groups of SP_GROUP contigs with fixed-size tables, or -- when a
    // segment holds more contigs than one group (cfg-4 / cfg-5) -- ONE table per segment sized for its contigs
    std::vector<int64_t> sp_tab_off(std::max<int64_t>(S, 1), 0);
    std::vector<int32_t> sp_tab_mask(std::max<int64_t>(S, 1), bs::SP_SLOTS - 1);
    {
        int64_t max_seg_contigs = 0;
        for (int64_t s = 0; s < S; s++) max_seg_contigs = std::max(max_seg_contigs, seg_cs[s + 1] - seg_cs[s]);
        sp_big = max_seg_contigs > bs::SP_GROUP;
        if (const char *env = std::getenv("BS_STARTPOS_BIG")) sp_big = env[0] == '1';  // tests: force a geometry
        int64_t slots = 0, groups = 0;
        for (int64_t s = 0; s < S; s++) {
            const int64_t nc = seg_cs[s + 1] - seg_cs[s];
            sp_tab_off[s] = slots;
            if (sp_big) {
                int64_t sl = 64;
                while (sl < 2 * nc) sl <<= 1;
                sp_tab_mask[s] = (int32_t)(sl - 1);
                slots += sl + 1;
                groups += 1;
            } else {
                const int64_t g = std::max<int64_t>(1, bs::startpos_groups_of(nc));
                slots += g * (bs::SP_SLOTS + 1);
                groups += g;
            }
        }
        sp_slots = slots;
        sp_bitmap_words = sp_big ? S * (int64_t)bs::SP_BITMAP_WORDS_BIG : groups * (int64_t)bs::SP_BITMAP_WORDS;
    }
    if (max_read > 0x3fffffff) return fail(ctx, BS_ERR_INVALID, "read longer than 2^30");
    W = (int)std::max<int64_t>(1, (max_read + 31) / 32);
    max_read_len = max_read;
    tile_mode = (e.flags & (BS_PLACE_SCAN | BS_PLACE_TILE)) != 0;

    // placement order of the index kernel: longest contigs first
    std::vector<int32_t> order;
    std::vector<bs::PlaceItem> items;
    // work order of the persistent kernels: the long contigs first, longest first (they set the
    // tail), then the rest in input order so that the blocks running at one time share few
    // segments (read index and reads of those segments stay in L2)
    order.resize((size_t)C);
    std::iota(order.begin(), order.end(), 0);
    {
        const int64_t long_len = 8192;
        auto len_of = [&](int32_t x) { return ctg_off[x + 1] - ctg_off[x]; };
        auto mid = std::stable_partition(order.begin(), order.end(), [&](int32_t x) { return len_of(x) > long_len; });
        std::stable_sort(order.begin(), mid, [&](int32_t x, int32_t y) { return len_of(x) > len_of(y); });
    }
    // the KS-A stage runs short contigs (most of a velvet-style set) through its own kernel
    std::vector<int32_t> order_small, order_large;
    for (int32_t c : order) {
        const int64_t nwin = ctg_off[c + 1] - ctg_off[c] - kmer + 1;
        (nwin <= bs::KS_SMALL_MAX ? order_small : order_large).push_back(c);
    }
    n_small = (int64_t)order_small.size();
    n_large = (int64_t)order_large.size();
    // k_break_score's share when the long-contig KS-A kernel scores on its way (fused()): contigs below FUSE_MIN_LEN
    std::vector<int32_t> order_score;
    if (const char *env = std::getenv("BS_FUSE_MIN_LEN")) fuse_min_len = std::max<int64_t>(std::atoll(env), bs::KS_SMALL_MAX + bs::MAXK + 1);
    for (int32_t c : order)
        if (ctg_off[c + 1] - ctg_off[c] < fuse_min_len) order_score.push_back(c);
    n_score = (int64_t)order_score.size();
    if (tile_mode) {
        // tile placement: one work item per (contig, read chunk)
        tile_len = (int)std::min<int64_t>(kMaxTile, std::max<int64_t>(32, (max_ctg + 31) / 32 * 32));
        hash_size = 64;
        while (hash_size < tile_len) hash_size <<= 1;
        items.reserve((size_t)C);
        const int64_t target_items = (int64_t)ctx->sm_count * 8;
        int64_t max_chunk = 1;
        for (int64_t c = 0; c < C; c++) {
            const int32_t s = ctg_seg[c];
            const int64_t r0 = seg_rs[s], r1 = seg_rs[s + 1];
            const int64_t nr = r1 - r0;
            int64_t nchunks = 1;
            if (C < target_items) nchunks = std::min((target_items + C - 1) / C, std::max<int64_t>(1, nr / kMinChunk));
            nchunks = std::max(nchunks, (nr + kMaxChunk - 1) / kMaxChunk);
            nchunks = std::max<int64_t>(nchunks, 1);
            const int64_t per = (nr + nchunks - 1) / nchunks;
            for (int64_t k = 0; k < nchunks; k++) {
                bs::PlaceItem it;
                it.contig = (int32_t)c;
                it.seg = s;
                it.read_begin = r0 + k * per;
                it.read_end = std::min(r1, it.read_begin + per);
                if (k > 0 && it.read_begin >= it.read_end) break;
                max_chunk = std::max(max_chunk, it.read_end - it.read_begin);
                items.push_back(it);
            }
        }
        found_words = (int)((max_chunk + 31) / 32);
        place_smem = bs::place_smem_bytes(tile_len, hash_size, W, found_words);
        if (place_smem > ctx->smem_optin) return fail(ctx, BS_ERR_INVALID, "placement tile needs %zu bytes of shared memory", place_smem);
    }

    std::vector<int64_t> pd_off, pos_off;
    if (e.want_pd) {
        pd_off.resize(C + 1);
        for (int64_t c = 0; c <= C; c++) pd_off[c] = res->path_prob_dist_off[ch.c0 + c] - res->path_prob_dist_off[ch.c0];
    }
    if (e.want_pos) {
        pos_off.resize(C + 1);
        for (int64_t c = 0; c <= C; c++) pos_off[c] = res->pos_off[ch.c0 + c] - res->pos_off[ch.c0];
    }
    std::vector<int64_t> roff_local;
    if (e.roff) {
        roff_local.resize(N + 1);
        for (int64_t n = 0; n <= N; n++) roff_local[n] = e.roff[ch.r0 + n] - rd_b0;
    }

    MetaBuilder mb;
    const size_t o_ctg_off = mb.add(ctg_off.data(), (size_t)C + 1);
    const size_t o_ctg_woff = mb.add(ctg_woff.data(), (size_t)C + 1);
    const size_t o_ctg_seg = mb.add(ctg_seg.data(), (size_t)C);
    const size_t o_tr_off = mb.add(tr_off.data(), (size_t)S + 1);
    const size_t o_tr_woff = mb.add(tr_woff.data(), (size_t)S + 1);
    const size_t o_seg_rs = mb.add(seg_rs.data(), (size_t)S + 1);
    const size_t o_seg_cs = mb.add(seg_cs.data(), (size_t)S + 1);
    const size_t o_seed = mb.add(seed_len.data(), (size_t)S);
    const size_t o_tab_off = mb.add(tab_off.data(), (size_t)S);
    const size_t o_tab_mask = mb.add(tab_mask.data(), (size_t)S);
    const size_t o_sp_tab_off = mb.add(sp_tab_off.data(), (size_t)S);
    const size_t o_sp_tab_mask = mb.add(sp_tab_mask.data(), (size_t)S);
    const size_t o_order = mb.add(order.data(), order.size());
    const size_t o_order_small = mb.add(order_small.data(), order_small.size());
    const size_t o_order_large = mb.add(order_large.data(), order_large.size());
    const size_t o_order_score = mb.add(order_score.data(), order_score.size());
    const size_t o_items = mb.add(items.data(), items.size());
    const size_t o_pd_off = e.want_pd ? mb.add(pd_off.data(), (size_t)C + 1) : 0;
    const size_t o_pos_off = e.want_pos ? mb.add(pos_off.data(), (size_t)C + 1) : 0;
    const size_t o_roff = e.roff ? mb.add(roff_local.data(), (size_t)N + 1) : 0;
    // compositional scoring: the base contigs as a contig set of their own (one segment) and the parts of the scaffolds
    size_t o_base_off = 0, o_base_woff = 0, o_base_seg = 0, o_base_order = 0, o_base_pos_off = 0, o_part_start = 0, o_part_base = 0,
           o_part_ov = 0, o_part_dst = 0, o_part_jid = 0;
    int64_t base_words = 0;
    if (comp) {
        const bs_scaffold_set *cp = e.comp;
        n_base = cp->n_base;
        const int64_t n_parts = cp->scaffold_part_start[cp->n_scaffolds];
        // distinct junctions (previous base contig, next base contig, overlap) whose crossing reads can be found once, in the
        // text  previous[L - span ..] + next[overlap .. overlap + span)  (span = read length - 1): both contigs at least span
        // long, reads of ONE length (then every occurrence in that text crosses the junction, and the leftmost one the read
        // index placement reports is the leftmost crossing one).  Other junctions are probed per scaffold (k_place_compose).
        std::vector<int32_t> part_jid((size_t)std::max<int64_t>(n_parts, 1), -1);
        std::vector<int64_t> junc_len;
        junc_chars.clear();
        {
            const int64_t span = max_read_len - 1;
            const char *jenv = std::getenv("BS_COMPOSE_JUNCTIONS");  // tests: 0 probes every junction per scaffold
            if (!e.roff && span >= 1 && N > 0 && !(jenv && jenv[0] == '0')) {
                // (base contig pair -> junction: a direct table for the usual handful of base contigs -- one overlap per pair is
                // the rule, further ones go to the map -- so that the 10^5 parts of a cfg-4 set cost a table look-up each)
                std::unordered_map<uint64_t, int32_t> seen;
                const bool direct = n_base <= 2048;
                std::vector<int32_t> pair_jid(direct ? (size_t)(n_base * n_base) : 0, -1), pair_ov(direct ? (size_t)(n_base * n_base) : 0, 0);
                // (bounded by memory alone -- the lists are [contig][read] arrays -- so that which junctions have a list does not
                // depend on how many scaffolds travel in the call)
                const int64_t max_junc = std::max<int64_t>(0, ((int64_t)1 << 28) / std::max<int64_t>(N, 1) - n_base);
                for (int64_t c = 0; c < cp->n_scaffolds; c++)
                    for (int64_t g = cp->scaffold_part_start[c] + 1; g < cp->scaffold_part_start[c + 1]; g++) {
                        const int64_t bp = cp->part_base[g - 1], bn = cp->part_base[g], ov = cp->part_overlap[g];
                        const int64_t Lp = cp->base_off[bp + 1] - cp->base_off[bp], Ln = cp->base_off[bn + 1] - cp->base_off[bn];
                        if (Lp < span || Ln < span || ov >= 65536 || bp >= (1 << 24) || bn >= (1 << 24)) continue;
                        int32_t jid = -1;
                        const size_t slot = direct ? (size_t)(bp * n_base + bn) : 0;
                        const uint64_t key = ((uint64_t)bp << 40) | ((uint64_t)bn << 16) | (uint64_t)ov;
                        if (direct && pair_jid[slot] >= 0 && pair_ov[slot] == (int32_t)ov) jid = pair_jid[slot];
                        else if (!direct || pair_jid[slot] >= 0) {
                            auto it = seen.find(key);
                            if (it != seen.end()) jid = it->second;
                        }
                        if (jid < 0) {
                            if ((int64_t)junc_len.size() >= max_junc) continue;  // (the lists are [contig][read] arrays: bounded)
                            jid = (int32_t)junc_len.size();
                            if (direct && pair_jid[slot] < 0) { pair_jid[slot] = jid; pair_ov[slot] = (int32_t)ov; }
                            else seen.emplace(key, jid);
                            const int64_t head = std::min<int64_t>(Ln - ov, span);
                            junc_chars.insert(junc_chars.end(), cp->base_chars + cp->base_off[bp + 1] - span, cp->base_chars + cp->base_off[bp + 1]);
                            junc_chars.insert(junc_chars.end(), cp->base_chars + cp->base_off[bn] + ov, cp->base_chars + cp->base_off[bn] + ov + head);
                            junc_len.push_back(span + head);
                        }
                        part_jid[(size_t)g] = (int32_t)n_base + jid;
                    }
            }
        }
        n_ext = n_base + (int64_t)junc_len.size();
        std::vector<int64_t> base_off((size_t)n_ext + 1), base_woff((size_t)n_ext + 1), base_pos_off((size_t)n_ext + 1);
        std::vector<int32_t> base_seg((size_t)std::max<int64_t>(n_ext, 1), 0), base_order((size_t)n_ext);
        base_woff[0] = 0;
        for (int64_t i = 0; i <= n_base; i++) base_off[i] = cp->base_off[i] - cp->base_off[0];
        for (int64_t i = n_base; i < n_ext; i++) base_off[i + 1] = base_off[i] + junc_len[(size_t)(i - n_base)];
        for (int64_t i = 0; i <= n_ext; i++) base_pos_off[i] = i * N;
        for (int64_t i = 0; i < n_ext; i++) base_woff[i + 1] = base_woff[i] + (base_off[i + 1] - base_off[i] + 31) / 32 + 2;
        std::iota(base_order.begin(), base_order.end(), 0);
        std::stable_sort(base_order.begin(), base_order.end(), [&](int32_t x, int32_t y) { return base_off[x + 1] - base_off[x] > base_off[y + 1] - base_off[y]; });
        ext_bytes = base_off[n_ext];
        base_bytes = base_off[n_base];
        base_words = base_woff[n_ext];
        {   // no kernel looks at the text of a contig made of ACGT only (every byte-comparison path starts from a byte outside
            // ACGT in the contig, the read or the truth, which then cannot match): such a set is kept as packed words alone
            const char *bc = cp->base_chars + cp->base_off[0];
            bool pure = true;
            for (int64_t i = 0; pure && i < base_bytes; i++) pure = bc[i] == 'A' || bc[i] == 'C' || bc[i] == 'G' || bc[i] == 'T';
            const char *env = std::getenv("BS_COMPOSE_TEXT");  // tests: 1 composes the text anyway
            comp_text = !pure || e.want_lev || (env && env[0] == '1');
        }
        base_w_elems = ext_bytes + n_ext;
        o_base_off = mb.add(base_off.data(), base_off.size());
        o_base_woff = mb.add(base_woff.data(), base_woff.size());
        o_base_seg = mb.add(base_seg.data(), base_seg.size());
        o_base_order = mb.add(base_order.data(), base_order.size());
        o_base_pos_off = mb.add(base_pos_off.data(), base_pos_off.size());
        o_part_start = mb.add(cp->scaffold_part_start, (size_t)cp->n_scaffolds + 1);
        o_part_base = mb.add(cp->part_base, (size_t)n_parts);
        o_part_ov = mb.add(cp->part_overlap, (size_t)n_parts);
        o_part_dst = mb.add(e.part_dst, (size_t)n_parts);
        o_part_jid = mb.add(part_jid.data(), (size_t)n_parts);
    }

    if (mb.bytes.size() > ws.h_meta_cap) {
        if (ws.h_meta) cudaFreeHost(ws.h_meta);
        ws.h_meta = nullptr;
        ws.h_meta_cap = 0;
        const size_t want = mb.bytes.size() * 2 + 4096;
        if (cudaHostAlloc(&ws.h_meta, want, cudaHostAllocDefault) != cudaSuccess)
            return fail(ctx, BS_ERR_ALLOC, "cudaHostAlloc(%zu) failed", want);
        ws.h_meta_cap = want;
    }
    std::memcpy(ws.h_meta, mb.bytes.data(), mb.bytes.size());

    // ---------------- device buffers ----------------
    w_elems = ctg_bytes + C;
    pd_elems = e.want_pd ? pd_off[C] : 0;
    pos_elems = e.want_pos ? pos_off[C] : 0;
    n_items = items.size();
    BS_TRY(ensure(ctx, ws.meta, mb.bytes.size()));
    if (!e.dev_chars) {
        BS_TRY(ensure(ctx, ws.read_chars, (size_t)read_bytes + 32));
        BS_TRY(ensure(ctx, ws.tr_chars, (size_t)tr_bytes + 32));
    }
    if (!e.dev_chars || (comp && comp_text)) BS_TRY(ensure(ctx, ws.ctg_chars, (size_t)ctg_bytes + 32));
    if (comp) {
        BS_TRY(ensure(ctx, ws.base_chars, (size_t)ext_bytes + 32));
        BS_TRY(ensure(ctx, ws.base_words, (size_t)base_words * 8 + 8));
        BS_TRY(ensure(ctx, ws.base_mask, (size_t)base_words * 4 + 8));
        BS_TRY(ensure(ctx, ws.base_pos, (size_t)std::max<int64_t>(n_ext * N, 1) * 4));
        BS_TRY(ensure(ctx, ws.base_w, (size_t)std::max<int64_t>(base_w_elems, 1) * 4));
        BS_TRY(ensure(ctx, ws.base_total, (size_t)std::max<int64_t>(n_ext, 1) * 4));
        BS_TRY(ensure(ctx, ws.base_hits, (size_t)std::max<int64_t>(n_ext * N, 1) * 8));
        BS_TRY(ensure(ctx, ws.base_cnt, (size_t)std::max<int64_t>(n_ext, 1) * 4));
        BS_TRY(ensure(ctx, ws.base_di, (size_t)std::max<int64_t>(n_ext * N, 1) * 4));
        BS_TRY(ensure(ctx, ws.base_rank, (size_t)std::max<int64_t>(ext_bytes, 1) * 2 + 16));
    }
    BS_TRY(ensure(ctx, ws.rwords, (size_t)std::max<int64_t>(N, 1) * W * 8));
    BS_TRY(ensure(ctx, ws.rflags, (size_t)std::max<int64_t>(N, 1) + 8));
    BS_TRY(ensure(ctx, ws.cwords, (size_t)ctg_woff[C] * 8 + 8));
    BS_TRY(ensure(ctx, ws.cmask, (size_t)ctg_woff[C] * 4 + 8));
    BS_TRY(ensure(ctx, ws.twords, (size_t)tr_woff[S] * 8 + 8));
    BS_TRY(ensure(ctx, ws.tmask, (size_t)tr_woff[S] * 4 + 8));
    if (e.w_out || e.w_in) {
        w_ptr = res->weights + (b->contig_off[ch.c0] - b->contig_off[0]) + ch.c0;
        total_ptr = res->weights_total + ch.c0;
    } else {
        if (!e.comp_score) BS_TRY(ensure(ctx, ws.w, (size_t)w_elems * 4));  // (scored inside the compositional placement: no position weights)
        BS_TRY(ensure(ctx, ws.total, (size_t)C * 4));
        w_ptr = e.comp_score ? nullptr : (int32_t *)ws.w.p;
        total_ptr = (int32_t *)ws.total.p;
    }
    if (!tile_mode) {
        BS_TRY(ensure(ctx, ws.head, (size_t)std::max<int64_t>(head_total, 1) * 4));
        BS_TRY(ensure(ctx, ws.next, (size_t)std::max<int64_t>(N, 1) * 8));
        BS_TRY(ensure(ctx, ws.odd_head, (size_t)std::max<int64_t>(S, 1) * 4));
    }
    if (!e.dev_res) {
        BS_TRY(ensure(ctx, ws.out_i32, (size_t)4 * C * 4));
        BS_TRY(ensure(ctx, ws.out_f64, (size_t)10 * C * 8));
        if (e.want_pd) BS_TRY(ensure(ctx, ws.pd, (size_t)std::max<int64_t>(pd_elems, 1) * 8));
        if (e.want_pd && e.second) BS_TRY(ensure(ctx, ws.pd2, (size_t)std::max<int64_t>(pd_elems, 1) * 8));
        if (e.want_hist) BS_TRY(ensure(ctx, ws.hist, (size_t)C * (T + 1) * 4));
        if (e.want_pos) BS_TRY(ensure(ctx, ws.pos, (size_t)std::max<int64_t>(pos_elems, 1) * 4));
    }

    unsigned char *dm = (unsigned char *)ws.meta.p;
    d_ctg_off = (const int64_t *)(dm + o_ctg_off);
    d_ctg_woff = (const int64_t *)(dm + o_ctg_woff);
    d_ctg_seg = (const int32_t *)(dm + o_ctg_seg);
    d_tr_off = (const int64_t *)(dm + o_tr_off);
    d_tr_woff = (const int64_t *)(dm + o_tr_woff);
    d_seg_rs = (const int64_t *)(dm + o_seg_rs);
    d_seg_cs = (const int64_t *)(dm + o_seg_cs);
    d_seed = (const int32_t *)(dm + o_seed);
    d_tab_off = (const int64_t *)(dm + o_tab_off);
    d_tab_mask = (const int32_t *)(dm + o_tab_mask);
    d_sp_tab_off = (const int64_t *)(dm + o_sp_tab_off);
    d_sp_tab_mask = (const int32_t *)(dm + o_sp_tab_mask);
    d_order = (const int32_t *)(dm + o_order);
    d_order_small = (const int32_t *)(dm + o_order_small);
    d_order_large = (const int32_t *)(dm + o_order_large);
    d_order_score = (const int32_t *)(dm + o_order_score);
    d_items = (const bs::PlaceItem *)(dm + o_items);
    d_pd_off = e.want_pd ? (const int64_t *)(dm + o_pd_off) : nullptr;
    d_pos_off = e.want_pos ? (const int64_t *)(dm + o_pos_off) : nullptr;
    d_roff = e.roff ? (const int64_t *)(dm + o_roff) : nullptr;
    if (comp) {
        d_base_off = (const int64_t *)(dm + o_base_off); d_base_woff = (const int64_t *)(dm + o_base_woff);
        d_base_seg = (const int32_t *)(dm + o_base_seg); d_base_order = (const int32_t *)(dm + o_base_order);
        d_base_pos_off = (const int64_t *)(dm + o_base_pos_off);
        sparts.part_start = (const int64_t *)(dm + o_part_start); sparts.part_base = (const int32_t *)(dm + o_part_base);
        sparts.part_ov = (const int32_t *)(dm + o_part_ov); sparts.part_dst = (const int32_t *)(dm + o_part_dst);
        d_part_jid = (const int32_t *)(dm + o_part_jid);
        bset = bs::SeqSet{(const uint8_t *)ws.base_chars.p, d_base_off, d_base_woff, (uint64_t *)ws.base_words.p, (uint32_t *)ws.base_mask.p, n_ext, base_words};
    }

    // ---------------- H2D (copy stream) ----------------
    {
        StageTimer tm(ctx, ST_H2D, ctx->copy_stream);
        cudaStream_t cs = ctx->copy_stream;
        BS_CUDA(cudaMemcpyAsync(ws.meta.p, ws.h_meta, mb.bytes.size(), cudaMemcpyHostToDevice, cs));
        if (e.dev_chars) {
            d_rchars = (const uint8_t *)e.read_chars + rd_b0;
            d_cchars = (const uint8_t *)b->contig_chars + ctg_b0;
            d_tchars = (const uint8_t *)b->truth_chars + tr_b0;
        } else {
            if (read_bytes) BS_CUDA(cudaMemcpyAsync(ws.read_chars.p, e.read_chars + rd_b0, (size_t)read_bytes, cudaMemcpyHostToDevice, cs));
            if (ctg_bytes && !comp) BS_CUDA(cudaMemcpyAsync(ws.ctg_chars.p, b->contig_chars + ctg_b0, (size_t)ctg_bytes, cudaMemcpyHostToDevice, cs));
            if (tr_bytes) BS_CUDA(cudaMemcpyAsync(ws.tr_chars.p, b->truth_chars + tr_b0, (size_t)tr_bytes, cudaMemcpyHostToDevice, cs));
            d_rchars = (const uint8_t *)ws.read_chars.p;
            d_cchars = (const uint8_t *)ws.ctg_chars.p;
            d_tchars = (const uint8_t *)ws.tr_chars.p;
        }
        if (comp) {  // only the base contigs cross PCIe; k_compose_text writes the scaffold texts into the workspace
            if (base_bytes) BS_CUDA(cudaMemcpyAsync(ws.base_chars.p, e.comp->base_chars + e.comp->base_off[0], (size_t)base_bytes, cudaMemcpyHostToDevice, cs));
            if (!junc_chars.empty())  // (pageable source: staged before the call returns)
                BS_CUDA(cudaMemcpyAsync((char *)ws.base_chars.p + base_bytes, junc_chars.data(), junc_chars.size(), cudaMemcpyHostToDevice, cs));
            d_cchars = comp_text ? (const uint8_t *)ws.ctg_chars.p : nullptr;
        }
    }
    BS_CUDA(cudaEventRecord(ws.ev_h2d, ctx->copy_stream));
    BS_CUDA(cudaStreamWaitEvent(st, ws.ev_h2d, 0));

    // result destinations on the device
    ks_a = e.want_ks && res->ks_stat_prob_dist;
    ks_b = e.want_ks && res->ks_stat_path_freq;
    if (e.dev_res) {
        o_len = res->sequence_len ? res->sequence_len + ch.c0 : nullptr;
        o_breaks = res->kmer_breaks ? res->kmer_breaks + ch.c0 : nullptr;
        o_startpos = e.want_sp ? res->path_prob_dist_startpos + ch.c0 : nullptr;
        o_lev = e.want_lev ? res->lev_dist_vs_true + ch.c0 : nullptr;
        o_score = res->bp_score ? res->bp_score + ch.c0 : nullptr;
        o_norm = res->bp_score_norm_by_break_freqs ? res->bp_score_norm_by_break_freqs + ch.c0 : nullptr;
        o_bylen = res->bp_score_norm_by_len ? res->bp_score_norm_by_len + ch.c0 : nullptr;
        o_ksa = ks_a ? res->ks_stat_prob_dist + ch.c0 : nullptr;
        o_ksb = ks_b ? res->ks_stat_path_freq + ch.c0 : nullptr;
        if (e.want_pd) o_pd = res->path_prob_dist + res->path_prob_dist_off[ch.c0];
        if (e.second) {
            o_score2 = res->bp_score2 ? res->bp_score2 + ch.c0 : nullptr;
            o_norm2 = res->bp_score_norm_by_break_freqs2 ? res->bp_score_norm_by_break_freqs2 + ch.c0 : nullptr;
            o_bylen2 = res->bp_score_norm_by_len2 ? res->bp_score_norm_by_len2 + ch.c0 : nullptr;
            o_ksa2 = (e.want_ks && res->ks_stat_prob_dist2) ? res->ks_stat_prob_dist2 + ch.c0 : nullptr;
            o_ksb2 = (e.want_ks && res->ks_stat_path_freq2) ? res->ks_stat_path_freq2 + ch.c0 : nullptr;
            if (e.want_pd) o_pd2 = res->path_prob_dist2 + res->path_prob_dist_off[ch.c0];
        }
        if (e.want_hist) o_hist = res->hist + ch.c0 * (T + 1);
        if (e.want_pos) o_pos = res->pos + res->pos_off[ch.c0];
    } else {
        int32_t *i32 = (int32_t *)ws.out_i32.p;
        double *f64 = (double *)ws.out_f64.p;
        o_len = i32; o_breaks = i32 + C; o_startpos = e.want_sp ? i32 + 2 * C : nullptr;
        o_lev = e.want_lev ? i32 + 3 * C : nullptr;
        o_score = f64; o_norm = f64 + C; o_bylen = f64 + 2 * C;
        o_ksa = ks_a ? f64 + 3 * C : nullptr;
        o_ksb = ks_b ? f64 + 4 * C : nullptr;
        if (e.want_pd) o_pd = (double *)ws.pd.p;
        if (e.second) {
            o_score2 = f64 + 5 * C; o_norm2 = f64 + 6 * C; o_bylen2 = f64 + 7 * C;
            o_ksa2 = (e.want_ks && res->ks_stat_prob_dist2) ? f64 + 8 * C : nullptr;
            o_ksb2 = (e.want_ks && res->ks_stat_path_freq2) ? f64 + 9 * C : nullptr;
            if (e.want_pd) o_pd2 = (double *)ws.pd2.p;
        }
        if (e.want_hist) o_hist = (int32_t *)ws.hist.p;
        if (e.want_pos) o_pos = (int32_t *)ws.pos.p;
    }

    if (!e.w_in) {
        if (w_ptr) BS_CUDA(cudaMemsetAsync(w_ptr, 0, (size_t)w_elems * 4, st));
        BS_CUDA(cudaMemsetAsync(total_ptr, 0, (size_t)C * 4, st));
    }
    if (comp) {
        BS_CUDA(cudaMemsetAsync(ws.base_w.p, 0, (size_t)std::max<int64_t>(base_w_elems, 1) * 4, st));
        BS_CUDA(cudaMemsetAsync(ws.base_total.p, 0, (size_t)std::max<int64_t>(n_ext, 1) * 4, st));
        BS_CUDA(cudaMemsetAsync(ws.base_cnt.p, 0, (size_t)std::max<int64_t>(n_ext, 1) * 4, st));
        BS_CUDA(cudaMemsetAsync(ws.base_pos.p, 0xff, (size_t)std::max<int64_t>(n_ext * N, 1) * 4, st));
    }
    BS_TRY(ensure(ctx, ctx->d_counters, 64));
    BS_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, 64, st));  // work counters of the persistent kernels
    if (o_hist) BS_CUDA(cudaMemsetAsync(o_hist, 0, (size_t)C * (T + 1) * 4, st));
    if (o_pos && pos_elems > 0) BS_CUDA(cudaMemsetAsync(o_pos, 0xff, (size_t)pos_elems * 4, st));

    // ---------------- kernels (compute stream) ----------------
    cs = bs::SeqSet{d_cchars, d_ctg_off, d_ctg_woff, (uint64_t *)ws.cwords.p, (uint32_t *)ws.cmask.p, C, ctg_woff[C]};
    ts = bs::SeqSet{d_tchars, d_tr_off, d_tr_woff, (uint64_t *)ws.twords.p, (uint32_t *)ws.tmask.p, S, tr_woff[S]};
    rs = bs::ReadSet{d_rchars, d_roff, e.rlen, W, N, (uint64_t *)ws.rwords.p, (uint8_t *)ws.rflags.p};
    std::memset(&ix, 0, sizeof(ix));
    if (!tile_mode) {
        ix.head = (uint32_t *)ws.head.p; ix.next = (uint2 *)ws.next.p; ix.odd_head = (uint32_t *)ws.odd_head.p;
        ix.tab_off = d_tab_off; ix.tab_mask = d_tab_mask; ix.seed_len = d_seed; ix.seg_read_start = d_seg_rs; ix.n_seg = (int32_t)S;
    }
    return BS_OK;
}

int ChunkRun::pack() {
    {
        StageTimer tm(ctx, ST_PACK, st);
        if (e.comp) {  // base contigs packed as their own set; scaffold texts written from them, then packed like any contig
            if (n_base > 0) {
                BS_LAUNCH(bs::k_pack_seqs, grid_for(bset.total_words, kPackThreads, grid_cap), kPackThreads, 0, st, bset);
                ctx->launches++;
            }
            if (comp_text) {
                bs::ComposeTextArgs ca;
                ca.sp = sparts; ca.base_chars = bset.chars; ca.base_off = d_base_off; ca.ctg_off = d_ctg_off; ca.ctg_chars = (uint8_t *)ws.ctg_chars.p;
                ca.n_scaffolds = C;
                BS_LAUNCH(bs::k_compose_text, (unsigned)std::min<int64_t>(C, (int64_t)ctx->sm_count * 8), kPackThreads, 0, st, ca);
            } else {
                bs::ComposeWordsArgs cw;
                cw.sp = sparts; cw.base_woff = d_base_woff; cw.base_words = bset.words; cw.base_mask = bset.mask;
                cw.ctg_off = d_ctg_off; cw.ctg_woff = d_ctg_woff; cw.ctg_words = cs.words; cw.ctg_mask = cs.mask; cw.n_scaffolds = C;
                BS_LAUNCH(bs::k_compose_words, (unsigned)std::min<int64_t>(C, (int64_t)ctx->sm_count * 8), kPackThreads, 0, st, cw);
            }
            ctx->launches++;
        }
        if (!e.comp || comp_text) {
            BS_LAUNCH(bs::k_pack_seqs, grid_for(cs.total_words, kPackThreads, grid_cap), kPackThreads, 0, st, cs);
            ctx->launches++;
        }
        if ((e.want_ks || e.want_sp || e.want_lev) && !e.w_out) {
            BS_LAUNCH(bs::k_pack_seqs, grid_for(ts.total_words, kPackThreads, grid_cap), kPackThreads, 0, st, ts);
            ctx->launches++;
        }
        if (N > 0 && !e.w_in) {
            if (!tile_mode) {
                BS_CUDA(cudaMemsetAsync(ws.head.p, 0, (size_t)head_total * 4, st));
                BS_CUDA(cudaMemsetAsync(ws.odd_head.p, 0, (size_t)S * 4, st));
            }
            const int tile_reads = (!e.roff && e.rlen >= 1) ? bs::packb_tile_reads(e.rlen) : 0;
            const char *bulk_env = std::getenv("BS_PACK_BULK");  // tests: 0 keeps the register-staged kernel
            if (tile_reads > 0 && bs::packb_stage_bytes(e.rlen, tile_reads) <= bs::PACKB_MAX_STAGE && !(bulk_env && bulk_env[0] == '0')) {
                // reads of one length: ASCII staged by bulk asynchronous copies (TMA unit), two stages per block
                const size_t smem = bs::packb_smem_bytes(e.rlen, tile_reads);
                BS_CUDA(cudaFuncSetAttribute(bs::k_pack_reads_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                const int64_t tiles = (N + tile_reads - 1) / tile_reads;
                const int per_sm = blocks_per_sm(bs::k_pack_reads_bulk, bs::PACKB_THREADS, smem);
                BS_LAUNCH(bs::k_pack_reads_bulk, (unsigned)std::min<int64_t>(tiles, (int64_t)ctx->sm_count * per_sm), bs::PACKB_THREADS, smem, st, rs, ix, tile_reads);
            } else if (!e.roff && e.rlen >= 1) {
                BS_CUDA(cudaMemsetAsync(ws.rflags.p, 0, (size_t)N + 8, st));
                const int64_t tiles = (N * W + bs::PACK_THREADS - 1) / bs::PACK_THREADS;
                BS_LAUNCH(bs::k_pack_reads_uniform, (unsigned)std::min<int64_t>(tiles, (int64_t)ctx->sm_count * blocks_per_sm(bs::k_pack_reads_uniform, bs::PACK_THREADS, 0)), bs::PACK_THREADS, 0, st, rs, ix);
            } else {
                BS_LAUNCH(bs::k_pack_reads, grid_for(N, kPackThreads, grid_cap), kPackThreads, 0, st, rs, ix);
            }
            ctx->launches++;
        }
    }
    return BS_OK;
}

int ChunkRun::place() {
    {
        StageTimer tm(ctx, ST_PLACE, st);
        if (N > 0 && C > 0 && !tile_mode) {
            // per-block scratch row of leftmost positions, all POS_INF between launches
            const size_t smem = bs::place_index_smem_bytes(kHitCap, kPlaceIxThreads);
            int place_per_sm = blocks_per_sm(bs::k_place_index<false>, kPlaceIxThreads, smem);
            if (const char *env = std::getenv("BS_PLACE_BLOCKS_PER_SM")) place_per_sm = std::max(1, std::min(place_per_sm, std::atoi(env)));  // tuning
            const int nblk = (int)std::min<int64_t>(C, (int64_t)ctx->sm_count * place_per_sm);
            const int64_t dense_stride = (max_seg_reads + 31) / 32 * 32;
            int64_t budget = (int64_t)4 << 30;  // bytes of placement scratch a full grid may take as dense rows
            if (const char *env = std::getenv("BS_PLACE_SCRATCH_MB")) budget = (int64_t)std::atoll(env) << 20;  // tests: 0 forces the hashed scratch
            const bool hashed = (int64_t)nblk * dense_stride * 8 > budget;
            bs::PlaceIxArgs pa;
            pa.order = d_order; pa.n_items = (int32_t)C; pa.work_counter = (int32_t *)ctx->d_counters.p;
            pa.ctg_off = d_ctg_off; pa.ctg_woff = d_ctg_woff; pa.ctg_words = cs.words; pa.ctg_mask = cs.mask;
            pa.ctg_chars = d_cchars; pa.ctg_seg = d_ctg_seg;
            pa.reads = rs; pa.ix = ix;
            pa.w = w_ptr; pa.total = total_ptr;
            pa.pos = o_pos; pa.pos_off = d_pos_off;
            pa.hit_cap = kHitCap; pa.keys = nullptr; pa.overflow = nullptr;
            if (!hashed) {
                const size_t need = (size_t)nblk * (size_t)dense_stride;
                if (need > ctx->best_elems || ctx->best_dirty) {
                    const size_t elems = std::max(need, ctx->best_elems);
                    BS_TRY(ensure(ctx, ctx->d_best, elems * 4));
                    ctx->best_elems = elems;
                    BS_CUDA(cudaMemsetAsync(ctx->d_best.p, 0x7f, elems * 4, st));
                }
                BS_TRY(ensure(ctx, ctx->d_hits_ovf, need * 4));
                ctx->best_dirty = true;  // cleared when the call ends without an error
                pa.best = (uint32_t *)ctx->d_best.p; pa.best_stride = dense_stride;
                pa.hits_ovf = (uint32_t *)ctx->d_hits_ovf.p;
                BS_CUDA(cudaFuncSetAttribute(bs::k_place_index<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                BS_LAUNCH(bs::k_place_index<false>, (unsigned)nblk, kPlaceIxThreads, smem, st, pa);
                ctx->launches++;
            } else {
                // one huge segment (cfg-5: 10^8 reads): a dense row per resident block does not fit, so every
                // block keeps (read -> leftmost position) in an open-addressed table sized for the reads ONE
                // contig places: four slots per position of the longest contig to start with.  A contig that
                // fills half of its table raises `overflow`; the launch is then repeated with twice the table
                // (ends at the latest when half the table holds every read of the segment).
                auto pow2_at_least = [](int64_t v) { int64_t c = 1; while (c < v) c <<= 1; return c; };
                const int64_t cap_min = 4 * (int64_t)kPlaceIxThreads, cap_max = pow2_at_least(2 * max_seg_reads + 4 * (int64_t)kPlaceIxThreads);
                int64_t cap = pow2_at_least(4 * max_ctg);
                if (const char *env = std::getenv("BS_PLACE_HASH_CAP")) cap = pow2_at_least(std::atoll(env));  // tests: a small table reaches the repeat
                cap = std::min(std::max(cap, cap_min), std::max(cap_min, cap_max));
                int32_t *d_overflow = (int32_t *)ctx->d_counters.p + 15;  // (the persistent kernels' work counters use 0..7)
                for (;;) {
                    const size_t elems = (size_t)nblk * (size_t)cap;
                    BS_TRY(ensure(ctx, ctx->d_best, std::max(elems, ctx->best_elems) * 4));
                    ctx->best_elems = std::max(elems, ctx->best_elems);
                    BS_TRY(ensure(ctx, ctx->d_hits_ovf, elems * 4));
                    BS_TRY(ensure(ctx, ctx->d_keys, elems * 4));
                    ctx->best_dirty = true;
                    BS_CUDA(cudaMemsetAsync(ctx->d_best.p, 0x7f, elems * 4, st));
                    BS_CUDA(cudaMemsetAsync(ctx->d_keys.p, 0, elems * 4, st));
                    BS_CUDA(cudaMemsetAsync(d_overflow, 0, 4, st));
                    pa.best = (uint32_t *)ctx->d_best.p; pa.best_stride = cap;
                    pa.hits_ovf = (uint32_t *)ctx->d_hits_ovf.p; pa.keys = (uint32_t *)ctx->d_keys.p; pa.overflow = d_overflow;
                    BS_CUDA(cudaFuncSetAttribute(bs::k_place_index<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    BS_LAUNCH(bs::k_place_index<true>, (unsigned)nblk, kPlaceIxThreads, smem, st, pa);
                    ctx->launches++;
                    int32_t overflow = 0;
                    BS_CUDA(cudaMemcpyAsync(&overflow, d_overflow, 4, cudaMemcpyDeviceToHost, st));
                    BS_CUDA(cudaStreamSynchronize(st));
                    if (!overflow) break;
                    if (cap >= cap_max) return fail(ctx, BS_ERR_STATE, "placement scratch overflow at full capacity");
                    cap *= 2;
                    // the abandoned launch added some contigs' weights already: start over
                    BS_CUDA(cudaMemsetAsync(w_ptr, 0, (size_t)w_elems * 4, st));
                    BS_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, 4, st));
                }
            }
        } else if (N > 0 && n_items > 0) {
            bs::PlaceArgs pa;
            pa.items = d_items;
            pa.ctg_off = d_ctg_off; pa.ctg_woff = d_ctg_woff;
            pa.ctg_words = cs.words; pa.ctg_mask = cs.mask; pa.ctg_chars = d_cchars;
            pa.reads = rs;
            pa.seg_seed_len = d_seed; pa.seg_read_start = d_seg_rs;
            pa.w = w_ptr; pa.total = total_ptr;
            pa.pos = o_pos; pa.pos_off = d_pos_off;
            pa.tile_len = tile_len; pa.hash_size = hash_size; pa.found_words = found_words;
            pa.scan_mode = (e.flags & BS_PLACE_SCAN) ? 1 : 0;
            BS_CUDA(cudaFuncSetAttribute(bs::k_place, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)place_smem));
            BS_LAUNCH(bs::k_place, (unsigned)n_items, kPlaceThreads, place_smem, st, pa);
            ctx->launches++;
        }
    }
    return BS_OK;
}

// Scaffold sets (bs_score_scaffolds): the reads are placed in the BASE contigs (k_place_index with the dense leftmost
// positions out), then every scaffold takes the minimum over its parts and probes its junction windows (k_place_compose),
// which also scores the break k-mers unless position weights are needed (second table).
bool ChunkRun::ks_compose() const {
    // scaffold set, first table, no path_prob_dist out, ranks fit 16 bits, the rank histogram fits shared memory
    const KsCache &k = ctx->ks;
    if (!e.comp || !ks_a || e.want_pd || kmer < 1 || kmer > bs::MAXK || k.R_x > 65535 || N <= 0 || n_base <= 0) return false;
    const char *env = std::getenv("BS_COMPOSE_KS");  // tests: 0 keeps k_prob_dist_ks on the composed texts
    if (env && env[0] == '0') return false;
    const bool packed = max_ctg - kmer + 1 < 65536;
    const size_t hist_bytes = (size_t)(bs::hist_phys_words(k.R_x, packed) + bs::hist_ranges(k.R_x, packed)) * 4;
    return hist_bytes + 8192 <= ctx->smem_optin;
}

int ChunkRun::place_composed() {
    StageTimer tm(ctx, ST_PLACE, st);
    if (C <= 0) return BS_OK;
    if (N > 0 && n_base > 0) {
        const size_t smem = bs::place_index_smem_bytes(kHitCap, kPlaceIxThreads);
        const int per_sm = blocks_per_sm(bs::k_place_index<false>, kPlaceIxThreads, smem);
        const int nblk = (int)std::min<int64_t>(n_ext, (int64_t)ctx->sm_count * per_sm);
        const int64_t dense_stride = (max_seg_reads + 31) / 32 * 32;
        bs::PlaceIxArgs pa;
        pa.order = d_base_order; pa.n_items = (int32_t)n_ext; pa.work_counter = (int32_t *)ctx->d_counters.p + 11;
        pa.ctg_off = d_base_off; pa.ctg_woff = d_base_woff; pa.ctg_words = bset.words; pa.ctg_mask = bset.mask;
        pa.ctg_chars = bset.chars; pa.ctg_seg = d_base_seg;
        pa.reads = rs; pa.ix = ix;
        pa.w = (int32_t *)ws.base_w.p; pa.total = (int32_t *)ws.base_total.p;
        pa.pos = (int32_t *)ws.base_pos.p; pa.pos_off = d_base_pos_off;
        pa.hit_cap = kHitCap; pa.keys = nullptr; pa.overflow = nullptr;
        const size_t need = (size_t)nblk * (size_t)dense_stride;
        if (need > ctx->best_elems || ctx->best_dirty) {
            const size_t elems = std::max(need, ctx->best_elems);
            BS_TRY(ensure(ctx, ctx->d_best, elems * 4));
            ctx->best_elems = elems;
            BS_CUDA(cudaMemsetAsync(ctx->d_best.p, 0x7f, elems * 4, st));
        }
        BS_TRY(ensure(ctx, ctx->d_hits_ovf, need * 4));
        ctx->best_dirty = true;  // cleared when the call ends without an error
        pa.best = (uint32_t *)ctx->d_best.p; pa.best_stride = dense_stride;
        pa.hits_ovf = (uint32_t *)ctx->d_hits_ovf.p;
        BS_CUDA(cudaFuncSetAttribute(bs::k_place_index<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BS_LAUNCH(bs::k_place_index<false>, (unsigned)nblk, kPlaceIxThreads, smem, st, pa);
        ctx->launches++;
    }
    if (n_base > 0) {  // per base contig, once: the list of reads placed in it and the break 8-mer of its interior positions
        bs::BaseSideArgs ba;
        ba.base_pos = (const int32_t *)ws.base_pos.p; ba.n_base = n_ext; ba.n_reads = N;
        ba.hits = (uint2 *)ws.base_hits.p; ba.hit_di = (int32_t *)ws.base_di.p; ba.cnt = (int32_t *)ws.base_cnt.p;
        ba.base_off = d_base_off; ba.base_woff = d_base_woff; ba.base_words = bset.words; ba.base_mask = bset.mask;
        ba.kmer = kmer;
        const bool ranks = ks_compose();
        ba.win = ranks ? (const bs::WinEntry *)ctx->ks.win.p : nullptr; ba.rank_zero = ctx->ks.rank_zero;
        ba.base_rank = ranks ? (uint16_t *)ws.base_rank.p : nullptr;
        BS_LAUNCH(bs::k_base_side, (unsigned)n_ext, kScoreThreads, 0, st, ba);
        ctx->launches++;
    }
    bs::PlaceComposeArgs ca;
    std::memset(&ca, 0, sizeof(ca));
    ca.p.order = d_order; ca.p.n_items = (int32_t)C; ca.p.work_counter = (int32_t *)ctx->d_counters.p + 12;
    ca.p.ctg_off = d_ctg_off; ca.p.ctg_woff = d_ctg_woff; ca.p.ctg_words = cs.words; ca.p.ctg_mask = cs.mask;
    ca.p.ctg_chars = d_cchars; ca.p.ctg_seg = d_ctg_seg;
    ca.p.reads = rs; ca.p.ix = ix;
    ca.p.w = w_ptr; ca.p.total = total_ptr; ca.p.pos = o_pos; ca.p.pos_off = d_pos_off;
    ca.sp = sparts; ca.part_jid = d_part_jid; ca.n_reads = N; ca.max_read_len = (int32_t)max_read_len;
    ca.base_hits = (const uint2 *)ws.base_hits.p; ca.base_cnt = (const int32_t *)ws.base_cnt.p; ca.base_hit_di = (const int32_t *)ws.base_di.p;
    // dynamic shared memory of a block: (scored here) the hash table of the break k-mers, sized for the reads one scaffold can
    // place so that KS-B never needs the global scratch rows; then the block's row of leftmost positions per read, while two
    // blocks per SM fit -- else the row lives in global memory (L2)
    // (an SM holds 228 KB of shared memory, 1 KB of it reserved per resident block; the kernel's static part goes off as well)
    const size_t static_smem = sizeof(bs::ScoreSharedCore) + 7 * (size_t)(2 * bs::COMPOSE_PART_CHUNK + 1) * 4 + 256;
    const size_t budget = ((size_t)228 * 1024) / 2 - 1024 - static_smem;
    int hash_slots = 0;
    if (e.comp_score) {
        hash_slots = bs::HASH_SLOTS;
        while (hash_slots < 32768 && (int64_t)hash_slots * 2 / 3 < std::min<int64_t>(N, 32767)) hash_slots <<= 1;
        while (hash_slots > bs::HASH_SLOTS && (size_t)hash_slots * 4 > budget) hash_slots >>= 1;
        if (const char *env = std::getenv("BS_COMPOSE_HASH_SLOTS")) {  // tests: a small table sends scaffolds through the global scratch rows
            hash_slots = 64;
            while (hash_slots < std::atoi(env) && hash_slots < 32768) hash_slots <<= 1;
        }
    }
    ca.hash_slots = hash_slots;
    const size_t hash_bytes = (size_t)hash_slots * 4;
    const size_t row_bytes = (size_t)std::max<int64_t>(N, 1) * 4;
    const char *rows_env = std::getenv("BS_COMPOSE_ROWS");  // tests: "global" keeps the rows out of shared memory
    const bool row_smem = hash_bytes + row_bytes <= budget && !(rows_env && rows_env[0] == 'g');
    const size_t smem = hash_bytes + (row_smem ? row_bytes : 0);
    auto launch = [&](auto kern) -> int {
        BS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int nblk = (int)std::min<int64_t>(C, (int64_t)ctx->sm_count * blocks_per_sm(kern, kComposeThreads, smem));
        if (!row_smem) {
            ca.row_stride = (N + 31) / 32 * 32;
            BS_TRY(ensure(ctx, ctx->d_hits_ovf, (size_t)nblk * (size_t)ca.row_stride * 4));  // (free again: the base placement is ahead on the same stream)
            ca.rows = (uint32_t *)ctx->d_hits_ovf.p;
        }
        if (e.comp_score) BS_TRY(score_args(0, nblk, ca.sc));
        BS_LAUNCH(kern, (unsigned)nblk, kComposeThreads, smem, st, ca);
        ctx->launches++;
        return BS_OK;
    };
    if (e.comp_score && row_smem) BS_TRY(launch(bs::k_place_compose<true, true>));
    else if (e.comp_score) BS_TRY(launch(bs::k_place_compose<true, false>));
    else if (row_smem) BS_TRY(launch(bs::k_place_compose<false, true>));
    else BS_TRY(launch(bs::k_place_compose<false, false>));
    return BS_OK;
}

int ChunkRun::spectrum() {
    if (e.want_ks) {
        const KsCache &k = ctx->ks;
        const int R_y = std::max(k.R_y, 1);
        BS_TRY(ensure(ctx, ws.ycnt, (size_t)S * R_y * 4));
        if (ks_a) BS_TRY(ensure(ctx, ws.yx, (size_t)S * k.R_x * sizeof(bs::LeLt)));
        StageTimer tm(ctx, ST_SPECTRUM, st);
        const size_t sp_smem = (size_t)((k.R_y + 1) / 2) * 4;
        bs::SpectrumArgs sp;
        sp.tr_off = d_tr_off; sp.tr_woff = d_tr_woff; sp.tr_words = ts.words; sp.tr_mask = ts.mask;
        sp.rank_y = (const int32_t *)k.rank_y.p; sp.ycnt = (int32_t *)ws.ycnt.p;
        sp.R_y = R_y; sp.kmer = kmer; sp.blocks_per_seg = 1;
        sp.lelt = ks_a ? (const bs::LeLt *)k.lelt.p : nullptr; sp.yx = ks_a ? (bs::LeLt *)ws.yx.p : nullptr; sp.R_x = k.R_x;
        if (k.R_y > 0 && kmer <= bs::MAXK && max_tr - kmer + 1 < 65536 && sp_smem + 1024 <= ctx->smem_optin) {
            // histogram and prefix sum in shared memory; with the rank table beside them when it fits
            // (ranks below 65 535, k-mers up to 8): one block per SM walking over the segments
            const size_t tab_smem = sp_smem + ((size_t)2 << (2 * kmer));
            const char *tab_env = std::getenv("BS_SPECTRUM_TABLE");  // tests: 0 keeps the rank table in global memory
            if (kmer <= 8 && k.R_y < 65535 && tab_smem + 1024 <= ctx->smem_optin && !(tab_env && tab_env[0] == '0')) {
                BS_CUDA(cudaFuncSetAttribute(bs::k_truth_spectrum_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_smem));
                BS_LAUNCH(bs::k_truth_spectrum_smem<true>, (unsigned)std::min<int64_t>(S, ctx->sm_count), kSpectrumTabThreads, tab_smem, st, sp, (int)S);
            } else {
                BS_CUDA(cudaFuncSetAttribute(bs::k_truth_spectrum_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp_smem));
                BS_LAUNCH(bs::k_truth_spectrum_smem<false>, (unsigned)S, kSpectrumThreads, sp_smem, st, sp, (int)S);
            }
            ctx->launches++;
        } else {
            BS_CUDA(cudaMemsetAsync(ws.ycnt.p, 0, (size_t)S * R_y * 4, st));
            if (k.R_y > 0) {
                // enough blocks for the whole machine whatever the number of segments (cfg-5: ONE truth of 100 Mb)
                const int64_t want_blocks = ((int64_t)ctx->sm_count * 8 + S - 1) / std::max<int64_t>(S, 1);
                sp.blocks_per_seg = (int)std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(64, want_blocks), (max_tr + kScoreThreads * 8 - 1) / (kScoreThreads * 8)));
                BS_LAUNCH(bs::k_truth_spectrum, (unsigned)(S * sp.blocks_per_seg), kScoreThreads, 0, st, sp);
                ctx->launches++;
                BS_LAUNCH(bs::k_row_cumsum, (unsigned)S, kScoreThreads, kScoreThreads * 8 + 16, st, (int32_t *)ws.ycnt.p, R_y);
                ctx->launches++;
            }
            if (ks_a) {
                BS_LAUNCH(bs::k_yx_gather, grid_for(S * (int64_t)k.R_x, kScoreThreads, grid_cap), kScoreThreads, 0, st, sp, S);
                ctx->launches++;
            }
        }
    }
    return BS_OK;
}

bool ChunkRun::fused(int which) const {
    const bool ksa = which ? o_ksa2 != nullptr : ks_a;
    if (which == 0 && ks_compose()) return false;  // (k_ks_compose takes the stage: nothing scores on the way)
    // contigs of at least bs::FUSE_MIN_LEN bases are then scored by the long-contig KS-A kernel, the others by k_break_score
    const char *env = std::getenv("BS_FUSE_SCORE");  // tests / tuning: 0 keeps k_break_score for every contig
    return kmer == 8 && !e.want_hist && !e.comp_score && (e.want_pd || ksa) && !(env && env[0] == '0');
}

int ChunkRun::score_args(int which, int64_t rows, bs::ScoreArgs &sa) {
    const KsCache &k = which ? ctx->ks2 : ctx->ks;
    const bool ksb = which ? o_ksb2 != nullptr : ks_b;
    std::memset(&sa, 0, sizeof(sa));
    sa.order = d_order; sa.work_counter = (int32_t *)ctx->d_counters.p + (which ? 5 : 2);
    sa.ctg_off = d_ctg_off; sa.ctg_woff = d_ctg_woff; sa.ctg_words = cs.words; sa.ctg_mask = cs.mask; sa.ctg_seg = d_ctg_seg;
    sa.w = w_ptr; sa.total = total_ptr;
    sa.tab = (const bs::TabEntry *)(which ? ctx->d_tab2.p : ctx->d_tab.p);
    sa.kmer = kmer; sa.T = (int32_t)T; sa.n_contigs = C;
    if (which == 0) {
        sa.sequence_len = o_len; sa.bp_score = o_score; sa.norm_by_break_freqs = o_norm; sa.norm_by_len = o_bylen;
        sa.kmer_breaks = o_breaks; sa.hist = o_hist;
    } else {  // lengths, break counts and the histogram do not depend on the table
        sa.bp_score = o_score2; sa.norm_by_break_freqs = o_norm2; sa.norm_by_len = o_bylen2;
    }
    if (ksb) {
        rows = std::max<int64_t>(rows, (int64_t)ctx->sm_count * 8);  // (k_break_score and the fused KS-A kernel take turns on the same rows)
        const size_t scratch_bytes = (size_t)rows * (T + 1) * 4;
        if (ctx->d_scratch.cap < scratch_bytes || !ctx->d_scratch.p) {
            BS_TRY(ensure(ctx, ctx->d_scratch, scratch_bytes));
            BS_CUDA(cudaMemsetAsync(ctx->d_scratch.p, 0, ctx->d_scratch.cap, st));
        }
        BS_TRY(ensure(ctx, ctx->d_ovf, (size_t)rows * bs::OVF_CAP * 4));
        sa.ks_b = which ? o_ksb2 : o_ksb; sa.yv = (const double *)k.yv.p; sa.ycum = (const int32_t *)ws.ycnt.p; sa.R_y = k.R_y;
        sa.zero_le = k.zero_le; sa.zero_lt = k.zero_lt; sa.y_max = k.y_max;
        sa.scratch = (int32_t *)ctx->d_scratch.p; sa.ovf_cnt = (int32_t *)ctx->d_ovf.p; sa.status = (int32_t *)ctx->d_status.p;
    }
    return BS_OK;
}

int ChunkRun::score(int which) {
    if (e.comp_score) return BS_OK;  // done inside k_place_compose
    const bool fz = fused(which);  // the long-contig KS-A kernel scores contigs of FUSE_MIN_LEN bases and more on its way
    const int64_t n = fz ? n_score : C;
    if (n > 0) {
        // scores (+ histogram, + KS of the normalised break histogram); after the truth spectrum
        StageTimer tm(ctx, ST_SCORE, st);
        const int nblk = (int)std::min<int64_t>(n, (int64_t)ctx->sm_count * blocks_per_sm(bs::k_break_score, kBreakScoreThreads, 0));
        bs::ScoreArgs sa;
        BS_TRY(score_args(which, nblk, sa));
        if (fz) { sa.order = d_order_score; sa.n_contigs = n_score; }
        BS_LAUNCH(bs::k_break_score, (unsigned)nblk, kBreakScoreThreads, 0, st, sa);
        ctx->launches++;
    }
    return BS_OK;
}

int ChunkRun::prob_dist(int which) {
    const bool ksa = which ? o_ksa2 != nullptr : ks_a;
    if (e.want_pd || ksa) {
        StageTimer tm(ctx, ST_PROBDIST, st);
        const KsCache &k = which ? ctx->ks2 : ctx->ks;
        bs::ProbDistArgs pa;
        pa.order = d_order; pa.work_counter = (int32_t *)ctx->d_counters.p + (which ? 4 : 1);
        pa.ctg_off = d_ctg_off; pa.ctg_woff = d_ctg_woff; pa.ctg_words = cs.words; pa.ctg_mask = cs.mask;
        pa.ctg_seg = d_ctg_seg;
        pa.win = (kmer >= 1 && kmer <= bs::MAXK) ? (const bs::WinEntry *)k.win.p : nullptr;
        pa.yx = ksa ? (const bs::LeLt *)(which ? ws.yx2.p : ws.yx.p) : nullptr;
        pa.ycum = ksa ? (const int32_t *)ws.ycnt.p : nullptr;
        pa.R_x = k.R_x; pa.R_y = k.R_y; pa.rank_zero = k.rank_zero;
        pa.kmer = kmer; pa.n_contigs = C;
        pa.prob_dist = which ? o_pd2 : o_pd; pa.pd_off = d_pd_off; pa.ks = ksa ? (which ? o_ksa2 : o_ksa) : nullptr;
        // rank histogram in shared memory: 16-bit counters (two per word) when no contig has 65 536
        // windows -- real table: 32 897 ranks = 66 KB, three blocks per SM -- else 32-bit counters,
        // else (all-distinct tables with long contigs) a per-block global scratch that stays in L2
        const bool packed = max_ctg - kmer + 1 < 65536;
        pa.hist_words = ksa ? bs::hist_phys_words(k.R_x, packed) : 0;
        pa.n_ranges = ksa ? bs::hist_ranges(k.R_x, packed) : 0;
        const size_t hist_bytes = (size_t)(pa.hist_words + pa.n_ranges) * 4;
        const bool in_smem = hist_bytes + 2048 <= ctx->smem_optin;
        const size_t smem = in_smem ? hist_bytes : 0;
        const int per_sm = std::max<int>(1, std::min<int>(2048 / kKsThreads, (int)((ctx->smem_optin + 1024) / (smem + 1024))));
        const int nblk = (int)std::min<int64_t>(C, (int64_t)ctx->sm_count * per_sm);
        pa.rank_scratch = nullptr;
        if (ksa && !in_smem) {
            BS_TRY(ensure(ctx, ctx->d_rank_scratch, (size_t)nblk * hist_bytes));
            pa.rank_scratch = (uint32_t *)ctx->d_rank_scratch.p;
        }
        // short contigs (<= KS_SMALL_MAX windows) sort their ranks by ranges in a small shared-memory
        // footprint; long ones keep the table-wide rank histogram
        const size_t small_smem = bs::ks_small_smem_bytes(k.R_x);
        const bool use_small = ksa && n_small > 0 && small_smem + 1024 <= ctx->smem_optin;
        const bool fuse = fused(which);
        auto small_kernel = bs::k_prob_dist_ks_small;
        int nb_small = 0;
        if (use_small) {
            BS_CUDA(cudaFuncSetAttribute(small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem));
            nb_small = (int)std::min<int64_t>(n_small, (int64_t)ctx->sm_count * blocks_per_sm(small_kernel, bs::KS_SMALL_THREADS, small_smem));
        }
        std::memset(&pa.sc, 0, sizeof(pa.sc));
        pa.fuse_min_len = fuse_min_len;
        if (which == 0 && ks_compose()) {
            bs::KsComposeArgs ka;
            ka.pd = pa; ka.pd.prob_dist = nullptr;
            ka.sp = sparts; ka.base_rank = (const uint16_t *)ws.base_rank.p; ka.base_off = d_base_off;
            auto launch_kc = [&](auto kern) -> int {
                BS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes));
                const int nb = (int)std::min<int64_t>(C, (int64_t)ctx->sm_count * blocks_per_sm(kern, kKsThreads, hist_bytes));
                BS_LAUNCH(kern, (unsigned)nb, kKsThreads, hist_bytes, st, ka);
                ctx->launches++;
                return BS_OK;
            };
            return packed ? launch_kc(bs::k_ks_compose<true>) : launch_kc(bs::k_ks_compose<false>);
        }
        if (fuse) BS_TRY(score_args(which, std::max(nb_small, nblk), pa.sc));  // (a scratch row per block of the wider launch)
        if (use_small) {
            bs::ProbDistArgs ps = pa;
            ps.order = d_order_small; ps.n_contigs = n_small; ps.work_counter = (int32_t *)ctx->d_counters.p + (which ? 7 : 6);
            BS_LAUNCH(small_kernel, (unsigned)nb_small, bs::KS_SMALL_THREADS, small_smem, st, ps);
            ctx->launches++;
            pa.order = d_order_large; pa.n_contigs = n_large;
        }
        auto launch = [&](auto kern) -> int {
            BS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            BS_LAUNCH(kern, (unsigned)std::min<int64_t>(nblk, std::max<int64_t>(pa.n_contigs, 1)), kKsThreads, smem, st, pa);
            return BS_OK;
        };
        if (pa.n_contigs > 0) {
            if (fuse) {
                if (packed && in_smem) BS_TRY(launch(bs::k_prob_dist_ks<true, true, true>));
                else if (packed) BS_TRY(launch(bs::k_prob_dist_ks<true, false, true>));
                else if (in_smem) BS_TRY(launch(bs::k_prob_dist_ks<false, true, true>));
                else BS_TRY(launch(bs::k_prob_dist_ks<false, false, true>));
            } else {
                if (packed && in_smem) BS_TRY(launch(bs::k_prob_dist_ks<true, true, false>));
                else if (packed) BS_TRY(launch(bs::k_prob_dist_ks<true, false, false>));
                else if (in_smem) BS_TRY(launch(bs::k_prob_dist_ks<false, true, false>));
                else BS_TRY(launch(bs::k_prob_dist_ks<false, false, false>));
            }
            ctx->launches++;
        }
    }
    return BS_OK;
}

// The second scoring table over the same rows (the R driver's "random" pass): the placement, the break
// counts, the contig offsets and the truth's cumulative counts are shared; only the table-dependent
// stages run again.
int ChunkRun::second_table() {
    if (!e.second) return BS_OK;
    if (o_ksa2) {  // counts <= / < every x value of the SECOND table, from the truth's cumulative counts
        const KsCache &k2 = ctx->ks2;
        BS_TRY(ensure(ctx, ws.yx2, (size_t)S * k2.R_x * sizeof(bs::LeLt)));
        bs::SpectrumArgs sp;
        std::memset(&sp, 0, sizeof(sp));
        sp.ycnt = (int32_t *)ws.ycnt.p; sp.R_y = std::max(ctx->ks.R_y, 1);
        sp.lelt = (const bs::LeLt *)k2.lelt.p; sp.yx = (bs::LeLt *)ws.yx2.p; sp.R_x = k2.R_x;
        StageTimer tm(ctx, ST_SPECTRUM, st);
        BS_LAUNCH(bs::k_yx_gather, grid_for(S * (int64_t)k2.R_x, kScoreThreads, grid_cap), kScoreThreads, 0, st, sp, S);
        ctx->launches++;
    }
    BS_TRY(score(1));
    return prob_dist(1);
}

int ChunkRun::startpos() {
    if (e.want_sp || e.want_lev) {
        StageTimer tm(ctx, ST_STARTPOS, st);
        BS_TRY(ensure(ctx, ws.spbest, (size_t)C * 4));
        BS_CUDA(cudaMemsetAsync(ws.spbest.p, 0x7f, (size_t)C * 4, st));
        if (e.want_lev) BS_TRY(ensure(ctx, ws.exact, (size_t)C * 4));
        bs::StartposArgs sa;
        sa.ctg_off = d_ctg_off; sa.ctg_woff = d_ctg_woff; sa.ctg_words = cs.words; sa.ctg_mask = cs.mask; sa.ctg_chars = d_cchars;
        sa.ctg_seg = d_ctg_seg; sa.seg_contig_start = d_seg_cs;
        sa.tr_off = d_tr_off; sa.tr_woff = d_tr_woff; sa.tr_words = ts.words; sa.tr_mask = ts.mask;
        sa.tr_chars = d_tchars; sa.total = total_ptr; sa.n_contigs = C; sa.n_seg = (int32_t)S;
        sa.best = (uint32_t *)ws.spbest.p; sa.startpos = o_startpos;
        sa.exact = e.want_lev ? (int32_t *)ws.exact.p : nullptr; sa.search_all = e.want_lev ? 1 : 0;
        // seed tables + prefix bitmaps of the contig groups, built in their own kernel
        sa.seg_tab_off = d_sp_tab_off; sa.seg_tab_mask = d_sp_tab_mask; sa.big = sp_big ? 1 : 0;
        BS_TRY(ensure(ctx, ws.sp_key, (size_t)sp_slots * 8));
        BS_TRY(ensure(ctx, ws.sp_head, (size_t)sp_slots * 4));
        BS_TRY(ensure(ctx, ws.sp_next, (size_t)C * 4));
        BS_TRY(ensure(ctx, ws.sp_bitmap, (size_t)sp_bitmap_words * 4));
        int64_t q_cap = std::max<int64_t>(4 * C, 1 << 16);
        if (const char *env = std::getenv("BS_STARTPOS_QCAP")) q_cap = std::max<int64_t>(1, std::atoll(env));  // tests: a tiny queue reaches the in-scan fallback
        BS_TRY(ensure(ctx, ws.sp_queue, (size_t)q_cap * 8));
        BS_CUDA(cudaMemsetAsync(ws.sp_key.p, 0xff, (size_t)sp_slots * 8, st));
        BS_CUDA(cudaMemsetAsync(ws.sp_head.p, 0, (size_t)sp_slots * 4, st));
        BS_CUDA(cudaMemsetAsync(ws.sp_bitmap.p, 0, (size_t)sp_bitmap_words * 4, st));
        sa.tab_key = (unsigned long long *)ws.sp_key.p; sa.tab_head = (int32_t *)ws.sp_head.p; sa.tab_next = (int32_t *)ws.sp_next.p;
        sa.bitmap = (uint32_t *)ws.sp_bitmap.p;
        sa.queue = (uint2 *)ws.sp_queue.p; sa.q_count = (int32_t *)ctx->d_counters.p + 8; sa.q_cap = (int32_t)std::min<int64_t>(q_cap, 0x7fffffff);
        BS_LAUNCH(bs::k_startpos_build, grid_for(C, kStartposThreads, grid_cap), kStartposThreads, 0, st, sa);
        ctx->launches++;
        // the truth streams past the bitmaps: enough blocks to fill the machine even for a single segment and to keep
        // the last wave short, at least 2048 positions each (BIG: one fat block per SM and a few waves, every block
        // copies 128 KB of bitmap first)
        if (!sp_big) {
            int64_t splits = ((int64_t)ctx->sm_count * 16 + S - 1) / std::max<int64_t>(S, 1);
            splits = std::max<int64_t>(1, std::min<int64_t>(splits, (max_tr + 2047) / 2048));
            sa.splits = (int32_t)splits;
            BS_LAUNCH(bs::k_startpos_scan<false>, (unsigned)(S * splits), kStartposThreads, 0, st, sa);
        } else {
            int64_t splits = ((int64_t)ctx->sm_count * 4 + S - 1) / std::max<int64_t>(S, 1);
            splits = std::max<int64_t>(1, std::min<int64_t>(splits, (max_tr + 65535) / 65536));
            sa.splits = (int32_t)splits;
            const size_t bm_smem = (size_t)bs::SP_BITMAP_WORDS_BIG * 4;
            BS_CUDA(cudaFuncSetAttribute(bs::k_startpos_scan<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bm_smem));
            BS_LAUNCH(bs::k_startpos_scan<true>, (unsigned)(S * splits), kStartposBigThreads, bm_smem, st, sa);
        }
        ctx->launches++;
        // queued (contig, position) candidates: one warp each
        BS_LAUNCH(bs::k_startpos_verify, (unsigned)std::min<int64_t>((q_cap * 32 + kStartposThreads - 1) / kStartposThreads, (int64_t)ctx->sm_count * 8),
                  kStartposThreads, 0, st, sa);
        ctx->launches++;
        BS_LAUNCH(bs::k_startpos, (unsigned)std::min<int64_t>(C, grid_cap), kStartposThreads, 0, st, sa);
        ctx->launches++;
    }
    return BS_OK;
}

int ChunkRun::lev() {
    if (e.want_lev) {
        // infix edit distance contig vs truth: one warp per contig, longest first
        StageTimer tm(ctx, ST_LEV, st);
        const int warps_per_block = kLevThreads / 32;
        int nblk = (int)std::min<int64_t>((C + warps_per_block - 1) / warps_per_block,
                                          (int64_t)ctx->sm_count * blocks_per_sm(bs::k_lev_infix, kLevThreads, 0));
        bs::LevArgs la;
        la.order = d_order; la.work_counter = (int32_t *)ctx->d_counters.p + 3;
        la.ctg_off = d_ctg_off; la.ctg_woff = d_ctg_woff; la.ctg_words = cs.words; la.ctg_mask = cs.mask; la.ctg_chars = d_cchars;
        la.ctg_seg = d_ctg_seg; la.tr_off = d_tr_off; la.tr_woff = d_tr_woff; la.tr_words = ts.words; la.tr_mask = ts.mask;
        la.tr_chars = d_tchars; la.exact = (const int32_t *)ws.exact.p; la.n_contigs = C; la.lev = o_lev;
        la.hbuf = nullptr; la.hbuf_stride = 0;
        if (max_ctg > 32 * 64) {  // chunk-boundary deltas of contigs longer than one warp of blocks
            const int64_t stride = (max_tr + 63) / 64 * 64;
            const int64_t budget = (int64_t)4 << 30;
            nblk = (int)std::max<int64_t>(1, std::min<int64_t>(nblk, budget / std::max<int64_t>(stride * warps_per_block, 1)));
            BS_TRY(ensure(ctx, ctx->d_hbuf, (size_t)nblk * warps_per_block * stride));
            la.hbuf = (int8_t *)ctx->d_hbuf.p; la.hbuf_stride = stride;
        }
        BS_LAUNCH(bs::k_lev_bound, (unsigned)grid_for(C * 32, kLevThreads, grid_cap), kLevThreads, 0, st, la);
        ctx->launches++;
        BS_LAUNCH(bs::k_lev_infix, (unsigned)nblk, kLevThreads, 0, st, la);
        ctx->launches++;
    }
    return BS_OK;
}

int ChunkRun::results() {
    BS_CUDA(cudaGetLastError());
    BS_CUDA(cudaEventRecord(ws.ev_compute, st));

    // ---------------- results ----------------
    if (e.w_out) {  // phase 1 of a two-phase call: the weights already sit in the caller's device arrays
        BS_CUDA(cudaEventRecord(ws.ev_d2h, st));
        ws.in_flight = true;
        return BS_OK;
    }
    if (e.dev_res) {
        if (res->lev_dist_vs_true && !e.want_lev) BS_CUDA(cudaMemsetAsync(res->lev_dist_vs_true + ch.c0, 0, (size_t)C * 4, st));
        if (!e.want_sp && res->path_prob_dist_startpos) BS_CUDA(cudaMemsetAsync(res->path_prob_dist_startpos + ch.c0, 0, (size_t)C * 4, st));
        BS_CUDA(cudaEventRecord(ws.ev_d2h, st));
        ws.in_flight = true;
        return BS_OK;
    }
    {
        cudaStream_t os = ctx->out_stream;
        BS_CUDA(cudaStreamWaitEvent(os, ws.ev_compute, 0));
        StageTimer tm(ctx, ST_D2H, os);
        auto d2h = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
            if (!dst || !src || !bytes) return cudaSuccess;
            return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, os);
        };
        auto at = [](auto *p, int64_t off) -> void * { return p ? (void *)(p + off) : nullptr; };
        BS_CUDA(d2h(at(res->sequence_len, ch.c0), o_len, (size_t)C * 4));
        BS_CUDA(d2h(at(res->kmer_breaks, ch.c0), o_breaks, (size_t)C * 4));
        BS_CUDA(d2h(at(res->path_prob_dist_startpos, ch.c0), o_startpos, (size_t)C * 4));
        BS_CUDA(d2h(at(res->lev_dist_vs_true, ch.c0), o_lev, (size_t)C * 4));
        BS_CUDA(d2h(at(res->bp_score, ch.c0), o_score, (size_t)C * 8));
        BS_CUDA(d2h(at(res->bp_score_norm_by_break_freqs, ch.c0), o_norm, (size_t)C * 8));
        BS_CUDA(d2h(at(res->bp_score_norm_by_len, ch.c0), o_bylen, (size_t)C * 8));
        BS_CUDA(d2h(at(res->ks_stat_prob_dist, ch.c0), o_ksa, (size_t)C * 8));
        BS_CUDA(d2h(at(res->ks_stat_path_freq, ch.c0), o_ksb, (size_t)C * 8));
        if (e.want_pd) BS_CUDA(d2h(at(res->path_prob_dist, res->path_prob_dist_off[ch.c0]), o_pd, (size_t)pd_elems * 8));
        if (e.second) {
            BS_CUDA(d2h(at(res->bp_score2, ch.c0), o_score2, (size_t)C * 8));
            BS_CUDA(d2h(at(res->bp_score_norm_by_break_freqs2, ch.c0), o_norm2, (size_t)C * 8));
            BS_CUDA(d2h(at(res->bp_score_norm_by_len2, ch.c0), o_bylen2, (size_t)C * 8));
            BS_CUDA(d2h(at(res->ks_stat_prob_dist2, ch.c0), o_ksa2, (size_t)C * 8));
            BS_CUDA(d2h(at(res->ks_stat_path_freq2, ch.c0), o_ksb2, (size_t)C * 8));
            if (e.want_pd) BS_CUDA(d2h(at(res->path_prob_dist2, res->path_prob_dist_off[ch.c0]), o_pd2, (size_t)pd_elems * 8));
        }
        if (e.want_hist) BS_CUDA(d2h(at(res->hist, ch.c0 * (T + 1)), o_hist, (size_t)C * (T + 1) * 4));
        if (e.want_pos) BS_CUDA(d2h(at(res->pos, res->pos_off[ch.c0]), o_pos, (size_t)pos_elems * 4));
    }
    BS_CUDA(cudaEventRecord(ws.ev_d2h, ctx->out_stream));
    ws.in_flight = true;
    return BS_OK;
}

// ---- one chunk: metadata, H2D, kernels, D2H --------------------------------------------------
// BS_TRACE=1: host wall time of every stage's queueing code on stderr (diagnostic; the device times are the
// CUDA-event stage timers of bs_ctx_last_timings)
struct HostTrace {
    bool on;
    std::chrono::steady_clock::time_point t;
    HostTrace() : on(std::getenv("BS_TRACE") != nullptr), t(std::chrono::steady_clock::now()) {}
    void lap(const char *what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[bs trace] %-12s %9.3f ms host\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

struct NvtxScope {
    explicit NvtxScope(const char *name) { BS_NVTX_PUSH(name); }
    ~NvtxScope() { BS_NVTX_POP(); }
};

int run_chunk(bs_ctx *ctx, Workspace &ws, const CallEnv &e, const Chunk &ch) {
    NvtxScope range("bs:chunk");
    ChunkRun r(ctx, ws, e, ch);
    HostTrace tr;
    BS_TRY(r.prepare());
    tr.lap("prepare");
    BS_TRY(r.pack());
    tr.lap("pack");
    if (e.comp_score) BS_TRY(r.spectrum());  // (the compositional placement scores on its way: KS-B needs the truth side first)
    if (!e.w_in) BS_TRY(e.comp ? r.place_composed() : r.place());
    tr.lap("place");
    if (e.w_out) return r.results();  // phase 1 of a two-phase call: the weights are the result
    if (!e.comp_score) BS_TRY(r.spectrum());
    BS_TRY(r.score(0));
    BS_TRY(r.prob_dist(0));
    BS_TRY(r.second_table());
    tr.lap("score+ks");
    BS_TRY(r.startpos());
    BS_TRY(r.lev());
    tr.lap("startpos+lev");
    const int rc = r.results();
    tr.lap("results");
    return rc;
}

}  // namespace

extern "C" {

int bs_abi_version(void) { return BS_ABI_VERSION; }

const char *bs_last_error(const bs_ctx *ctx) { return ctx ? ctx->err : g_create_error; }

int bs_ctx_create(int device, bs_ctx **out) {
    bs_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, BS_ERR_INVALID, "bs_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(nullptr, BS_ERR_NO_DEVICE, "no usable CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev)
        return fail(nullptr, BS_ERR_NO_DEVICE, "device %d out of range (found %d)", device, ndev);
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, BS_ERR_CUDA, "cannot open device %d: %s", device, cudaGetErrorString(e));
#ifndef BS_CPU_EMUL
    if (prop.major != 10)
        return fail(nullptr, BS_ERR_NO_DEVICE, "device %d (%s) is sm_%d%d; this build holds sm_100a code only",
                    device, prop.name, prop.major, prop.minor);
#endif
    ctx = new (std::nothrow) bs_ctx();
    if (!ctx) return fail(nullptr, BS_ERR_ALLOC, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->out_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < kWorkspaces; i++)
        ok = cudaEventCreateWithFlags(&ctx->ws[i].ev_h2d, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ws[i].ev_compute, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ws[i].ev_d2h, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        fail(nullptr, BS_ERR_CUDA, "stream/event creation failed");
        bs_ctx_destroy(ctx);
        return BS_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    if (const char *env = std::getenv("BS_CHUNK_KB")) {  // tuning / tests: ASCII bytes per pipeline chunk
        const long kb = std::atol(env);
        if (kb > 0) ctx->chunk_bytes_host = ctx->chunk_bytes_dev = (int64_t)kb << 10;
    }
    *out = ctx;
    return BS_OK;
}

void bs_ctx_destroy(bs_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    DevBuf *bufs[] = {&ctx->d_tab, &ctx->d_tab2, &ctx->ks2.win, &ctx->ks2.rank_y, &ctx->ks2.lelt, &ctx->ks2.yv, &ctx->ks.win, &ctx->ks.rank_y, &ctx->ks.lelt, &ctx->ks.yv, &ctx->d_best, &ctx->d_hits_ovf, &ctx->d_keys, &ctx->d_scratch, &ctx->d_ovf, &ctx->d_status,
                      &ctx->d_rank_scratch, &ctx->d_counters, &ctx->d_hbuf, &ctx->sim_meta, &ctx->sim_chars,
                      &ctx->sim_words, &ctx->sim_mask, &ctx->sim_cdf, &ctx->sim_starts, &ctx->sim_kept, &ctx->sim_reads};
    for (DevBuf *b : bufs) release(*b);
    for (Workspace &w : ctx->ws) {
        DevBuf *wb[] = {&w.meta, &w.read_chars, &w.read_off, &w.ctg_chars, &w.tr_chars, &w.rwords, &w.rflags, &w.cwords,
                        &w.cmask, &w.twords, &w.tmask, &w.w, &w.total, &w.ycnt, &w.yx, &w.yx2, &w.head, &w.next, &w.odd_head, &w.spbest, &w.exact, &w.sp_key, &w.sp_head, &w.sp_next, &w.sp_bitmap, &w.sp_queue,
                        &w.out_i32, &w.out_f64, &w.pd, &w.pd2, &w.hist, &w.pos,
                        &w.base_chars, &w.base_words, &w.base_mask, &w.base_pos, &w.base_w, &w.base_total, &w.base_hits, &w.base_cnt, &w.base_di, &w.base_rank};
        for (DevBuf *b : wb) release(*b);
        if (w.h_meta) cudaFreeHost(w.h_meta);
        if (w.ev_h2d) cudaEventDestroy(w.ev_h2d);
        if (w.ev_compute) cudaEventDestroy(w.ev_compute);
        if (w.ev_d2h) cudaEventDestroy(w.ev_d2h);
    }
    for (TimedSpan &t : ctx->spans) {
        if (t.a) cudaEventDestroy(t.a);
        if (t.b) cudaEventDestroy(t.b);
    }
    if (ctx->h_multi) cudaFreeHost(ctx->h_multi);
    release(ctx->m_reads);
    release(ctx->m_ctgs);
    release(ctx->m_truth);
    if (ctx->ev_share) cudaEventDestroy(ctx->ev_share);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->out_stream) cudaStreamDestroy(ctx->out_stream);
    delete ctx;
}

int bs_ctx_set_stream(bs_ctx *ctx, void *cuda_stream) {
    if (!ctx) return BS_ERR_INVALID;
    cudaSetDevice(ctx->device);
    BS_TRY(sync_all(ctx));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return BS_OK;
}

int bs_ctx_synchronize(bs_ctx *ctx) {
    if (!ctx) return BS_ERR_INVALID;
    cudaSetDevice(ctx->device);
    BS_TRY(sync_all(ctx));
    if (ctx->status_pending) {  // KS-B tally overflow of an asynchronous (device-result) call: report it here
        int32_t status = 0;
        ctx->status_pending = false;
        BS_CUDA(cudaMemcpyAsync(&status, ctx->d_status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        BS_CUDA(cudaStreamSynchronize(ctx->stream));
        if (status) return fail(ctx, BS_ERR_INVALID, "ks_stat_path_freq of an earlier device-result call: more than %d table rows with a count >= %d in one contig", bs::OVF_CAP, bs::CC_DENSE);
    }
    return BS_OK;
}

int64_t bs_ctx_launch_count(const bs_ctx *ctx) { return ctx ? ctx->launches : 0; }

int bs_ctx_set_poll(bs_ctx *ctx, bs_poll_fn poll, void *user) {
    if (!ctx) return BS_ERR_INVALID;
    ctx->poll = poll;
    ctx->poll_user = poll ? user : nullptr;
    return BS_OK;
}

int bs_ctx_enable_timing(bs_ctx *ctx, int on) {
    if (!ctx) return BS_ERR_INVALID;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    ctx->timing = on != 0;
    ctx->spans_used = 0;
    return BS_OK;
}

// per-stage device time (ms) summed over the chunks of every scoring call since timing was enabled
// or last read (reading resets the sums; no synchronisation is forced between calls, so a caller
// can time K back-to-back calls and divide); order: h2d, pack, place, score, spectrum,
// prob_dist_ks, ks_path_freq, startpos, d2h, lev.  Returns the number written.
int bs_ctx_last_timings(bs_ctx *ctx, double *ms, int n) {
    if (!ctx || !ms) return 0;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    double acc[ST_COUNT];
    bool used[ST_COUNT] = {};
    for (int i = 0; i < ST_COUNT; i++) acc[i] = 0.0;
    for (size_t i = 0; i < ctx->spans_used; i++) {
        float t = 0.f;
        const TimedSpan &sp = ctx->spans[i];
        if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) { acc[sp.stage] += t; used[sp.stage] = true; }
    }
    int k = 0;
    for (; k < n && k < ST_COUNT; k++) ms[k] = (ctx->timing && used[k]) ? acc[k] : -1.0;
    ctx->spans_used = 0;
    return k;
}

double bs_ctx_last_place_ms(bs_ctx *ctx) {
    double ms[ST_COUNT];
    if (bs_ctx_last_timings(ctx, ms, ST_COUNT) <= ST_PLACE) return -1.0;
    return ms[ST_PLACE];
}

void *bs_host_alloc(int64_t bytes) {
    void *p = nullptr;
    if (bytes < 0) return nullptr;
    if (cudaHostAlloc(&p, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void bs_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

static int bs_set_table_impl(bs_ctx *ctx, const char *kmer_chars, const int64_t *kmer_off, const double *prob, int64_t n);

int bs_set_table(bs_ctx *ctx, const char *kmer_chars, const int64_t *kmer_off, const double *prob, int64_t n) {
    if (!ctx) return BS_ERR_INVALID;
    try {
        return bs_set_table_impl(ctx, kmer_chars, kmer_off, prob, n);
    } catch (const std::bad_alloc &) {
        return fail(ctx, BS_ERR_ALLOC, "bs_set_table: out of host memory");
    } catch (...) {
        return fail(ctx, BS_ERR_STATE, "bs_set_table: unexpected exception");
    }
}

static int bs_set_table_impl(bs_ctx *ctx, const char *kmer_chars, const int64_t *kmer_off, const double *prob, int64_t n) {
    if (!ctx) return BS_ERR_INVALID;
    if (n < 0 || (n > 0 && (!kmer_chars || !kmer_off || !prob))) return fail(ctx, BS_ERR_INVALID, "bs_set_table: NULL argument");
    if (n > 0x7fffffff) return fail(ctx, BS_ERR_INVALID, "bs_set_table: too many rows");
    cudaSetDevice(ctx->device);
    std::vector<double> pd(bs::DENSE_SIZE, 0.0);
    std::vector<int32_t> rd(bs::DENSE_SIZE, -1);
    for (int64_t i = 0; i < n; i++) {
        const int64_t len = kmer_off[i + 1] - kmer_off[i];
        if (len < 1 || len > bs::MAXK)
            return fail(ctx, BS_ERR_TABLE, "table row %lld: k-mer length %lld outside 1..8", (long long)i, (long long)len);
        int code = 0;
        for (int64_t j = 0; j < len; j++) {
            const unsigned char ch = (unsigned char)kmer_chars[kmer_off[i] + j];
            if (!bs::base_valid(ch))
                return fail(ctx, BS_ERR_TABLE, "table row %lld holds a byte outside ACGT", (long long)i);
            code = code * 4 + (int)bs::base_code(ch);
        }
        if (std::isnan(prob[i])) return fail(ctx, BS_ERR_TABLE, "table row %lld: probability is NaN", (long long)i);
        const int di = bs::dense_offset((int)len) + code;
        pd[di] = prob[i];
        rd[di] = (int32_t)i;  // a repeated key overrides the earlier row
    }
    BS_TRY(sync_all(ctx));
    ctx->prob_dense.swap(pd);
    ctx->row_dense.swap(rd);
    // a truth-side table is indexed by the PREVIOUS table's key -> row mapping: it goes with the old table
    // (callers that want one set it again with bs_set_truth_table, as the Python and Rcpp front-ends do)
    ctx->tprob_dense = ctx->prob_dense;
    ctx->has_truth_table = false;
    ctx->T = n;
    ctx->has_table = true;
    ctx->table_version++;
    ctx->has_table2 = false;  // rows may have changed: a second table has to be set again
    return upload_table(ctx, ctx->prob_dense, ctx->d_tab);
}

static int bs_set_second_table_impl(bs_ctx *ctx, const double *prob, int64_t n);

int bs_set_second_table(bs_ctx *ctx, const double *prob, int64_t n) {
    if (!ctx) return BS_ERR_INVALID;
    try {
        return bs_set_second_table_impl(ctx, prob, n);
    } catch (const std::bad_alloc &) {
        return fail(ctx, BS_ERR_ALLOC, "bs_set_second_table: out of host memory");
    } catch (...) {
        return fail(ctx, BS_ERR_STATE, "bs_set_second_table: unexpected exception");
    }
}

static int bs_set_second_table_impl(bs_ctx *ctx, const double *prob, int64_t n) {
    if (!ctx) return BS_ERR_INVALID;
    if (!ctx->has_table) return fail(ctx, BS_ERR_STATE, "bs_set_second_table before bs_set_table");
    if (!prob) { ctx->has_table2 = false; return BS_OK; }
    if (n != ctx->T) return fail(ctx, BS_ERR_INVALID, "second table has %lld rows, scoring table %lld", (long long)n, (long long)ctx->T);
    cudaSetDevice(ctx->device);
    BS_TRY(sync_all(ctx));
    std::vector<double> t(bs::DENSE_SIZE, 0.0);
    for (int di = 0; di < bs::DENSE_SIZE; di++) {
        const int32_t r = ctx->row_dense[di];
        if (r >= 0) {
            if (std::isnan(prob[r])) return fail(ctx, BS_ERR_TABLE, "second table row %d is NaN", r);
            t[di] = prob[r];
        }
    }
    ctx->prob2_dense.swap(t);
    ctx->has_table2 = true;
    ctx->ks2.kmer = -1;  // rank maps are rebuilt on the next call
    return upload_table(ctx, ctx->prob2_dense, ctx->d_tab2);
}

static int bs_set_truth_table_impl(bs_ctx *ctx, const double *prob, int64_t n);

int bs_set_truth_table(bs_ctx *ctx, const double *prob, int64_t n) {
    if (!ctx) return BS_ERR_INVALID;
    try {
        return bs_set_truth_table_impl(ctx, prob, n);
    } catch (const std::bad_alloc &) {
        return fail(ctx, BS_ERR_ALLOC, "bs_set_truth_table: out of host memory");
    } catch (...) {
        return fail(ctx, BS_ERR_STATE, "bs_set_truth_table: unexpected exception");
    }
}

static int bs_set_truth_table_impl(bs_ctx *ctx, const double *prob, int64_t n) {
    if (!ctx) return BS_ERR_INVALID;
    if (!ctx->has_table) return fail(ctx, BS_ERR_STATE, "bs_set_truth_table before bs_set_table");
    if (!prob) {
        ctx->tprob_dense = ctx->prob_dense;
        ctx->has_truth_table = false;
        ctx->table_version++;
        return BS_OK;
    }
    if (n != ctx->T) return fail(ctx, BS_ERR_INVALID, "truth table has %lld rows, scoring table %lld", (long long)n, (long long)ctx->T);
    std::vector<double> t(bs::DENSE_SIZE, 0.0);
    for (int di = 0; di < bs::DENSE_SIZE; di++) {
        const int32_t r = ctx->row_dense[di];
        if (r >= 0) {
            if (std::isnan(prob[r])) return fail(ctx, BS_ERR_TABLE, "truth table row %d is NaN", r);
            t[di] = prob[r];
        }
    }
    ctx->tprob_dense.swap(t);
    ctx->has_truth_table = true;
    ctx->table_version++;
    return BS_OK;
}

static int score_batch_impl(bs_ctx *ctx, const bs_batch *b, int kmer, uint32_t flags, bs_result *res,
                            const bs_scaffold_set *comp = nullptr, const int32_t *part_dst = nullptr);

// nothing throws across the boundary: host allocation failures of the std containers used inside become a status
int bs_score_batch(bs_ctx *ctx, const bs_batch *b, int kmer, uint32_t flags, bs_result *res) {
    if (!ctx) return BS_ERR_INVALID;
    try {
        return score_batch_impl(ctx, b, kmer, flags, res);
    } catch (const std::bad_alloc &) {
        return fail(ctx, BS_ERR_ALLOC, "bs_score_batch: out of host memory");
    } catch (const std::exception &ex) {
        return fail(ctx, BS_ERR_STATE, "bs_score_batch: %s", ex.what());
    } catch (...) {
        return fail(ctx, BS_ERR_STATE, "bs_score_batch: unknown exception");
    }
}

static int score_batch_impl(bs_ctx *ctx, const bs_batch *b, int kmer, uint32_t flags, bs_result *res,
                            const bs_scaffold_set *comp, const int32_t *part_dst) {
    if (!b || !res) return fail(ctx, BS_ERR_INVALID, "bs_score_batch: NULL batch or result");
    if (!ctx->has_table) return fail(ctx, BS_ERR_STATE, "bs_score_batch before bs_set_table");
    if (kmer < 1) return fail(ctx, BS_ERR_INVALID, "kmer must be >= 1 (got %d)", kmer);
    const int64_t S = b->n_segments, N = b->n_reads, C = b->n_contigs;
    if (S < 0 || N < 0 || C < 0) return fail(ctx, BS_ERR_INVALID, "negative counts");
    if (C > 0x7fffffff || S > 0x7fffffff || N > 0xfffffff0ll) return fail(ctx, BS_ERR_INVALID, "too many contigs/segments/reads for one call");
    BS_TRY(check_offsets(ctx, "contig", b->contig_off, C));
    BS_TRY(check_offsets(ctx, "truth", b->truth_off, S));
    BS_TRY(check_offsets(ctx, "segment read-start", b->seg_read_start, S));
    BS_TRY(check_offsets(ctx, "segment contig-start", b->seg_contig_start, S));
    if (b->seg_read_start[0] != 0 || b->seg_read_start[S] != N || b->seg_contig_start[0] != 0 || b->seg_contig_start[S] != C)
        return fail(ctx, BS_ERR_INVALID, "segment start arrays do not cover all reads/contigs");
    if (b->read_off) BS_TRY(check_offsets(ctx, "read", b->read_off, N));
    else if (b->read_len < 0) return fail(ctx, BS_ERR_INVALID, "read_len < 0");
    if ((flags & BS_WANT_PROB_DIST) && (!res->path_prob_dist || !res->path_prob_dist_off))
        return fail(ctx, BS_ERR_INVALID, "BS_WANT_PROB_DIST needs path_prob_dist and path_prob_dist_off");
    if ((flags & BS_WANT_POS) && (!res->pos || !res->pos_off)) return fail(ctx, BS_ERR_INVALID, "BS_WANT_POS needs pos and pos_off");
    if ((flags & BS_WANT_HIST) && !res->hist) return fail(ctx, BS_ERR_INVALID, "BS_WANT_HIST needs hist");
    for (int64_t c = 0; c < C; c++)
        if (b->contig_off[c + 1] - b->contig_off[c] > 0x7f000000ll) return fail(ctx, BS_ERR_INVALID, "contig %lld longer than 2^31", (long long)c);
    for (int64_t s = 0; s < S; s++)
        if (b->truth_off[s + 1] - b->truth_off[s] > 0x7f000000ll) return fail(ctx, BS_ERR_INVALID, "truth %lld longer than 2^31", (long long)s);
    if (C == 0) return BS_OK;
    const int64_t read_total = b->read_off ? b->read_off[N] : N * (int64_t)b->read_len;
    if (N > 0 && !b->read_chars && read_total > 0) return fail(ctx, BS_ERR_INVALID, "read_chars is NULL");
    if (b->contig_off[C] > 0 && !b->contig_chars && !comp) return fail(ctx, BS_ERR_INVALID, "contig_chars is NULL");
    if (b->truth_off[S] > 0 && !b->truth_chars) return fail(ctx, BS_ERR_INVALID, "truth_chars is NULL");
    cudaSetDevice(ctx->device);

    CallEnv e;
    e.b = b; e.res = res; e.kmer = kmer; e.flags = flags; e.T = ctx->T;
    e.dev_chars = (flags & BS_DEVICE_CHARS) != 0;
    e.dev_res = (flags & BS_DEVICE_RESULT) != 0;
    e.want_ks = (flags & BS_WANT_KS) && (res->ks_stat_prob_dist || res->ks_stat_path_freq);
    e.want_pd = (flags & BS_WANT_PROB_DIST) != 0;
    e.want_pos = (flags & BS_WANT_POS) != 0;
    e.want_hist = (flags & BS_WANT_HIST) != 0;
    e.want_sp = (flags & BS_WANT_STARTPOS) && res->path_prob_dist_startpos;
    e.want_lev = (flags & BS_WANT_LEV) && res->lev_dist_vs_true;
    e.second = (flags & BS_WANT_SECOND_TABLE) != 0;
    e.w_out = (flags & BS_WEIGHTS_OUT) != 0;
    e.w_in = (flags & BS_WEIGHTS_IN) != 0;
    if (e.w_out && e.w_in) return fail(ctx, BS_ERR_INVALID, "BS_WEIGHTS_OUT and BS_WEIGHTS_IN exclude each other");
    if ((e.w_out || e.w_in) && (!res->weights || !res->weights_total))
        return fail(ctx, BS_ERR_INVALID, "BS_WEIGHTS_OUT / BS_WEIGHTS_IN need result->weights and result->weights_total (device pointers)");
    if (e.w_in && (flags & BS_WANT_POS)) return fail(ctx, BS_ERR_INVALID, "BS_WANT_POS needs the placement: not with BS_WEIGHTS_IN");
    if (e.w_out) {  // nothing but the placement runs
        e.want_ks = e.want_pd = e.want_hist = e.want_sp = e.want_lev = e.second = false;
    }
    if (e.second && !ctx->has_table2) return fail(ctx, BS_ERR_STATE, "BS_WANT_SECOND_TABLE before bs_set_second_table");
    if (comp) {
        if (e.w_out || e.w_in || (flags & (BS_PLACE_SCAN | BS_PLACE_TILE)))
            return fail(ctx, BS_ERR_INVALID, "bs_score_scaffolds: BS_WEIGHTS_* and BS_PLACE_SCAN / BS_PLACE_TILE do not apply to a scaffold set");
        if (S != 1) return fail(ctx, BS_ERR_INVALID, "bs_score_scaffolds: one segment per call");
        e.comp = comp;
        e.part_dst = part_dst;
        // the second table needs the position weights (its scores come from k_break_score): then the compositional kernel
        // writes them and the scoring stages run as for any contig set
        const char *env = std::getenv("BS_COMPOSE_SCORE");  // tests: 0 keeps the weights path
        e.comp_score = !e.second && !(env && env[0] == '0');
    }
    if (e.second && e.want_pd && !res->path_prob_dist2) return fail(ctx, BS_ERR_INVALID, "BS_WANT_SECOND_TABLE with BS_WANT_PROB_DIST needs path_prob_dist2");
    e.read_chars = b->read_chars;
    e.roff = b->read_off;
    e.rlen = b->read_len;
    // reads given with offsets but all of one length (dense by construction): the fast uniform path
    if (e.roff && N > 0) {
        const int64_t l0 = e.roff[1] - e.roff[0];
        bool uni = l0 >= 1 && l0 <= 0x3fffffff;
        for (int64_t n = 1; uni && n < N; n++) uni = (e.roff[n + 1] - e.roff[n]) == l0;
        if (uni) {
            e.read_chars = b->read_chars + e.roff[0];
            e.roff = nullptr;
            e.rlen = (int32_t)l0;
        }
    }
    if (e.want_ks || e.want_pd) BS_TRY(prepare_ks(ctx, kmer));
    if (e.second && (e.want_ks || e.want_pd)) BS_TRY(prepare_ks(ctx, kmer, ctx->ks2, ctx->prob2_dense));
    const bool any_ksb = e.want_ks && (res->ks_stat_path_freq || (e.second && res->ks_stat_path_freq2));
    if (any_ksb) {
        BS_TRY(ensure(ctx, ctx->d_status, 16));
        // (not cleared between device-result calls: an overflow of an earlier asynchronous call is reported by the
        // next synchronising call -- bs_ctx_synchronize or a host-result score)
        if (!ctx->status_pending) BS_CUDA(cudaMemsetAsync(ctx->d_status.p, 0, 16, ctx->stream));
    }

    // ---------------- chunks of whole segments ----------------
    // inputs on the device but results going home: medium chunks, so that the D2H of one chunk hides
    // behind the kernels of the next
    const int64_t target = !e.dev_chars ? ctx->chunk_bytes_host
                           : (e.dev_res ? ctx->chunk_bytes_dev : std::min<int64_t>(ctx->chunk_bytes_dev, (int64_t)384 << 20));
    std::vector<Chunk> chunks;
    {
        auto seg_bytes_of = [&](int64_t s) {
            return read_byte_begin(e, b->seg_read_start[s + 1]) - read_byte_begin(e, b->seg_read_start[s]) +
                   (b->contig_off[b->seg_contig_start[s + 1]] - b->contig_off[b->seg_contig_start[s]]) +
                   (b->truth_off[s + 1] - b->truth_off[s]);
        };
        int64_t total_bytes = 0;
        for (int64_t s = 0; s < S; s++) total_bytes += seg_bytes_of(s);
        // host inputs: small chunks at both ends of a long call (the first copy and the last kernels and
        // result copy are the only parts of the pipeline that nothing overlaps), full-size ones between
        const bool ramp = !e.dev_chars && total_bytes > 4 * target;
        int64_t s0 = 0, bytes = 0, done = 0;
        for (int64_t s = 0; s < S; s++) {
            const int64_t seg_bytes = seg_bytes_of(s);
            int64_t limit = target;
            if (ramp) {
                const int64_t left = total_bytes - done;
                if (done < target / 4 || left <= target / 2) limit = target / 4;
                else if (done < target || left <= 3 * target / 2) limit = target / 2;
            }
            if (s > s0 && bytes + seg_bytes > limit) {
                chunks.push_back({s0, s, b->seg_read_start[s0], b->seg_read_start[s], b->seg_contig_start[s0], b->seg_contig_start[s]});
                s0 = s;
                done += bytes;
                bytes = 0;
            }
            bytes += seg_bytes;
        }
        chunks.push_back({s0, S, b->seg_read_start[s0], b->seg_read_start[S], b->seg_contig_start[s0], b->seg_contig_start[S]});
    }
    bool first_chunk = true;
    for (const Chunk &ch : chunks) {
        if (ch.c1 == ch.c0) continue;  // segments without contigs produce nothing
        if (!first_chunk && ctx->poll && ctx->poll(ctx->poll_user)) {
            // the caller wants out (R: user interrupt): nothing new is queued, what is in flight finishes
            sync_all(ctx);
            for (Workspace &w : ctx->ws) w.in_flight = false;
            return fail(ctx, BS_ERR_INTERRUPTED, "interrupted by the caller's poll callback after segment %lld of %lld",
                        (long long)ch.s0, (long long)S);
        }
        first_chunk = false;
        BS_TRY(run_chunk(ctx, ctx->ws[ctx->ws_cursor++ % kWorkspaces], e, ch));
    }
    if (e.dev_res) {
        // asynchronous: the caller orders later work on the compute stream.  The workspaces stay marked
        // in flight: whoever takes one next (this call's successor included) first waits for the event
        // recorded behind its last kernel, so staging memory is never rewritten under a running chunk.
        ctx->best_dirty = false;
        if (any_ksb) ctx->status_pending = true;  // read by bs_ctx_synchronize / the next host-result call
        return BS_OK;
    }
    int32_t status = 0;
    if (any_ksb) BS_CUDA(cudaMemcpyAsync(&status, ctx->d_status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->status_pending = false;
    BS_TRY(sync_all(ctx));
    for (Workspace &w : ctx->ws) w.in_flight = false;
    ctx->best_dirty = false;
    if (res->lev_dist_vs_true && !e.want_lev) std::memset(res->lev_dist_vs_true, 0, (size_t)C * 4);
    if (!e.want_sp && res->path_prob_dist_startpos) std::memset(res->path_prob_dist_startpos, 0, (size_t)C * 4);
    if (status) return fail(ctx, BS_ERR_INVALID, "ks_stat_path_freq: more than %d table rows with a count >= %d in one contig", bs::OVF_CAP, bs::CC_DENSE);
    return BS_OK;
}

// full validation of a scaffold set (bs_assemble.cpp)
int bs_scaffold_validate(const bs_scaffold_set *set, int64_t *lengths, int32_t *part_dst, char *err, size_t errn);

static int score_scaffolds_impl(bs_ctx *ctx, const bs_scaffold_set *set, const char *read_chars, const int64_t *read_off, int64_t n_reads,
                                int32_t read_len, const char *truth, int64_t truth_len, int kmer, uint32_t flags, bs_result *result) {
    if (!set || !result) return fail(ctx, BS_ERR_INVALID, "bs_score_scaffolds: NULL set or result");
    if (truth_len < 0 || n_reads < 0) return fail(ctx, BS_ERR_INVALID, "bs_score_scaffolds: negative size");
    const int64_t C = set->n_scaffolds;
    if (C < 0) return fail(ctx, BS_ERR_INVALID, "bs_score_scaffolds: negative scaffold count");
    if (C == 0) return BS_OK;
    const int64_t n_parts = set->scaffold_part_start ? set->scaffold_part_start[C] : 0;
    std::vector<int64_t> off((size_t)C + 1, 0);
    std::vector<int32_t> part_dst((size_t)std::max<int64_t>(n_parts, 1));
    {
        std::vector<int64_t> len((size_t)C);
        char msg[256] = "";
        if (bs_scaffold_validate(set, len.data(), part_dst.data(), msg, sizeof msg) != BS_OK) return fail(ctx, BS_ERR_INVALID, "%s", msg);
        for (int64_t c = 0; c < C; c++) off[(size_t)c + 1] = off[(size_t)c] + len[(size_t)c];
    }
    const int64_t truth_off[2] = {0, truth_len};
    const int64_t seg_r[2] = {0, n_reads}, seg_c[2] = {0, C};
    bs_batch b;
    std::memset(&b, 0, sizeof(b));
    b.n_segments = 1; b.n_reads = n_reads; b.n_contigs = C;
    b.read_chars = read_chars; b.read_off = read_off; b.read_len = read_len;
    b.contig_chars = nullptr; b.contig_off = off.data();
    b.truth_chars = truth; b.truth_off = truth_off;
    b.seg_read_start = seg_r; b.seg_contig_start = seg_c;
    // The lists of the from-the-parts path are [base contig][read] arrays: past 2^31 cells (a cfg-5 sized read set against
    // hundreds of base contigs) the texts are materialised instead and scored the ordinary way -- same results.
    int64_t max_cells = (int64_t)1 << 31;
    if (const char *env = std::getenv("BS_COMPOSE_MAX_CELLS")) max_cells = std::atoll(env);  // tests
    if (set->n_base * n_reads > max_cells) {
        if (flags & (BS_WEIGHTS_OUT | BS_WEIGHTS_IN)) return fail(ctx, BS_ERR_INVALID, "bs_score_scaffolds: BS_WEIGHTS_* do not apply to a scaffold set");
        std::vector<char> chars((size_t)off[(size_t)C] + 1);
        std::vector<int64_t> off2((size_t)C + 1);
        if (bs_scaffold_texts(set, chars.data(), off2.data()) != BS_OK) return fail(ctx, BS_ERR_INVALID, "%s", bs_assemble_last_error());
        b.contig_chars = chars.data();
        if (flags & BS_DEVICE_CHARS) {  // reads and truth are on the device: the texts go there too
            cudaSetDevice(ctx->device);
            BS_TRY(sync_all(ctx));  // (m_ctgs may still be read by an earlier asynchronous call)
            BS_TRY(ensure(ctx, ctx->m_ctgs, chars.size() + 32));
            BS_CUDA(cudaMemcpyAsync(ctx->m_ctgs.p, chars.data(), (size_t)off[(size_t)C], cudaMemcpyHostToDevice, ctx->stream));
            BS_CUDA(cudaStreamSynchronize(ctx->stream));
            b.contig_chars = (const char *)ctx->m_ctgs.p;
        }
        return score_batch_impl(ctx, &b, kmer, flags, result);
    }
    return score_batch_impl(ctx, &b, kmer, flags, result, set, part_dst.data());
}

int bs_score_scaffolds(bs_ctx *ctx, const bs_scaffold_set *set, const char *read_chars, const int64_t *read_off, int64_t n_reads,
                       int32_t read_len, const char *truth, int64_t truth_len, int kmer, uint32_t flags, bs_result *result) {
    if (!ctx) return BS_ERR_INVALID;
    try {
        return score_scaffolds_impl(ctx, set, read_chars, read_off, n_reads, read_len, truth, truth_len, kmer, flags, result);
    } catch (const std::bad_alloc &) {
        return fail(ctx, BS_ERR_ALLOC, "bs_score_scaffolds: out of host memory");
    } catch (const std::exception &ex) {
        return fail(ctx, BS_ERR_STATE, "bs_score_scaffolds: %s", ex.what());
    } catch (...) {
        return fail(ctx, BS_ERR_STATE, "bs_score_scaffolds: unknown exception");
    }
}

int bs_score(bs_ctx *ctx, const char *contig_chars, const int64_t *contig_off, int64_t n_contigs,
             const char *read_chars, const int64_t *read_off, int64_t n_reads,
             const char *truth, int64_t truth_len, int kmer, uint32_t flags, bs_result *result) {
    if (!ctx) return BS_ERR_INVALID;
    if (truth_len < 0) return fail(ctx, BS_ERR_INVALID, "truth_len < 0");
    const int64_t truth_off[2] = {0, truth_len};
    const int64_t seg_r[2] = {0, n_reads}, seg_c[2] = {0, n_contigs};
    bs_batch b;
    std::memset(&b, 0, sizeof(b));
    b.n_segments = 1; b.n_reads = n_reads; b.n_contigs = n_contigs;
    b.read_chars = read_chars; b.read_off = read_off; b.read_len = 0;
    b.contig_chars = contig_chars; b.contig_off = contig_off;
    b.truth_chars = truth; b.truth_off = truth_off;
    b.seg_read_start = seg_r; b.seg_contig_start = seg_c;
    if (!read_off && n_reads > 0) return fail(ctx, BS_ERR_INVALID, "bs_score: read_off is NULL");
    return bs_score_batch(ctx, &b, kmer, flags, result);
}

// ---- one segment over several contexts (GPUs) from one process --------------------------------
namespace {

// the share of one context: its contigs (ascending input indices), their flattened text and private result arrays
struct MultiShard {
    std::vector<int64_t> idx, off, pd_off, pos_off;
    char *chars = nullptr;            // the owning context's pinned staging buffer
    std::vector<int32_t> i32[4];      // sequence_len, kmer_breaks, startpos, lev
    std::vector<double> f64[10];      // score, norm, by_len, ks_a, ks_b and the same for the second table
    std::vector<double> pd, pd2;
    std::vector<int32_t> hist, pos;
    bs_result r;
    int rc = BS_OK;
};

}  // namespace

static int score_multi_impl(bs_ctx *const *ctxs, int n_ctx, const char *contig_chars, const int64_t *contig_off, int64_t n_contigs,
                            const char *read_chars, const int64_t *read_off, int64_t n_reads,
                            const char *truth, int64_t truth_len, int kmer, uint32_t flags, bs_result *res);

int bs_score_multi(bs_ctx *const *ctxs, int n_ctx, const char *contig_chars, const int64_t *contig_off, int64_t n_contigs,
                   const char *read_chars, const int64_t *read_off, int64_t n_reads,
                   const char *truth, int64_t truth_len, int kmer, uint32_t flags, bs_result *res) {
    if (!ctxs || n_ctx < 1 || !ctxs[0]) return BS_ERR_INVALID;
    try {
        return score_multi_impl(ctxs, n_ctx, contig_chars, contig_off, n_contigs, read_chars, read_off, n_reads, truth, truth_len, kmer, flags, res);
    } catch (const std::bad_alloc &) {
        return fail(ctxs[0], BS_ERR_ALLOC, "bs_score_multi: out of host memory");
    } catch (const std::exception &ex) {  // e.g. std::system_error from std::thread
        return fail(ctxs[0], BS_ERR_STATE, "bs_score_multi: %s", ex.what());
    } catch (...) {
        return fail(ctxs[0], BS_ERR_STATE, "bs_score_multi: unknown exception");
    }
}

static int score_multi_impl(bs_ctx *const *ctxs, int n_ctx, const char *contig_chars, const int64_t *contig_off, int64_t n_contigs,
                            const char *read_chars, const int64_t *read_off, int64_t n_reads,
                            const char *truth, int64_t truth_len, int kmer, uint32_t flags, bs_result *res) {
    bs_ctx *ctx = ctxs[0];
    for (int k = 0; k < n_ctx; k++) {
        if (!ctxs[k]) return fail(ctx, BS_ERR_INVALID, "bs_score_multi: context %d is NULL", k);
        for (int j = 0; j < k; j++)  // a context is single-threaded: the same one twice would be driven by two threads
            if (ctxs[j] == ctxs[k]) return fail(ctx, BS_ERR_INVALID, "bs_score_multi: context %d is the same as context %d", k, j);
    }
    if (flags & (BS_DEVICE_CHARS | BS_DEVICE_RESULT)) return fail(ctx, BS_ERR_INVALID, "bs_score_multi takes host buffers only");
    if (!res) return fail(ctx, BS_ERR_INVALID, "bs_score_multi: NULL result");
    if (n_contigs < 0 || n_reads < 0) return fail(ctx, BS_ERR_INVALID, "negative counts");
    if (n_ctx == 1 || n_contigs == 0)
        return bs_score(ctx, contig_chars, contig_off, n_contigs, read_chars, read_off, n_reads, truth, truth_len, kmer, flags, res);
    BS_TRY(check_offsets(ctx, "contig", contig_off, n_contigs));
    const bool want_pd = (flags & BS_WANT_PROB_DIST) != 0, want_pos = (flags & BS_WANT_POS) != 0, want_hist = (flags & BS_WANT_HIST) != 0;
    const bool second = (flags & BS_WANT_SECOND_TABLE) != 0;
    if (want_pd && (!res->path_prob_dist || !res->path_prob_dist_off)) return fail(ctx, BS_ERR_INVALID, "BS_WANT_PROB_DIST needs path_prob_dist and path_prob_dist_off");
    if (want_pd && second && !res->path_prob_dist2) return fail(ctx, BS_ERR_INVALID, "BS_WANT_SECOND_TABLE with BS_WANT_PROB_DIST needs path_prob_dist2");
    if (want_pos && (!res->pos || !res->pos_off)) return fail(ctx, BS_ERR_INVALID, "BS_WANT_POS needs pos and pos_off");
    if (want_hist && !res->hist) return fail(ctx, BS_ERR_INVALID, "BS_WANT_HIST needs hist");
    const int64_t T = ctx->T;
    for (int k = 1; k < n_ctx; k++)
        if (!ctxs[k]->has_table || ctxs[k]->T != T) return fail(ctx, BS_ERR_STATE, "bs_score_multi: context %d does not hold the same table", k);

    // longest-processing-time-first over contig lengths; a context keeps its contigs in input order
    std::vector<int64_t> order((size_t)n_contigs);
    std::iota(order.begin(), order.end(), (int64_t)0);
    auto len_of = [&](int64_t c) { return contig_off[c + 1] - contig_off[c]; };
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return len_of(a) > len_of(b); });
    std::vector<MultiShard> sh((size_t)n_ctx);
    {
        std::vector<int64_t> load((size_t)n_ctx, 0);
        for (int64_t c : order) {
            const int k = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            sh[k].idx.push_back(c);
            load[k] += len_of(c) + 1;
        }
    }
    double *const f64_dst[10] = {res->bp_score, res->bp_score_norm_by_break_freqs, res->bp_score_norm_by_len, res->ks_stat_prob_dist,
                                 res->ks_stat_path_freq, res->bp_score2, res->bp_score_norm_by_break_freqs2,
                                 res->bp_score_norm_by_len2, res->ks_stat_prob_dist2, res->ks_stat_path_freq2};
    int32_t *const i32_dst[4] = {res->sequence_len, res->kmer_breaks, res->path_prob_dist_startpos, res->lev_dist_vs_true};
    for (MultiShard &m : sh) {
        std::sort(m.idx.begin(), m.idx.end());
        const size_t n = m.idx.size();
        m.off.assign(n + 1, 0);
        m.pd_off.assign(n + 1, 0);
        m.pos_off.assign(n + 1, 0);
        for (size_t i = 0; i < n; i++) {
            const int64_t L = len_of(m.idx[i]);
            m.off[i + 1] = m.off[i] + L;
            m.pd_off[i + 1] = m.pd_off[i] + std::max<int64_t>(L - kmer + 1, 0);
            m.pos_off[i + 1] = m.pos_off[i] + n_reads;
        }
        {
            bs_ctx *own = ctxs[&m - sh.data()];
            const size_t need = (size_t)m.off[n] + 1;
            if (need > own->h_multi_cap) {
                cudaSetDevice(own->device);
                if (own->h_multi) cudaFreeHost(own->h_multi);
                own->h_multi = nullptr;
                own->h_multi_cap = 0;
                const size_t want = need + need / 4 + 4096;
                if (cudaHostAlloc((void **)&own->h_multi, want, cudaHostAllocPortable) != cudaSuccess)
                    return fail(ctx, BS_ERR_ALLOC, "bs_score_multi: cudaHostAlloc(%zu) failed", want);
                own->h_multi_cap = want;
            }
            m.chars = own->h_multi;
        }
        for (size_t i = 0; i < n; i++)
            if (len_of(m.idx[i])) std::memcpy(m.chars + m.off[i], contig_chars + contig_off[m.idx[i]], (size_t)len_of(m.idx[i]));
        std::memset(&m.r, 0, sizeof(m.r));
        for (int j = 0; j < 4; j++) if (i32_dst[j]) m.i32[j].assign(n + 1, 0);
        for (int j = 0; j < 10; j++) if (f64_dst[j]) m.f64[j].assign(n + 1, 0.0);
        m.r.sequence_len = i32_dst[0] ? m.i32[0].data() : nullptr;
        m.r.kmer_breaks = i32_dst[1] ? m.i32[1].data() : nullptr;
        m.r.path_prob_dist_startpos = i32_dst[2] ? m.i32[2].data() : nullptr;
        m.r.lev_dist_vs_true = i32_dst[3] ? m.i32[3].data() : nullptr;
        double **f64_src[10] = {&m.r.bp_score, &m.r.bp_score_norm_by_break_freqs, &m.r.bp_score_norm_by_len, &m.r.ks_stat_prob_dist,
                                &m.r.ks_stat_path_freq, &m.r.bp_score2, &m.r.bp_score_norm_by_break_freqs2,
                                &m.r.bp_score_norm_by_len2, &m.r.ks_stat_prob_dist2, &m.r.ks_stat_path_freq2};
        for (int j = 0; j < 10; j++) *f64_src[j] = f64_dst[j] ? m.f64[j].data() : nullptr;
        if (want_pd) {
            m.pd.assign((size_t)m.pd_off[n] + 1, 0.0);
            m.r.path_prob_dist = m.pd.data();
            m.r.path_prob_dist_off = m.pd_off.data();
            if (second) { m.pd2.assign((size_t)m.pd_off[n] + 1, 0.0); m.r.path_prob_dist2 = m.pd2.data(); }
        }
        if (want_hist) { m.hist.assign(n * (size_t)(T + 1) + 1, 0); m.r.hist = m.hist.data(); }
        if (want_pos) { m.pos.assign((size_t)m.pos_off[n] + 1, 0); m.r.pos = m.pos.data(); m.r.pos_off = m.pos_off.data(); }
    }

    // A large read set is replicated over NVLink instead of PCIe: context 0 copies it in once, the others get it by peer
    // copies along a binary tree (round r: the 2^r contexts that hold it each feed one that does not), everything
    // queued here on the contexts' copy streams and ordered by events.  BS_MULTI_P2P_MB: smallest read set (MB) that
    // takes this way (default 64; a negative value switches it off).
    bool share = false;
    std::vector<int64_t> read_off0;
#ifndef BS_CPU_EMUL
    {
        const int64_t read_bytes = n_reads > 0 ? read_off[n_reads] - read_off[0] : 0;
        int64_t thresh_mb = 64;
        if (const char *env = std::getenv("BS_MULTI_P2P_MB")) thresh_mb = std::atoll(env);
        share = n_reads > 0 && thresh_mb >= 0 && read_bytes >= (thresh_mb << 20);
        if (share) {
            if (read_off[0] != 0) {  // offsets relative to the staged copy
                read_off0.resize((size_t)n_reads + 1);
                for (int64_t n = 0; n <= n_reads; n++) read_off0[(size_t)n] = read_off[n] - read_off[0];
            }
            for (int k = 0; k < n_ctx; k++) {
                bs_ctx *c = ctxs[k];
                cudaSetDevice(c->device);
                BS_TRY(ensure(c, c->m_reads, (size_t)read_bytes + 32));
                if (!c->ev_share && cudaEventCreateWithFlags(&c->ev_share, cudaEventDisableTiming) != cudaSuccess)
                    return fail(ctx, BS_ERR_CUDA, "bs_score_multi: event creation failed on context %d", k);
            }
            cudaSetDevice(ctx->device);
            BS_CUDA(cudaMemcpyAsync(ctx->m_reads.p, read_chars + read_off[0], (size_t)read_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            BS_CUDA(cudaEventRecord(ctx->ev_share, ctx->copy_stream));
            for (int have = 1; have < n_ctx; have *= 2) {
                for (int src = 0; src < have && src + have < n_ctx; src++) {
                    bs_ctx *from = ctxs[src], *to = ctxs[src + have];
                    cudaSetDevice(to->device);
                    if (to->device != from->device) {
                        const cudaError_t pe = cudaDeviceEnablePeerAccess(from->device, 0);  // (already enabled / unsupported: the copy is staged)
                        if (pe != cudaSuccess) cudaGetLastError();
                    }
                    BS_CUDA(cudaStreamWaitEvent(to->copy_stream, from->ev_share, 0));
                    BS_CUDA(cudaMemcpyPeerAsync(to->m_reads.p, to->device, from->m_reads.p, from->device, (size_t)read_bytes, to->copy_stream));
                    BS_CUDA(cudaEventRecord(to->ev_share, to->copy_stream));
                }
            }
            if (std::getenv("BS_TRACE")) {  // diagnostic: when the last GPU holds the reads
                const auto t0 = std::chrono::steady_clock::now();
                for (int k = 0; k < n_ctx; k++) cudaEventSynchronize(ctxs[k]->ev_share);
                std::fprintf(stderr, "[bs trace] reads on %d GPUs   %9.3f ms after queueing (%lld bytes)\n", n_ctx,
                             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), (long long)read_bytes);
            }
        }
    }
#endif
    // one host thread per context with work (a context is single-threaded, contexts are independent)
    auto run = [&](int k) {  // (bs_score catches everything itself: nothing escapes a worker thread)
        MultiShard &m = sh[k];
        if (m.idx.empty()) return;
#ifndef BS_CPU_EMUL
        if (share) {  // this context's contigs and the truth go up beside the reads; the call then takes device pointers
            bs_ctx *c = ctxs[k];
            cudaSetDevice(c->device);
            const size_t cb = (size_t)m.off[m.idx.size()], tb = (size_t)truth_len;
            m.rc = ensure(c, c->m_ctgs, cb + 32);
            if (m.rc == BS_OK) m.rc = ensure(c, c->m_truth, tb + 32);
            if (m.rc != BS_OK) return;
            cudaError_t ce = cudaSuccess;
            if (cb) ce = cudaMemcpyAsync(c->m_ctgs.p, m.chars, cb, cudaMemcpyHostToDevice, c->copy_stream);
            if (ce == cudaSuccess && tb) ce = cudaMemcpyAsync(c->m_truth.p, truth, tb, cudaMemcpyHostToDevice, c->copy_stream);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->copy_stream);  // (the reads arrive on the same stream)
            if (ce != cudaSuccess) { m.rc = fail(c, BS_ERR_CUDA, "bs_score_multi: staging failed: %s", cudaGetErrorString(ce)); return; }
            m.rc = bs_score(c, (const char *)c->m_ctgs.p, m.off.data(), (int64_t)m.idx.size(), (const char *)c->m_reads.p,
                            read_off0.empty() ? read_off : read_off0.data(), n_reads,
                            (const char *)c->m_truth.p, truth_len, kmer, flags | BS_DEVICE_CHARS, &m.r);
            return;
        }
#endif
        m.rc = bs_score(ctxs[k], m.chars, m.off.data(), (int64_t)m.idx.size(), read_chars, read_off, n_reads, truth, truth_len,
                        kmer, flags, &m.r);
    };
#ifdef BS_CPU_EMUL
    for (int k = 0; k < n_ctx; k++) run(k);  // (the CPU emulation keeps __shared__ arrays in statics: one launch at a time)
#else
    {
        std::vector<std::thread> threads;
        threads.reserve((size_t)n_ctx);
        int started = 1;
        try {
            for (; started < n_ctx; started++) threads.emplace_back(run, started);
        } catch (...) {  // could not start a thread: its share (and the rest) runs here
        }
        run(0);
        for (int k = started; k < n_ctx; k++) run(k);
        for (std::thread &t : threads) t.join();
    }
#endif
    for (int k = 0; k < n_ctx; k++)
        if (sh[k].rc != BS_OK) {
            if (k != 0) std::snprintf(ctx->err, sizeof(ctx->err), "context %d: %.480s", k, ctxs[k]->err);
            return sh[k].rc;
        }

    // back into input order
    for (const MultiShard &m : sh) {
        for (size_t i = 0; i < m.idx.size(); i++) {
            const int64_t c = m.idx[i];
            for (int j = 0; j < 4; j++) if (i32_dst[j]) i32_dst[j][c] = m.i32[j][i];
            for (int j = 0; j < 10; j++) if (f64_dst[j] && (j < 5 || second)) f64_dst[j][c] = m.f64[j][i];
            if (want_pd) {
                const size_t nb = (size_t)(m.pd_off[i + 1] - m.pd_off[i]) * sizeof(double);
                if (nb) {
                    std::memcpy(res->path_prob_dist + res->path_prob_dist_off[c], m.pd.data() + m.pd_off[i], nb);
                    if (second) std::memcpy(res->path_prob_dist2 + res->path_prob_dist_off[c], m.pd2.data() + m.pd_off[i], nb);
                }
            }
            if (want_hist) std::memcpy(res->hist + (size_t)c * (size_t)(T + 1), m.hist.data() + i * (size_t)(T + 1), (size_t)(T + 1) * 4);
            if (want_pos && n_reads) std::memcpy(res->pos + res->pos_off[c], m.pos.data() + m.pos_off[i], (size_t)n_reads * 4);
        }
    }
    return BS_OK;
}

int64_t bs_simulate_capacity(const int64_t *truth_off, int64_t n_segments, int32_t read_len, double coverage) {
    if (!truth_off || n_segments < 0 || read_len < 1 || !(coverage >= 0.0)) return -1;
    int64_t draws = 0;
    for (int64_t s = 0; s < n_segments; s++) {
        const int64_t L = truth_off[s + 1] - truth_off[s];
        draws += (int64_t)std::ceil(coverage * (double)L / (double)read_len);  // lib/GenerateReads.R:302
    }
    return draws * (int64_t)read_len;
}

static int bs_simulate_reads_impl(bs_ctx *ctx, const char *truth_chars, const int64_t *truth_off, int64_t n_segments, int32_t read_len,
                      double coverage, int kmer, uint64_t seed, uint32_t flags, char *reads_out, int64_t reads_capacity,
                      int64_t *seg_read_start);

int bs_simulate_reads(bs_ctx *ctx, const char *truth_chars, const int64_t *truth_off, int64_t n_segments, int32_t read_len,
                      double coverage, int kmer, uint64_t seed, uint32_t flags, char *reads_out, int64_t reads_capacity,
                      int64_t *seg_read_start) {
    if (!ctx) return BS_ERR_INVALID;
    try {
        return bs_simulate_reads_impl(ctx, truth_chars, truth_off, n_segments, read_len, coverage, kmer, seed, flags, reads_out, reads_capacity, seg_read_start);
    } catch (const std::bad_alloc &) {
        return fail(ctx, BS_ERR_ALLOC, "bs_simulate_reads: out of host memory");
    } catch (...) {
        return fail(ctx, BS_ERR_STATE, "bs_simulate_reads: unexpected exception");
    }
}

static int bs_simulate_reads_impl(bs_ctx *ctx, const char *truth_chars, const int64_t *truth_off, int64_t n_segments, int32_t read_len,
                      double coverage, int kmer, uint64_t seed, uint32_t flags, char *reads_out, int64_t reads_capacity,
                      int64_t *seg_read_start) {
    if (!ctx) return BS_ERR_INVALID;
    if (!ctx->has_table) return fail(ctx, BS_ERR_STATE, "bs_simulate_reads before bs_set_table");
    if (!truth_off || !seg_read_start || n_segments < 0) return fail(ctx, BS_ERR_INVALID, "bs_simulate_reads: NULL argument");
    if (read_len < 1 || !(coverage >= 0.0)) return fail(ctx, BS_ERR_INVALID, "read_len must be >= 1 and coverage >= 0");
    if (kmer < 1 || kmer > bs::MAXK) return fail(ctx, BS_ERR_INVALID, "kmer must be 1..8 (got %d)", kmer);
    const int64_t S = n_segments;
    BS_TRY(check_offsets(ctx, "truth", truth_off, S));
    seg_read_start[0] = 0;
    if (S == 0) return BS_OK;
    const int64_t tr_bytes = truth_off[S] - truth_off[0];
    if (tr_bytes > 0 && !truth_chars) return fail(ctx, BS_ERR_INVALID, "truth_chars is NULL");
    std::vector<int64_t> tr_off(S + 1), tr_woff(S + 1), draw_off(S + 1);
    tr_woff[0] = 0;
    draw_off[0] = 0;
    for (int64_t s = 0; s <= S; s++) tr_off[s] = truth_off[s] - truth_off[0];
    for (int64_t s = 0; s < S; s++) {
        const int64_t L = tr_off[s + 1] - tr_off[s];
        if (L > 0x7f000000ll) return fail(ctx, BS_ERR_INVALID, "truth %lld longer than 2^31", (long long)s);
        tr_woff[s + 1] = tr_woff[s] + (L + 31) / 32 + 2;
        draw_off[s + 1] = draw_off[s] + (int64_t)std::ceil(coverage * (double)L / (double)read_len);
    }
    const int64_t n_draws = draw_off[S];
    if (n_draws * (int64_t)read_len > reads_capacity)
        return fail(ctx, BS_ERR_INVALID, "reads_out holds %lld bytes, %lld needed (bs_simulate_capacity)", (long long)reads_capacity,
                    (long long)(n_draws * (int64_t)read_len));
    if (n_draws > 0 && !reads_out) return fail(ctx, BS_ERR_INVALID, "reads_out is NULL");
    cudaSetDevice(ctx->device);
    BS_TRY(prepare_ks(ctx, kmer));
    cudaStream_t st = ctx->stream;
    const bool dev_chars = (flags & BS_DEVICE_CHARS) != 0, dev_out = (flags & BS_DEVICE_RESULT) != 0;

    MetaBuilder mb;
    const size_t o_tr_off = mb.add(tr_off.data(), (size_t)S + 1);
    const size_t o_tr_woff = mb.add(tr_woff.data(), (size_t)S + 1);
    const size_t o_draw_off = mb.add(draw_off.data(), (size_t)S + 1);
    const size_t o_srs = mb.add(draw_off.data(), (size_t)S + 1);  // placeholder, rewritten below
    BS_TRY(ensure(ctx, ctx->sim_meta, mb.bytes.size()));
    if (!dev_chars) BS_TRY(ensure(ctx, ctx->sim_chars, (size_t)tr_bytes + 32));
    BS_TRY(ensure(ctx, ctx->sim_words, (size_t)tr_woff[S] * 8 + 8));
    BS_TRY(ensure(ctx, ctx->sim_mask, (size_t)tr_woff[S] * 4 + 8));
    BS_TRY(ensure(ctx, ctx->sim_cdf, (size_t)std::max<int64_t>(tr_bytes, 1) * 8));
    BS_TRY(ensure(ctx, ctx->sim_starts, (size_t)std::max<int64_t>(n_draws, 1) * 4));
    BS_TRY(ensure(ctx, ctx->sim_kept, (size_t)S * 4));
    if (!dev_out) BS_TRY(ensure(ctx, ctx->sim_reads, (size_t)std::max<int64_t>(n_draws * read_len, 1)));
    BS_CUDA(cudaMemcpyAsync(ctx->sim_meta.p, mb.bytes.data(), mb.bytes.size(), cudaMemcpyHostToDevice, st));
    const uint8_t *d_tchars = (const uint8_t *)truth_chars + truth_off[0];
    if (!dev_chars) {
        if (tr_bytes) BS_CUDA(cudaMemcpyAsync(ctx->sim_chars.p, truth_chars + truth_off[0], (size_t)tr_bytes, cudaMemcpyHostToDevice, st));
        d_tchars = (const uint8_t *)ctx->sim_chars.p;
    }
    BS_CUDA(cudaStreamSynchronize(st));  // mb.bytes is pageable and reused below
    unsigned char *dm = (unsigned char *)ctx->sim_meta.p;
    const int grid_cap = ctx->sm_count * 32;
    bs::SeqSet ts{d_tchars, (const int64_t *)(dm + o_tr_off), (const int64_t *)(dm + o_tr_woff), (uint64_t *)ctx->sim_words.p,
                  (uint32_t *)ctx->sim_mask.p, S, tr_woff[S]};
    bs::SimArgs sa;
    sa.tr_off = ts.off; sa.tr_woff = ts.woff; sa.tr_words = ts.words; sa.tr_mask = ts.mask; sa.tr_chars = d_tchars;
    sa.win = (const bs::WinEntry *)ctx->ks.win.p; sa.kmer = kmer; sa.read_len = read_len; sa.n_seg = S; sa.seg_base = 0; sa.seed = seed;
    sa.draw_off = (const int64_t *)(dm + o_draw_off); sa.cdf = (double *)ctx->sim_cdf.p; sa.starts = (int32_t *)ctx->sim_starts.p;
    sa.kept = (int32_t *)ctx->sim_kept.p; sa.seg_read_start = (const int64_t *)(dm + o_srs);
    sa.reads = dev_out ? (uint8_t *)reads_out : (uint8_t *)ctx->sim_reads.p;
    BS_LAUNCH(bs::k_pack_seqs, grid_for(ts.total_words, kPackThreads, grid_cap), kPackThreads, 0, st, ts);
    BS_LAUNCH(bs::k_sim_cdf, (unsigned)S, kSimThreads, 0, st, sa);
    if (n_draws > 0) BS_LAUNCH(bs::k_sim_draw, grid_for(n_draws, kSimThreads, grid_cap), kSimThreads, 0, st, sa);
    BS_LAUNCH(bs::k_sim_compact, (unsigned)S, kSimThreads, 0, st, sa);
    ctx->launches += 3 + (n_draws > 0 ? 1 : 0);
    std::vector<int32_t> kept((size_t)S);
    BS_CUDA(cudaMemcpyAsync(kept.data(), ctx->sim_kept.p, (size_t)S * 4, cudaMemcpyDeviceToHost, st));
    BS_CUDA(cudaStreamSynchronize(st));
    for (int64_t s = 0; s < S; s++) seg_read_start[s + 1] = seg_read_start[s] + kept[(size_t)s];
    const int64_t n_reads = seg_read_start[S];
    BS_CUDA(cudaMemcpyAsync(dm + o_srs, seg_read_start, (size_t)(S + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_reads > 0) {
        BS_LAUNCH(bs::k_sim_emit, grid_for(n_reads * 32, kSimThreads, grid_cap), kSimThreads, 0, st, sa);
        ctx->launches++;
        if (!dev_out) BS_CUDA(cudaMemcpyAsync(reads_out, ctx->sim_reads.p, (size_t)(n_reads * read_len), cudaMemcpyDeviceToHost, st));
    }
    BS_CUDA(cudaGetLastError());
    BS_CUDA(cudaStreamSynchronize(st));  // seg_read_start is the caller's (pageable) array; results are complete on return
    return BS_OK;
}

}  // extern "C"
