// C-ABI of the breakage scorer (include/breakscore.h): context, resident tables, work buffers
// and the per-call orchestration of the kernels in bs_kernels.cuh.  Host C++ only talks to
// the CUDA runtime; nothing here depends on torch, Python or R.
#ifdef BS_CPU_EMUL
#include "cuda_emul.h"  // tests/emul: CPU emulation used by the CPU-only test-suite, never shipped
#endif
#include "bs_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/breakscore.h"

namespace {

char g_create_error[512] = "";

#ifdef BS_CPU_EMUL
constexpr int kPlaceThreads = 64;
constexpr int kScoreThreads = 64;
constexpr int kKsThreads = 64;
constexpr int kPackThreads = 64;
#else
constexpr int kPlaceThreads = 256;
constexpr int kScoreThreads = 256;
constexpr int kKsThreads = 1024;
constexpr int kPackThreads = 256;
#endif
constexpr int kMaxTile = 8192;        // contig positions per shared-memory tile
constexpr int64_t kMaxChunk = 32768;  // reads per placement work item (bitmask of 4 KB)
constexpr int64_t kMinChunk = 2048;

enum Stage { ST_H2D, ST_PACK, ST_PLACE, ST_SCORE, ST_SPECTRUM, ST_PROBDIST, ST_PATHFREQ, ST_STARTPOS, ST_D2H, ST_COUNT };

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct KsCache {
    int kmer = -1;
    uint64_t table_version = 0;
    int R_x = 0, R_y = 0, rank_zero = 0;
    DevBuf rank_x, rank_y, le_idx, lt_idx, yv;
};

}  // namespace

struct bs_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    char err[512] = "";
    int64_t launches = 0;

    // table
    bool has_table = false;
    uint64_t table_version = 0;
    int64_t T = 0;
    std::vector<double> prob_dense;    // [DENSE_SIZE]
    std::vector<double> tprob_dense;   // [DENSE_SIZE] truth-side values (follows prob_dense unless set)
    std::vector<int32_t> row_dense;    // [DENSE_SIZE]
    bool has_truth_table = false;
    DevBuf d_tab_prob, d_tab_row;
    KsCache ks;

    // staging + work buffers (grow-only)
    void *h_meta = nullptr;
    size_t h_meta_cap = 0;
    cudaEvent_t meta_done = nullptr;
    bool meta_pending = false;
    DevBuf d_meta, d_read_chars, d_read_off, d_ctg_chars, d_tr_chars;
    DevBuf d_rwords, d_rflags, d_cwords, d_cmask, d_twords, d_tmask;
    DevBuf d_w, d_total, d_ycnt, d_scratch, d_ovf, d_status, d_rank_scratch;
    DevBuf d_out_i32, d_out_f64, d_pd, d_hist, d_pos;

    // timing
    bool timing = false;
    cudaEvent_t ev[ST_COUNT][2] = {};
    bool ev_used[ST_COUNT] = {};
    double last_ms[ST_COUNT] = {};
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

// host-side assembly of the small per-call metadata arrays into one pinned block, one H2D
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
    Stage st;
    StageTimer(bs_ctx *c, Stage s) : ctx(c), st(s) {
        if (ctx->timing) { cudaEventRecord(ctx->ev[st][0], ctx->stream); }
    }
    ~StageTimer() {
        if (ctx->timing) { cudaEventRecord(ctx->ev[st][1], ctx->stream); ctx->ev_used[st] = true; }
    }
};

int grid_for(int64_t work_items, int threads, int cap) {
    int64_t g = (work_items + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

int upload_table(bs_ctx *ctx) {
    int rc;
    if ((rc = ensure(ctx, ctx->d_tab_prob, bs::DENSE_SIZE * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, ctx->d_tab_row, bs::DENSE_SIZE * sizeof(int32_t)))) return rc;
    BS_CUDA(cudaMemcpyAsync(ctx->d_tab_prob.p, ctx->prob_dense.data(), bs::DENSE_SIZE * sizeof(double),
                            cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaMemcpyAsync(ctx->d_tab_row.p, ctx->row_dense.data(), bs::DENSE_SIZE * sizeof(int32_t),
                            cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaStreamSynchronize(ctx->stream));
    return BS_OK;
}

// sorted distinct values + rank maps for the KS statistics of window length kmer
int prepare_ks(bs_ctx *ctx, int kmer) {
    KsCache &k = ctx->ks;
    if (k.kmer == kmer && k.table_version == ctx->table_version) return BS_OK;
    const bool in_range = kmer >= 1 && kmer <= bs::MAXK;
    const int ncode = in_range ? 1 << (2 * kmer) : 1;
    const int doff = in_range ? bs::dense_offset(kmer) : 0;
    std::vector<double> xv{0.0}, yv;
    if (in_range) {
        for (int c = 0; c < ncode; c++)
            if (ctx->row_dense[doff + c] >= 0) {
                xv.push_back(ctx->prob_dense[doff + c]);
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
            rank_x[c] = (int32_t)(std::lower_bound(xv.begin(), xv.end(), ctx->prob_dense[doff + c]) - xv.begin());
            rank_y[c] = (int32_t)(std::lower_bound(yv.begin(), yv.end(), ctx->tprob_dense[doff + c]) - yv.begin());
        }
    }
    std::vector<int32_t> le(xv.size()), lt(xv.size());
    for (size_t i = 0; i < xv.size(); i++) {
        le[i] = (int32_t)(std::upper_bound(yv.begin(), yv.end(), xv[i]) - yv.begin()) - 1;
        lt[i] = (int32_t)(std::lower_bound(yv.begin(), yv.end(), xv[i]) - yv.begin()) - 1;
    }
    if (yv.empty()) yv.push_back(0.0);  // keep the device array non-empty; R_y stays 0
    int rc;
    if ((rc = ensure(ctx, k.rank_x, rank_x.size() * 4))) return rc;
    if ((rc = ensure(ctx, k.rank_y, rank_y.size() * 4))) return rc;
    if ((rc = ensure(ctx, k.le_idx, le.size() * 4))) return rc;
    if ((rc = ensure(ctx, k.lt_idx, lt.size() * 4))) return rc;
    if ((rc = ensure(ctx, k.yv, yv.size() * 8))) return rc;
    BS_CUDA(cudaMemcpyAsync(k.rank_x.p, rank_x.data(), rank_x.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaMemcpyAsync(k.rank_y.p, rank_y.data(), rank_y.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaMemcpyAsync(k.le_idx.p, le.data(), le.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaMemcpyAsync(k.lt_idx.p, lt.data(), lt.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaMemcpyAsync(k.yv.p, yv.data(), yv.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaStreamSynchronize(ctx->stream));
    k.kmer = kmer;
    k.table_version = ctx->table_version;
    k.R_x = (int)xv.size();
    // number of distinct truth-side values actually present
    int ry = 0;
    for (int c = 0; c < ncode; c++) ry = std::max(ry, rank_y[c] + 1);
    k.R_y = ry;
    k.rank_zero = rank_zero;
    return BS_OK;
}

int check_offsets(bs_ctx *ctx, const char *what, const int64_t *off, int64_t n) {
    if (!off) return fail(ctx, BS_ERR_INVALID, "%s offsets are NULL", what);
    if (off[0] < 0) return fail(ctx, BS_ERR_INVALID, "%s offsets start below 0", what);
    for (int64_t i = 0; i < n; i++)
        if (off[i + 1] < off[i]) return fail(ctx, BS_ERR_INVALID, "%s offsets are not monotone at %lld", what, (long long)i);
    return BS_OK;
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
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->meta_done)) != cudaSuccess) {
        fail(nullptr, BS_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
        delete ctx;
        return BS_ERR_CUDA;
    }
    for (int s = 0; s < ST_COUNT; s++)
        for (int j = 0; j < 2; j++) cudaEventCreate(&ctx->ev[s][j]);
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return BS_OK;
}

void bs_ctx_destroy(bs_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->d_tab_prob, &ctx->d_tab_row, &ctx->ks.rank_x, &ctx->ks.rank_y, &ctx->ks.le_idx,
                      &ctx->ks.lt_idx, &ctx->ks.yv, &ctx->d_meta, &ctx->d_read_chars, &ctx->d_read_off,
                      &ctx->d_ctg_chars, &ctx->d_tr_chars, &ctx->d_rwords, &ctx->d_rflags, &ctx->d_cwords,
                      &ctx->d_cmask, &ctx->d_twords, &ctx->d_tmask, &ctx->d_w, &ctx->d_total, &ctx->d_ycnt,
                      &ctx->d_scratch, &ctx->d_ovf, &ctx->d_status, &ctx->d_rank_scratch, &ctx->d_out_i32, &ctx->d_out_f64,
                      &ctx->d_pd, &ctx->d_hist, &ctx->d_pos};
    for (DevBuf *b : bufs) release(*b);
    if (ctx->h_meta) cudaFreeHost(ctx->h_meta);
    for (int s = 0; s < ST_COUNT; s++)
        for (int j = 0; j < 2; j++)
            if (ctx->ev[s][j]) cudaEventDestroy(ctx->ev[s][j]);
    if (ctx->meta_done) cudaEventDestroy(ctx->meta_done);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int bs_ctx_set_stream(bs_ctx *ctx, void *cuda_stream) {
    if (!ctx) return BS_ERR_INVALID;
    cudaSetDevice(ctx->device);
    BS_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return BS_OK;
}

int bs_ctx_synchronize(bs_ctx *ctx) {
    if (!ctx) return BS_ERR_INVALID;
    cudaSetDevice(ctx->device);
    BS_CUDA(cudaStreamSynchronize(ctx->stream));
    return BS_OK;
}

int64_t bs_ctx_launch_count(const bs_ctx *ctx) { return ctx ? ctx->launches : 0; }

int bs_ctx_enable_timing(bs_ctx *ctx, int on) {
    if (!ctx) return BS_ERR_INVALID;
    ctx->timing = on != 0;
    return BS_OK;
}

// per-stage device time (ms) of the last scoring call; order: h2d, pack, place, score, spectrum,
// prob_dist_ks, ks_path_freq, startpos, d2h.  Returns the number of stages written.
int bs_ctx_last_timings(bs_ctx *ctx, double *ms, int n) {
    if (!ctx || !ms) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    int k = 0;
    for (; k < n && k < ST_COUNT; k++) {
        float t = -1.f;
        if (ctx->timing && ctx->ev_used[k]) {
            if (cudaEventElapsedTime(&t, ctx->ev[k][0], ctx->ev[k][1]) != cudaSuccess) t = -1.f;
        }
        ms[k] = t;
    }
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

int bs_set_table(bs_ctx *ctx, const char *kmer_chars, const int64_t *kmer_off, const double *prob, int64_t n) {
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
    ctx->prob_dense.swap(pd);
    ctx->row_dense.swap(rd);
    if (!ctx->has_truth_table || (int64_t)ctx->T != n) {
        ctx->tprob_dense = ctx->prob_dense;
        ctx->has_truth_table = false;
    }
    ctx->T = n;
    ctx->has_table = true;
    ctx->table_version++;
    return upload_table(ctx);
}

int bs_set_truth_table(bs_ctx *ctx, const double *prob, int64_t n) {
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

int bs_score_batch(bs_ctx *ctx, const bs_batch *b, int kmer, uint32_t flags, bs_result *res) {
    if (!ctx) return BS_ERR_INVALID;
    if (!b || !res) return fail(ctx, BS_ERR_INVALID, "bs_score_batch: NULL batch or result");
    if (!ctx->has_table) return fail(ctx, BS_ERR_STATE, "bs_score_batch before bs_set_table");
    if (kmer < 1) return fail(ctx, BS_ERR_INVALID, "kmer must be >= 1 (got %d)", kmer);
    const int64_t S = b->n_segments, N = b->n_reads, C = b->n_contigs;
    if (S < 0 || N < 0 || C < 0) return fail(ctx, BS_ERR_INVALID, "negative counts");
    if (C > 0x7fffffff || S > 0x7fffffff) return fail(ctx, BS_ERR_INVALID, "too many contigs/segments for one call");
    int rc;
    if ((rc = check_offsets(ctx, "contig", b->contig_off, C))) return rc;
    if ((rc = check_offsets(ctx, "truth", b->truth_off, S))) return rc;
    if ((rc = check_offsets(ctx, "segment read-start", b->seg_read_start, S))) return rc;
    if ((rc = check_offsets(ctx, "segment contig-start", b->seg_contig_start, S))) return rc;
    if (b->seg_read_start[0] != 0 || b->seg_read_start[S] != N || b->seg_contig_start[0] != 0 || b->seg_contig_start[S] != C)
        return fail(ctx, BS_ERR_INVALID, "segment start arrays do not cover all reads/contigs");
    if (b->read_off) { if ((rc = check_offsets(ctx, "read", b->read_off, N))) return rc; }
    else if (b->read_len < 0) return fail(ctx, BS_ERR_INVALID, "read_len < 0");
    if ((flags & BS_WANT_PROB_DIST) && (!res->path_prob_dist || !res->path_prob_dist_off))
        return fail(ctx, BS_ERR_INVALID, "BS_WANT_PROB_DIST needs path_prob_dist and path_prob_dist_off");
    if ((flags & BS_WANT_POS) && (!res->pos || !res->pos_off)) return fail(ctx, BS_ERR_INVALID, "BS_WANT_POS needs pos and pos_off");
    if ((flags & BS_WANT_HIST) && !res->hist) return fail(ctx, BS_ERR_INVALID, "BS_WANT_HIST needs hist");
    if (C == 0) return BS_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    for (int i = 0; i < ST_COUNT; i++) ctx->ev_used[i] = false;
    const bool dev_chars = (flags & BS_DEVICE_CHARS) != 0;
    const bool dev_res = (flags & BS_DEVICE_RESULT) != 0;
    const bool want_ks = (flags & BS_WANT_KS) && (res->ks_stat_prob_dist || res->ks_stat_path_freq);
    const bool want_pd = (flags & BS_WANT_PROB_DIST) != 0;
    const bool want_pos = (flags & BS_WANT_POS) != 0;
    const bool want_hist = (flags & BS_WANT_HIST) != 0;
    const bool want_sp = (flags & BS_WANT_STARTPOS) && res->path_prob_dist_startpos;
    const int64_t T = ctx->T;

    // ---------------- host metadata ----------------
    const int64_t read_bytes = b->read_off ? b->read_off[N] : N * (int64_t)b->read_len;
    const int64_t ctg_bytes = b->contig_off[C];
    const int64_t tr_bytes = b->truth_off[S];
    if (N > 0 && !b->read_chars && read_bytes > 0) return fail(ctx, BS_ERR_INVALID, "read_chars is NULL");
    if (ctg_bytes > 0 && !b->contig_chars) return fail(ctx, BS_ERR_INVALID, "contig_chars is NULL");
    if (tr_bytes > 0 && !b->truth_chars) return fail(ctx, BS_ERR_INVALID, "truth_chars is NULL");

    std::vector<int64_t> ctg_woff(C + 1), tr_woff(S + 1);
    std::vector<int32_t> ctg_seg(C), seed_len(std::max<int64_t>(S, 1));
    int64_t max_ctg = 0;
    ctg_woff[0] = 0;
    for (int64_t c = 0; c < C; c++) {
        const int64_t L = b->contig_off[c + 1] - b->contig_off[c];
        if (L > 0x7ffffff0) return fail(ctx, BS_ERR_INVALID, "contig %lld longer than 2^31", (long long)c);
        max_ctg = std::max(max_ctg, L);
        ctg_woff[c + 1] = ctg_woff[c] + (L + 31) / 32 + 2;
    }
    tr_woff[0] = 0;
    for (int64_t s = 0; s < S; s++) {
        const int64_t L = b->truth_off[s + 1] - b->truth_off[s];
        if (L > 0x7ffffff0) return fail(ctx, BS_ERR_INVALID, "truth %lld longer than 2^31", (long long)s);
        tr_woff[s + 1] = tr_woff[s] + (L + 31) / 32 + 2;
        for (int64_t c = b->seg_contig_start[s]; c < b->seg_contig_start[s + 1]; c++) ctg_seg[c] = (int32_t)s;
    }
    int64_t max_read = 0;
    if (b->read_off) {
        for (int64_t s = 0; s < S; s++) {
            int64_t mn = 32;
            for (int64_t n = b->seg_read_start[s]; n < b->seg_read_start[s + 1]; n++) {
                const int64_t l = b->read_off[n + 1] - b->read_off[n];
                max_read = std::max(max_read, l);
                if (l > 0 && l < mn) mn = l;
            }
            seed_len[s] = (int32_t)mn;
        }
    } else {
        max_read = b->read_len;
        for (int64_t s = 0; s < S; s++) seed_len[s] = (int32_t)std::min<int64_t>(32, std::max<int64_t>(1, b->read_len));
    }
    if (max_read > 0x3fffffff) return fail(ctx, BS_ERR_INVALID, "read longer than 2^30");
    const int W = (int)std::max<int64_t>(1, (max_read + 31) / 32);

    // placement work items: one per (contig, read chunk)
    int tile_len = (int)std::min<int64_t>(kMaxTile, std::max<int64_t>(32, (max_ctg + 31) / 32 * 32));
    int hash_size = 64;
    while (hash_size < tile_len) hash_size <<= 1;
    std::vector<bs::PlaceItem> items;
    items.reserve((size_t)C);
    const int64_t target_items = (int64_t)ctx->sm_count * 8;
    int64_t max_chunk = 1;
    for (int64_t c = 0; c < C; c++) {
        const int32_t s = ctg_seg[c];
        const int64_t r0 = b->seg_read_start[s], r1 = b->seg_read_start[s + 1];
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
    const int found_words = (int)((max_chunk + 31) / 32);
    const size_t place_smem = bs::place_smem_bytes(tile_len, hash_size, W, found_words);
    if (place_smem > ctx->smem_optin) return fail(ctx, BS_ERR_INVALID, "placement tile needs %zu bytes of shared memory", place_smem);

    std::vector<int64_t> w_dummy;
    MetaBuilder mb;
    const size_t o_ctg_off = mb.add(b->contig_off, (size_t)C + 1);
    const size_t o_ctg_woff = mb.add(ctg_woff.data(), (size_t)C + 1);
    const size_t o_ctg_seg = mb.add(ctg_seg.data(), (size_t)C);
    const size_t o_tr_off = mb.add(b->truth_off, (size_t)S + 1);
    const size_t o_tr_woff = mb.add(tr_woff.data(), (size_t)S + 1);
    const size_t o_seg_rs = mb.add(b->seg_read_start, (size_t)S + 1);
    const size_t o_seed = mb.add(seed_len.data(), (size_t)S);
    const size_t o_items = mb.add(items.data(), items.size());
    const size_t o_pd_off = want_pd ? mb.add(res->path_prob_dist_off, (size_t)C + 1) : 0;
    const size_t o_pos_off = want_pos ? mb.add(res->pos_off, (size_t)C + 1) : 0;

    // previous call's metadata copy must have left the staging block
    if (ctx->meta_pending) { BS_CUDA(cudaEventSynchronize(ctx->meta_done)); ctx->meta_pending = false; }
    if (mb.bytes.size() > ctx->h_meta_cap) {
        if (ctx->h_meta) cudaFreeHost(ctx->h_meta);
        ctx->h_meta = nullptr;
        ctx->h_meta_cap = 0;
        const size_t want = mb.bytes.size() * 2 + 4096;
        if (cudaHostAlloc(&ctx->h_meta, want, cudaHostAllocDefault) != cudaSuccess)
            return fail(ctx, BS_ERR_ALLOC, "cudaHostAlloc(%zu) failed", want);
        ctx->h_meta_cap = want;
    }
    std::memcpy(ctx->h_meta, mb.bytes.data(), mb.bytes.size());

    // ---------------- device buffers ----------------
    const int64_t n_out_i32 = 3 * C, n_out_f64 = 5 * C;
    const int64_t w_elems = ctg_bytes + C;
    const int ks_R_y_max = 65537;
    (void)ks_R_y_max;
    if ((rc = ensure(ctx, ctx->d_meta, mb.bytes.size()))) return rc;
    if (!dev_chars) {
        if ((rc = ensure(ctx, ctx->d_read_chars, (size_t)read_bytes + 16))) return rc;
        if ((rc = ensure(ctx, ctx->d_ctg_chars, (size_t)ctg_bytes + 16))) return rc;
        if ((rc = ensure(ctx, ctx->d_tr_chars, (size_t)tr_bytes + 16))) return rc;
    }
    if (b->read_off) { if ((rc = ensure(ctx, ctx->d_read_off, (size_t)(N + 1) * 8))) return rc; }
    if ((rc = ensure(ctx, ctx->d_rwords, (size_t)std::max<int64_t>(N, 1) * W * 8))) return rc;
    if ((rc = ensure(ctx, ctx->d_rflags, (size_t)std::max<int64_t>(N, 1)))) return rc;
    if ((rc = ensure(ctx, ctx->d_cwords, (size_t)ctg_woff[C] * 8))) return rc;
    if ((rc = ensure(ctx, ctx->d_cmask, (size_t)ctg_woff[C] * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_twords, (size_t)tr_woff[S] * 8 + 8))) return rc;
    if ((rc = ensure(ctx, ctx->d_tmask, (size_t)tr_woff[S] * 4 + 8))) return rc;
    if ((rc = ensure(ctx, ctx->d_w, (size_t)w_elems * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_total, (size_t)C * 4))) return rc;
    if (!dev_res) {
        if ((rc = ensure(ctx, ctx->d_out_i32, (size_t)n_out_i32 * 4))) return rc;
        if ((rc = ensure(ctx, ctx->d_out_f64, (size_t)n_out_f64 * 8))) return rc;
        if (want_pd) { if ((rc = ensure(ctx, ctx->d_pd, (size_t)std::max<int64_t>(res->path_prob_dist_off[C], 1) * 8))) return rc; }
        if (want_hist) { if ((rc = ensure(ctx, ctx->d_hist, (size_t)C * (T + 1) * 4))) return rc; }
        if (want_pos) { if ((rc = ensure(ctx, ctx->d_pos, (size_t)std::max<int64_t>(res->pos_off[C], 1) * 4))) return rc; }
    }
    if (want_ks && (rc = prepare_ks(ctx, kmer))) return rc;

    unsigned char *dm = (unsigned char *)ctx->d_meta.p;
    const int64_t *d_ctg_off = (const int64_t *)(dm + o_ctg_off);
    const int64_t *d_ctg_woff = (const int64_t *)(dm + o_ctg_woff);
    const int32_t *d_ctg_seg = (const int32_t *)(dm + o_ctg_seg);
    const int64_t *d_tr_off = (const int64_t *)(dm + o_tr_off);
    const int64_t *d_tr_woff = (const int64_t *)(dm + o_tr_woff);
    const int64_t *d_seg_rs = (const int64_t *)(dm + o_seg_rs);
    const int32_t *d_seed = (const int32_t *)(dm + o_seed);
    const bs::PlaceItem *d_items = (const bs::PlaceItem *)(dm + o_items);
    const int64_t *d_pd_off = want_pd ? (const int64_t *)(dm + o_pd_off) : nullptr;
    const int64_t *d_pos_off = want_pos ? (const int64_t *)(dm + o_pos_off) : nullptr;

    const uint8_t *d_rchars, *d_cchars, *d_tchars;
    {
        StageTimer tm(ctx, ST_H2D);
        BS_CUDA(cudaMemcpyAsync(ctx->d_meta.p, ctx->h_meta, mb.bytes.size(), cudaMemcpyHostToDevice, st));
        BS_CUDA(cudaEventRecord(ctx->meta_done, st));
        ctx->meta_pending = true;
        if (dev_chars) {
            d_rchars = (const uint8_t *)b->read_chars;
            d_cchars = (const uint8_t *)b->contig_chars;
            d_tchars = (const uint8_t *)b->truth_chars;
        } else {
            if (read_bytes) BS_CUDA(cudaMemcpyAsync(ctx->d_read_chars.p, b->read_chars, (size_t)read_bytes, cudaMemcpyHostToDevice, st));
            if (ctg_bytes) BS_CUDA(cudaMemcpyAsync(ctx->d_ctg_chars.p, b->contig_chars, (size_t)ctg_bytes, cudaMemcpyHostToDevice, st));
            if (tr_bytes) BS_CUDA(cudaMemcpyAsync(ctx->d_tr_chars.p, b->truth_chars, (size_t)tr_bytes, cudaMemcpyHostToDevice, st));
            d_rchars = (const uint8_t *)ctx->d_read_chars.p;
            d_cchars = (const uint8_t *)ctx->d_ctg_chars.p;
            d_tchars = (const uint8_t *)ctx->d_tr_chars.p;
        }
        if (b->read_off) BS_CUDA(cudaMemcpyAsync(ctx->d_read_off.p, b->read_off, (size_t)(N + 1) * 8, cudaMemcpyHostToDevice, st));
    }

    // result destinations on the device
    int32_t *o_len, *o_breaks, *o_startpos, *o_hist = nullptr, *o_pos = nullptr;
    double *o_score, *o_norm, *o_bylen, *o_ksa, *o_ksb, *o_pd = nullptr;
    if (dev_res) {
        o_len = res->sequence_len; o_breaks = res->kmer_breaks; o_startpos = want_sp ? res->path_prob_dist_startpos : nullptr;
        o_score = res->bp_score; o_norm = res->bp_score_norm_by_break_freqs; o_bylen = res->bp_score_norm_by_len;
        o_ksa = want_ks ? res->ks_stat_prob_dist : nullptr; o_ksb = want_ks ? res->ks_stat_path_freq : nullptr;
        if (want_pd) o_pd = res->path_prob_dist;
        if (want_hist) o_hist = res->hist;
        if (want_pos) o_pos = res->pos;
    } else {
        int32_t *i32 = (int32_t *)ctx->d_out_i32.p;
        double *f64 = (double *)ctx->d_out_f64.p;
        o_len = i32; o_breaks = i32 + C; o_startpos = want_sp ? i32 + 2 * C : nullptr;
        o_score = f64; o_norm = f64 + C; o_bylen = f64 + 2 * C;
        o_ksa = (want_ks && res->ks_stat_prob_dist) ? f64 + 3 * C : nullptr;
        o_ksb = (want_ks && res->ks_stat_path_freq) ? f64 + 4 * C : nullptr;
        if (want_pd) o_pd = (double *)ctx->d_pd.p;
        if (want_hist) o_hist = (int32_t *)ctx->d_hist.p;
        if (want_pos) o_pos = (int32_t *)ctx->d_pos.p;
    }

    BS_CUDA(cudaMemsetAsync(ctx->d_w.p, 0, (size_t)w_elems * 4, st));
    BS_CUDA(cudaMemsetAsync(ctx->d_total.p, 0, (size_t)C * 4, st));
    if (o_hist) BS_CUDA(cudaMemsetAsync(o_hist, 0, (size_t)C * (T + 1) * 4, st));
    if (o_pos && res->pos_off[C] > 0) BS_CUDA(cudaMemsetAsync(o_pos, 0xff, (size_t)res->pos_off[C] * 4, st));

    // ---------------- kernels ----------------
    const int grid_cap = ctx->sm_count * 32;
    bs::SeqSet cs{d_cchars, d_ctg_off, d_ctg_woff, (uint64_t *)ctx->d_cwords.p, (uint32_t *)ctx->d_cmask.p, C, ctg_woff[C]};
    bs::SeqSet ts{d_tchars, d_tr_off, d_tr_woff, (uint64_t *)ctx->d_twords.p, (uint32_t *)ctx->d_tmask.p, S, tr_woff[S]};
    bs::ReadSet rs{d_rchars, b->read_off ? (const int64_t *)ctx->d_read_off.p : nullptr, b->read_len, W, N,
                   (uint64_t *)ctx->d_rwords.p, (uint8_t *)ctx->d_rflags.p};
    {
        StageTimer tm(ctx, ST_PACK);
        BS_LAUNCH(bs::k_pack_seqs, grid_for(cs.total_words, kPackThreads, grid_cap), kPackThreads, 0, st, cs);
        ctx->launches++;
        if (want_ks || want_sp) {
            BS_LAUNCH(bs::k_pack_seqs, grid_for(ts.total_words, kPackThreads, grid_cap), kPackThreads, 0, st, ts);
            ctx->launches++;
        }
        if (N > 0) {
            BS_LAUNCH(bs::k_pack_reads, grid_for(N, kPackThreads, grid_cap), kPackThreads, 0, st, rs);
            ctx->launches++;
        }
    }
    {
        StageTimer tm(ctx, ST_PLACE);
        if (N > 0 && !items.empty()) {
            bs::PlaceArgs pa;
            pa.items = d_items;
            pa.ctg_off = d_ctg_off; pa.ctg_woff = d_ctg_woff;
            pa.ctg_words = cs.words; pa.ctg_mask = cs.mask; pa.ctg_chars = d_cchars;
            pa.reads = rs;
            pa.seg_seed_len = d_seed; pa.seg_read_start = d_seg_rs;
            pa.w = (int32_t *)ctx->d_w.p; pa.total = (int32_t *)ctx->d_total.p;
            pa.pos = o_pos; pa.pos_off = d_pos_off;
            pa.tile_len = tile_len; pa.hash_size = hash_size; pa.found_words = found_words;
            pa.scan_mode = (flags & BS_PLACE_SCAN) ? 1 : 0;
            BS_CUDA(cudaFuncSetAttribute(bs::k_place, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)place_smem));
            BS_LAUNCH(bs::k_place, (unsigned)items.size(), kPlaceThreads, place_smem, st, pa);
            ctx->launches++;
        }
    }
    {
        StageTimer tm(ctx, ST_SCORE);
        bs::ScoreArgs sa;
        sa.ctg_off = d_ctg_off; sa.ctg_woff = d_ctg_woff; sa.ctg_words = cs.words; sa.ctg_mask = cs.mask;
        sa.w = (const int32_t *)ctx->d_w.p; sa.total = (const int32_t *)ctx->d_total.p;
        sa.tab_prob = (const double *)ctx->d_tab_prob.p; sa.tab_row = (const int32_t *)ctx->d_tab_row.p;
        sa.kmer = kmer; sa.T = (int32_t)T; sa.n_contigs = C;
        sa.sequence_len = o_len; sa.bp_score = o_score; sa.norm_by_break_freqs = o_norm; sa.norm_by_len = o_bylen;
        sa.kmer_breaks = o_breaks; sa.hist = o_hist;
        BS_LAUNCH(bs::k_break_score, (unsigned)std::min<int64_t>(C, grid_cap), kScoreThreads, kScoreThreads * 8 + 16, st, sa);
        ctx->launches++;
    }
    if (want_ks) {
        const KsCache &k = ctx->ks;
        const int R_y = std::max(k.R_y, 1);
        if ((rc = ensure(ctx, ctx->d_ycnt, (size_t)S * R_y * 4))) return rc;
        {
            StageTimer tm(ctx, ST_SPECTRUM);
            BS_CUDA(cudaMemsetAsync(ctx->d_ycnt.p, 0, (size_t)S * R_y * 4, st));
            if (k.R_y > 0) {
                bs::SpectrumArgs sp;
                sp.tr_off = d_tr_off; sp.tr_woff = d_tr_woff; sp.tr_words = ts.words; sp.tr_mask = ts.mask;
                sp.rank_y = (const int32_t *)k.rank_y.p; sp.ycnt = (int32_t *)ctx->d_ycnt.p;
                sp.R_y = R_y; sp.kmer = kmer;
                int64_t max_tr = 0;
                for (int64_t s = 0; s < S; s++) max_tr = std::max(max_tr, b->truth_off[s + 1] - b->truth_off[s]);
                sp.blocks_per_seg = (int)std::max<int64_t>(1, std::min<int64_t>(64, (max_tr + kScoreThreads * 8 - 1) / (kScoreThreads * 8)));
                BS_LAUNCH(bs::k_truth_spectrum, (unsigned)(S * sp.blocks_per_seg), kScoreThreads, 0, st, sp);
                ctx->launches++;
                BS_LAUNCH(bs::k_row_cumsum, (unsigned)S, kScoreThreads, kScoreThreads * 8 + 16, st, (int32_t *)ctx->d_ycnt.p, R_y);
                ctx->launches++;
            }
        }
    }
    if (want_pd || (want_ks && o_ksa)) {
        StageTimer tm(ctx, ST_PROBDIST);
        const KsCache &k = ctx->ks;
        bs::ProbDistArgs pa;
        pa.ctg_off = d_ctg_off; pa.ctg_woff = d_ctg_woff; pa.ctg_words = cs.words; pa.ctg_mask = cs.mask;
        pa.ctg_seg = d_ctg_seg;
        pa.tab_prob = (const double *)ctx->d_tab_prob.p; pa.tab_row = (const int32_t *)ctx->d_tab_row.p;
        const bool ks_a = want_ks && o_ksa;
        pa.rank_x = ks_a ? (const int32_t *)k.rank_x.p : nullptr;
        pa.le_idx = ks_a ? (const int32_t *)k.le_idx.p : nullptr;
        pa.lt_idx = ks_a ? (const int32_t *)k.lt_idx.p : nullptr;
        pa.ycum = ks_a ? (const int32_t *)ctx->d_ycnt.p : nullptr;
        pa.R_x = ks_a ? k.R_x : 0; pa.R_y = ks_a ? k.R_y : 0; pa.rank_zero = ks_a ? k.rank_zero : 0;
        pa.kmer = kmer; pa.n_contigs = C;
        pa.prob_dist = o_pd; pa.pd_off = d_pd_off; pa.ks = ks_a ? o_ksa : nullptr;
        // rank histogram: shared memory when it fits (real table: 32 897 ranks = 129 KB), else a
        // per-block global scratch that stays in L2 (all-distinct tables: 65 537 ranks)
        bool in_smem = ks_a && bs::probdist_smem_bytes(pa.R_x, kKsThreads, true) + 1024 <= ctx->smem_optin;
        const size_t smem = bs::probdist_smem_bytes(pa.R_x, kKsThreads, in_smem);
        BS_CUDA(cudaFuncSetAttribute(bs::k_prob_dist_ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = in_smem ? std::max<int>(1, (int)(ctx->smem_optin / (smem + 1024))) : 2;
        const int nblk = (int)std::min<int64_t>(C, (int64_t)ctx->sm_count * per_sm);
        pa.rank_scratch = nullptr;
        if (ks_a && !in_smem) {
            if ((rc = ensure(ctx, ctx->d_rank_scratch, (size_t)nblk * pa.R_x * 4))) return rc;
            pa.rank_scratch = (int32_t *)ctx->d_rank_scratch.p;
        }
        BS_LAUNCH(bs::k_prob_dist_ks, (unsigned)nblk, kKsThreads, smem, st, pa);
        ctx->launches++;
    }
    if (want_ks && o_ksb) {
        StageTimer tm(ctx, ST_PATHFREQ);
        const KsCache &k = ctx->ks;
        const int nblk = (int)std::min<int64_t>(C, (int64_t)ctx->sm_count * 4);
        const size_t scratch_bytes = (size_t)nblk * (T + 1) * 4;
        if (ctx->d_scratch.cap < scratch_bytes || !ctx->d_scratch.p) {
            if ((rc = ensure(ctx, ctx->d_scratch, scratch_bytes))) return rc;
            BS_CUDA(cudaMemsetAsync(ctx->d_scratch.p, 0, ctx->d_scratch.cap, st));
        }
        if ((rc = ensure(ctx, ctx->d_ovf, (size_t)nblk * bs::OVF_CAP * 4))) return rc;
        if ((rc = ensure(ctx, ctx->d_status, 16))) return rc;
        BS_CUDA(cudaMemsetAsync(ctx->d_status.p, 0, 16, st));
        bs::PathFreqArgs fa;
        fa.ctg_off = d_ctg_off; fa.ctg_woff = d_ctg_woff; fa.ctg_words = cs.words; fa.ctg_mask = cs.mask;
        fa.ctg_seg = d_ctg_seg; fa.w = (const int32_t *)ctx->d_w.p; fa.total = (const int32_t *)ctx->d_total.p;
        fa.tab_row = (const int32_t *)ctx->d_tab_row.p; fa.yv = (const double *)k.yv.p; fa.ycum = (const int32_t *)ctx->d_ycnt.p;
        fa.scratch = (int32_t *)ctx->d_scratch.p; fa.ovf_cnt = (int32_t *)ctx->d_ovf.p; fa.status = (int32_t *)ctx->d_status.p;
        fa.R_y = k.R_y; fa.kmer = kmer; fa.T = (int32_t)T; fa.n_contigs = C; fa.ks = o_ksb;
        BS_LAUNCH(bs::k_ks_path_freq, (unsigned)nblk, kScoreThreads, kScoreThreads * 8 + bs::CC_DENSE * 4 + 16, st, fa);
        ctx->launches++;
    }
    if (want_sp) {
        StageTimer tm(ctx, ST_STARTPOS);
        bs::StartposArgs sa;
        sa.ctg_off = d_ctg_off; sa.ctg_woff = d_ctg_woff; sa.ctg_words = cs.words; sa.ctg_mask = cs.mask; sa.ctg_chars = d_cchars;
        sa.ctg_seg = d_ctg_seg; sa.tr_off = d_tr_off; sa.tr_woff = d_tr_woff; sa.tr_words = ts.words; sa.tr_mask = ts.mask;
        sa.tr_chars = d_tchars; sa.total = (const int32_t *)ctx->d_total.p; sa.n_contigs = C; sa.startpos = o_startpos;
        BS_LAUNCH(bs::k_startpos, (unsigned)std::min<int64_t>(C, grid_cap), kScoreThreads, 0, st, sa);
        ctx->launches++;
    }
    BS_CUDA(cudaGetLastError());

    // ---------------- results ----------------
    if (dev_res) {
        if (res->lev_dist_vs_true) BS_CUDA(cudaMemsetAsync(res->lev_dist_vs_true, 0, (size_t)C * 4, st));
        return BS_OK;  // asynchronous: the caller orders later work on the same stream
    }
    {
        StageTimer tm(ctx, ST_D2H);
        auto d2h = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
            if (!dst || !src || !bytes) return cudaSuccess;
            return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
        };
        BS_CUDA(d2h(res->sequence_len, o_len, (size_t)C * 4));
        BS_CUDA(d2h(res->kmer_breaks, o_breaks, (size_t)C * 4));
        BS_CUDA(d2h(res->path_prob_dist_startpos, o_startpos, (size_t)C * 4));
        BS_CUDA(d2h(res->bp_score, o_score, (size_t)C * 8));
        BS_CUDA(d2h(res->bp_score_norm_by_break_freqs, o_norm, (size_t)C * 8));
        BS_CUDA(d2h(res->bp_score_norm_by_len, o_bylen, (size_t)C * 8));
        BS_CUDA(d2h(res->ks_stat_prob_dist, o_ksa, (size_t)C * 8));
        BS_CUDA(d2h(res->ks_stat_path_freq, o_ksb, (size_t)C * 8));
        if (want_pd) BS_CUDA(d2h(res->path_prob_dist, o_pd, (size_t)res->path_prob_dist_off[C] * 8));
        if (want_hist) BS_CUDA(d2h(res->hist, o_hist, (size_t)C * (T + 1) * 4));
        if (want_pos) BS_CUDA(d2h(res->pos, o_pos, (size_t)res->pos_off[C] * 4));
    }
    int32_t status = 0;
    if (want_ks && o_ksb) BS_CUDA(cudaMemcpyAsync(&status, ctx->d_status.p, 4, cudaMemcpyDeviceToHost, st));
    BS_CUDA(cudaStreamSynchronize(st));
    ctx->meta_pending = false;
    if (res->lev_dist_vs_true) std::memset(res->lev_dist_vs_true, 0, (size_t)C * 4);
    if (want_sp == false && res->path_prob_dist_startpos) std::memset(res->path_prob_dist_startpos, 0, (size_t)C * 4);
    if (status) return fail(ctx, BS_ERR_INVALID, "ks_stat_path_freq: more than %d table rows with a count >= %d in one contig", bs::OVF_CAP, bs::CC_DENSE);
    return BS_OK;
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

}  // extern "C"
