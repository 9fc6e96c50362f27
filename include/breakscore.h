/*
 * breakscore.h -- C-ABI of the B200-native breakage scorer.
 *
 * Drop-in boundary for ONE path of SahakyanLab/GenomeAssembler_dev: the Rcpp-exported
 *     Rcpp::List calc_breakscore(path, sequencing_reads, true_solution, kmer, bp_kmer, bp_prob)
 * of upstream lib/BreakageScorer.cpp:185-191 (called from lib/DeNovoAssembler.R:348-355).
 * The entry points below are what an Rcpp / ctypes / cgo binding for that path binds; every
 * argument is a plain pointer or size, strings travel as flat buffers + offsets
 * (string i of a set = chars[off[i] .. off[i+1])).  The library links cudart only: no torch,
 * no Python, no R.  There is NO CPU fallback: without a usable sm_100 GPU, bs_ctx_create fails.
 *
 * Upstream item replaced                              | entry point
 * ----------------------------------------------------+-----------------------------------------
 * bp_matrix build, lib/BreakageScorer.cpp:194-197     | bs_set_table
 * kmer_from_seq table, lib/GenerateReads.R:243-259    | bs_set_truth_table (KS y-side)
 * calc_breakscore body, lib/BreakageScorer.cpp:200-353| bs_score (one segment) / bs_score_batch
 * ks.test statistic, lib/DeNovoAssembler.R:416-424    | BS_WANT_KS outputs of the same calls
 * second pass with rep(1/n, n), DeNovoAssembler.R:325-333 | bs_set_second_table + BS_WANT_SECOND_TABLE (one call, one placement)
 * Rcpp::stop / R error                                | int status + bs_last_error
 * assemble_contigs, lib/BreakageScorer.cpp:79-174     | bs_assemble_contigs (host; candidate generator)
 * assemble_contigs + calc_breakscore on its output    | bs_assemble_scaffolds + bs_score_scaffolds: the candidates stay a list of
 *   (lib/DeNovoAssembler.R:343-355)                   |   (base contig, overlap) parts and are scored from their parts
 * generate_sequencing_reads, lib/GenerateReads.R:234-379 | bs_simulate_reads (device; the step before the scorer)
 */
#ifndef BREAKSCORE_H
#define BREAKSCORE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BS_ABI_VERSION 5

/* the library is built with -fvisibility=hidden: only these entry points are exported */
#if defined(__GNUC__)
#define BS_API __attribute__((visibility("default")))
#else
#define BS_API
#endif

typedef struct bs_ctx bs_ctx;

/* status codes (0 = ok).  Nothing throws across this boundary. */
enum {
    BS_OK = 0,
    BS_ERR_INVALID = 1,   /* bad argument (NULL pointer, negative size, offsets not monotone ...) */
    BS_ERR_CUDA = 2,      /* a CUDA runtime call failed; bs_last_error has the text */
    BS_ERR_TABLE = 3,     /* table key outside the supported domain (ACGT, length 1..8) */
    BS_ERR_ALLOC = 4,     /* host or device allocation failed */
    BS_ERR_NO_DEVICE = 5, /* no usable GPU: there is no CPU fallback */
    BS_ERR_STATE = 6,     /* e.g. scoring before bs_set_table */
    BS_ERR_INTERRUPTED = 7 /* the poll callback of bs_ctx_set_poll asked the call to stop */
};

/* flags of bs_score / bs_score_batch */
#define BS_WANT_PROB_DIST 0x001u /* fill path_prob_dist (lib/BreakageScorer.cpp:200-215,352) */
#define BS_WANT_KS        0x002u /* fill ks_stat_prob_dist and ks_stat_path_freq */
#define BS_WANT_HIST      0x004u /* fill hist: dense per-contig break-k-mer counts (parity/debug) */
#define BS_WANT_POS       0x008u /* fill pos: leftmost match of every read in every contig (parity/debug) */
#define BS_WANT_STARTPOS  0x010u /* fill path_prob_dist_startpos (lib/BreakageScorer.cpp:273-274) */
#define BS_WANT_LEV       0x020u /* fill lev_dist_vs_true: infix edit distance contig vs truth (lib/BreakageScorer.cpp:41-55,339) */
#define BS_WANT_SECOND_TABLE 0x040u /* also fill the *2 members of bs_result from the table of bs_set_second_table */
#define BS_PLACE_SCAN     0x100u /* placement by exhaustive all-pairs scan of the contig tile in shared memory (same results) */
#define BS_PLACE_TILE     0x800u /* placement by a seed index over the contig tile, reads streamed past it (same results) */
#define BS_DEVICE_CHARS   0x200u /* read_chars / contig_chars / truth_chars are DEVICE pointers */
#define BS_DEVICE_RESULT  0x400u /* every non-NULL pointer in bs_result is a DEVICE pointer */
/* Two-phase scoring, for ONE job whose READS are sharded over several GPUs (cfg-5: the read set is the big input).
 * The leftmost placement of a read in a contig does not depend on any other read (lib/BreakageScorer.cpp:235-243),
 * so the per-position break counts of disjoint read sets ADD.  Phase 1, every GPU with its own reads and ALL contigs:
 * BS_WEIGHTS_OUT stops after the placement and leaves the position weights in result->weights / weights_total.
 * The caller sums both arrays over the GPUs (NCCL all-reduce).  Phase 2, any split of the contigs: BS_WEIGHTS_IN skips
 * read packing and placement and scores from the given weights (reads of the batch are ignored; n_reads may be 0).
 * Same results as one call with all the reads, bit for bit. */
#define BS_WEIGHTS_OUT    0x1000u
#define BS_WEIGHTS_IN     0x2000u
#define BS_DEFAULT_FLAGS (BS_WANT_PROB_DIST | BS_WANT_KS | BS_WANT_STARTPOS)

/*
 * Input of a scoring call: n_segments independent experiments (upstream: one calc_breakscore
 * call each).  Segment s owns reads [seg_read_start[s], seg_read_start[s+1]) and contigs
 * [seg_contig_start[s], seg_contig_start[s+1]) and the truth string s.
 * All offset / start arrays are HOST memory (metadata).  The three char buffers are host
 * memory unless BS_DEVICE_CHARS is set.
 * Reads: if read_off == NULL every read has read_len bytes and read_chars is dense
 * (read i at i*read_len); otherwise read_off[n_reads+1] gives arbitrary lengths.
 */
typedef struct {
    int64_t n_segments;
    int64_t n_reads;
    int64_t n_contigs;
    const char *read_chars;
    const int64_t *read_off; /* may be NULL (uniform length) */
    int32_t read_len;        /* used when read_off == NULL */
    const char *contig_chars;
    const int64_t *contig_off;       /* [n_contigs+1] */
    const char *truth_chars;
    const int64_t *truth_off;        /* [n_segments+1] */
    const int64_t *seg_read_start;   /* [n_segments+1] */
    const int64_t *seg_contig_start; /* [n_segments+1] */
} bs_batch;

/*
 * Output of a scoring call; arrays are caller-allocated, any pointer may be NULL, all are in
 * input contig order (upstream keeps input order: lib/BreakageScorer.cpp:308-315).
 * Names follow the upstream R list (lib/BreakageScorer.cpp:343-353).
 */
typedef struct {
    int32_t *sequence_len;                  /* [C] */
    double *bp_score;                       /* [C] sum prob*count               (:286-287) */
    double *bp_score_norm_by_break_freqs;   /* [C] sum prob*(count/total)       (:290-292) */
    double *bp_score_norm_by_len;           /* [C] bp_score / sequence_len      (:302-303) */
    int32_t *kmer_breaks;                   /* [C] total_breaks                 (:298)     */
    int32_t *path_prob_dist_startpos;       /* [C] truth.find(contig) if any read hit else 0 */
    int32_t *lev_dist_vs_true;              /* [C] with BS_WANT_LEV: smallest edit distance of the contig to any substring of the truth
                                             *     (edlib EDLIB_MODE_HW semantics); without it: 0 */
    double *ks_stat_prob_dist;              /* [C] KS(path_prob_dist, truth window probabilities) */
    double *ks_stat_path_freq;              /* [C] KS(count/total over table rows, same y); NaN if no breaks */
    double *path_prob_dist;                 /* flat; contig c at path_prob_dist_off[c], max(L_c-kmer+1,0) values */
    const int64_t *path_prob_dist_off;      /* HOST [C+1], required with BS_WANT_PROB_DIST */
    int32_t *hist;                          /* [C][T+1], row-major; bin T = key not in table */
    int32_t *pos;                           /* contig c: n_reads(segment of c) entries at pos_off[c]; -1 = no match */
    const int64_t *pos_off;                 /* HOST [C+1], required with BS_WANT_POS */
    /* BS_WANT_SECOND_TABLE: the table-dependent members again, for the second scoring table, from the
     * same placement (the R driver scores every experiment twice, real table then uniform table:
     * lib/DeNovoAssembler.R:325-333).  Same layout as their namesakes; any may be NULL. */
    double *bp_score2;
    double *bp_score_norm_by_break_freqs2;
    double *bp_score_norm_by_len2;
    double *ks_stat_prob_dist2;
    double *ks_stat_path_freq2;
    double *path_prob_dist2;                /* uses path_prob_dist_off */
    /* BS_WEIGHTS_OUT / BS_WEIGHTS_IN: always DEVICE pointers.  weights: int32, contig c (index within this call) owns
     * L_c + 1 entries starting at (contig_off[c] - contig_off[0]) + c; entry p = number of reads whose leftmost match
     * in the contig is at position p.  weights_total: int32 [C], reads placed per contig (kmer_breaks). */
    int32_t *weights;
    int32_t *weights_total;
} bs_result;

/* context: owns the device, its streams, the resident tables and grow-only work buffers.
 * A context is NOT thread-safe: one host thread at a time (R calls from its single main thread);
 * use one context per thread / per GPU otherwise. */
BS_API int bs_ctx_create(int device, bs_ctx **out);
BS_API void bs_ctx_destroy(bs_ctx *ctx);
/* last error text of ctx (or of the last failed bs_ctx_create when ctx == NULL) */
BS_API const char *bs_last_error(const bs_ctx *ctx);
/* run the kernels on a caller-owned CUDA stream (cudaStream_t as void*); NULL = the context's own
 * stream (so the legacy default stream, handle 0, cannot be selected: create a stream).  Copies of
 * host inputs / results use two internal streams ordered against it by events. */
BS_API int bs_ctx_set_stream(bs_ctx *ctx, void *cuda_stream);
/* block until everything queued by this context has finished */
BS_API int bs_ctx_synchronize(bs_ctx *ctx);
/* number of kernel launches issued by this context so far */
BS_API int64_t bs_ctx_launch_count(const bs_ctx *ctx);
/* per-stage CUDA-event timing (off by default).  bs_ctx_last_timings blocks on the streams and
 * writes up to n stage times (ms, -1 = stage did not run), summed over every scoring call since
 * timing was enabled or last read (the read resets the sums), in the order
 * h2d, pack, place, score (includes the KS of the break histogram), truth_spectrum, prob_dist_ks,
 * ks_path_freq (unused since it was merged into score: always -1), startpos, d2h, lev; returns how
 * many it wrote. */
#define BS_N_STAGES 10
BS_API int bs_ctx_enable_timing(bs_ctx *ctx, int on);
BS_API int bs_ctx_last_timings(bs_ctx *ctx, double *ms, int n);
/* Interrupt poll for long calls (upstream's R driver is interactive: Rcpp::checkUserInterrupt is the idiom,
 * SURVEY.md 8b).  `poll(user)` is called on the CALLING thread between the pipeline chunks of a scoring call
 * (never from another thread, never inside a kernel launch sequence); a non-zero return makes the call stop
 * queueing chunks, wait for the ones in flight and return BS_ERR_INTERRUPTED (result arrays are then partly
 * filled and must not be used).  The context stays usable.  poll == NULL removes the callback. */
typedef int (*bs_poll_fn)(void *user);
BS_API int bs_ctx_set_poll(bs_ctx *ctx, bs_poll_fn poll, void *user);
/* device-event time (ms) of the placement kernel of the last scoring call, -1 if none */
BS_API double bs_ctx_last_place_ms(bs_ctx *ctx);

/* scoring table: n (k-mer, probability) pairs.  Keys: ACGT strings of length 1..8; a repeated
 * key overrides the earlier row (map assignment, lib/BreakageScorer.cpp:195-197).  Setting a table removes
 * a truth-side table and a second table set earlier (both are indexed by the rows of the table they were
 * set against): set them again afterwards. */
BS_API int bs_set_table(bs_ctx *ctx, const char *kmer_chars, const int64_t *kmer_off, const double *prob, int64_t n);
/* probabilities used for the truth-side distribution of the KS statistics (same keys/rows as the
 * scoring table).  prob == NULL: follow the scoring table. */
BS_API int bs_set_truth_table(bs_ctx *ctx, const double *prob, int64_t n);
/* a second scoring table over the SAME rows (prob[i] belongs to row i of bs_set_table); scored in the
 * same call with BS_WANT_SECOND_TABLE.  prob == NULL removes it; bs_set_table removes it too. */
BS_API int bs_set_second_table(bs_ctx *ctx, const double *prob, int64_t n);

BS_API int bs_score_batch(bs_ctx *ctx, const bs_batch *batch, int kmer, uint32_t flags, bs_result *result);

/* one segment, the upstream argument list flattened */
BS_API int bs_score(bs_ctx *ctx, const char *contig_chars, const int64_t *contig_off, int64_t n_contigs,
             const char *read_chars, const int64_t *read_off, int64_t n_reads,
             const char *truth, int64_t truth_len, int kmer, uint32_t flags, bs_result *result);

/*
 * One segment with many candidate contigs / scaffolds, scored by SEVERAL contexts at once -- one per GPU of the box,
 * driven from ONE process and one calling thread, which is how upstream's R driver runs.  Contigs
 * are dealt out to the contexts longest first (LPT), reads and truth go to every context (replicated input), every
 * context runs bs_score on its share from its own host thread, and the results come back in INPUT order: the same
 * bytes as one bs_score call on one context (per-contig reductions have a fixed order).  Every context must hold the
 * same tables (bs_set_table ... on each).  Host buffers only (BS_DEVICE_CHARS / BS_DEVICE_RESULT are refused).
 * On failure the status of the first failing context is returned and its text is copied to bs_last_error(ctxs[0]).
 */
BS_API int bs_score_multi(bs_ctx *const *ctxs, int n_ctx, const char *contig_chars, const int64_t *contig_off, int64_t n_contigs,
             const char *read_chars, const int64_t *read_off, int64_t n_reads,
             const char *truth, int64_t truth_len, int kmer, uint32_t flags, bs_result *result);

/*
 * Scaffold explosion on the host (upstream assemble_contigs, lib/BreakageScorer.cpp:79-174): the
 * generator of the candidate set that bs_score then scores.  n_shuffles: upstream uses 20000;
 * n_threads <= 0: all host cores.  The result is an opaque list of strings (longest first, upstream
 * order), read with bs_string_list_size / _bytes / _copy (chars may be NULL to get offsets only;
 * off has size+1 entries) and released with bs_string_list_free.  Needs no GPU.
 */
typedef struct bs_string_list bs_string_list;
BS_API int bs_assemble_contigs(const char *contig_chars, const int64_t *contig_off, int64_t n_contigs, int dbg_kmer,
                               int seed, int n_shuffles, int n_threads, bs_string_list **out);
BS_API const char *bs_assemble_last_error(void);
BS_API int64_t bs_string_list_size(const bs_string_list *l);
BS_API int64_t bs_string_list_bytes(const bs_string_list *l);
BS_API void bs_string_list_copy(const bs_string_list *l, char *chars, int64_t *off);
BS_API void bs_string_list_free(bs_string_list *l);

/*
 * COMPOSITIONAL scoring of a scaffold set.  The candidates upstream scores after assemble_contigs
 * (lib/BreakageScorer.cpp:105-171) are chains of the same few base contigs, each glued to the scaffold built so far over
 * a suffix/prefix overlap (:118-131).  A scaffold is given here as its PARTS instead of its text:
 *     text(scaffold) = base[part 0] + base[part 1][overlap 1 ..] + base[part 2][overlap 2 ..] + ...
 * where the first overlap_i bases of part i must equal the last overlap_i bases of the text built so far (checked; it is
 * what makes every base contig a whole substring of the scaffold).  bs_score_scaffolds returns exactly what bs_score
 * returns for the materialised texts (integers bit for bit, fp64 sums to 1e-9: another summation order), but the reads
 * are placed ONCE per base contig (and once per distinct junction: the text around it); a scaffold's leftmost positions
 * are then the minimum over the lists of its parts and junctions -- O(reads placed in its parts) per scaffold instead of
 * O(scaffold length) index probes -- and the break k-mers and both KS statistics come from the parts as well.
 * The scaffold texts never cross PCIe (cfg-4: 50 kb of base contigs instead of 324 MB).
 * All arrays of bs_scaffold_set are HOST memory.  One segment per call.  flags: as bs_score; BS_DEVICE_CHARS applies to
 * read_chars / truth only; BS_PLACE_SCAN / BS_PLACE_TILE / BS_WEIGHTS_* are refused.  read_off == NULL: every read has
 * read_len bytes.  path_prob_dist_off / pos_off refer to the scaffold lengths (bs_scaffold_lengths).  The lists are
 * [base contig][read] arrays: with more than 2^31 such cells the texts are materialised and scored like any contig set.
 */
typedef struct {
    int64_t n_base;
    const char *base_chars;
    const int64_t *base_off;             /* [n_base+1] */
    int64_t n_scaffolds;
    const int64_t *scaffold_part_start;  /* [n_scaffolds+1]; scaffold s = parts [start[s], start[s+1]), at least one */
    const int32_t *part_base;            /* base contig of a part */
    const int32_t *part_overlap;         /* bases of the part already present at the end of the scaffold so far; 0 for a first part, < length of the part */
} bs_scaffold_set;

BS_API int bs_score_scaffolds(bs_ctx *ctx, const bs_scaffold_set *set, const char *read_chars, const int64_t *read_off, int64_t n_reads,
                              int32_t read_len, const char *truth, int64_t truth_len, int kmer, uint32_t flags, bs_result *result);
/* lengths[n_scaffolds] of the scaffold texts; returns BS_ERR_INVALID for a malformed set (needs no GPU, no context) */
BS_API int bs_scaffold_lengths(const bs_scaffold_set *set, int64_t *lengths);
/* the texts themselves: off[n_scaffolds+1], chars (may be NULL: offsets only) -- for callers that want the strings upstream returns */
BS_API int bs_scaffold_texts(const bs_scaffold_set *set, char *chars, int64_t *off);
/*
 * bs_assemble_contigs that also reports HOW every scaffold was glued together: the same strings in the same (upstream)
 * order in *out, and their parts in *parts (one composition per string: the first one found).  Read the composition
 * with bs_scaffold_list_parts (n_parts total) / _copy (part_start[n+1], part_base[n_parts], part_overlap[n_parts]).
 */
typedef struct bs_scaffold_list bs_scaffold_list;
BS_API int bs_assemble_scaffolds(const char *contig_chars, const int64_t *contig_off, int64_t n_contigs, int dbg_kmer,
                                 int seed, int n_shuffles, int n_threads, bs_string_list **out, bs_scaffold_list **parts);
BS_API int64_t bs_scaffold_list_size(const bs_scaffold_list *l);
BS_API int64_t bs_scaffold_list_parts(const bs_scaffold_list *l);
BS_API void bs_scaffold_list_copy(const bs_scaffold_list *l, int64_t *part_start, int32_t *part_base, int32_t *part_overlap);
BS_API void bs_scaffold_list_free(bs_scaffold_list *l);

/*
 * Read simulation on the device (upstream lib/GenerateReads.R:243-259,302-313,368-379): for every
 * truth, ceil(coverage * L / read_len) start positions drawn with replacement with probability
 * proportional to the scoring table's probability of the kmer-window starting there; draws whose
 * read would overrun the truth are dropped; reads are the substrings (read_len bytes each, dense,
 * segment after segment).  seg_read_start (HOST, [n_segments+1]) receives the read ranges.
 * flags: BS_DEVICE_CHARS = truth_chars is a device pointer, BS_DEVICE_RESULT = reads_out is a device
 * pointer (so the reads can go straight into bs_score_batch without crossing PCIe).  The random
 * stream is this library's (seeded, reproducible), not R's: the law is upstream's, the sample is not.
 * bs_simulate_capacity returns the bytes reads_out must hold (the no-drop upper bound).
 */
BS_API int64_t bs_simulate_capacity(const int64_t *truth_off, int64_t n_segments, int32_t read_len, double coverage);
BS_API int bs_simulate_reads(bs_ctx *ctx, const char *truth_chars, const int64_t *truth_off, int64_t n_segments,
                             int32_t read_len, double coverage, int kmer, uint64_t seed, uint32_t flags,
                             char *reads_out, int64_t reads_capacity, int64_t *seg_read_start);

/* pinned host memory for staging buffers (cudaHostAlloc / cudaFreeHost) */
BS_API void *bs_host_alloc(int64_t bytes);
BS_API void bs_host_free(void *p);

BS_API int bs_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BREAKSCORE_H */
