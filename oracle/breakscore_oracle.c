/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the reference breakage scorer.
 *
 * This file restates, in plain C, the algorithm of the upstream calc_breakscore
 * (lib/BreakageScorer.cpp:185-353) plus the R-side Kolmogorov-Smirnov statistic that the
 * upstream driver computes from its output (lib/DeNovoAssembler.R:416-424, stats::ks.test).
 * It exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can
 * check the CUDA path; nothing under genomeassembler_dev_b200/ may import, link or execute
 * it, and it is never what is measured as the product.
 *
 * Pinning: the upstream repository ships no tests, golden vectors or fixtures for this path
 * (SURVEY.md section 4), so the restatement is pinned against the upstream code itself,
 * compiled verbatim by oracle/Makefile into oracle/_ref/ (tests/test_oracle_vs_ref.py), and
 * against the known-answer vectors of SURVEY.md appendix A.4 (tests/golden/).  The KS
 * statistic restates R's stats::ks.test (R is not installed here): PARITY UNPINNED by any
 * upstream test; cross-checked against scipy.stats.ks_2samp in tests/test_oracle_ks.py.
 * lev_dist_vs_true (edlib, un-vendored upstream) is likewise PARITY UNPINNED.
 *
 * What the oracle exposes beyond the upstream return list (internal state of the upstream
 * loop, needed for bit-exact checks): per (contig, read) leftmost match position and the
 * per-contig break-k-mer histogram over the table rows (+1 "not in table" bin).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK 0
#define ORACLE_ERR_TABLE 3 /* table key not ACGT / longer than 8: outside the supported domain */
#define ORACLE_ERR_ALLOC 4

#define MAXK 8

/* dense index of an ACGT string of length len (lexicographic, A<C<G<T), or -1 */
static int64_t kmer_code(const char *s, int64_t len) {
    int64_t code = 0;
    for (int64_t i = 0; i < len; i++) {
        int b;
        switch (s[i]) {
        case 'A': b = 0; break;
        case 'C': b = 1; break;
        case 'G': b = 2; break;
        case 'T': b = 3; break;
        default: return -1;
        }
        code = code * 4 + b;
    }
    return code;
}

typedef struct {
    /* row[len][code] = LAST table row holding that key (map assignment semantics of
     * lib/BreakageScorer.cpp:195-197: a repeated key overwrites), or -1 */
    int32_t *row[MAXK + 1];
} table_index;

static void table_free(table_index *t) {
    for (int k = 0; k <= MAXK; k++) free(t->row[k]);
}

static int table_build(table_index *t, const char *kmer_chars, const int64_t *kmer_off, int64_t n) {
    memset(t, 0, sizeof(*t));
    for (int k = 1; k <= MAXK; k++) {
        int64_t sz = (int64_t)1 << (2 * k);
        t->row[k] = (int32_t *)malloc(sizeof(int32_t) * (size_t)sz);
        if (!t->row[k]) return ORACLE_ERR_ALLOC;
        for (int64_t i = 0; i < sz; i++) t->row[k][i] = -1;
    }
    for (int64_t i = 0; i < n; i++) {
        int64_t len = kmer_off[i + 1] - kmer_off[i];
        if (len < 1 || len > MAXK) return ORACLE_ERR_TABLE;
        int64_t code = kmer_code(kmer_chars + kmer_off[i], len);
        if (code < 0) return ORACLE_ERR_TABLE;
        t->row[len][code] = (int32_t)i;
    }
    return ORACLE_OK;
}

/* table row of the string s[0..len) or -1 ("not in table") */
static int32_t table_lookup(const table_index *t, const char *s, int64_t len) {
    if (len < 1 || len > MAXK) return -1;
    int64_t code = kmer_code(s, len);
    if (code < 0) return -1;
    return t->row[len][code];
}

/* std::string::find semantics: leftmost exact occurrence, "" matches at 0, else -1 */
static int64_t str_find(const char *hay, int64_t hay_len, const char *needle, int64_t needle_len) {
    if (needle_len == 0) return 0;
    if (needle_len > hay_len) return -1;
    const char *p = (const char *)memmem(hay, (size_t)hay_len, needle, (size_t)needle_len);
    return p ? (int64_t)(p - hay) : -1;
}

static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/*
 * Two-sample, two-sided KS statistic as computed by R's stats::ks.test (statistic only):
 *   x <- x[!is.na(x)]; y <- y[!is.na(y)]; w <- c(x, y)
 *   z <- cumsum(ifelse(order(w) <= n.x, 1/n.x, -1/n.y))
 *   if (ties) z <- z[c(which(diff(sort(w)) != 0), n.x + n.y)]
 *   max(abs(z))
 * order() is stable, so among equal values every x precedes every y; the cumulative sum is
 * carried out in floating point in exactly that order, as R does.
 * Returns NaN when either sample is empty after NA removal (R stops with an error there).
 */
double oracle_ks_statistic(const double *x_in, int64_t nx_in, const double *y_in, int64_t ny_in) {
    double *x = (double *)malloc(sizeof(double) * (size_t)(nx_in > 0 ? nx_in : 1));
    double *y = (double *)malloc(sizeof(double) * (size_t)(ny_in > 0 ? ny_in : 1));
    if (!x || !y) { free(x); free(y); return NAN; }
    int64_t nx = 0, ny = 0;
    for (int64_t i = 0; i < nx_in; i++) if (!isnan(x_in[i])) x[nx++] = x_in[i];
    for (int64_t i = 0; i < ny_in; i++) if (!isnan(y_in[i])) y[ny++] = y_in[i];
    if (nx < 1 || ny < 1) { free(x); free(y); return NAN; }
    qsort(x, (size_t)nx, sizeof(double), cmp_double);
    qsort(y, (size_t)ny, sizeof(double), cmp_double);
    const double dx = 1.0 / (double)nx, dy = -1.0 / (double)ny;
    double z = 0.0, best = 0.0;
    int64_t i = 0, j = 0;
    while (i < nx || j < ny) {
        double v;
        if (j >= ny || (i < nx && x[i] <= y[j])) { v = x[i++]; z += dx; }
        else { v = y[j++]; z += dy; }
        /* keep z only at the last element of each group of equal pooled values */
        double next;
        int has_next = 1;
        if (i < nx && j < ny) next = (x[i] <= y[j]) ? x[i] : y[j];
        else if (i < nx) next = x[i];
        else if (j < ny) next = y[j];
        else { has_next = 0; next = 0.0; }
        if (!has_next || next != v) { double a = fabs(z); if (a > best) best = a; }
    }
    free(x); free(y);
    return best;
}

/* plain O(n*m) infix (edlib "HW") edit distance: whole query vs any substring of target.
 * Restates the intent of lib/BreakageScorer.cpp:41-55; PARITY UNPINNED (edlib is absent). */
static int32_t infix_edit_distance(const char *q, int64_t n, const char *t, int64_t m) {
    int32_t *prev = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
    int32_t *cur = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
    if (!prev || !cur) { free(prev); free(cur); return 0; }
    for (int64_t i = 0; i <= n; i++) prev[i] = (int32_t)i;
    int32_t best = prev[n];
    for (int64_t j = 1; j <= m; j++) {
        cur[0] = 0;
        for (int64_t i = 1; i <= n; i++) {
            int32_t sub = prev[i - 1] + (q[i - 1] != t[j - 1]);
            int32_t gap = (prev[i] < cur[i - 1] ? prev[i] : cur[i - 1]) + 1;
            cur[i] = sub < gap ? sub : gap;
        }
        if (cur[n] < best) best = cur[n];
        int32_t *tmp = prev; prev = cur; cur = tmp;
    }
    free(prev); free(cur);
    return best;
}

/*
 * The scorer.  Flat-buffer layout: string i of a set is chars[off[i] .. off[i+1]).
 *   prob        scoring table (bp_prob), aligned with the table rows
 *   truth_prob  table used for the truth-side distribution of the KS statistic (R keeps the
 *               real probabilities there even in the "random" pass, lib/DeNovoAssembler.R:326-333);
 *               NULL = same as prob
 * Every output pointer may be NULL.  Outputs, all in input contig order
 * (lib/BreakageScorer.cpp:308-315, sort disabled):
 *   sequence_len[C], bp_score[C], norm_by_break_freqs[C], norm_by_len[C], kmer_breaks[C],
 *   startpos[C], prob_dist_flat (at prob_dist_off[c], L_c-kmer+1 doubles each),
 *   pos[C*N]  leftmost match of read n in contig c or -1      (lib/BreakageScorer.cpp:241)
 *   hist[C*(T+1)] break counts per table row, last bin = key not in table (:267-271)
 *   ks_prob_dist[C]  KS(x = path_prob_dist, y = truth window probabilities)  ("KS-A")
 *   ks_path_freq[C]  KS(x = hist[0..T)/kmer_breaks, y = same)                ("KS-B", variant
 *                    lib/DeNovoAssembler.cpp:395-420; NaN when kmer_breaks == 0)
 *   lev[C]  infix edit distance contig vs truth (only computed when non-NULL; O(L_c*L))
 * Domain restrictions (loud errors, not silent differences): table keys must be ACGT strings
 * of length 1..8.  Contigs shorter than kmer give an empty prob_dist (upstream: size_t
 * underflow / exception at lib/BreakageScorer.cpp:205-209).
 */
int oracle_calc_breakscore(const char *contig_chars, const int64_t *contig_off, int64_t n_contigs,
                           const char *read_chars, const int64_t *read_off, int64_t n_reads,
                           const char *truth, int64_t truth_len, int kmer,
                           const char *kmer_chars, const int64_t *kmer_off, const double *prob,
                           int64_t n_table, const double *truth_prob,
                           int32_t *sequence_len, double *bp_score, double *norm_by_break_freqs,
                           double *norm_by_len, int32_t *kmer_breaks, int32_t *startpos,
                           double *prob_dist_flat, const int64_t *prob_dist_off,
                           int32_t *pos_out, int32_t *hist_out, double *ks_prob_dist,
                           double *ks_path_freq, int32_t *lev) {
    table_index tab;
    int rc = table_build(&tab, kmer_chars, kmer_off, n_table);
    if (rc != ORACLE_OK) { table_free(&tab); return rc; }
    if (!truth_prob) truth_prob = prob;
    const int64_t T = n_table;

    /* truth-side distribution: kmer_from_seq of lib/GenerateReads.R:243-259 (NA = not in table) */
    int64_t ny_all = truth_len - kmer + 1;
    if (ny_all < 0 || kmer < 1) ny_all = 0;
    double *yv = (double *)malloc(sizeof(double) * (size_t)(ny_all > 0 ? ny_all : 1));
    int32_t *hist = (int32_t *)calloc((size_t)(T + 1), sizeof(int32_t));
    double *freq = (double *)malloc(sizeof(double) * (size_t)(T > 0 ? T : 1));
    if (!yv || !hist || !freq) { free(yv); free(hist); free(freq); table_free(&tab); return ORACLE_ERR_ALLOC; }
    int64_t ny = 0;
    if (ks_prob_dist || ks_path_freq) {
        for (int64_t p = 0; p < ny_all; p++) {
            int32_t r = table_lookup(&tab, truth + p, kmer);
            if (r >= 0) yv[ny++] = truth_prob[r];
        }
    }

    for (int64_t c = 0; c < n_contigs; c++) {
        const char *ctg = contig_chars + contig_off[c];
        const int64_t L = contig_off[c + 1] - contig_off[c];

        /* rolling-window probabilities, lib/BreakageScorer.cpp:200-215 */
        int64_t nwin = L - kmer + 1;
        if (nwin < 0 || kmer < 1) nwin = 0;
        double *pd_tmp = NULL;
        if (prob_dist_flat || ks_prob_dist) {
            pd_tmp = prob_dist_flat ? prob_dist_flat + prob_dist_off[c]
                                    : (double *)malloc(sizeof(double) * (size_t)(nwin > 0 ? nwin : 1));
            for (int64_t p = 0; p < nwin; p++) {
                int32_t r = table_lookup(&tab, ctg + p, kmer);
                pd_tmp[p] = (r >= 0) ? prob[r] : 0.0; /* operator[] default-inserts {0.0, 0} */
            }
        }

        /* placement + break-k-mer histogram, lib/BreakageScorer.cpp:231-276 */
        memset(hist, 0, sizeof(int32_t) * (size_t)(T + 1));
        int64_t total = 0;
        int any_hit = 0;
        for (int64_t n = 0; n < n_reads; n++) {
            const char *rd = read_chars + read_off[n];
            const int64_t rl = read_off[n + 1] - read_off[n];
            int64_t pos = str_find(ctg, L, rd, rl);
            if (pos_out) pos_out[c * n_reads + n] = (int32_t)pos;
            if (pos < 0) continue;
            int64_t start = pos - kmer / 2;
            if (start < 0) start = 0;
            int64_t e = 8;
            if (start == 0) {
                if (pos == 1) e = 2;
                else if (pos == 2) e = 4;
                else if (pos == 3) e = 6;
            }
            if (start + e > L) e = L - start; /* substr clamps at the end of the contig */
            int32_t r = table_lookup(&tab, ctg + start, e);
            hist[r >= 0 ? r : T] += 1; /* one per read copy == multiplicity of the unique read */
            total += 1;
            any_hit = 1;
        }

        /* weighted sums, lib/BreakageScorer.cpp:279-303 (row order; upstream uses hash order) */
        double score = 0.0, norm = 0.0;
        for (int64_t r = 0; r < T; r++) {
            /* rows shadowed by a later duplicate key never receive counts */
            if (hist[r] != 0) {
                score += prob[r] * (double)hist[r];
                norm += prob[r] * ((double)hist[r] / (double)total);
            }
        }
        if (sequence_len) sequence_len[c] = (int32_t)L;
        if (bp_score) bp_score[c] = score;
        if (norm_by_break_freqs) norm_by_break_freqs[c] = norm;
        if (norm_by_len) norm_by_len[c] = score / (double)(int32_t)L;
        if (kmer_breaks) kmer_breaks[c] = (int32_t)total;
        /* only assigned inside the hit branch, else stays 0: lib/BreakageScorer.cpp:229,273-274 */
        if (startpos) startpos[c] = any_hit ? (int32_t)str_find(truth, truth_len, ctg, L) : 0;
        if (hist_out) memcpy(hist_out + c * (T + 1), hist, sizeof(int32_t) * (size_t)(T + 1));

        if (ks_prob_dist) ks_prob_dist[c] = oracle_ks_statistic(pd_tmp, nwin, yv, ny);
        if (ks_path_freq) {
            if (total == 0) ks_path_freq[c] = NAN;
            else {
                for (int64_t r = 0; r < T; r++) freq[r] = (double)hist[r] / (double)total;
                ks_path_freq[c] = oracle_ks_statistic(freq, T, yv, ny);
            }
        }
        if (lev) lev[c] = infix_edit_distance(ctg, L, truth, truth_len);
        if (pd_tmp && !prob_dist_flat) free(pd_tmp);
    }
    free(yv); free(hist); free(freq);
    table_free(&tab);
    return ORACLE_OK;
}
