// TEST INFRASTRUCTURE ONLY (oracle). C-ABI wrapper around the UNMODIFIED reference scorer.
//
// This translation unit #includes the upstream file lib/BreakageScorer.cpp where it lies under
// /root/reference (never copied into this repository) and exposes its calc_breakscore
// (upstream lib/BreakageScorer.cpp:185-353) through one extern "C" function taking flat
// buffers, so that tests/ and bench.py's CPU-baseline legs can call the real reference
// through ctypes.  Built by oracle/Makefile into oracle/_ref/libref_breakscore.so (and
// ..._noedit.so with the edit distance stubbed, for timing).  The product path never loads it.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#ifndef BS_REFERENCE_SRC
#error "define BS_REFERENCE_SRC to the path of the upstream lib/BreakageScorer.cpp"
#endif
#include BS_REFERENCE_SRC

namespace {
std::vector<std::string> unflatten(const char *chars, const int64_t *off, int64_t n) {
    std::vector<std::string> v((size_t)n);
    for (int64_t i = 0; i < n; i++) v[(size_t)i].assign(chars + off[i], (size_t)(off[i + 1] - off[i]));
    return v;
}
}  // namespace

extern "C" {

// Returns 0 on success, 1 if the reference threw.  prob_dist_flat receives the concatenated
// path_prob_dist vectors; prob_dist_off[C+1] (caller-computed: prefix sums of L_c-kmer+1) says
// where each one goes.  Any output pointer may be NULL.
int ref_calc_breakscore(const char *contig_chars, const int64_t *contig_off, int64_t n_contigs,
                        const char *read_chars, const int64_t *read_off, int64_t n_reads,
                        const char *truth, int64_t truth_len, int kmer,
                        const char *kmer_chars, const int64_t *kmer_off, const double *prob,
                        int64_t n_table,
                        int32_t *sequence_len, double *bp_score, double *bp_score_norm_by_break_freqs,
                        double *bp_score_norm_by_len, int32_t *kmer_breaks, int32_t *lev_dist_vs_true,
                        int32_t *path_prob_dist_startpos, double *prob_dist_flat,
                        const int64_t *prob_dist_off) {
    try {
        std::vector<std::string> path = unflatten(contig_chars, contig_off, n_contigs);
        std::vector<std::string> reads = unflatten(read_chars, read_off, n_reads);
        std::string true_solution(truth, (size_t)truth_len);
        std::vector<std::string> bp_kmer = unflatten(kmer_chars, kmer_off, n_table);
        std::vector<double> bp_prob(prob, prob + n_table);

        Rcpp::List res = calc_breakscore(path, reads, true_solution, kmer, bp_kmer, bp_prob);

        auto ints = [&](const char *name, int32_t *dst) {
            if (!dst) return;
            const auto &v = std::get<std::vector<int>>(res[name]);
            for (size_t i = 0; i < v.size(); i++) dst[i] = v[i];
        };
        auto dbls = [&](const char *name, double *dst) {
            if (!dst) return;
            const auto &v = std::get<std::vector<double>>(res[name]);
            std::memcpy(dst, v.data(), v.size() * sizeof(double));
        };
        ints("sequence_len", sequence_len);
        dbls("bp_score", bp_score);
        dbls("bp_score_norm_by_break_freqs", bp_score_norm_by_break_freqs);
        dbls("bp_score_norm_by_len", bp_score_norm_by_len);
        ints("kmer_breaks", kmer_breaks);
        ints("lev_dist_vs_true", lev_dist_vs_true);
        ints("path_prob_dist_startpos", path_prob_dist_startpos);
        if (prob_dist_flat) {
            const auto &pd = std::get<std::vector<std::vector<double>>>(res["path_prob_dist"]);
            for (size_t i = 0; i < pd.size(); i++) {
                int64_t want = prob_dist_off[i + 1] - prob_dist_off[i];
                if ((int64_t)pd[i].size() != want) return 2;
                std::memcpy(prob_dist_flat + prob_dist_off[i], pd[i].data(), pd[i].size() * sizeof(double));
            }
        }
        return 0;
    } catch (const Rcpp::internal::InterruptedException &) {
        return 3;  // (only the B200 glue raises it: a user interrupt seen by its poll callback)
    } catch (...) {
        return 1;
    }
}

// test hook: what Ctrl-C does in an R session -- the next R_CheckUserInterrupt of the stand-in header fires
void ref_raise_interrupt(void) { shim_pending_interrupt() = 1; }

#ifndef BS_DRIVER_NO_ASSEMBLE
// Upstream assemble_contigs (lib/BreakageScorer.cpp:79-174): returns the number of scaffolds
// and, if out_chars != NULL, writes them '\n'-separated (out_cap bytes available).
int64_t ref_assemble_contigs(const char *contig_chars, const int64_t *contig_off, int64_t n_contigs,
                             int dbg_kmer, int seed, char *out_chars, int64_t out_cap,
                             int64_t *out_bytes) {
    try {
        std::vector<std::string> contigs = unflatten(contig_chars, contig_off, n_contigs);
        std::vector<std::string> res = assemble_contigs(contigs, dbg_kmer, seed);
        int64_t need = 0;
        for (const auto &s : res) need += (int64_t)s.size() + 1;
        if (out_bytes) *out_bytes = need;
        if (out_chars && need <= out_cap) {
            char *p = out_chars;
            for (const auto &s : res) {
                std::memcpy(p, s.data(), s.size());
                p += s.size();
                *p++ = '\n';
            }
        }
        return (int64_t)res.size();
    } catch (...) {
        return -1;
    }
}
#endif

}  // extern "C"
