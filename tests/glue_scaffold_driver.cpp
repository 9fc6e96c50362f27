// TEST INFRASTRUCTURE ONLY.  C wrapper around the Rcpp glue's assemble_and_score export (rcpp/BreakageScorer.cpp), compiled
// with the stand-in Rcpp.h of oracle/shim next to oracle/ref_driver.cpp (which #includes the glue): tests/test_rcpp_glue.py.
#include <Rcpp.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

Rcpp::List assemble_and_score(const std::vector<std::string> &velvet_contigs, const int &dbg_kmer, const int &seed,
                              const std::vector<std::string> &sequencing_reads, const std::string &true_solution, const int &kmer,
                              const std::vector<std::string> &bp_kmer, const std::vector<double> &bp_prob);

namespace {
std::vector<std::string> unflat(const char *chars, const int64_t *off, int64_t n) {
    std::vector<std::string> v((size_t)n);
    for (int64_t i = 0; i < n; i++) v[(size_t)i].assign(chars + off[i], (size_t)(off[i + 1] - off[i]));
    return v;
}
}  // namespace

// returns the number of scaffolds (-1: the glue threw); with cap >= that number fills the per-scaffold arrays and, if
// seq_chars != NULL and seq_cap suffices, the '\n'-separated "sequence" member
extern "C" int64_t glue_assemble_and_score(const char *contig_chars, const int64_t *contig_off, int64_t n_contigs, int dbg_kmer, int seed,
                                           const char *read_chars, const int64_t *read_off, int64_t n_reads, const char *truth,
                                           int64_t truth_len, int kmer, const char *kmer_chars, const int64_t *kmer_off,
                                           const double *prob, int64_t n_table, int64_t cap, int32_t *sequence_len, int32_t *kmer_breaks,
                                           int32_t *startpos, int32_t *lev, double *bp_score, double *ks_a, double *ks_b,
                                           char *seq_chars, int64_t seq_cap, int64_t *seq_bytes) {
    try {
        Rcpp::List res = assemble_and_score(unflat(contig_chars, contig_off, n_contigs), dbg_kmer, seed, unflat(read_chars, read_off, n_reads),
                                            std::string(truth, (size_t)truth_len), kmer, unflat(kmer_chars, kmer_off, n_table),
                                            std::vector<double>(prob, prob + n_table));
        const auto &seq = std::get<std::vector<std::string>>(res["sequence"]);
        const int64_t n = (int64_t)seq.size();
        int64_t need = 0;
        for (const auto &s : seq) need += (int64_t)s.size() + 1;
        if (seq_bytes) *seq_bytes = need;
        if (cap < n) return n;
        auto ints = [&](const char *name, int32_t *dst) {
            const auto &v = std::get<std::vector<int>>(res[name]);
            for (size_t i = 0; i < v.size(); i++) dst[i] = v[i];
        };
        auto dbls = [&](const char *name, double *dst) {
            const auto &v = std::get<std::vector<double>>(res[name]);
            std::memcpy(dst, v.data(), v.size() * sizeof(double));
        };
        ints("sequence_len", sequence_len); ints("kmer_breaks", kmer_breaks); ints("path_prob_dist_startpos", startpos);
        ints("lev_dist_vs_true", lev); dbls("bp_score", bp_score); dbls("ks_stat_prob_dist", ks_a); dbls("ks_stat_path_freq", ks_b);
        if (seq_chars && need <= seq_cap) {
            char *p = seq_chars;
            for (const auto &s : seq) { std::memcpy(p, s.data(), s.size()); p += s.size(); *p++ = '\n'; }
        }
        return n;
    } catch (...) {
        return -1;
    }
}
