// [[Rcpp::plugins("cpp17")]]
// Drop-in replacement for the scorer half of upstream lib/BreakageScorer.cpp: the same
// Rcpp-exported entry point, argument list and returned list (upstream :185-191, :343-353), with
// the bodies replaced by calls into libbreakscore.so (include/breakscore.h).  Both exports of the
// upstream file are here: assemble_contigs (:79-174, the scaffold explosion, native host code in the
// library) and calc_breakscore (the scorer, CUDA) -- plus assemble_and_score, the two in one call for the place
// where the R driver runs them back to back (lib/DeNovoAssembler.R:343-355): the candidates then stay a list of
// (base contig, overlap) parts and are scored from their parts (bs_score_scaffolds).
//
// Build from R:
//   Sys.setenv(PKG_CXXFLAGS = "-I<repo>/include", PKG_LIBS = "-L<repo>/genomeassembler_dev_b200 -lbreakscore -Wl,-rpath,<repo>/genomeassembler_dev_b200")
//   Rcpp::sourceCpp("BreakageScorer.cpp")
// No CPU fallback: without a B200 the call stops with the library's error text.
#include <Rcpp.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "breakscore.h"

namespace {

// process-lifetime device context: table, work buffers and pinned staging survive across calls
struct Session {
    bs_ctx *ctx = nullptr;
    std::vector<bs_ctx *> all;  // ctx first, then one context per further GPU of BREAKSCORE_DEVICES
    std::vector<std::string> table_kmer;
    std::vector<double> table_prob;
    std::vector<double> truth_prob;  // optional: table behind the KS truth side (set_breakscore_truth_prob)
    bool truth_dirty = false;
    char *pinned = nullptr;
    int64_t pinned_cap = 0;
    ~Session() {
        if (pinned) bs_host_free(pinned);
        for (size_t k = 1; k < all.size(); k++) bs_ctx_destroy(all[k]);
        if (ctx) bs_ctx_destroy(ctx);
    }
};

// User interrupts (Ctrl-C / Esc) during a long call: the library polls this between its pipeline chunks, on R's
// main thread.  R_CheckUserInterrupt leaves by longjmp, so it runs under R_ToplevelExec and nothing unwinds through
// the C-ABI; the library then drains its streams and returns BS_ERR_INTERRUPTED, which check() turns into the
// exception Rcpp reports to R as an interrupt.
void check_interrupt(void *) { R_CheckUserInterrupt(); }
int poll_user_interrupt(void *) { return R_ToplevelExec(check_interrupt, nullptr) == FALSE; }

Session &session() {
    static Session s;
    if (!s.ctx) {
        // BREAKSCORE_DEVICES=0,1,...,7: the contigs of every call are shared out over these GPUs (bs_score_multi; R is one
        // process, so one process drives them all); else BREAKSCORE_DEVICE=<n> (default 0), one GPU
        std::vector<int> devs;
        if (const char *list = std::getenv("BREAKSCORE_DEVICES")) {
            for (const char *p = list; *p;) {
                char *end = nullptr;
                const long d = std::strtol(p, &end, 10);
                if (end == p) break;
                devs.push_back((int)d);
                p = *end == ',' ? end + 1 : end;
            }
        }
        if (devs.empty()) {
            const char *dev = std::getenv("BREAKSCORE_DEVICE");
            devs.push_back(dev ? std::atoi(dev) : 0);
        }
        if (bs_ctx_create(devs[0], &s.ctx) != BS_OK) Rcpp::stop(bs_last_error(nullptr));
        bs_ctx_set_poll(s.ctx, poll_user_interrupt, nullptr);
        s.all.assign(1, s.ctx);
        for (size_t k = 1; k < devs.size(); k++) {
            bs_ctx *more = nullptr;
            if (bs_ctx_create(devs[k], &more) != BS_OK) Rcpp::stop(bs_last_error(nullptr));
            s.all.push_back(more);
        }
    }
    return s;
}

void check(Session &s, int rc) {
    if (rc == BS_ERR_INTERRUPTED) throw Rcpp::internal::InterruptedException();
    if (rc != BS_OK) Rcpp::stop(bs_last_error(s.ctx));
}

// strings -> one flat buffer (written at dst) + offsets
int64_t flat_size(const std::vector<std::string> &v) {
    int64_t n = 0;
    for (const auto &x : v) n += (int64_t)x.size();
    return n;
}
char *flatten(const std::vector<std::string> &v, char *dst, std::vector<int64_t> &off) {
    off.resize(v.size() + 1);
    int64_t o = 0;
    for (size_t i = 0; i < v.size(); i++) {
        off[i] = o;
        std::memcpy(dst + o, v[i].data(), v[i].size());
        o += (int64_t)v[i].size();
    }
    off[v.size()] = o;
    return dst + o;
}

}  // namespace

// Same name, arguments and result as upstream lib/BreakageScorer.cpp:79-83: every scaffold reachable by
// greedy suffix/prefix merging over 20 000 seeded shuffles of the contig list, duplicates removed,
// longest first (identical strings in identical order; multi-threaded host code, no GPU needed).
// [[Rcpp::export]]
std::vector<std::string> assemble_contigs(
    const std::vector<std::string> &velvet_contigs,
    const int &dbg_kmer,
    const int &seed) {
    std::vector<char> chars((size_t)flat_size(velvet_contigs) + 1);
    std::vector<int64_t> off;
    flatten(velvet_contigs, chars.data(), off);
    bs_string_list *list = nullptr;
    if (bs_assemble_contigs(chars.data(), off.data(), (int64_t)velvet_contigs.size(), dbg_kmer, seed, 20000, 0, &list) != BS_OK)
        Rcpp::stop(bs_assemble_last_error());
    const int64_t n = bs_string_list_size(list);
    std::vector<char> out_chars((size_t)bs_string_list_bytes(list) + 1);
    std::vector<int64_t> out_off((size_t)n + 1);
    bs_string_list_copy(list, out_chars.data(), out_off.data());
    bs_string_list_free(list);
    std::vector<std::string> res((size_t)n);
    for (int64_t i = 0; i < n; i++) res[(size_t)i].assign(out_chars.data() + out_off[i], (size_t)(out_off[i + 1] - out_off[i]));
    return res;
}

// Optional: probabilities used for the truth-side distribution of the KS columns.  The upstream
// driver keeps the REAL table there even in its "random" pass (lib/DeNovoAssembler.R:326-333);
// call this once with df_prob$all$prob to reproduce that.  An empty vector resets to "same as bp_prob".
// [[Rcpp::export]]
void set_breakscore_truth_prob(const std::vector<double> &prob) {
    Session &s = session();
    s.truth_prob = prob;
    s.truth_dirty = true;
}

namespace {
// calc_breakscore's body; `set` != NULL: `path` holds the texts of a scaffold set given as parts, scored from the parts
Rcpp::List score_impl(
    const std::vector<std::string> &path,
    const std::vector<std::string> &sequencing_reads,
    const std::string &true_solution,
    const int &kmer,
    const std::vector<std::string> &bp_kmer,
    const std::vector<double> &bp_prob,
    const bs_scaffold_set *set) {
    Session &s = session();
    if (bp_kmer.size() != bp_prob.size()) Rcpp::stop("bp_kmer and bp_prob differ in length");

    // table: uploaded only when it changed (the driver passes the same one for every call of a pass)
    if (bp_prob != s.table_prob || bp_kmer != s.table_kmer) {
        std::vector<char> chars((size_t)flat_size(bp_kmer) + 1);
        std::vector<int64_t> off;
        flatten(bp_kmer, chars.data(), off);
        for (bs_ctx *c : s.all)
            if (bs_set_table(c, chars.data(), off.data(), bp_prob.data(), (int64_t)bp_prob.size()) != BS_OK) Rcpp::stop(bs_last_error(c));
        s.table_kmer = bp_kmer;
        s.table_prob = bp_prob;
        s.truth_dirty = true;
    }
    if (s.truth_dirty) {
        const bool own = s.truth_prob.size() == bp_prob.size();
        for (bs_ctx *c : s.all)
            if (bs_set_truth_table(c, own ? s.truth_prob.data() : nullptr, own ? (int64_t)s.truth_prob.size() : 0) != BS_OK)
                Rcpp::stop(bs_last_error(c));
        s.truth_dirty = false;
    }

    // inputs: flattened straight into pinned memory (the copy Rcpp's conversion forces anyway)
    const int64_t need = (set ? 0 : flat_size(path)) + flat_size(sequencing_reads) + (int64_t)true_solution.size() + 64;
    if (need > s.pinned_cap) {
        if (s.pinned) bs_host_free(s.pinned);
        s.pinned = (char *)bs_host_alloc(need + need / 4);
        s.pinned_cap = s.pinned ? need + need / 4 : 0;
        if (!s.pinned) Rcpp::stop("pinned host allocation failed");
    }
    std::vector<int64_t> ctg_off, read_off;
    char *ctg_chars = s.pinned;
    char *read_chars = set ? ctg_chars : flatten(path, ctg_chars, ctg_off);
    char *truth_chars = flatten(sequencing_reads, read_chars, read_off);
    std::memcpy(truth_chars, true_solution.data(), true_solution.size());

    const int64_t C = (int64_t)path.size();
    std::vector<int> sequence_len(C), kmer_breaks(C), lev_dist(C), startpos(C);
    std::vector<double> bp_score(C), norm_freq(C), norm_len(C), ks_a(C), ks_b(C);
    std::vector<int64_t> pd_off(C + 1, 0);
    for (int64_t c = 0; c < C; c++) {
        const int64_t n = (int64_t)path[c].size() - kmer + 1;
        pd_off[c + 1] = pd_off[c] + (n > 0 ? n : 0);
    }
    std::vector<double> pd_flat((size_t)pd_off[C] + 1);
    // optional (BREAKSCORE_PATH_FREQ=1): the per-row break frequencies that the drifted R driver expects
    // under the name path_freq (lib/DeNovoAssembler.cpp:352,395-420); 69 904 doubles per contig, so off by default
    const char *pf_env = std::getenv("BREAKSCORE_PATH_FREQ");
    const bool want_path_freq = pf_env && pf_env[0] == '1';
    const int64_t T = (int64_t)bp_prob.size();
    std::vector<int32_t> hist(want_path_freq ? (size_t)(C * (T + 1)) : 0);

    bs_result r;
    std::memset(&r, 0, sizeof(r));
    r.sequence_len = sequence_len.data();
    r.bp_score = bp_score.data();
    r.bp_score_norm_by_break_freqs = norm_freq.data();
    r.bp_score_norm_by_len = norm_len.data();
    r.kmer_breaks = kmer_breaks.data();
    r.path_prob_dist_startpos = startpos.data();
    r.lev_dist_vs_true = lev_dist.data();
    r.ks_stat_prob_dist = ks_a.data();
    r.ks_stat_path_freq = ks_b.data();
    r.path_prob_dist = pd_flat.data();
    r.path_prob_dist_off = pd_off.data();
    if (want_path_freq) r.hist = hist.data();
    const uint32_t flags = BS_DEFAULT_FLAGS | BS_WANT_LEV | (want_path_freq ? BS_WANT_HIST : 0u);
    if (set)  // (one GPU: the first context)
        check(s, bs_score_scaffolds(s.ctx, set, read_chars, read_off.data(), (int64_t)sequencing_reads.size(), 0, truth_chars,
                                    (int64_t)true_solution.size(), kmer, flags, &r));
    else
        check(s, bs_score_multi(s.all.data(), (int)s.all.size(), ctg_chars, ctg_off.data(), C, read_chars, read_off.data(),
                                (int64_t)sequencing_reads.size(), truth_chars, (int64_t)true_solution.size(), kmer, flags, &r));

    std::vector<std::vector<double>> path_prob_dist((size_t)C);
    for (int64_t c = 0; c < C; c++) path_prob_dist[c].assign(pd_flat.begin() + pd_off[c], pd_flat.begin() + pd_off[c + 1]);

    // always C entries (R turns the list into a data.table: every member needs length C); empty unless asked for
    std::vector<std::vector<double>> path_freq((size_t)C);
    if (want_path_freq) {
        for (int64_t c = 0; c < C; c++) {
            path_freq[c].resize((size_t)T);
            for (int64_t t = 0; t < T; t++)  // count / total_breaks; 0/0 = NaN like upstream when nothing was placed
                path_freq[c][t] = (double)hist[(size_t)(c * (T + 1) + t)] / (double)kmer_breaks[c];
        }
    }

    // the upstream list (lib/BreakageScorer.cpp:343-353), input order, plus the KS statistics of
    // lib/DeNovoAssembler.R:416-424 computed on the device, plus the member names the R driver written
    // against lib/DeNovoAssembler.cpp looks for (path_freq_startpos; path_freq when asked for)
    return Rcpp::List::create(
        Rcpp::Named("sequence") = path,
        Rcpp::Named("sequence_len") = sequence_len,
        Rcpp::Named("bp_score") = bp_score,
        Rcpp::Named("bp_score_norm_by_break_freqs") = norm_freq,
        Rcpp::Named("bp_score_norm_by_len") = norm_len,
        Rcpp::Named("kmer_breaks") = kmer_breaks,
        Rcpp::Named("lev_dist_vs_true") = lev_dist,
        Rcpp::Named("path_prob_dist_startpos") = startpos,
        Rcpp::Named("path_prob_dist") = Rcpp::wrap(path_prob_dist),
        Rcpp::Named("ks_stat_prob_dist") = ks_a,
        Rcpp::Named("ks_stat_path_freq") = ks_b,
        Rcpp::Named("path_freq_startpos") = startpos,
        Rcpp::Named("path_freq") = Rcpp::wrap(path_freq));
}
}  // namespace

// [[Rcpp::export]]
Rcpp::List calc_breakscore(
    const std::vector<std::string> &path,
    const std::vector<std::string> &sequencing_reads,
    const std::string &true_solution,
    const int &kmer,
    const std::vector<std::string> &bp_kmer,
    const std::vector<double> &bp_prob) {
    return score_impl(path, sequencing_reads, true_solution, kmer, bp_kmer, bp_prob, nullptr);
}

// assemble_contigs(velvet_contigs, dbg_kmer, seed) followed by calc_breakscore(<its result>, ...) in one call: the same
// list (its "sequence" member is what assemble_contigs returns, same strings in the same order), but the candidates are
// scored from their parts -- reads placed once per velvet contig, scaffold texts never copied to the device.
// [[Rcpp::export]]
Rcpp::List assemble_and_score(
    const std::vector<std::string> &velvet_contigs,
    const int &dbg_kmer,
    const int &seed,
    const std::vector<std::string> &sequencing_reads,
    const std::string &true_solution,
    const int &kmer,
    const std::vector<std::string> &bp_kmer,
    const std::vector<double> &bp_prob) {
    std::vector<char> chars((size_t)flat_size(velvet_contigs) + 1);
    std::vector<int64_t> off;
    flatten(velvet_contigs, chars.data(), off);
    bs_string_list *list = nullptr;
    bs_scaffold_list *parts = nullptr;
    if (bs_assemble_scaffolds(chars.data(), off.data(), (int64_t)velvet_contigs.size(), dbg_kmer, seed, 20000, 0, &list, &parts) != BS_OK)
        Rcpp::stop(bs_assemble_last_error());
    const int64_t n = bs_string_list_size(list);
    std::vector<char> out_chars((size_t)bs_string_list_bytes(list) + 1);
    std::vector<int64_t> out_off((size_t)n + 1), part_start((size_t)n + 1);
    std::vector<int32_t> part_base((size_t)bs_scaffold_list_parts(parts) + 1), part_overlap(part_base.size());
    bs_string_list_copy(list, out_chars.data(), out_off.data());
    bs_scaffold_list_copy(parts, part_start.data(), part_base.data(), part_overlap.data());
    bs_string_list_free(list);
    bs_scaffold_list_free(parts);
    std::vector<std::string> path((size_t)n);
    for (int64_t i = 0; i < n; i++) path[(size_t)i].assign(out_chars.data() + out_off[i], (size_t)(out_off[i + 1] - out_off[i]));
    bs_scaffold_set set;
    set.n_base = (int64_t)velvet_contigs.size(); set.base_chars = chars.data(); set.base_off = off.data();
    set.n_scaffolds = n; set.scaffold_part_start = part_start.data(); set.part_base = part_base.data(); set.part_overlap = part_overlap.data();
    return score_impl(path, sequencing_reads, true_solution, kmer, bp_kmer, bp_prob, &set);
}
