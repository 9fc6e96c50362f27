// Scaffold explosion on the host: the candidate generator that feeds the scorer
// (upstream assemble_contigs, lib/BreakageScorer.cpp:79-174).
//
// Semantics kept: `n_shuffles` (upstream: 20 000) permutations of the contig list drawn from ONE
// std::mt19937(seed) stream with std::shuffle; for every permutation and every overlap length
// k = dbg_kmer-1 ... 1, repeated sweeps "i ascending, j descending: if contig i differs from contig j
// and the last k characters of i equal the first k of j, append j (without its first k characters)
// to i and empty j", until a sweep merges nothing; all scaffolds of all permutations, duplicates
// removed, longest first.
//
// What is different from upstream is only how it is computed: the permutations are index vectors
// (std::shuffle consumes the engine the same way for any element type), suffix/prefix tests are
// memcmp in place (upstream allocates two substrings per (i, j, k)), the permutations are worked
// on by all host cores, and identical permutations results are deduplicated through a hash set
// before the final sort.  The final order reproduces upstream's: lexicographic sort, unique, then
// std::sort by length (the same libstdc++ algorithm on the same sequence gives the same tie order).
#include <algorithm>
#include <exception>
#include <new>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../../include/breakscore.h"

namespace {

struct StringList {
    std::vector<std::string> items;
};

thread_local char g_assemble_error[256] = "";

// one permutation: greedy suffix/prefix merging, exactly the sweep order of upstream :105-143
void merge_permutation(std::vector<std::string> &contigs, int dbg_kmer) {
    for (int k = dbg_kmer - 1; k > 0; k--) {
        bool len_changed = true;
        while (len_changed) {
            const size_t before = contigs.size();
            for (size_t i = 0; i < contigs.size(); i++) {
                if (contigs[i].empty()) continue;
                for (size_t jj = contigs.size(); jj-- > 0;) {
                    std::string &a = contigs[i];
                    const std::string &b = contigs[jj];
                    if (a.size() == b.size() && a == b) continue;     // equal strings never merge (also i == j)
                    if (b.size() < (size_t)k) continue;               // its prefix is shorter than k: cannot equal the suffix
                    // a.size() >= k is guaranteed by the caller's length check
                    if (std::memcmp(a.data() + a.size() - k, b.data(), (size_t)k) != 0) continue;
                    a.append(b, (size_t)k, std::string::npos);
                    contigs[jj].clear();
                }
            }
            contigs.erase(std::remove_if(contigs.begin(), contigs.end(), [](const std::string &s) { return s.empty(); }),
                          contigs.end());
            len_changed = before != contigs.size();
        }
    }
}

}  // namespace

extern "C" {

const char *bs_assemble_last_error(void) { return g_assemble_error; }

static int assemble_impl(const char *contig_chars, const int64_t *contig_off, int64_t n_contigs, int dbg_kmer, int seed,
                         int n_shuffles, int n_threads, bs_string_list **out);

// nothing throws across the boundary (bad_alloc of the containers, system_error of std::thread)
int bs_assemble_contigs(const char *contig_chars, const int64_t *contig_off, int64_t n_contigs, int dbg_kmer, int seed,
                        int n_shuffles, int n_threads, bs_string_list **out) {
    try {
        return assemble_impl(contig_chars, contig_off, n_contigs, dbg_kmer, seed, n_shuffles, n_threads, out);
    } catch (const std::bad_alloc &) {
        std::snprintf(g_assemble_error, sizeof g_assemble_error, "bs_assemble_contigs: out of host memory");
        return BS_ERR_ALLOC;
    } catch (const std::exception &ex) {
        std::snprintf(g_assemble_error, sizeof g_assemble_error, "bs_assemble_contigs: %.200s", ex.what());
        return BS_ERR_STATE;
    } catch (...) {
        std::snprintf(g_assemble_error, sizeof g_assemble_error, "bs_assemble_contigs: unknown exception");
        return BS_ERR_STATE;
    }
}

static int assemble_impl(const char *contig_chars, const int64_t *contig_off, int64_t n_contigs, int dbg_kmer, int seed,
                         int n_shuffles, int n_threads, bs_string_list **out) {
    g_assemble_error[0] = 0;
    if (!out || !contig_off || n_contigs < 0 || n_shuffles < 0 || (n_contigs > 0 && !contig_chars)) {
        std::snprintf(g_assemble_error, sizeof g_assemble_error, "bs_assemble_contigs: bad argument");
        return BS_ERR_INVALID;
    }
    *out = nullptr;
    std::vector<std::string> input((size_t)n_contigs);
    bool all_equal = true;
    size_t min_len = SIZE_MAX;
    for (int64_t i = 0; i < n_contigs; i++) {
        if (contig_off[i + 1] < contig_off[i]) {
            std::snprintf(g_assemble_error, sizeof g_assemble_error, "contig offsets are not monotone");
            return BS_ERR_INVALID;
        }
        input[(size_t)i].assign(contig_chars + contig_off[i], (size_t)(contig_off[i + 1] - contig_off[i]));
        if (input[(size_t)i] != input[0]) all_equal = false;
        if (!input[(size_t)i].empty()) min_len = std::min(min_len, input[(size_t)i].size());
    }
    // upstream takes substr(size - k) of contig i for k = dbg_kmer-1 first: a shorter non-empty contig
    // makes std::string::substr throw there (an R error); the same inputs are an error here
    if (!all_equal && dbg_kmer > 1 && min_len != SIZE_MAX && min_len < (size_t)(dbg_kmer - 1)) {
        std::snprintf(g_assemble_error, sizeof g_assemble_error,
                      "a contig is shorter than dbg_kmer-1 = %d (upstream throws std::out_of_range here)", dbg_kmer - 1);
        return BS_ERR_INVALID;
    }

    // the permutations, from one engine stream like upstream :86-94
    std::vector<std::vector<int32_t>> perms((size_t)n_shuffles);
    {
        std::mt19937 engine((std::mt19937::result_type)seed);
        std::vector<int32_t> ids((size_t)n_contigs);
        for (int s = 0; s < n_shuffles; s++) {
            for (int64_t i = 0; i < n_contigs; i++) ids[(size_t)i] = (int32_t)i;
            std::shuffle(ids.begin(), ids.end(), engine);
            perms[(size_t)s] = ids;
        }
    }
    // identical permutations (of identical strings) give identical scaffolds: work on distinct ones only
    std::sort(perms.begin(), perms.end());
    perms.erase(std::unique(perms.begin(), perms.end()), perms.end());

    std::unordered_set<std::string> found;
    std::mutex mu;
    std::atomic<size_t> next_perm{0};
    unsigned nt = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
    nt = (unsigned)std::min<size_t>(nt, std::max<size_t>(perms.size(), 1));
    std::atomic<bool> failed{false};  // an exception inside a worker (bad_alloc): reported after the join, never escapes a thread
    auto worker = [&]() {
        try {
            std::unordered_set<std::string> local;
            std::vector<std::string> contigs;
            for (;;) {
                const size_t p = next_perm.fetch_add(1);
                if (p >= perms.size()) break;
                contigs.clear();
                for (int32_t id : perms[p]) contigs.push_back(input[(size_t)id]);
                merge_permutation(contigs, dbg_kmer);
                for (auto &s : contigs) local.insert(std::move(s));
            }
            std::lock_guard<std::mutex> lock(mu);
            for (auto &s : local) found.insert(s);
        } catch (...) {
            failed = true;
            next_perm = perms.size();  // the other workers stop at their next permutation
        }
    };
    {
        std::vector<std::thread> pool;
        pool.reserve(nt);
        try {
            for (unsigned t = 1; t < nt; t++) pool.emplace_back(worker);
        } catch (...) {  // fewer threads than asked for: the ones that started share the work
        }
        worker();
        for (auto &t : pool) t.join();
    }
    if (failed) {
        std::snprintf(g_assemble_error, sizeof g_assemble_error, "bs_assemble_contigs: out of host memory in a worker");
        return BS_ERR_ALLOC;
    }

    StringList *res = new (std::nothrow) StringList();
    if (!res) return BS_ERR_ALLOC;
    res->items.assign(found.begin(), found.end());
    std::sort(res->items.begin(), res->items.end());  // upstream remove_duplicates: sort + unique
    res->items.erase(std::unique(res->items.begin(), res->items.end()), res->items.end());
    std::sort(res->items.begin(), res->items.end(),
              [](const std::string &a, const std::string &b) -> bool { return a.length() > b.length(); });
    *out = reinterpret_cast<bs_string_list *>(res);
    return BS_OK;
}

int64_t bs_string_list_size(const bs_string_list *l) { return l ? (int64_t)reinterpret_cast<const StringList *>(l)->items.size() : 0; }

int64_t bs_string_list_bytes(const bs_string_list *l) {
    int64_t n = 0;
    if (l) for (const auto &s : reinterpret_cast<const StringList *>(l)->items) n += (int64_t)s.size();
    return n;
}

void bs_string_list_copy(const bs_string_list *l, char *chars, int64_t *off) {
    if (!l) return;
    int64_t o = 0, i = 0;
    for (const auto &s : reinterpret_cast<const StringList *>(l)->items) {
        off[i++] = o;
        if (chars) std::memcpy(chars + o, s.data(), s.size());
        o += (int64_t)s.size();
    }
    off[i] = o;
}

void bs_string_list_free(bs_string_list *l) { delete reinterpret_cast<StringList *>(l); }

}  // extern "C"
