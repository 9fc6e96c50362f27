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
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/breakscore.h"

namespace {

struct StringList {
    std::vector<std::string> items;
};

// how a scaffold was glued together: (base contig, overlap with the scaffold so far) per part
struct Part {
    int32_t base, overlap;
};
typedef std::vector<Part> Composition;
struct ScaffoldList {
    std::vector<int64_t> part_start;
    std::vector<int32_t> part_base, part_overlap;
};

thread_local char g_assemble_error[256] = "";

// one permutation: greedy suffix/prefix merging, exactly the sweep order of upstream :105-143
// comps (optional): the parts of every string of `contigs`, kept in step with it
void merge_permutation(std::vector<std::string> &contigs, int dbg_kmer, std::vector<Composition> *comps = nullptr) {
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
                    if (comps) {  // j's parts follow i's; its first part now overlaps the end of i by k
                        Composition &ci = (*comps)[i], &cj = (*comps)[jj];
                        const size_t at = ci.size();
                        ci.insert(ci.end(), cj.begin(), cj.end());
                        ci[at].overlap = k;
                        cj.clear();
                    }
                }
            }
            if (comps) {  // (an input contig that is empty to begin with has parts but no text: it goes with its string)
                size_t o = 0;
                for (size_t i = 0; i < contigs.size(); i++)
                    if (!contigs[i].empty()) { if (o != i) (*comps)[o] = std::move((*comps)[i]); o++; }
                comps->resize(o);
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
                         int n_shuffles, int n_threads, bs_string_list **out, bs_scaffold_list **parts = nullptr);

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

int bs_assemble_scaffolds(const char *contig_chars, const int64_t *contig_off, int64_t n_contigs, int dbg_kmer, int seed,
                          int n_shuffles, int n_threads, bs_string_list **out, bs_scaffold_list **parts) {
    if (!parts) {
        std::snprintf(g_assemble_error, sizeof g_assemble_error, "bs_assemble_scaffolds: parts is NULL");
        return BS_ERR_INVALID;
    }
    try {
        return assemble_impl(contig_chars, contig_off, n_contigs, dbg_kmer, seed, n_shuffles, n_threads, out, parts);
    } catch (const std::bad_alloc &) {
        std::snprintf(g_assemble_error, sizeof g_assemble_error, "bs_assemble_scaffolds: out of host memory");
        return BS_ERR_ALLOC;
    } catch (const std::exception &ex) {
        std::snprintf(g_assemble_error, sizeof g_assemble_error, "bs_assemble_scaffolds: %.200s", ex.what());
        return BS_ERR_STATE;
    } catch (...) {
        std::snprintf(g_assemble_error, sizeof g_assemble_error, "bs_assemble_scaffolds: unknown exception");
        return BS_ERR_STATE;
    }
}

static int assemble_impl(const char *contig_chars, const int64_t *contig_off, int64_t n_contigs, int dbg_kmer, int seed,
                         int n_shuffles, int n_threads, bs_string_list **out, bs_scaffold_list **parts) {
    g_assemble_error[0] = 0;
    if (parts) *parts = nullptr;
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
    std::unordered_map<std::string, Composition> found_comp;  // (with parts) one composition per distinct scaffold: the first one seen
    std::mutex mu;
    std::atomic<size_t> next_perm{0};
    unsigned nt = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
    nt = (unsigned)std::min<size_t>(nt, std::max<size_t>(perms.size(), 1));
    std::atomic<bool> failed{false};  // an exception inside a worker (bad_alloc): reported after the join, never escapes a thread
    // which of several equal-text compositions is kept must not depend on threads or timing: the smallest one
    auto keep_smallest = [](std::unordered_map<std::string, Composition> &m, const std::string &text, Composition &&comp) {
        auto less = [](const Composition &x, const Composition &y) {
            return std::lexicographical_compare(x.begin(), x.end(), y.begin(), y.end(), [](const Part &a, const Part &b) {
                return a.base != b.base ? a.base < b.base : a.overlap < b.overlap; });
        };
        auto it = m.find(text);
        if (it == m.end()) m.emplace(text, std::move(comp));
        else if (less(comp, it->second)) it->second = std::move(comp);
    };
    auto worker = [&]() {
        try {
            std::unordered_set<std::string> local;
            std::unordered_map<std::string, Composition> local_comp;
            std::vector<std::string> contigs;
            std::vector<Composition> comps;
            for (;;) {
                const size_t p = next_perm.fetch_add(1);
                if (p >= perms.size()) break;
                contigs.clear();
                for (int32_t id : perms[p]) contigs.push_back(input[(size_t)id]);
                if (parts) {
                    comps.clear();
                    for (int32_t id : perms[p]) comps.push_back(Composition{Part{id, 0}});
                    merge_permutation(contigs, dbg_kmer, &comps);
                    for (size_t i = 0; i < contigs.size(); i++) keep_smallest(local_comp, contigs[i], std::move(comps[i]));
                } else {
                    merge_permutation(contigs, dbg_kmer);
                    for (auto &s : contigs) local.insert(std::move(s));
                }
            }
            std::lock_guard<std::mutex> lock(mu);
            for (auto &s : local) found.insert(s);
            for (auto &kv : local_comp) {
                found.insert(kv.first);
                keep_smallest(found_comp, kv.first, std::move(kv.second));
            }
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
    if (parts) {
        ScaffoldList *sl = new (std::nothrow) ScaffoldList();
        if (!sl) { delete res; return BS_ERR_ALLOC; }
        sl->part_start.push_back(0);
        for (const std::string &s : res->items) {
            for (const Part &pt : found_comp.at(s)) { sl->part_base.push_back(pt.base); sl->part_overlap.push_back(pt.overlap); }
            sl->part_start.push_back((int64_t)sl->part_base.size());
        }
        *parts = reinterpret_cast<bs_scaffold_list *>(sl);
    }
    *out = reinterpret_cast<bs_string_list *>(res);
    return BS_OK;
}

int64_t bs_scaffold_list_size(const bs_scaffold_list *l) { return l ? (int64_t)reinterpret_cast<const ScaffoldList *>(l)->part_start.size() - 1 : 0; }
int64_t bs_scaffold_list_parts(const bs_scaffold_list *l) { return l ? (int64_t)reinterpret_cast<const ScaffoldList *>(l)->part_base.size() : 0; }
void bs_scaffold_list_copy(const bs_scaffold_list *l, int64_t *part_start, int32_t *part_base, int32_t *part_overlap) {
    if (!l) return;
    const ScaffoldList *sl = reinterpret_cast<const ScaffoldList *>(l);
    if (part_start) std::memcpy(part_start, sl->part_start.data(), sl->part_start.size() * sizeof(int64_t));
    if (part_base && !sl->part_base.empty()) std::memcpy(part_base, sl->part_base.data(), sl->part_base.size() * sizeof(int32_t));
    if (part_overlap && !sl->part_overlap.empty()) std::memcpy(part_overlap, sl->part_overlap.data(), sl->part_overlap.size() * sizeof(int32_t));
}
void bs_scaffold_list_free(bs_scaffold_list *l) { delete reinterpret_cast<ScaffoldList *>(l); }

// ---- scaffold sets given as parts (include/breakscore.h: bs_scaffold_set) ----

}  // extern "C"

// Full validation of a scaffold set (shared with bs_score_scaffolds in bs_api.cu; not exported): counts, indices, overlaps
// in range, starts of the parts ascending, and every overlap a true suffix/prefix match of the text built so far.
// lengths [n_scaffolds] and part_dst [parts] (scaffold position of the first base a part adds) are optional outputs.
extern "C" int bs_scaffold_validate(const bs_scaffold_set *set, int64_t *lengths, int32_t *part_dst, char *err, size_t errn) {
    auto bad = [&](const char *what, long long a, long long b) {
        std::snprintf(err, errn, "scaffold set: %s (%lld, %lld)", what, a, b);
        return (int)BS_ERR_INVALID;
    };
    if (!set) return bad("NULL set", 0, 0);
    if (set->n_base < 0 || set->n_scaffolds < 0 || set->n_base > 0x7fffffff || set->n_scaffolds > 0x7fffffff) return bad("bad counts", set->n_base, set->n_scaffolds);
    if (set->n_scaffolds == 0) return BS_OK;
    if (!set->base_off || !set->scaffold_part_start || !set->part_base || !set->part_overlap) return bad("NULL array", 0, 0);
    for (int64_t b = 0; b < set->n_base; b++)
        if (set->base_off[b + 1] < set->base_off[b]) return bad("base offsets are not monotone at", b, 0);
    if (set->n_base > 0 && set->base_off[set->n_base] > set->base_off[0] && !set->base_chars) return bad("base_chars is NULL", 0, 0);
    if (set->scaffold_part_start[0] != 0) return bad("scaffold_part_start[0] != 0", set->scaffold_part_start[0], 0);
    for (int64_t c = 0; c < set->n_scaffolds; c++) {
        const int64_t p0 = set->scaffold_part_start[c], p1 = set->scaffold_part_start[c + 1];
        if (p1 <= p0) return bad("scaffold without parts", c, p1 - p0);
        int64_t len = 0, prev_start = 0;
        for (int64_t i = p0; i < p1; i++) {
            const int64_t b = set->part_base[i], ov = set->part_overlap[i];
            if (b < 0 || b >= set->n_base) return bad("part names a base contig out of range", i, b);
            const int64_t Lb = set->base_off[b + 1] - set->base_off[b];
            if (ov < 0 || ov >= Lb || (i == p0 && ov != 0)) return bad("overlap out of range (0 for a first part, below the part's length)", i, ov);
            if (ov > len) return bad("overlap longer than the scaffold built so far", i, ov);
            if (i > p0 && len - ov < prev_start) return bad("a part starts before its predecessor", i, ov);
            // the last ov bases of the text so far == the first ov bases of the part: walk back over the parts
            const char *pc = set->base_chars + set->base_off[b];
            int64_t rem = ov;
            for (int64_t k = i - 1; rem > 0 && k >= p0; k--) {
                const int64_t bk = set->part_base[k], ovk = set->part_overlap[k];
                const char *kc = set->base_chars + set->base_off[bk];
                const int64_t Lk = set->base_off[bk + 1] - set->base_off[bk];
                const int64_t take = std::min(rem, Lk - ovk);  // bases that part k added
                if (std::memcmp(kc + Lk - take, pc + rem - take, (size_t)take) != 0) return bad("overlap is not a suffix/prefix match", i, ov);
                rem -= take;
            }
            prev_start = len - ov;
            if (part_dst) part_dst[i] = (int32_t)len;
            len += Lb - ov;
            if (len > 0x7f000000ll) return bad("scaffold longer than 2^31", c, len);
        }
        if (lengths) lengths[c] = len;
    }
    return BS_OK;
}

extern "C" {

int bs_scaffold_lengths(const bs_scaffold_set *set, int64_t *lengths) {
    if (!lengths) return BS_ERR_INVALID;
    return bs_scaffold_validate(set, lengths, nullptr, g_assemble_error, sizeof g_assemble_error);
}

int bs_scaffold_texts(const bs_scaffold_set *set, char *chars, int64_t *off) {
    if (!set || !off) return BS_ERR_INVALID;
    try {
        std::vector<int64_t> len((size_t)std::max<int64_t>(set->n_scaffolds, 1));
        const int rc = bs_scaffold_validate(set, len.data(), nullptr, g_assemble_error, sizeof g_assemble_error);
        if (rc != BS_OK) return rc;
        off[0] = 0;
        for (int64_t c = 0; c < set->n_scaffolds; c++) off[c + 1] = off[c] + len[(size_t)c];
        if (!chars) return BS_OK;
        for (int64_t c = 0; c < set->n_scaffolds; c++) {
            char *o = chars + off[c];
            for (int64_t i = set->scaffold_part_start[c]; i < set->scaffold_part_start[c + 1]; i++) {
                const int64_t b = set->part_base[i], ov = set->part_overlap[i];
                const int64_t n = set->base_off[b + 1] - set->base_off[b] - ov;
                std::memcpy(o, set->base_chars + set->base_off[b] + ov, (size_t)n);
                o += n;
            }
        }
        return BS_OK;
    } catch (...) {
        std::snprintf(g_assemble_error, sizeof g_assemble_error, "bs_scaffold_texts: out of host memory");
        return BS_ERR_ALLOC;
    }
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
