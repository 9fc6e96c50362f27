// TEST INFRASTRUCTURE ONLY (oracle). Stand-in for <Rcpp.h>, written from scratch.
//
// The reference hot path (lib/BreakageScorer.cpp in the upstream repo) uses exactly three
// pieces of Rcpp: Rcpp::Named("x") = value, Rcpp::List::create(...) and Rcpp::wrap(x)
// (upstream lib/BreakageScorer.cpp:343-353).  R and Rcpp are not installed in this image, so
// the verbatim reference build (oracle/Makefile -> oracle/_ref/) compiles against this
// header instead.  Nothing under genomeassembler_dev_b200/ may include it.
#pragma once
#include <string>
#include <utility>
#include <variant>
#include <vector>
#include <stdexcept>

namespace Rcpp {

using Value = std::variant<std::vector<std::string>,
                           std::vector<int>,
                           std::vector<double>,
                           std::vector<std::vector<double>>>;

struct NamedValue {
    std::string name;
    Value value;
};

// Rcpp::Named("name") = x  ->  NamedValue{name, x}
struct Named {
    std::string name;
    explicit Named(const char *n) : name(n) {}
    template <class T>
    NamedValue operator=(const T &v) const { return NamedValue{name, Value(v)}; }
};

template <class T>
inline T wrap(const T &v) { return v; }

class List {
public:
    std::vector<NamedValue> items;

    template <class... Args>
    static List create(Args &&...args) {
        List l;
        (l.items.push_back(std::forward<Args>(args)), ...);
        return l;
    }

    const Value &operator[](const std::string &name) const {
        for (const auto &it : items)
            if (it.name == name) return it.value;
        throw std::out_of_range("Rcpp shim: no list member named " + name);
    }
};

inline void stop(const std::string &msg) { throw std::runtime_error(msg); }

// what Rcpp's END_RCPP turns into an R interrupt condition (used by the B200 glue, rcpp/BreakageScorer.cpp)
namespace internal {
struct InterruptedException {};
}  // namespace internal

}  // namespace Rcpp

// stand-ins for the two entry points of R's C API the B200 glue uses to poll for a user interrupt: the test
// driver raises the flag through shim_pending_interrupt(); R_CheckUserInterrupt "longjmps" (here: throws) when
// it is set and R_ToplevelExec reports that as FALSE, as in R.
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
inline int &shim_pending_interrupt() { static int flag = 0; return flag; }
struct ShimLongjmp {};
inline void R_CheckUserInterrupt(void) {
    if (shim_pending_interrupt()) { shim_pending_interrupt() = 0; throw ShimLongjmp{}; }
}
inline Rboolean R_ToplevelExec(void (*fun)(void *), void *data) {
    try { fun(data); } catch (const ShimLongjmp &) { return FALSE; }
    return TRUE;
}
