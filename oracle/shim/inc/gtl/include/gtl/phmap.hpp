// TEST INFRASTRUCTURE ONLY (oracle). Stand-in for gtl/phmap.hpp (greg7mdp/gtl, un-vendored
// upstream, no version pinned in upstream README.md:25-43).  The reference uses
// gtl::flat_hash_map only through operator[], range-for and an initializer-list constructor
// (upstream lib/BreakageScorer.cpp:22,194,200,218), all of which std::unordered_map provides.
// Iteration order differs from gtl's, which changes the fp64 summation order of
// lib/BreakageScorer.cpp:279-297 at the 1e-16 relative level only (counts are unaffected).
#pragma once
#include <unordered_map>

namespace gtl {
template <class K, class V>
using flat_hash_map = std::unordered_map<K, V>;
}
