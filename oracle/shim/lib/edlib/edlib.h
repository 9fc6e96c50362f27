// TEST INFRASTRUCTURE ONLY (oracle). Stand-in for edlib.h (Martinsos/edlib, un-vendored
// upstream; README.md:27 says ">= 1.2.7", no pin => lev_dist_vs_true is PARITY UNPINNED).
// Reached through the quoted include "../lib/edlib/edlib.h" of upstream
// lib/BreakageScorer.cpp:11 resolved against -I oracle/shim/inc.
//
// Only the call shape of upstream lib/BreakageScorer.cpp:43-50 is provided.  The distance is
// a plain O(n*m) semi-global / global / prefix DP (two rolling rows) so that small cases
// have a meaningful value; define BS_EDLIB_STUB to make it return 0 (used when the
// reference is timed as a CPU baseline: edit distance is outside the scored path).
#pragma once
#include <algorithm>
#include <cstddef>
#include <vector>

#define EDLIB_STATUS_OK 0
#define EDLIB_STATUS_ERROR 1

typedef enum { EDLIB_MODE_NW, EDLIB_MODE_SHW, EDLIB_MODE_HW } EdlibAlignMode;
typedef enum { EDLIB_TASK_DISTANCE, EDLIB_TASK_LOC, EDLIB_TASK_PATH } EdlibAlignTask;

typedef struct {
    char first;
    char second;
} EdlibEqualityPair;

typedef struct {
    int k;
    EdlibAlignMode mode;
    EdlibAlignTask task;
    const EdlibEqualityPair *additionalEqualities;
    int additionalEqualitiesLength;
} EdlibAlignConfig;

typedef struct {
    int status;
    int editDistance;
} EdlibAlignResult;

static inline EdlibAlignConfig edlibNewAlignConfig(int k, EdlibAlignMode mode, EdlibAlignTask task,
                                                   const EdlibEqualityPair *eq, int neq) {
    EdlibAlignConfig c;
    c.k = k;
    c.mode = mode;
    c.task = task;
    c.additionalEqualities = eq;
    c.additionalEqualitiesLength = neq;
    return c;
}

static inline EdlibAlignResult edlibAlign(const char *query, int queryLength, const char *target,
                                          int targetLength, const EdlibAlignConfig config) {
    EdlibAlignResult res;
    res.status = EDLIB_STATUS_OK;
    res.editDistance = 0;
#ifndef BS_EDLIB_STUB
    const int n = queryLength, m = targetLength;
    // column-wise DP over the target; D[i] = cost of aligning query[0..i) ending at this column
    std::vector<int> prev(n + 1), cur(n + 1);
    for (int i = 0; i <= n; i++) prev[i] = i;
    int best = prev[n];
    for (int j = 1; j <= m; j++) {
        cur[0] = (config.mode == EDLIB_MODE_HW) ? 0 : j;  // free start in target for HW
        for (int i = 1; i <= n; i++) {
            int sub = prev[i - 1] + (query[i - 1] != target[j - 1]);
            cur[i] = std::min(sub, std::min(prev[i], cur[i - 1]) + 1);
        }
        if (cur[n] < best) best = cur[n];
        std::swap(prev, cur);
    }
    // NW: must consume the whole target; SHW/HW: free end in target
    res.editDistance = (config.mode == EDLIB_MODE_NW) ? prev[n] : best;
    if (config.k >= 0 && res.editDistance > config.k) res.editDistance = -1;
#else
    (void)query; (void)queryLength; (void)target; (void)targetLength; (void)config;
#endif
    return res;
}
