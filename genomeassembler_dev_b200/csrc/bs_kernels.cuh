// All CUDA kernels of the breakage scorer (sm_100a), by stage.  Integer compare-and-count work: no
// tensor cores; the rules that matter are coalesced streaming, shared-memory staging and grids
// sized to the 148 SMs.
//
// The same sources compile under tests/emul/cuda_emul.h (BS_CPU_EMUL) so that the device algorithm
// can be checked against the oracle on a machine without a GPU; that build is test infrastructure,
// not a fallback.
#pragma once
#include "bs_seq.cuh"       // descriptors, 2-bit packing, read index
#include "bs_place.cuh"     // placement (read index; contig-tile index; all-pairs scan)
#include "bs_score.cuh"     // break k-mers -> sums, histogram, KS-B
#include "bs_ks.cuh"        // truth spectrum, path_prob_dist, KS-A
#include "bs_compose.cuh"   // scaffold sets scored from their parts (base contigs + junction windows)
#include "bs_startpos.cuh"  // contig-in-truth offset
#include "bs_lev.cuh"       // infix edit distance (lev_dist_vs_true)
#include "bs_simulate.cuh"  // read simulation (the step before the scorer)
