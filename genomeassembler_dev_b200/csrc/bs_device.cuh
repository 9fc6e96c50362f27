// Device-side common ground of the breakage scorer's kernels (sm_100a).  Integer compare-and-count work: no tensor
// cores; the rules that matter are coalesced streaming, shared-memory staging of the contig
// and grids sized to the 148 SMs.
//
// The same source compiles under tests/emul/cuda_emul.h (BS_CPU_EMUL) so that the device
// algorithm can be checked against the oracle on a machine without a GPU; that build is test
// infrastructure, not a fallback.
#pragma once
#include "bs_common.h"

#ifndef BS_CPU_EMUL
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned char *bs_dyn_smem() {
    extern __shared__ __align__(16) unsigned char bs_smem_raw[];
    return bs_smem_raw;
}
#define BS_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

namespace bs {

constexpr unsigned FULL_MASK = 0xffffffffu;

}  // namespace bs
