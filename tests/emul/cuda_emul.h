// TEST INFRASTRUCTURE ONLY.  A minimal CPU emulation of the subset of CUDA that
// genomeassembler_dev_b200/csrc uses, so that the DEVICE ALGORITHM (the same kernel source,
// compiled with g++ -DBS_CPU_EMUL) can be checked against the oracle on machines without a
// GPU (this build container, CPU-only CI).  It is NOT a fallback: the product package only
// ever loads the nvcc-built libbreakscore.so and fails loudly without a GPU; this header is
// reachable only from tests/emul/ and the library it produces is loaded only by
// tests/ (see tests/conftest.py, fixture `emul_lib`).
//
// Model: blocks run one after another; the threads of a block are real std::threads;
// __syncthreads is a std::barrier; warp collectives (__ballot_sync, __shfl_*_sync) rendezvous
// the 32 threads of a warp on a per-warp barrier.  blockDim.x must be a multiple of 32.
#pragma once
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct bs_emul_uint3 {
    unsigned x, y, z;
};

namespace bs_emul {

struct WarpState {
    std::barrier<> bar{32};
    std::atomic<uint32_t> ballot[2];
    uint64_t xchg[32];
    WarpState() { ballot[0] = 0; ballot[1] = 0; }
};

struct BlockState {
    std::barrier<> bar;
    std::vector<std::unique_ptr<WarpState>> warps;
    std::atomic<int> or_slot[2];
    std::vector<unsigned char> smem;
    BlockState(unsigned nthreads, size_t smem_bytes) : bar(nthreads), smem(smem_bytes + 64) {
        for (unsigned w = 0; w < nthreads / 32; w++) warps.emplace_back(new WarpState());
        or_slot[0] = 0;
        or_slot[1] = 0;
    }
};

inline thread_local BlockState *t_blk = nullptr;
inline thread_local WarpState *t_warp = nullptr;
inline thread_local unsigned t_lane = 0;
inline thread_local unsigned t_ballot_phase = 0;
inline thread_local unsigned t_or_phase = 0;

}  // namespace bs_emul

inline thread_local bs_emul_uint3 threadIdx{0, 0, 0};
inline thread_local bs_emul_uint3 blockIdx{0, 0, 0};
inline thread_local dim3 blockDim{1, 1, 1};
inline thread_local dim3 gridDim{1, 1, 1};

namespace bs_emul {

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F &&fn) {
    const unsigned nthreads = block.x * block.y * block.z;
    if (nthreads % 32 != 0) {
        std::fprintf(stderr, "bs_emul: blockDim must be a multiple of 32\n");
        std::abort();
    }
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                BlockState blk(nthreads, smem_bytes);
                std::vector<std::thread> ths;
                ths.reserve(nthreads);
                for (unsigned t = 0; t < nthreads; t++) {
                    ths.emplace_back([&, t]() {
                        threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                        blockIdx = {bx, by, bz};
                        blockDim = block;
                        gridDim = grid;
                        t_blk = &blk;
                        t_warp = blk.warps[t / 32].get();
                        t_lane = t % 32;
                        t_ballot_phase = 0;
                        t_or_phase = 0;
                        fn();
                    });
                }
                for (auto &th : ths) th.join();
            }
}

}  // namespace bs_emul

// ---- dynamic shared memory -------------------------------------------------------------------
inline unsigned char *bs_dyn_smem() {
    uintptr_t p = (uintptr_t)bs_emul::t_blk->smem.data();
    return (unsigned char *)((p + 15) & ~(uintptr_t)15);
}

// ---- barriers -------------------------------------------------------------------------------
inline void __syncthreads() { bs_emul::t_blk->bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { bs_emul::t_warp->bar.arrive_and_wait(); }
inline int __syncthreads_or(int pred) {
    auto *b = bs_emul::t_blk;
    unsigned ph = bs_emul::t_or_phase++ & 1u;
    if (pred) b->or_slot[ph].fetch_or(1);
    b->bar.arrive_and_wait();
    int r = b->or_slot[ph].load();
    b->bar.arrive_and_wait();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) b->or_slot[ph].store(0);
    return r;
}

// ---- warp collectives -------------------------------------------------------------------------
inline unsigned __ballot_sync(unsigned, int pred) {
    auto *w = bs_emul::t_warp;
    unsigned ph = bs_emul::t_ballot_phase++ & 1u;
    if (pred) w->ballot[ph].fetch_or(1u << bs_emul::t_lane);
    w->bar.arrive_and_wait();
    unsigned r = w->ballot[ph].load();
    w->bar.arrive_and_wait();
    if (bs_emul::t_lane == 0) w->ballot[ph].store(0);
    return r;
}
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
template <class T>
inline T bs_emul_shfl(T v, unsigned src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    auto *w = bs_emul::t_warp;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w->xchg[bs_emul::t_lane] = raw;
    w->bar.arrive_and_wait();
    uint64_t got = w->xchg[src_lane & 31u];
    w->bar.arrive_and_wait();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src) { return bs_emul_shfl(v, (unsigned)src); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m) { return bs_emul_shfl(v, bs_emul::t_lane ^ (unsigned)m); }
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    unsigned s = bs_emul::t_lane + d;
    return bs_emul_shfl(v, s < 32 ? s : bs_emul::t_lane);
}

// ---- atomics (any address space is plain host memory here) -----------------------------------
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicMin(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned atomicMin(unsigned *p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline int atomicMax(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}

// ---- intrinsics ---------------------------------------------------------------------------------
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
template <class T>
inline T __ldg(const T *p) { return *p; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    unsigned l = bs_emul::t_lane;
    return bs_emul_shfl(v, l >= d ? l - d : l);
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) {
    shift &= 31u;
    return shift ? (hi << shift) | (lo >> (32u - shift)) : hi;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
    shift &= 31u;
    return shift ? (lo >> shift) | (hi << (32u - shift)) : lo;
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)((v >> (8 * (sel & 7))) & 0xff);
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
struct uint4 {
    unsigned x, y, z, w;
};
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct uint2 {
    unsigned x, y;
};
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- the slice of the runtime API that bs_api.cu uses ------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef struct bs_emul_event { double t; } *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaHostAllocDefault = 0, cudaHostAllocPortable = 1, cudaStreamNonBlocking = 1, cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaEventDisableTiming = 2 };
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount;
    size_t sharedMemPerBlockOptin, totalGlobalMem;
};
inline const char *cudaGetErrorString(cudaError_t e) { return e ? "emulated failure" : "no error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    std::memset(p, 0, sizeof(*p));
    std::strcpy(p->name, "CPU emulation (tests only)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 4;
    p->sharedMemPerBlockOptin = 227 * 1024; p->totalGlobalMem = (size_t)8 << 30;
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new bs_emul_event{0.0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class K>
inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
template <class K>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) { *n = 2; return cudaSuccess; }

#define BS_LAUNCH(kern, grid, block, smem, stream, ...) \
    bs_emul::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kern(__VA_ARGS__); })  // synchronous: references stay valid
