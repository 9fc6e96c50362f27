// Read simulation on the device (upstream lib/GenerateReads.R:243-259,302-313,368-379):
// start positions are drawn WITH replacement from the rolling k-mer windows of the truth with
// probability proportional to the table probability of the window, ceil(coverage * L / read_len)
// draws per segment, draws whose read would overrun the segment are dropped, reads are the
// substrings.  The random stream is a counter-based hash of (seed, segment, draw) -- not R's
// Mersenne Twister -- so the reads follow upstream's LAW, not upstream's sample.
#pragma once
#include "bs_ks.cuh"

namespace bs {

struct SimArgs {
    const int64_t *tr_off;    // [S+1] chars of the truths
    const int64_t *tr_woff;   // [S+1] packed words
    const uint64_t *tr_words;
    const uint32_t *tr_mask;
    const uint8_t *tr_chars;
    const WinEntry *win;      // [4^kmer] window -> table probability
    int32_t kmer;
    int32_t read_len;
    int64_t n_seg;
    int64_t seg_base;         // global index of the first segment (enters the random stream)
    uint64_t seed;
    const int64_t *draw_off;  // [S+1] draws of the segments
    double *cdf;              // [sum L] inclusive cumulative window probability, per segment
    int32_t *starts;          // [sum draws] kept start or -1; compacted in place by k_sim_compact
    int32_t *kept;            // [S] reads kept per segment
    const int64_t *seg_read_start;  // [S+1] (k_sim_emit)
    uint8_t *reads;           // [n_reads * read_len]
};

// uniform double in [0, 1) from (seed, segment, draw): splitmix64 finaliser over a mixed counter
__device__ __forceinline__ double sim_uniform(uint64_t seed, uint64_t seg, uint64_t draw) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (seg + 1) + 0xD1B54A32D192ED03ull * (draw + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// per segment: cdf[p] = sum of the table probabilities of the windows at 0..p (one block per
// segment; tiles of 4 windows per thread, carried in position order: the same sums on any GPU)
__global__ void __launch_bounds__(256) k_sim_cdf(SimArgs a) {
    __shared__ double s_w[32];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const int64_t s = blockIdx.x;
    const int64_t L = a.tr_off[s + 1] - a.tr_off[s];
    const uint64_t *gw = a.tr_words + a.tr_woff[s];
    const uint32_t *gm = a.tr_mask + a.tr_woff[s];
    double *cdf = a.cdf + a.tr_off[s];
    const int64_t nwin = L - a.kmer + 1;
    const int kshift = 64 - 2 * a.kmer;
    const uint32_t kbits = keep_bits(a.kmer);
    double carry = 0.0;
    for (int64_t p0 = 0; p0 < L; p0 += 4 * (int64_t)nthr) {
        const int64_t pb = p0 + 4 * (int64_t)tid;
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        if (pb < nwin) {
            const int64_t wi = pb >> 5;
            const uint32_t o = (uint32_t)(pb & 31);
            const uint64_t w0 = gw[wi], w1 = gw[wi + 1];
            const uint32_t m0 = gm[wi], m1 = gm[wi + 1];
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (pb + u < nwin && !(window32(m0, m1, o + u) & kbits)) v[u] = a.win[window64(w0, w1, o + u) >> kshift].prob;
        }
        v[1] += v[0]; v[2] += v[1]; v[3] += v[2];
        // exclusive prefix of the per-thread sums over the block
        double incl = v[3];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= d) incl += o;
        }
        __syncthreads();
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        double base = carry + incl - v[3];
        double tile_total = 0.0;
        for (int w = 0; w < nwarp; w++) {
            if (w < warp) base += s_w[w];
            tile_total += s_w[w];
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (pb + u < L) cdf[pb + u] = base + v[u];
        carry += tile_total;
    }
}

// one thread per draw: inverse-CDF sampling; -1 for a draw whose read would overrun the segment
__global__ void k_sim_draw(SimArgs a) {
    const int64_t total = a.draw_off[a.n_seg];
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < total; d += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = a.n_seg - 1;  // segment of the draw
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (a.draw_off[mid] <= d) lo = mid; else hi = mid - 1;
        }
        const int64_t s = lo;
        const int64_t L = a.tr_off[s + 1] - a.tr_off[s];
        const int64_t nwin = L - a.kmer + 1;
        int32_t start = -1;
        if (nwin > 0) {
            const double *cdf = a.cdf + a.tr_off[s];
            const double mass = cdf[nwin - 1];
            if (mass > 0.0) {
                const double t = sim_uniform(a.seed, (uint64_t)(a.seg_base + s), (uint64_t)(d - a.draw_off[s])) * mass;
                int64_t l2 = 0, h2 = nwin - 1;  // first window with cdf > t
                while (l2 < h2) {
                    const int64_t mid = (l2 + h2) >> 1;
                    if (cdf[mid] > t) h2 = mid; else l2 = mid + 1;
                }
                if (l2 + a.read_len <= L) start = (int32_t)l2;
            }
        }
        a.starts[d] = start;
    }
}

// per segment: kept draws moved to the front of the segment's slots, in draw order
__global__ void __launch_bounds__(256) k_sim_compact(SimArgs a) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const int64_t s = blockIdx.x;
    int32_t *st = a.starts + a.draw_off[s];
    const int64_t n = a.draw_off[s + 1] - a.draw_off[s];
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int64_t i0 = 0; i0 < n; i0 += nthr) {
        const int64_t i = i0 + tid;
        const int32_t v = i < n ? st[i] : -1;
        const unsigned m = __ballot_sync(FULL_MASK, v >= 0);
        const int before = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        int base = s_carry;
        for (int w = 0; w < warp; w++) base += s_w[w];
        int tile = 0;
        for (int w = 0; w < nwarp; w++) tile += s_w[w];
        __syncthreads();  // every read of this tile's slots precedes the writes below (dst <= src)
        if (v >= 0) st[base + before] = v;
        if (tid == 0) s_carry += tile;
        __syncthreads();
    }
    if (tid == 0) a.kept[s] = s_carry;
}

// one warp per read: the substring of the truth
__global__ void k_sim_emit(SimArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t n_reads = a.seg_read_start[a.n_seg];
    for (int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < n_reads; j += warps) {
        int64_t lo = 0, hi = a.n_seg - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (a.seg_read_start[mid] <= j) lo = mid; else hi = mid - 1;
        }
        const int64_t s = lo;
        const int32_t start = a.starts[a.draw_off[s] + (j - a.seg_read_start[s])];
        const uint8_t *src = a.tr_chars + a.tr_off[s] + start;
        uint8_t *dst = a.reads + j * (int64_t)a.read_len;
        for (int b = lane; b < a.read_len; b += 32) dst[b] = src[b];
    }
}

}  // namespace bs
