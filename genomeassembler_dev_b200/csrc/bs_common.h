// Shared host/device helpers of the breakage scorer: 2-bit packing conventions, window
// extraction, the dense k-mer index space and the break-k-mer rule of the reference
// (upstream lib/BreakageScorer.cpp:244-267).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) || defined(BS_CPU_EMUL)
#define BS_HD __host__ __device__ __forceinline__
#else
#define BS_HD inline
#endif
#if defined(__CUDACC__) && !defined(BS_CPU_EMUL)
#define BS_NOINLINE __noinline__
#else
#define BS_NOINLINE
#endif

namespace bs {

// ---- packing convention ---------------------------------------------------------------
// Sequence words are uint64, 32 bases each, BIG-endian inside the word: base i of the word
// sits at bits [62-2i, 63-2i].  Codes are lexicographic: A=0 C=1 G=2 T=3, so the top 2k bits
// of a window ARE the row index of that k-mer inside its 4^k table.  Any other byte packs as
// code 0 and sets its bit in the parallel validity mask (uint32 per 32 bases, base i at bit
// 31-i; 1 = not ACGT or beyond the end of the string).
constexpr int MAXK = 8;
// dense table space over all k-mer lengths 1..8: offset of length k is (4^k - 4) / 3
constexpr int DENSE_SIZE = (65536 * 4 - 4) / 3;  // 87380
BS_HD int dense_offset(int k) { return ((1 << (2 * k)) - 4) / 3; }

BS_HD uint32_t base_code(uint32_t c) { return ((c >> 1) ^ (c >> 2)) & 3u; }
BS_HD bool base_valid(uint32_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// 32 bases starting o (0..31) bases into w0, spilling into w1
BS_HD uint64_t window64(uint64_t w0, uint64_t w1, uint32_t o) {
    const uint32_t s = 2u * o;
    return (w0 << s) | ((w1 >> 1) >> (63u - s));
}
// 32 mask bits starting o (0..31) bits into m0
BS_HD uint32_t window32(uint32_t m0, uint32_t m1, uint32_t o) {
    return (m0 << o) | ((m1 >> 1) >> (31u - o));
}
// keep the first n (0..32) bases / mask bits of a window
BS_HD uint64_t keep_bases(int n) { return n >= 32 ? ~0ull : (n <= 0 ? 0ull : ~0ull << (64 - 2 * n)); }
BS_HD uint32_t keep_bits(int n) { return n >= 32 ? ~0u : (n <= 0 ? 0u : ~0u << (32 - n)); }

BS_HD uint32_t seed_hash(uint64_t s) {
    uint32_t lo = (uint32_t)s, hi = (uint32_t)(s >> 32);
    uint32_t x = lo * 0x9E3779B1u ^ hi * 0x85EBCA77u;
    x ^= x >> 15;
    x *= 0x2C1B3C6Du;
    x ^= x >> 13;
    return x;
}

// second, independent 32-bit mix of a seed: stored with a read in its index chain
BS_HD uint32_t seed_tag(uint64_t s) {
    uint32_t lo = (uint32_t)s, hi = (uint32_t)(s >> 32);
    uint32_t x = lo * 0xC2B2AE3Du + hi * 0x27D4EB2Fu;
    x ^= x >> 16;
    x *= 0x165667B1u;
    x ^= x >> 15;
    return x;
}

// ---- the break-k-mer rule --------------------------------------------------------------
// A read placed at `pos` of a contig of length L breaks the k-mer
//   contig.substr(start, e),  start = max(0, pos - kmer/2),  e = 8, except when start == 0:
//   pos == 1 -> 2, pos == 2 -> 4, pos == 3 -> 6   (upstream lib/BreakageScorer.cpp:247-262)
// clamped at the end of the contig (std::string::substr, :267).
struct BreakWindow {
    int64_t start;
    int len;
};
BS_HD BreakWindow break_window(int64_t pos, int kmer, int64_t L) {
    int64_t start = pos - kmer / 2;
    if (start < 0) start = 0;
    int e = 8;
    if (start == 0) {
        if (pos == 1) e = 2;
        else if (pos == 2) e = 4;
        else if (pos == 3) e = 6;
    }
    if (start + e > L) e = (int)(L - start);
    if (e < 0) e = 0;
    BreakWindow w;
    w.start = start;
    w.len = e;
    return w;
}

// dense index of the k-mer of length len (1..8) starting at base `start` of a packed string
// (words / mask are the string's own arrays), or -1 if any base is invalid or len is 0.
BS_HD int dense_index_at(const uint64_t *words, const uint32_t *mask, int64_t start, int len) {
    if (len < 1 || len > MAXK) return -1;
    const int64_t wi = start >> 5;
    const uint32_t o = (uint32_t)(start & 31);
    const uint32_t m = window32(mask[wi], mask[wi + 1], o) & keep_bits(len);
    if (m) return -1;
    const uint64_t win = window64(words[wi], words[wi + 1], o);
    return dense_offset(len) + (int)(win >> (64 - 2 * len));
}

}  // namespace bs
