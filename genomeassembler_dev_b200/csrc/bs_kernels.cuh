// CUDA kernels of the breakage scorer (sm_100a).  Integer compare-and-count work: no tensor
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

// ------------------------------------------------------------------------------------------
// descriptors (plain structs passed by value)
// ------------------------------------------------------------------------------------------

// a set of strings packed AoS: string i owns words [woff[i], woff[i+1]) = ceil(L_i/32) + 2
// (two pad words: sequence 0, mask all-ones) and the same range of mask words
struct SeqSet {
    const uint8_t *chars;  // ASCII
    const int64_t *off;    // [n+1] char offsets
    const int64_t *woff;   // [n+1] word offsets
    uint64_t *words;
    uint32_t *mask;
    int64_t n;
    int64_t total_words;
};

// reads packed AoS: word j of read i at words[i * W + j] (a placement candidate is verified from
// one or two 32-byte sectors)
struct ReadSet {
    const uint8_t *chars;
    const int64_t *off;  // NULL => uniform_len, dense
    int32_t uniform_len;
    int32_t W;  // words per read
    int64_t n;
    uint64_t *words;
    uint8_t *flags;  // bit0: read holds a byte outside ACGT
};

BS_HD int64_t read_begin(const ReadSet &r, int64_t i) { return r.off ? r.off[i] : i * (int64_t)r.uniform_len; }
BS_HD int32_t read_length(const ReadSet &r, int64_t i) { return r.off ? (int32_t)(r.off[i + 1] - r.off[i]) : r.uniform_len; }

// ------------------------------------------------------------------------------------------
// 2-bit packing
// ------------------------------------------------------------------------------------------

// one thread per packed word of a SeqSet
__global__ void k_pack_seqs(SeqSet s) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < s.total_words; wi += stride) {
        // string owning word wi: largest i with woff[i] <= wi
        int64_t lo = 0, hi = s.n - 1;
        while (lo < hi) {
            int64_t mid = (lo + hi + 1) >> 1;
            if (s.woff[mid] <= wi) lo = mid; else hi = mid - 1;
        }
        const int64_t c0 = s.off[lo];
        const int64_t L = s.off[lo + 1] - c0;
        const int64_t b0 = (wi - s.woff[lo]) * 32;
        uint64_t w = 0;
        uint32_t m = 0;
        for (int b = 0; b < 32; b++) {
            const int64_t p = b0 + b;
            uint32_t code = 0, bad = 1;
            if (p < L) {
                const uint32_t ch = s.chars[c0 + p];
                if (base_valid(ch)) { code = base_code(ch); bad = 0; }
            }
            w = (w << 2) | code;
            m = (m << 1) | bad;
        }
        s.words[wi] = w;
        s.mask[wi] = m;
    }
}

// ---- read index ------------------------------------------------------------------------------
// Per segment, a chained hash table over the reads' seeds (their first seed_len bases, seed_len =
// min(32, shortest read of the segment)): head[tab_off[s] + (hash(seed) & tab_mask[s])] -> read id
// + 1, next[read] -> next read id + 1 of the same bucket (0 ends the chain).  Reads whose seed
// cannot be packed (a byte outside ACGT in it, or an empty read) hang on odd_head[s] instead and
// are placed by byte comparison.  Built by the packing kernels, consumed by k_place_index.
struct ReadIndex {
    uint32_t *head;
    uint2 *next;              // [N] x = next read id + 1 of the bucket (0 ends the chain), y = seed tag
    uint32_t *odd_head;       // [S]
    const int64_t *tab_off;   // [S]
    const int32_t *tab_mask;  // [S] table size - 1 (size is a power of two)
    const int32_t *seed_len;  // [S]
    const int64_t *seg_read_start;  // [S+1]
    int32_t n_seg;
};

// segment owning read n: largest s with seg_read_start[s] <= n (empty segments are skipped)
__device__ __forceinline__ int segment_of_read(const ReadIndex &ix, int64_t n) {
    int lo = 0, hi = ix.n_seg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (ix.seg_read_start[mid] <= n) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ void index_insert(const ReadIndex &ix, int s, int64_t n, uint64_t word0, int len, bool seed_bad) {
    const int S = ix.seed_len[s];
    if (len == 0 || seed_bad) {
        ix.next[n] = make_uint2(atomicExch(&ix.odd_head[s], (uint32_t)n + 1u), 0u);
    } else {
        const uint64_t seed = word0 & keep_bases(S);
        const uint32_t h = seed_hash(seed) & (uint32_t)ix.tab_mask[s];
        ix.next[n] = make_uint2(atomicExch(&ix.head[ix.tab_off[s] + h], (uint32_t)n + 1u), seed_tag(seed));
    }
}

// general packing (reads of arbitrary lengths): one thread per read, all W words, the
// "has a byte outside ACGT" flag and the index insertion
__global__ void k_pack_reads(ReadSet r, ReadIndex ix) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < r.n; i += stride) {
        const int64_t c0 = read_begin(r, i);
        const int32_t len = read_length(r, i);
        uint32_t any_bad = 0, seed_bad = 0;
        uint64_t w0 = 0;
        for (int j = 0; j < r.W; j++) {
            uint64_t w = 0;
            for (int b = 0; b < 32; b++) {
                const int p = j * 32 + b;
                uint32_t code = 0;
                if (p < len) {
                    const uint32_t ch = r.chars[c0 + p];
                    if (base_valid(ch)) code = base_code(ch); else any_bad = 1;
                }
                w = (w << 2) | code;
            }
            r.words[i * r.W + j] = w;
            if (j == 0) { w0 = w; seed_bad = any_bad; }
        }
        r.flags[i] = (uint8_t)any_bad;
        if (ix.head) index_insert(ix, segment_of_read(ix, i), i, w0, len, seed_bad != 0);
    }
}

// ---- fast packing for reads of one common length (the simulated-read case) -------------------
// 16 ASCII bytes -> 32 bits of 2-bit codes, four bytes at a time inside a 32-bit register:
//   code  = ((c >> 1) ^ (c >> 2)) & 3 per byte, gathered into one byte by a multiply;
//   valid = the byte equals "ACGT"[code] (a byte-permute looks the expected letter up).
__device__ __forceinline__ uint32_t pack4(uint32_t x, uint32_t &diff) {
    const uint32_t t = ((x >> 1) ^ (x >> 2)) & 0x03030303u;
    uint32_t u = (t | (t >> 4)) & 0x00ff00ffu;
    u = u | (u >> 8);  // code of byte i in nibble i
    diff |= x ^ __byte_perm(0x54474341u, 0u, u & 0xffffu);
    return t * 0x40100401u;  // top byte = c0 c1 c2 c3 (first base in the most significant bits)
}
__device__ __forceinline__ uint32_t pack16(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t &diff) {
    const uint32_t g0 = pack4(x0, diff), g1 = pack4(x1, diff), g2 = pack4(x2, diff), g3 = pack4(x3, diff);
    const uint32_t hi = __byte_perm(g0, g1, 0x3700u);  // byte3 = g0.3, byte2 = g1.3
    const uint32_t lo = __byte_perm(g2, g3, 0x0037u);  // byte1 = g2.3, byte0 = g3.3
    return __byte_perm(hi, lo, 0x3254u);
}

constexpr int PACK_THREADS = 256;
constexpr int PACK_WARP_CELLS = 66;  // 16-byte cells a warp tile (32 words = at most 1024 bytes + alignment) can touch

// One WARP tile = 32 consecutive output words (AoS order: read-major).  Their source bytes are
// one contiguous span of the dense read buffer: staged by coalesced 16-byte loads, converted
// once per 16-byte cell into the warp's slice of shared memory, then each lane cuts its 32 bases
// out of two or three cells with funnel shifts.  Warps never wait for one another (__syncwarp
// only).  A cell holding any byte outside ACGT flags every read that overlaps it (conservative:
// flagged reads are verified by byte comparison, still exact).
__global__ void __launch_bounds__(PACK_THREADS) k_pack_reads_uniform(ReadSet r, ReadIndex ix) {
    __shared__ uint32_t s_code_all[PACK_THREADS / 32][PACK_WARP_CELLS];
    __shared__ uint32_t s_bad_all[PACK_THREADS / 32][PACK_WARP_CELLS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *s_code = s_code_all[warp], *s_bad = s_bad_all[warp];
    const int64_t n_words = r.n * r.W;
    const int64_t total_bytes = r.n * (int64_t)r.uniform_len;
    const uintptr_t base = (uintptr_t)r.chars;
    const int L = r.uniform_len;
    const uint32_t W = (uint32_t)r.W;
    // every warp owns a contiguous run of tiles: consecutive tiles read adjacent bytes and stay in
    // the same segment, and the next tile's loads can be issued before the current one is finished
    const int64_t warps_total = (int64_t)gridDim.x * (PACK_THREADS / 32);
    const int64_t tiles_total = (n_words + 31) / 32;
    const int64_t per_warp = (tiles_total + warps_total - 1) / warps_total;
    const int64_t wid = (int64_t)blockIdx.x * (PACK_THREADS / 32) + warp;
    int64_t k0 = wid * per_warp * 32;
    int64_t k_end = k0 + per_warp * 32;
    if (k_end > n_words) k_end = n_words;
    if (k0 >= k_end) return;
    // (read, word-in-read) of the tile's first word: one 64-bit division, then 32-bit increments
    int64_t n0 = k0 / W;
    uint32_t j0 = (uint32_t)(k0 - n0 * W);
    const uint32_t adv = 32u % W, adv_n = 32u / W;
    int seg = ix.head ? segment_of_read(ix, n0) : 0;

    // geometry of a tile and the (up to three) 16-byte cells this lane stages for it
    struct Tile { int64_t lo16; int n_cells; int ntile; };
    auto tile_of = [&](int64_t k, int64_t n, uint32_t j) {
        Tile t;
        t.ntile = (int)(k_end - k < 32 ? k_end - k : 32);
        const uint32_t xl = j + (uint32_t)t.ntile - 1;
        const uint32_t dnl = xl / W;
        const int64_t lo = n * L + (int64_t)j * 32;                          // first source byte of the tile
        int64_t hi = (n + dnl) * L + (int64_t)(xl - dnl * W) * 32 + 32;      // one past the last
        if (hi > total_bytes) hi = total_bytes;
        t.lo16 = lo - (int64_t)((base + (uintptr_t)lo) & 15);               // 16-byte aligned address, may be < 0
        t.n_cells = (int)((hi - t.lo16 + 15) >> 4);
        return t;
    };
    auto load_cell = [&](const Tile &t, int ci) {
        uint4 v = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
        if (ci < t.n_cells) {
            const int64_t cb = t.lo16 + 16 * (int64_t)ci;
            if (cb >= 0 && cb + 16 <= total_bytes) {
                v = *reinterpret_cast<const uint4 *>(r.chars + cb);
            } else {  // partly outside the buffer: bytes that do not exist read as 'A'
                uint32_t xs[4];
                for (int q = 0; q < 4; q++) {
                    uint32_t x = 0;
                    for (int b = 3; b >= 0; b--) {
                        const int64_t pb = cb + 4 * q + b;
                        x = (x << 8) | ((pb >= 0 && pb < total_bytes) ? (uint32_t)r.chars[pb] : (uint32_t)'A');
                    }
                    xs[q] = x;
                }
                v = make_uint4(xs[0], xs[1], xs[2], xs[3]);
            }
        }
        return v;
    };

    Tile t = tile_of(k0, n0, j0);
    uint4 c0 = load_cell(t, lane), c1 = load_cell(t, lane + 32), c2 = load_cell(t, lane + 64);
    for (;;) {
        // ---- convert this tile's cells into the warp's shared-memory slice ----
        __syncwarp();
        {
            uint32_t diff = 0;
            s_code[lane] = pack16(c0.x, c0.y, c0.z, c0.w, diff);
            s_bad[lane] = diff;
            diff = 0;
            s_code[lane + 32] = pack16(c1.x, c1.y, c1.z, c1.w, diff);
            s_bad[lane + 32] = diff;
            if (lane + 64 < PACK_WARP_CELLS) {
                diff = 0;
                s_code[lane + 64] = pack16(c2.x, c2.y, c2.z, c2.w, diff);
                s_bad[lane + 64] = diff;
            }
        }
        __syncwarp();
        // ---- issue the next tile's loads before cutting this tile's words ----
        const Tile cur = t;
        const int64_t k_cur = k0, n_cur = n0;
        const uint32_t j_cur = j0;
        k0 += 32;
        n0 += adv_n;
        j0 += adv;
        if (j0 >= W) { j0 -= W; n0++; }
        const bool more = k0 < k_end;
        if (more) {
            t = tile_of(k0, n0, j0);
            c0 = load_cell(t, lane);
            c1 = load_cell(t, lane + 32);
            c2 = load_cell(t, lane + 64);
        }
        if (lane < cur.ntile) {
            const uint32_t x = j_cur + (uint32_t)lane;
            const uint32_t dn = x / W;
            const int64_t n = n_cur + dn;
            const int j = (int)(x - dn * W);
            const int a = (int)(n * L + 32 * j - cur.lo16);  // byte offset inside the staged span (< 1100)
            const int ci = a >> 4;
            const uint32_t sh = 2u * (uint32_t)(a & 15);
            const int rem = L - 32 * j;  // bases of this word that belong to the read (may exceed 32)
            // cells past the staged span are only touched by bases beyond the read: clamp the index
            const int i1 = ci + 1 < cur.n_cells ? ci + 1 : cur.n_cells - 1, i2 = ci + 2 < cur.n_cells ? ci + 2 : cur.n_cells - 1;
            const uint32_t w0 = s_code[ci], w1 = s_code[i1], w2 = s_code[i2];
            const uint32_t o_hi = sh ? ((w0 << sh) | (w1 >> (32u - sh))) : w0;
            const uint32_t o_lo = sh ? ((w1 << sh) | (w2 >> (32u - sh))) : w1;
            const uint64_t word = (((uint64_t)o_hi << 32) | o_lo) & keep_bases(rem);
            r.words[k_cur + lane] = word;
            // cells overlapping this word's own bytes [a, a + min(rem, 32))
            const int last_cell = (a + (rem < 32 ? rem : 32) - 1) >> 4;
            uint32_t bad = s_bad[ci];
            if (last_cell >= ci + 1) bad |= s_bad[i1];
            if (last_cell >= ci + 2) bad |= s_bad[i2];
            if (bad) atomicOr(reinterpret_cast<unsigned *>(r.flags) + (n >> 2), 1u << (8 * (int)(n & 3)));
            if (j == 0 && ix.head) {
                int sg = seg;  // reads of a tile rarely span more than one segment
                while (sg + 1 < ix.n_seg && n >= ix.seg_read_start[sg + 1]) sg++;
                index_insert(ix, sg, n, word, L, bad != 0);
            }
        }
        if (!more) break;
        if (ix.head) while (seg + 1 < ix.n_seg && n0 >= ix.seg_read_start[seg + 1]) seg++;
    }
}

// ------------------------------------------------------------------------------------------
// placement: leftmost exact occurrence of every read of a segment in one contig
// (upstream lib/BreakageScorer.cpp:235-243, std::string::find)
// ------------------------------------------------------------------------------------------

struct PlaceItem {
    int32_t contig;
    int32_t seg;
    int64_t read_begin;  // global read indices [read_begin, read_end)
    int64_t read_end;
};

struct PlaceArgs {
    const PlaceItem *items;
    // contigs
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const uint8_t *ctg_chars;
    // reads
    ReadSet reads;
    const int32_t *seg_seed_len;    // [S] bases of the seed (1..32) = min read length of the segment, capped
    const int64_t *seg_read_start;  // [S+1]
    // outputs
    int32_t *w;      // position weights: contig c, position p at ctg_off[c] + c + p
    int32_t *total;  // [C] number of reads placed (kmer_breaks)
    int32_t *pos;    // optional: contig c, read n at pos_off[c] + (n - seg_read_start[seg])
    const int64_t *pos_off;
    int32_t tile_len;   // positions per shared-memory tile (multiple of 32, <= 65504)
    int32_t hash_size;  // power of two
    int32_t found_words;  // shared-memory words reserved for the "already placed" bitmask
    int32_t scan_mode;  // 1: exhaustive scan instead of the seed index
};

BS_HD size_t place_smem_bytes(int tile_len, int hash_size, int W, int found_words) {
    const size_t tile_words = (size_t)tile_len / 32 + W + 2;
    return tile_words * 8 + (size_t)hash_size * 4 + tile_words * 4 + (size_t)found_words * 4 + (size_t)tile_len * 2 + 16;
}

// does read n (all-ACGT, length len) occur at tile-relative position pr?
__device__ __forceinline__ bool match_packed(const uint64_t *s_seq, const uint32_t *s_mask, int pr,
                                             const ReadSet &rd, int64_t n, int len, uint64_t r0) {
    const int idx = pr >> 5;
    const uint32_t o = (uint32_t)(pr & 31);
    const int nw = (len + 31) >> 5;
    for (int j = 0; j < nw; j++) {
        const int rem = len - 32 * j;
        const uint64_t rw = (j == 0) ? r0 : __ldg(&rd.words[n * rd.W + j]);
        const uint64_t cw = window64(s_seq[idx + j], s_seq[idx + j + 1], o);
        if ((cw ^ rw) & keep_bases(rem)) return false;
        if (window32(s_mask[idx + j], s_mask[idx + j + 1], o) & keep_bits(rem)) return false;
    }
    return true;
}

__global__ void k_place(PlaceArgs a) {
    const PlaceItem it = a.items[blockIdx.x];
    const int c = it.contig;
    const int64_t coff = a.ctg_off[c];
    const int64_t L = a.ctg_off[c + 1] - coff;
    const int64_t nW = a.ctg_woff[c + 1] - a.ctg_woff[c];
    const uint64_t *gw = a.ctg_words + a.ctg_woff[c];
    const uint32_t *gm = a.ctg_mask + a.ctg_woff[c];
    const int S = a.seg_seed_len[it.seg];
    const int Lt = a.tile_len;
    const int H = a.hash_size;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int tile_words = Lt / 32 + a.reads.W + 2;

    unsigned char *sm = bs_dyn_smem();
    uint64_t *s_seq = (uint64_t *)sm;           sm += (size_t)tile_words * 8;
    uint32_t *s_head = (uint32_t *)sm;          sm += (size_t)H * 4;
    uint32_t *s_mask = (uint32_t *)sm;          sm += (size_t)tile_words * 4;
    uint32_t *s_found = (uint32_t *)sm;         sm += (size_t)a.found_words * 4;
    uint16_t *s_next = (uint16_t *)sm;

    const int64_t nchunk = it.read_end - it.read_begin;
    const bool multi_tile = L > Lt;
    if (multi_tile) {
        for (int i = tid; i < a.found_words; i += nthr) s_found[i] = 0;
    }
    int placed = 0;  // meaningful on lane 0 of each warp

    for (int64_t t0 = 0; t0 == 0 || t0 < L; t0 += Lt) {
        // ---- stage the tile: packed bases + validity mask (with read-length overhang) ----
        __syncthreads();
        for (int i = tid; i < tile_words; i += nthr) {
            const int64_t gi = (t0 >> 5) + i;
            uint64_t wv = 0;
            uint32_t mv = ~0u;
            if (gi < nW) { wv = gw[gi]; mv = gm[gi]; }
            s_seq[i] = wv;
            s_mask[i] = mv;
        }
        for (int i = tid; i < H; i += nthr) s_head[i] = 0;
        __syncthreads();
        // ---- seed index: hash of the S-base window at every position -> chain of positions ----
        if (!a.scan_mode) {
            for (int pr = tid; pr < Lt; pr += nthr) {
                if (t0 + pr + S > L) break;
                const int idx = pr >> 5;
                const uint32_t o = (uint32_t)(pr & 31);
                if (window32(s_mask[idx], s_mask[idx + 1], o) & keep_bits(S)) continue;  // window holds a non-ACGT base
                const uint64_t seed = window64(s_seq[idx], s_seq[idx + 1], o) & keep_bases(S);
                const uint32_t h = seed_hash(seed) & (uint32_t)(H - 1);
                const uint32_t old = atomicExch(&s_head[h], (uint32_t)pr + 1u);
                s_next[pr] = (uint16_t)old;
            }
        }
        __syncthreads();
        // ---- stream the reads of the segment past the tile ----
        for (int64_t base = 0; base < nchunk; base += nthr) {
            const int64_t k = base + tid;
            bool hit = false;
            if (k < nchunk && !(multi_tile && ((s_found[k >> 5] >> (k & 31)) & 1u))) {
                const int64_t n = it.read_begin + k;
                const int len = read_length(a.reads, n);
                int64_t best = -1;
                if (len == 0) {
                    if (t0 == 0) best = 0;  // std::string::find("") == 0
                } else if (a.reads.flags[n] & 1) {
                    // read with bytes outside ACGT: exact byte comparison against the contig text
                    const uint8_t *rc = a.reads.chars + read_begin(a.reads, n);
                    const uint8_t *cc = a.ctg_chars + coff;
                    int64_t pend = t0 + Lt;
                    if (pend > L - len + 1) pend = L - len + 1;
                    for (int64_t p = t0; p < pend; p++) {
                        int i = 0;
                        while (i < len && cc[p + i] == rc[i]) i++;
                        if (i == len) { best = p; break; }
                    }
                } else {
                    const uint64_t r0 = __ldg(&a.reads.words[n * a.reads.W]);
                    if (a.scan_mode) {
                        for (int pr = 0; pr < Lt; pr++) {
                            if (t0 + pr + len > L) break;
                            if (match_packed(s_seq, s_mask, pr, a.reads, n, len, r0)) { best = t0 + pr; break; }
                        }
                    } else {
                        const uint64_t seed = r0 & keep_bases(S);
                        const uint32_t h = seed_hash(seed) & (uint32_t)(H - 1);
                        int bestpr = 0x7fffffff;
                        for (uint32_t q = s_head[h]; q != 0; q = s_next[q - 1]) {
                            const int pr = (int)q - 1;
                            if (pr >= bestpr || t0 + pr + len > L) continue;
                            if (match_packed(s_seq, s_mask, pr, a.reads, n, len, r0)) bestpr = pr;
                        }
                        if (bestpr != 0x7fffffff) best = t0 + bestpr;
                    }
                }
                if (best >= 0) {
                    hit = true;
                    atomicAdd(&a.w[coff + c + best], 1);
                    if (a.pos) a.pos[a.pos_off[c] + (n - a.seg_read_start[it.seg])] = (int32_t)best;
                    if (multi_tile) atomicOr(&s_found[k >> 5], 1u << (k & 31));
                }
            }
            const unsigned m = __ballot_sync(FULL_MASK, hit);
            if ((tid & 31) == 0) placed += __popc(m);
        }
    }
    if ((tid & 31) == 0 && placed) atomicAdd(&a.total[c], placed);
}

// ------------------------------------------------------------------------------------------
// placement through the read index (default).  The contig streams past the segment's read
// index: every contig position looks its seed up, candidates are verified on the packed words
// (XOR of whole words under the validity mask), and the LEFTMOST position per (read, contig)
// wins through an atomicMin on a per-block scratch row -- std::string::find semantics of
// upstream lib/BreakageScorer.cpp:241.  Work per contig is O(L_c + hits) instead of O(U * L_c).
// ------------------------------------------------------------------------------------------

constexpr uint32_t POS_INF = 0x7f7f7f7fu;  // memset-able "no position yet"

struct PlaceIxArgs {
    const int32_t *order;  // [n_items] contig ids, longest first
    int32_t n_items;
    int32_t *work_counter;  // zero on entry
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const uint8_t *ctg_chars;
    const int32_t *ctg_seg;
    ReadSet reads;
    ReadIndex ix;
    uint32_t *best;       // [gridDim][best_stride], all POS_INF on entry and on exit
    int64_t best_stride;  // >= reads of the largest segment
    int32_t *w;           // position weights: contig c, position p at ctg_off[c] + c + p
    int32_t *total;       // [C] reads placed (kmer_breaks)
    int32_t *pos;         // optional
    const int64_t *pos_off;
    int32_t hit_cap;      // shared-memory list of reads placed in the current contig
};

// leftmost byte-exact occurrence of read n in contig text, or -1 (reads that cannot be packed)
__device__ __forceinline__ int64_t find_bytes(const uint8_t *cc, int64_t L, const uint8_t *rc, int len) {
    if (len == 0) return 0;  // std::string::find("") == 0
    for (int64_t p = 0; p + len <= L; p++) {
        int i = 0;
        while (i < len && cc[p + i] == rc[i]) i++;
        if (i == len) return p;
    }
    return -1;
}

// does read n (length len, packed words rw) equal the contig at position p?
__device__ __forceinline__ bool verify_at(const PlaceIxArgs &a, const uint64_t *gw, const uint32_t *gm, const uint8_t *cc,
                                          int64_t p, int64_t n, int len, uint64_t w0) {
    if (a.reads.flags[n] & 1) {  // bytes outside ACGT somewhere in (or next to) the read: compare the text
        const uint8_t *rc = a.reads.chars + read_begin(a.reads, n);
        for (int i = 0; i < len; i++)
            if (cc[p + i] != rc[i]) return false;
        return true;
    }
    const int64_t idx = p >> 5;
    const uint32_t o = (uint32_t)(p & 31);
    const int nw = (len + 31) >> 5;
    const uint64_t *rw = a.reads.words + n * a.reads.W;
    uint64_t c_lo = __ldg(&gw[idx]);
    uint32_t m_lo = __ldg(&gm[idx]);
    for (int j = 0; j < nw; j++) {
        const int rem = len - 32 * j;
        const uint64_t c_hi = __ldg(&gw[idx + j + 1]);
        const uint32_t m_hi = __ldg(&gm[idx + j + 1]);
        const uint64_t r = (j == 0) ? w0 : __ldg(&rw[j]);
        if ((window64(c_lo, c_hi, o) ^ r) & keep_bases(rem)) return false;
        if (window32(m_lo, m_hi, o) & keep_bits(rem)) return false;
        c_lo = c_hi;
        m_lo = m_hi;
    }
    return true;
}

constexpr int PLACE_CAND_CAP = 256;  // seed hits of one WARP iteration (128 positions) awaiting verification

BS_HD size_t place_index_smem_bytes(int hit_cap, int nthr) { return (size_t)hit_cap * 4 + (size_t)(nthr / 32) * PLACE_CAND_CAP * 8; }

__global__ void __launch_bounds__(256, 5) k_place_index(PlaceIxArgs a) {
    uint32_t *s_hits = (uint32_t *)bs_dyn_smem();
    __shared__ int s_item, s_nhit, s_placed, s_ncand[32];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    // every warp queues and verifies its own candidates: no block barrier inside the position loop
    uint2 *s_cand = (uint2 *)(s_hits + a.hit_cap) + warp * PLACE_CAND_CAP;  // (read id, contig position) with an equal seed tag
    uint32_t *best = a.best + (int64_t)blockIdx.x * a.best_stride;
    for (;;) {
        __syncthreads();
        if (tid == 0) { s_item = atomicAdd(a.work_counter, 1); s_nhit = 0; s_placed = 0; }
        if (lane == 0) s_ncand[warp] = 0;
        __syncthreads();
        const int item = s_item;
        if (item >= a.n_items) break;
        const int c = a.order[item];
        const int s = a.ctg_seg[c];
        const int64_t coff = a.ctg_off[c];
        const int64_t L = a.ctg_off[c + 1] - coff;
        const uint64_t *gw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *gm = a.ctg_mask + a.ctg_woff[c];
        const uint8_t *cc = a.ctg_chars + coff;
        const int64_t r0 = a.ix.seg_read_start[s];
        const int64_t n_seg_reads = a.ix.seg_read_start[s + 1] - r0;
        const int S = a.ix.seed_len[s];
        const uint32_t *head = a.ix.head + a.ix.tab_off[s];
        const uint32_t hmask = (uint32_t)a.ix.tab_mask[s];
        const uint64_t keepS = keep_bases(S);
        const uint32_t keepSm = keep_bits(S);

        // a read with an equal seed at position p: verify the rest, keep the leftmost position
        auto verify_and_record = [&](int64_t n, int64_t p) {
            const int len = read_length(a.reads, n);
            if (p + len > L) return;
            const uint64_t w0 = __ldg(&a.reads.words[n * a.reads.W]);
            if (!verify_at(a, gw, gm, cc, p, n, len, w0)) return;
            const uint32_t old = atomicMin(&best[n - r0], (uint32_t)p);
            if (old == POS_INF) {  // first time this read is seen in this contig
                const int slot = atomicAdd(&s_nhit, 1);
                if (slot < a.hit_cap) s_hits[slot] = (uint32_t)(n - r0);
            }
        };

        if (n_seg_reads > 0) {
            // ---- every contig position against the index.  Phase 1 (sparse, cheap): four positions
            // per thread in flight, bucket head -> chain entry; equal seed tags are queued per warp.
            // Phase 2 (dense): the warp's lanes verify one queued candidate each on the packed words. ----
            for (int64_t p0 = 0; p0 + S <= L; p0 += 4 * (int64_t)nthr) {
                // a thread takes four CONSECUTIVE positions (never straddling a word boundary)
                const int64_t pb = p0 + 4 * (int64_t)tid;
                uint32_t q4[4];
                uint64_t seed4[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { q4[u] = 0; seed4[u] = 0; }
                if (pb + S <= L) {
                    const int64_t idx = pb >> 5;
                    const uint32_t o = (uint32_t)(pb & 31);
                    const uint64_t cw0 = __ldg(&gw[idx]), cw1 = __ldg(&gw[idx + 1]);
                    const uint32_t cm0 = __ldg(&gm[idx]), cm1 = __ldg(&gm[idx + 1]);
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (pb + u + S <= L && !(window32(cm0, cm1, o + u) & keepSm)) {  // else: a non-ACGT base in the seed window
                            seed4[u] = window64(cw0, cw1, o + u) & keepS;
                            q4[u] = head[seed_hash(seed4[u]) & hmask];
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int64_t p = pb + u;
                    uint32_t q = q4[u];
                    const uint32_t tag = seed_tag(seed4[u]);
                    while (q != 0) {  // the chain entry carries a 32-bit tag of the read's seed: the packed
                        const int64_t n = (int64_t)q - 1;  // reads themselves are only touched in phase 2
                        const uint2 e = a.ix.next[n];
                        q = e.x;
                        if (e.y != tag) continue;
                        const int slot = atomicAdd(&s_ncand[warp], 1);
                        if (slot < PLACE_CAND_CAP) s_cand[slot] = make_uint2((uint32_t)n, (uint32_t)p);
                        else verify_and_record(n, p);  // queue full (long chains): verify in place
                    }
                }
                __syncwarp();
                const int nc = s_ncand[warp] < PLACE_CAND_CAP ? s_ncand[warp] : PLACE_CAND_CAP;
                for (int i = lane; i < nc; i += 32) {
                    const uint2 cd = s_cand[i];
                    verify_and_record((int64_t)cd.x, (int64_t)cd.y);
                }
                __syncwarp();
                if (lane == 0) s_ncand[warp] = 0;
                __syncwarp();
            }
            // ---- reads outside the index: text comparison, one read per thread ----
            int i = 0;
            for (uint32_t q = a.ix.odd_head[s]; q != 0; i++) {
                const int64_t n = (int64_t)q - 1;
                q = a.ix.next[n].x;
                if (i % nthr != tid) continue;
                const int64_t p = find_bytes(cc, L, a.reads.chars + read_begin(a.reads, n), read_length(a.reads, n));
                if (p < 0) continue;
                best[n - r0] = (uint32_t)p;  // each such read is visited by exactly one thread
                const int slot = atomicAdd(&s_nhit, 1);
                if (slot < a.hit_cap) s_hits[slot] = (uint32_t)(n - r0);
            }
        }
        __threadfence_block();
        __syncthreads();
        // ---- leftmost positions -> position weights ----
        const int nh = s_nhit;
        int placed = 0;
        if (nh <= a.hit_cap) {
            for (int h = tid; h < nh; h += nthr) {
                const uint32_t nl = s_hits[h];
                const uint32_t p = atomicExch(&best[nl], POS_INF);
                atomicAdd(&a.w[coff + c + p], 1);
                if (a.pos) a.pos[a.pos_off[c] + nl] = (int32_t)p;
                placed++;
            }
        } else {  // more distinct reads than the list holds: sweep the whole scratch row
            for (int64_t nl = tid; nl < n_seg_reads; nl += nthr) {
                const uint32_t p = atomicExch(&best[nl], POS_INF);
                if (p == POS_INF) continue;
                atomicAdd(&a.w[coff + c + p], 1);
                if (a.pos) a.pos[a.pos_off[c] + nl] = (int32_t)p;
                placed++;
            }
        }
        if (placed) atomicAdd(&s_placed, placed);
        __syncthreads();
        if (tid == 0) a.total[c] = s_placed;
    }
}

// ------------------------------------------------------------------------------------------
// block-level helpers
// ------------------------------------------------------------------------------------------

// sum of v over the block, fixed tree => bit-reproducible; result valid on every thread
__device__ __forceinline__ double block_sum(double v, double *s_red) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    __syncthreads();
    s_red[tid] = v;
    __syncthreads();
    for (int s = nthr >> 1; s > 0; s >>= 1) {
        if (tid < s) s_red[tid] += s_red[tid + s];
        __syncthreads();
    }
    return s_red[0];
}
__device__ __forceinline__ double block_max(double v, double *s_red) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    __syncthreads();
    s_red[tid] = v;
    __syncthreads();
    for (int s = nthr >> 1; s > 0; s >>= 1) {
        if (tid < s) { const double o = s_red[tid + s]; if (o > s_red[tid]) s_red[tid] = o; }
        __syncthreads();
    }
    return s_red[0];
}
// exclusive prefix of v over the block (thread order); s_scan holds blockDim ints
__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *s_scan, int64_t *total) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    __syncthreads();
    s_scan[tid] = v;
    __syncthreads();
    for (int d = 1; d < nthr; d <<= 1) {
        int64_t add = (tid >= d) ? s_scan[tid - d] : 0;
        __syncthreads();
        s_scan[tid] += add;
        __syncthreads();
    }
    if (total) *total = s_scan[nthr - 1];
    return s_scan[tid] - v;
}

// ------------------------------------------------------------------------------------------
// break k-mers -> probability-weighted sums (upstream lib/BreakageScorer.cpp:244-303), the
// optional dense histogram, and the KS statistic of the normalised break histogram (variant
// lib/DeNovoAssembler.cpp:395-420: x = count_row / total for every table row, y = truth
// distribution) -- one pass family over the position weights of a contig
// ------------------------------------------------------------------------------------------

// scoring table over the dense k-mer index space (all lengths 1..8), one 16-byte gather per break
struct alignas(16) TabEntry {
    double prob;
    int32_t row;
    int32_t pad;
};

constexpr int CC_DENSE = 4096;  // counts below this are tallied in a dense shared-memory array
constexpr int OVF_CAP = 4096;   // per-block capacity for larger counts

struct ScoreArgs {
    const int32_t *order;   // [C] contig ids, longest first
    int32_t *work_counter;  // zero on entry
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const int32_t *ctg_seg;
    const int32_t *w;
    const int32_t *total;
    const TabEntry *tab;     // [DENSE_SIZE] probability and table row of a dense k-mer index (row -1 if absent)
    int32_t kmer;
    int32_t T;
    int64_t n_contigs;
    int32_t *sequence_len;
    double *bp_score;
    double *norm_by_break_freqs;
    double *norm_by_len;
    int32_t *kmer_breaks;
    int32_t *hist;  // optional [C][T+1]
    // KS of the normalised break histogram (all NULL / 0 when not wanted)
    double *ks_b;          // [C]
    const double *yv;      // [R_y] sorted distinct truth-table values
    const int32_t *ycum;   // [S][R_y] inclusive cumulative counts of the truth distribution
    int32_t R_y;
    int32_t zero_le, zero_lt;  // indices into ycum of the last y value <= 0.0 / < 0.0 (-1: none)
    double y_max;          // largest truth-table value
    int32_t *scratch;      // [gridDim][T+1], all zero on entry and on exit
    int32_t *ovf_cnt;      // [gridDim][OVF_CAP] counts >= CC_DENSE (unsorted)
    int32_t *status;       // set to 1 if a block ran out of overflow space
};

// number of y values <= v (le) and < v (lt), from the cumulative counts
__device__ __forceinline__ void y_counts_at(const double *yv, const int32_t *ycum, int R_y, double y_max, double v,
                                            double *le, double *lt) {
    if (R_y <= 0) { *le = 0.0; *lt = 0.0; return; }
    if (v > y_max) { *le = *lt = (double)ycum[R_y - 1]; return; }  // the usual case: count/total >> any table value
    int lo = 0, hi = R_y;  // first index with yv > v
    while (lo < hi) { int mid = (lo + hi) >> 1; if (yv[mid] <= v) lo = mid + 1; else hi = mid; }
    const int ule = lo;  // #distinct <= v
    lo = 0; hi = R_y;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (yv[mid] < v) lo = mid + 1; else hi = mid; }
    const int ult = lo;  // #distinct < v
    *le = ule > 0 ? (double)ycum[ule - 1] : 0.0;
    *lt = ult > 0 ? (double)ycum[ult - 1] : 0.0;
}

// sum over the block in a fixed order (lane tree, then warps in order): bit-reproducible, the
// same for every GPU count; valid on thread 0
__device__ __forceinline__ double block_sum_fixed(double v, double *s_w) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(FULL_MASK, v, m);
    __syncthreads();
    if (lane == 0) s_w[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) for (int i = 0; i < nwarp; i++) t += s_w[i];
    return t;
}

__global__ void __launch_bounds__(256) k_break_score(ScoreArgs a) {
    __shared__ double s_w[32];
    __shared__ int32_t s_cc[CC_DENSE];  // rows having count j
    __shared__ int s_item, s_novf, s_maxc, s_nz;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const bool want_ks = a.ks_b != nullptr;
    int32_t *scratch = want_ks ? a.scratch + (int64_t)blockIdx.x * (a.T + 1) : nullptr;
    int32_t *ovf = want_ks ? a.ovf_cnt + (int64_t)blockIdx.x * OVF_CAP : nullptr;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    if (want_ks) for (int i = tid; i < CC_DENSE; i += nthr) s_cc[i] = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) { s_item = atomicAdd(a.work_counter, 1); s_novf = 0; s_maxc = 0; s_nz = 0; }
        __syncthreads();
        if (s_item >= a.n_contigs) break;
        const int64_t c = a.order[s_item];
        const int64_t coff = a.ctg_off[c];
        const int64_t L = a.ctg_off[c + 1] - coff;
        const uint64_t *gw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *gm = a.ctg_mask + a.ctg_woff[c];
        const int32_t *w = a.w + coff + c;
        const int32_t total = a.total[c];
        const int64_t np = L > 0 ? L : 1;
        double s1 = 0.0, s2 = 0.0;
        // pass 1: weighted sums in position order (+ histogram, + per-row counts for the KS); four
        // positions per thread in flight so that the table gathers overlap
        if (total != 0) {
            for (int64_t p0 = 0; p0 < np; p0 += 4 * (int64_t)nthr) {
                int32_t wv[4];
                TabEntry te[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int64_t p = p0 + (int64_t)u * nthr + tid;
                    wv[u] = p < np ? w[p] : 0;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    te[u].prob = 0.0;
                    te[u].row = -1;
                    if (wv[u] != 0) {
                        const BreakWindow bw = break_window(p0 + (int64_t)u * nthr + tid, a.kmer, L);
                        const int di = dense_index_at(gw, gm, bw.start, bw.len);
                        if (di >= 0) te[u] = a.tab[di];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (wv[u] == 0) continue;
                    const int32_t row = te[u].row;
                    if (row >= 0) {
                        s1 += te[u].prob * (double)wv[u];
                        s2 += te[u].prob * ((double)wv[u] / (double)total);
                        if (want_ks) atomicAdd(&scratch[row], wv[u]);
                    }
                    if (a.hist) atomicAdd(&a.hist[c * (int64_t)(a.T + 1) + (row >= 0 ? row : a.T)], wv[u]);
                }
            }
        }
        s1 = block_sum_fixed(s1, s_w);
        s2 = block_sum_fixed(s2, s_w);
        if (tid == 0) {
            if (a.sequence_len) a.sequence_len[c] = (int32_t)L;
            if (a.bp_score) a.bp_score[c] = s1;
            if (a.norm_by_break_freqs) a.norm_by_break_freqs[c] = s2;
            if (a.norm_by_len) a.norm_by_len[c] = s1 / (double)(int32_t)L;
            if (a.kmer_breaks) a.kmer_breaks[c] = total;
        }
        if (!want_ks) continue;
        if (total == 0) {  // 0/0 for every row: R drops the NaNs and ks.test stops on empty x
            if (tid == 0) a.ks_b[c] = qnan;
            continue;
        }
        __threadfence_block();
        __syncthreads();
        // pass 2: whoever swaps a row's count out first owns it; tally rows per count value
        for (int64_t p0 = 0; p0 < np; p0 += 4 * (int64_t)nthr) {
            int32_t wv[4], row[4], cnt[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t p = p0 + (int64_t)u * nthr + tid;
                wv[u] = p < np ? w[p] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                row[u] = -1;
                if (wv[u] != 0) {
                    const BreakWindow bw = break_window(p0 + (int64_t)u * nthr + tid, a.kmer, L);
                    const int di = dense_index_at(gw, gm, bw.start, bw.len);
                    if (di >= 0) row[u] = a.tab[di].row;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) cnt[u] = row[u] >= 0 ? atomicExch(&scratch[row[u]], 0) : 0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (cnt[u] == 0) continue;
                atomicAdd(&s_nz, 1);
                if (cnt[u] < CC_DENSE) { atomicAdd(&s_cc[cnt[u]], 1); atomicMax(&s_maxc, cnt[u]); }
                else {
                    const int slot = atomicAdd(&s_novf, 1);
                    if (slot < OVF_CAP) ovf[slot] = cnt[u]; else *a.status = 1;
                }
            }
        }
        __syncthreads();
        // the distinct x values are 0 and count/total for the few distinct counts: thread 0 walks them
        if (tid == 0) {
            const int32_t *ycum = a.ycum + (int64_t)a.ctg_seg[c] * a.R_y;
            const int64_t n_y = a.R_y > 0 ? ycum[a.R_y - 1] : 0;
            const int novf = s_novf < OVF_CAP ? s_novf : OVF_CAP;
            double d = 0.0;
            if (n_y > 0 && a.T > 0) {
                const double inx = (double)a.T, iny = (double)n_y;
                int64_t run = (int64_t)a.T - s_nz;  // rows never broken: x value 0
                if (run > 0) {
                    const double le = a.zero_le >= 0 ? (double)ycum[a.zero_le] : 0.0;
                    const double lt = a.zero_lt >= 0 ? (double)ycum[a.zero_lt] : 0.0;
                    double d1 = lt / iny, d2 = (double)run / inx - le / iny;
                    if (d1 < 0) d1 = -d1;
                    if (d2 < 0) d2 = -d2;
                    d = d1 > d2 ? d1 : d2;
                }
                for (int j = 1; j <= s_maxc; j++) {
                    const int32_t cnt = s_cc[j];
                    if (cnt == 0) continue;
                    s_cc[j] = 0;
                    double le, lt;
                    y_counts_at(a.yv, ycum, a.R_y, a.y_max, (double)j / (double)total, &le, &lt);
                    double d1 = (double)run / inx - lt / iny;
                    run += cnt;
                    double d2 = (double)run / inx - le / iny;
                    if (d1 < 0) d1 = -d1;
                    if (d2 < 0) d2 = -d2;
                    if (d1 > d) d = d1;
                    if (d2 > d) d = d2;
                }
                // counts >= CC_DENSE: few; walked in ascending order
                int32_t last = CC_DENSE - 1;
                for (int done = 0; done < novf;) {
                    int32_t cur = 0x7fffffff;
                    int mult = 0;
                    for (int i = 0; i < novf; i++) {
                        const int32_t v = ovf[i];
                        if (v > last && v < cur) { cur = v; mult = 1; }
                        else if (v == cur) mult++;
                    }
                    double le, lt;
                    y_counts_at(a.yv, ycum, a.R_y, a.y_max, (double)cur / (double)total, &le, &lt);
                    double d1 = (double)run / inx - lt / iny;
                    run += mult;
                    double d2 = (double)run / inx - le / iny;
                    if (d1 < 0) d1 = -d1;
                    if (d2 < 0) d2 = -d2;
                    if (d1 > d) d = d1;
                    if (d2 > d) d = d2;
                    last = cur;
                    done += mult;
                }
            } else {
                for (int j = 1; j <= s_maxc; j++) s_cc[j] = 0;
            }
            a.ks_b[c] = (n_y > 0 && a.T > 0) ? d : qnan;
        }
    }
}

// ------------------------------------------------------------------------------------------
// truth-side distribution for the KS statistics (kmer_from_seq, upstream
// lib/GenerateReads.R:243-259): per segment, counts of truth windows per distinct table value
// ------------------------------------------------------------------------------------------

// per x-value rank: index into the truth cumulative counts of the last y value <= / < that x value
// (table level), or the cumulative counts themselves (segment level, yx below)
struct LeLt {
    int32_t le, lt;
};

struct SpectrumArgs {
    const int64_t *tr_off;
    const int64_t *tr_woff;
    const uint64_t *tr_words;
    const uint32_t *tr_mask;
    const int32_t *rank_y;  // [4^kmer] rank of the window's truth-table value, -1 = not in table (NA)
    int32_t *ycnt;          // [S][R_y]
    const LeLt *lelt;       // [R_x] table-level indices (NULL: no yx output)
    LeLt *yx;               // [S][R_x] #{truth windows <= x value}, #{truth windows < x value} per x-value rank
    int32_t R_x;
    int32_t R_y;
    int32_t kmer;
    int32_t blocks_per_seg;
};

__global__ void k_truth_spectrum(SpectrumArgs a) {
    const int64_t s = blockIdx.x / a.blocks_per_seg;
    const int b = blockIdx.x % a.blocks_per_seg;
    const int64_t L = a.tr_off[s + 1] - a.tr_off[s];
    const uint64_t *gw = a.tr_words + a.tr_woff[s];
    const uint32_t *gm = a.tr_mask + a.tr_woff[s];
    const int64_t nwin = L - a.kmer + 1;
    const int doff = dense_offset(a.kmer);
    for (int64_t p = (int64_t)b * blockDim.x + threadIdx.x; p < nwin; p += (int64_t)a.blocks_per_seg * blockDim.x) {
        const int di = dense_index_at(gw, gm, p, a.kmer);
        if (di < 0) continue;
        const int32_t rk = a.rank_y[di - doff];
        if (rk >= 0) atomicAdd(&a.ycnt[s * a.R_y + rk], 1);
    }
}

// The same for truths of fewer than 65 536 windows, one block per segment: counts in shared
// memory (two 16-bit counters per word), cumulated in place, written out once, coalesced.
__global__ void __launch_bounds__(512, 3) k_truth_spectrum_smem(SpectrumArgs a) {
    uint32_t *s_h = (uint32_t *)bs_dyn_smem();
    __shared__ uint32_t s_wsum[32];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int64_t s = blockIdx.x;
    const int64_t L = a.tr_off[s + 1] - a.tr_off[s];
    const uint64_t *gw = a.tr_words + a.tr_woff[s];
    const uint32_t *gm = a.tr_mask + a.tr_woff[s];
    const int64_t nwin = L - a.kmer + 1;
    const int kshift = 64 - 2 * a.kmer;
    const uint32_t kbits = keep_bits(a.kmer);
    const int nword = (a.R_y + 1) >> 1;
    for (int i = tid; i < nword; i += nthr) s_h[i] = 0;
    __syncthreads();
    // a thread takes four CONSECUTIVE windows (never straddling a 32-base word boundary): the truth
    // words are loaded once per four windows and the four rank gathers are in flight together
    for (int64_t p0 = 0; p0 < nwin; p0 += 4 * (int64_t)nthr) {
        const int64_t pb = p0 + 4 * (int64_t)tid;
        if (pb >= nwin) continue;
        const int64_t wi = pb >> 5;
        const uint32_t o = (uint32_t)(pb & 31);
        const uint64_t w0 = __ldg(&gw[wi]), w1 = __ldg(&gw[wi + 1]);
        const uint32_t m0 = __ldg(&gm[wi]), m1 = __ldg(&gm[wi + 1]);
        int32_t rk[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            rk[u] = -1;
            if (pb + u < nwin && !(window32(m0, m1, o + u) & kbits)) rk[u] = __ldg(&a.rank_y[window64(w0, w1, o + u) >> kshift]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (rk[u] >= 0) atomicAdd(&s_h[rk[u] >> 1], 1u << (16 * (rk[u] & 1)));
    }
    __syncthreads();
    // inclusive prefix over ranks: every thread owns an even number of consecutive ranks
    int per = (a.R_y + nthr - 1) / nthr;
    per += per & 1;
    const int lo = tid * per < a.R_y ? tid * per : a.R_y;
    const int hi = lo + per < a.R_y ? lo + per : a.R_y;
    uint32_t sum = 0;
    for (int i = lo >> 1; i < (hi + 1) >> 1; i++) { const uint32_t w = s_h[i]; sum += (w & 0xffffu) + (w >> 16); }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL_MASK, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t run = incl - sum;
    for (int w = 0; w < warp; w++) run += s_wsum[w];
    for (int i = lo >> 1; i < (hi + 1) >> 1; i++) {
        const uint32_t w = s_h[i];
        const uint32_t c0 = run + (w & 0xffffu), c1 = c0 + (w >> 16);
        s_h[i] = c0 | (c1 << 16);  // totals stay below 65 536
        run = c1;
    }
    __syncthreads();
    int32_t *out = a.ycnt + s * a.R_y;
    for (int i = tid; i < a.R_y; i += nthr) out[i] = (int32_t)((s_h[i >> 1] >> (16 * (i & 1))) & 0xffffu);
    if (a.lelt) {  // the same counts looked up per x-value rank: the KS sweep then needs one gather, not two
        LeLt *yx = a.yx + s * a.R_x;
        for (int i = tid; i < a.R_x; i += nthr) {
            const LeLt q = a.lelt[i];
            LeLt o;
            o.le = q.le >= 0 ? (int32_t)((s_h[q.le >> 1] >> (16 * (q.le & 1))) & 0xffffu) : 0;
            o.lt = q.lt >= 0 ? (int32_t)((s_h[q.lt >> 1] >> (16 * (q.lt & 1))) & 0xffffu) : 0;
            yx[i] = o;
        }
    }
}

// yx from the cumulative counts in global memory (after k_truth_spectrum + k_row_cumsum)
__global__ void k_yx_gather(SpectrumArgs a, int64_t n_seg) {
    const int64_t total = n_seg * a.R_x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = t / a.R_x;
        const int i = (int)(t - s * a.R_x);
        const LeLt q = a.lelt[i];
        const int32_t *ycum = a.ycnt + s * a.R_y;
        LeLt o;
        o.le = q.le >= 0 ? ycum[q.le] : 0;
        o.lt = q.lt >= 0 ? ycum[q.lt] : 0;
        a.yx[t] = o;
    }
}

// in-place inclusive prefix sum of every row of a [rows][R] int32 matrix; one block per row
__global__ void k_row_cumsum(int32_t *m, int32_t R) {
    int64_t *s_scan = (int64_t *)bs_dyn_smem();
    int32_t *row = m + (int64_t)blockIdx.x * R;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int per = (R + nthr - 1) / nthr;
    const int lo = tid * per, hi = (lo + per < R) ? lo + per : R;
    int64_t sum = 0;
    for (int i = lo; i < hi; i++) sum += row[i];
    int64_t run = block_exclusive_scan(sum, s_scan, nullptr);
    for (int i = lo; i < hi; i++) { run += row[i]; row[i] = (int32_t)run; }
}

// ------------------------------------------------------------------------------------------
// rolling-window probabilities + KS statistic against the truth distribution
// (upstream lib/BreakageScorer.cpp:200-215 and lib/DeNovoAssembler.R:416-424)
// ------------------------------------------------------------------------------------------

// one entry per k-mer code of the window length: what a window contributes to path_prob_dist
// (0.0 when the k-mer is not a table row) and the rank of that value among the distinct x values
struct alignas(16) WinEntry {
    double prob;
    int32_t rank;
    int32_t pad;
};

struct ProbDistArgs {
    const int32_t *order;     // [C] contig ids, longest first
    int32_t *work_counter;    // zero on entry
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const int32_t *ctg_seg;  // [C] segment of a contig
    const WinEntry *win;     // [4^kmer], NULL when kmer is outside 1..8 (every window is "not in the table")
    const LeLt *yx;          // [S][R_x] truth windows <= / < every x value
    const int32_t *ycum;     // [S][R_y] inclusive cumulative counts of the truth distribution
    int32_t R_x, R_y;
    int32_t rank_zero;       // rank of the value 0.0 (windows not in the table)
    int32_t kmer;
    int64_t n_contigs;
    double *prob_dist;       // optional
    const int64_t *pd_off;
    double *ks;              // optional [C]
    uint32_t *rank_scratch;  // [gridDim][hist_words + n_ranges] global rank histogram when it does not fit shared memory, else NULL
    int32_t hist_words;      // 32-bit words of one rank histogram (hist_phys_words)
    int32_t n_ranges;        // hist_ranges: ranges of 32 histogram words
};

// Rank histogram layout.  PACKED: two 16-bit counters per 32-bit word (every contig of the launch
// has < 65536 windows), else one counter per word.  Words are grouped into RANGES of 32 (one
// thread sweeps one range) with one pad word after every range, which keeps a warp's
// simultaneous sweeps on different banks.  A bitmap word per range marks the non-empty words, so
// a sweep touches only those: its cost follows the contig's windows, not the number of ranks.
template <bool PACKED>
__device__ __forceinline__ int hist_logical_word(int r) { return PACKED ? r >> 1 : r; }
BS_HD int hist_logical_words(int R_x, bool packed) { return packed ? (R_x + 1) / 2 : R_x; }
BS_HD int hist_phys_words(int R_x, bool packed) {
    const int w = hist_logical_words(R_x, packed);
    return w + (w >> 5) + 2;
}
BS_HD int hist_ranges(int R_x, bool packed) { return (hist_logical_words(R_x, packed) + 31) / 32; }

template <bool PACKED, bool IN_SMEM>
__global__ void __launch_bounds__(768, 2) k_prob_dist_ks(ProbDistArgs a) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    __shared__ int64_t s_wsum[32];
    __shared__ int64_t s_wmax[32];
    __shared__ int s_item;
    // dynamic shared memory (or the global scratch row): histogram words, then the range bitmaps
    uint32_t *s_hist = IN_SMEM ? (uint32_t *)bs_dyn_smem() : a.rank_scratch + (int64_t)blockIdx.x * (a.hist_words + a.n_ranges);
    uint32_t *s_bm = s_hist + a.hist_words;
    const bool want_ks = a.ks != nullptr;
    const int kshift = 64 - 2 * a.kmer;
    const uint32_t kbits = keep_bits(a.kmer);
    if (want_ks) {
        for (int i = tid; i < a.hist_words + a.n_ranges; i += nthr) s_hist[i] = 0;
    }
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(a.work_counter, 1);
        __syncthreads();
        if (s_item >= a.n_contigs) break;
        const int64_t c = a.order[s_item];
        const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
        const uint64_t *gw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *gm = a.ctg_mask + a.ctg_woff[c];
        int64_t nwin = L - a.kmer + 1;
        if (nwin < 0) nwin = 0;
        double *pd = a.prob_dist ? a.prob_dist + a.pd_off[c] : nullptr;
        // truth side of this contig's segment, loaded early so that the latency hides behind the windows
        const int64_t seg = a.ctg_seg[c];
        const LeLt *yx = want_ks ? a.yx + seg * a.R_x : nullptr;
        const int64_t n_y = (want_ks && a.R_y > 0) ? a.ycum[seg * a.R_y + a.R_y - 1] : 0;
        // ---- windows: table value out, rank histogram in.  A thread takes four CONSECUTIVE
        // positions (they never straddle a 32-base word boundary), so the contig words are
        // loaded once per four windows; the four table gathers are in flight together ----
        for (int64_t p0 = 0; p0 < nwin; p0 += 4 * (int64_t)nthr) {
            const int64_t pb = p0 + 4 * (int64_t)tid;
            double val[4];
            int32_t rk[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { val[u] = 0.0; rk[u] = a.rank_zero; }
            if (pb < nwin && a.win) {
                const int64_t wi = pb >> 5;
                const uint32_t o = (uint32_t)(pb & 31);  // multiple of 4: o + 3 <= 31
                const uint64_t w0 = __ldg(&gw[wi]), w1 = __ldg(&gw[wi + 1]);
                const uint32_t m0 = __ldg(&gm[wi]), m1 = __ldg(&gm[wi + 1]);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (pb + u < nwin && !(window32(m0, m1, o + u) & kbits)) {
                        const WinEntry e = a.win[window64(w0, w1, o + u) >> kshift];  // one 16-byte gather
                        val[u] = e.prob;
                        rk[u] = e.rank;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (pb + u < nwin) {
                    if (pd) pd[pb + u] = val[u];
                    if (want_ks) {
                        const int lw = hist_logical_word<PACKED>(rk[u]);
                        atomicAdd(&s_hist[lw + (lw >> 5)], PACKED ? 1u << (16 * (rk[u] & 1)) : 1u);
                        atomicOr(&s_bm[lw >> 5], 1u << (lw & 31));
                    }
                }
            }
        }
        if (!want_ks) continue;
        __syncthreads();
        // ---- D = sup |F_x - F_y| over the pooled distinct values, evaluated at every x value that
        // is present: just below it (F_x of the previous step vs #{y < v}) and at it ----
        const bool defined = nwin > 0 && n_y > 0;
        // |F_x - F_y| = |run * n_y - ycount * nwin| / (nwin * n_y): the numerator is maximised in exact
        // 64-bit integers (both factors are below 2^31), one division at the end
        int64_t best = 0;
        int64_t carry = 0;  // windows in the ranges of earlier rounds (same value on every thread)
        for (int r0 = 0; r0 < a.n_ranges; r0 += nthr) {  // one range of 32 words per thread and round
            const int r = r0 + tid;
            const uint32_t m = r < a.n_ranges ? s_bm[r] : 0u;
            if (m) s_bm[r] = 0;
            const int w0 = 33 * r;  // physical index of the range's first word
            int64_t cnt_r = 0;
            for (uint32_t mm = m; mm; mm &= mm - 1) {
                const uint32_t w = s_hist[w0 + __ffs((int)mm) - 1];
                cnt_r += PACKED ? (w & 0xffffu) + (w >> 16) : w;
            }
            // exclusive prefix of the range counts over the block (+ what earlier rounds held)
            int64_t incl = cnt_r;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                const int64_t o = __shfl_up_sync(FULL_MASK, incl, dd);
                if (lane >= dd) incl += o;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            int64_t run = carry + incl - cnt_r;
            for (int w = 0; w < nwarp; w++) {
                if (w < warp) run += s_wsum[w];
                carry += s_wsum[w];
            }
            for (uint32_t mm = m; mm; mm &= mm - 1) {
                const int k = __ffs((int)mm) - 1;
                const uint32_t w = s_hist[w0 + k];
                s_hist[w0 + k] = 0;  // leave the histogram zeroed for the next contig
#pragma unroll
                for (int h = 0; h < (PACKED ? 2 : 1); h++) {
                    const uint32_t cnt = PACKED ? (w >> (16 * h)) & 0xffffu : w;
                    if (cnt == 0 || !defined) continue;
                    const int i = PACKED ? 64 * r + 2 * k + h : 32 * r + k;
                    const LeLt q = yx[i];  // one 8-byte gather
                    const int64_t lt = q.lt, le = q.le;
                    int64_t d1 = run * n_y - lt * nwin;
                    run += cnt;
                    int64_t d2 = run * n_y - le * nwin;
                    if (d1 < 0) d1 = -d1;
                    if (d2 < 0) d2 = -d2;
                    if (d1 > best) best = d1;
                    if (d2 > best) best = d2;
                }
            }
            __syncthreads();  // s_wsum is rewritten by the next round
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            const int64_t o = __shfl_xor_sync(FULL_MASK, best, m);
            if (o > best) best = o;
        }
        if (lane == 0) s_wmax[warp] = best;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < nwarp; w++) if (s_wmax[w] > best) best = s_wmax[w];
            a.ks[c] = defined ? (double)best / ((double)nwin * (double)n_y) : __longlong_as_double(0x7ff8000000000000ll);
        }
    }
}

// ------------------------------------------------------------------------------------------
// contig-in-truth offset (upstream lib/BreakageScorer.cpp:273-274): leftmost exact occurrence
// of the whole contig in its segment's truth, assigned only if at least one read was placed.
//
// k_startpos_index: the truth streams past a small shared-memory table of the segment's contig
// SEEDS (first 32 bases); a seed hit is verified by the whole warp on packed words; the leftmost
// verified position wins through atomicMin.  O(L_truth + sum L_c) per segment instead of
// O(C * L_truth).  Contigs without a packable seed (shorter than 32, or a non-ACGT byte in it)
// are left to k_startpos, which also turns the atomicMin scratch into the final values.
// ------------------------------------------------------------------------------------------

struct StartposArgs {
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const uint8_t *ctg_chars;
    const int32_t *ctg_seg;
    const int64_t *seg_contig_start;  // [S+1]
    const int64_t *tr_off;
    const int64_t *tr_woff;
    const uint64_t *tr_words;
    const uint32_t *tr_mask;
    const uint8_t *tr_chars;
    const int32_t *total;
    int64_t n_contigs;
    int32_t n_seg;
    int32_t splits;       // blocks per segment (each scans a slice of the truth)
    uint32_t *best;       // [C] leftmost verified position, POS_INF on entry (k_startpos_index writes, k_startpos reads)
    int32_t *startpos;    // optional: truth.find(contig) if a read was placed, else 0 (upstream semantics)
    int32_t *exact;       // optional: truth.find(contig) for EVERY contig (-1: not a substring); feeds the edit distance
    int32_t search_all;   // search also for contigs without a placed read (needed for `exact`)
};

constexpr int SP_SLOTS = 2048;  // seed table slots per pass
constexpr int SP_GROUP = 1024;  // contigs per pass

// a contig the seed index can look for: 32 valid leading bases, fits in the truth, had a read placed
__device__ __forceinline__ bool startpos_indexable(const StartposArgs &a, int64_t c, int64_t LT) {
    const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
    return L >= 32 && L <= LT && (a.search_all || a.total[c] != 0) && a.ctg_mask[a.ctg_woff[c]] == 0;
}

__global__ void __launch_bounds__(256) k_startpos_index(StartposArgs a) {
    __shared__ unsigned long long s_key[SP_SLOTS + 1];  // slot SP_SLOTS is reserved for the all-ones seed (= the empty marker)
    __shared__ int32_t s_head[SP_SLOTS + 1];  // local contig index + 1 of the first contig with that seed
    __shared__ int32_t s_next[SP_GROUP];  // next contig with the same seed
    const unsigned long long EMPTY = ~0ull;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    const int seg = blockIdx.x / a.splits, part = blockIdx.x % a.splits;
    const int64_t LT = a.tr_off[seg + 1] - a.tr_off[seg];
    const uint64_t *tw = a.tr_words + a.tr_woff[seg];
    const uint32_t *tm = a.tr_mask + a.tr_woff[seg];
    const uint8_t *tc = a.tr_chars + a.tr_off[seg];
    const int64_t nq = LT - 31;  // positions that can hold a 32-base seed
    if (nq <= 0) return;
    int64_t span = (nq + a.splits - 1) / a.splits;
    span = (span + 31) / 32 * 32;
    const int64_t q_begin = (int64_t)part * span;
    const int64_t q_end = q_begin + span < nq ? q_begin + span : nq;
    if (q_begin >= q_end) return;
    const int64_t c0 = a.seg_contig_start[seg], c1 = a.seg_contig_start[seg + 1];
    for (int64_t g0 = c0; g0 < c1; g0 += SP_GROUP) {
        const int gn = (int)(c1 - g0 < SP_GROUP ? c1 - g0 : SP_GROUP);
        __syncthreads();
        for (int i = tid; i <= SP_SLOTS; i += nthr) { s_head[i] = 0; s_key[i] = EMPTY; }
        __syncthreads();
        // parallel build: claim the seed's slot by compare-and-swap (linear probing), then push the
        // contig on the slot's chain
        for (int i = tid; i < gn; i += nthr) {
            const int64_t c = g0 + i;
            if (!startpos_indexable(a, c, LT)) continue;
            const unsigned long long key = a.ctg_words[a.ctg_woff[c]];
            uint32_t h = SP_SLOTS;
            if (key != EMPTY) {
                h = seed_hash(key) & (SP_SLOTS - 1);
                for (;;) {
                    const unsigned long long prev = atomicCAS(&s_key[h], EMPTY, key);
                    if (prev == EMPTY || prev == key) break;
                    h = (h + 1) & (SP_SLOTS - 1);
                }
            }
            s_next[i] = atomicExch(&s_head[h], i + 1);
        }
        __syncthreads();
        for (int64_t qb = q_begin; qb < q_end; qb += nthr) {
            const int64_t q = qb + tid;
            int cand = 0;
            if (q < q_end) {
                const int64_t idx = q >> 5;
                const uint32_t o = (uint32_t)(q & 31);
                if (window32(__ldg(&tm[idx]), __ldg(&tm[idx + 1]), o) == 0) {
                    const unsigned long long seed = window64(__ldg(&tw[idx]), __ldg(&tw[idx + 1]), o);
                    if (seed == EMPTY) cand = s_head[SP_SLOTS];
                    else {
                        uint32_t h = seed_hash(seed) & (SP_SLOTS - 1);
                        for (;;) {
                            const unsigned long long k = s_key[h];
                            if (k == EMPTY) break;
                            if (k == seed) { cand = s_head[h]; break; }
                            h = (h + 1) & (SP_SLOTS - 1);
                        }
                    }
                }
            }
            // seed hits of the warp, one after the other, each verified by all 32 lanes
            unsigned hits = __ballot_sync(FULL_MASK, cand != 0);
            while (hits) {
                const int src = __ffs((int)hits) - 1;
                hits &= hits - 1;
                const int64_t qc = __shfl_sync(FULL_MASK, q, src);
                int ci = __shfl_sync(FULL_MASK, cand, src);
                for (; ci != 0; ci = s_next[ci - 1]) {
                    const int64_t c = g0 + ci - 1;
                    const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
                    if (qc + L > LT) continue;
                    // a position further left may already be known; one lane reads so that the whole
                    // warp takes the same branch (the value can change under us)
                    uint32_t known = 0;
                    if (lane == 0) known = *(volatile uint32_t *)&a.best[c];
                    known = __shfl_sync(FULL_MASK, known, 0);
                    if ((uint32_t)qc >= known) continue;
                    const uint64_t *cw = a.ctg_words + a.ctg_woff[c];
                    const uint32_t *cm = a.ctg_mask + a.ctg_woff[c];
                    const int64_t idx = qc >> 5;
                    const uint32_t o = (uint32_t)(qc & 31);
                    const int64_t nw = (L + 31) >> 5;
                    bool ok = true, any_invalid = false;
                    for (int64_t j0 = 0; ok && j0 < nw; j0 += 32) {
                        const int64_t j = j0 + lane;
                        bool okl = true;
                        if (j < nw) {
                            const int rem = (L - 32 * j) < 32 ? (int)(L - 32 * j) : 32;
                            const uint32_t mj = cm[j];
                            // byte equality == equal 2-bit codes AND equal validity AND equal raw bytes where invalid
                            okl = ((window64(tw[idx + j], tw[idx + j + 1], o) ^ cw[j]) & keep_bases(rem)) == 0 &&
                                  ((window32(tm[idx + j], tm[idx + j + 1], o) ^ mj) & keep_bits(rem)) == 0;
                            any_invalid |= (mj & keep_bits(rem)) != 0;
                        }
                        ok = __ballot_sync(FULL_MASK, !okl) == 0;
                    }
                    if (ok && __ballot_sync(FULL_MASK, any_invalid) != 0) {
                        const uint8_t *cc = a.ctg_chars + a.ctg_off[c];
                        for (int64_t i0 = 0; ok && i0 < L; i0 += 32) {
                            const int64_t i = i0 + lane;
                            const bool okl = i < L ? tc[qc + i] == cc[i] : true;
                            ok = __ballot_sync(FULL_MASK, !okl) == 0;
                        }
                    }
                    if (ok && lane == 0) atomicMin(&a.best[c], (uint32_t)qc);
                }
            }
        }
    }
}

// final values, and the scan for the contigs the index cannot take
__global__ void k_startpos(StartposArgs a) {
    __shared__ int s_best;
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int64_t c = blockIdx.x; c < a.n_contigs; c += gridDim.x) {
        const int64_t L = a.ctg_off[c + 1] - a.ctg_off[c];
        const int64_t seg = a.ctg_seg[c];
        const int64_t LT = a.tr_off[seg + 1] - a.tr_off[seg];
        const bool placed = a.total[c] != 0;
        // upstream assigns truth.find(contig) only inside the read-hit branch, else the value stays 0
        auto publish = [&](int found) {
            if (a.startpos) a.startpos[c] = placed ? found : 0;
            if (a.exact) a.exact[c] = found;
        };
        if (!placed && !a.search_all) {
            if (tid == 0) publish(0);
            continue;
        }
        if (L == 0 || L > LT) {  // "" is found at 0; longer than the truth: npos
            if (tid == 0) publish(L == 0 ? 0 : -1);
            continue;
        }
        if (a.best && startpos_indexable(a, c, LT)) {
            if (tid == 0) { const uint32_t b = a.best[c]; publish(b == POS_INF ? -1 : (int32_t)b); }
            continue;
        }
        const uint64_t *cw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *cm = a.ctg_mask + a.ctg_woff[c];
        const uint8_t *cc = a.ctg_chars + a.ctg_off[c];
        const uint64_t *tw = a.tr_words + a.tr_woff[seg];
        const uint32_t *tm = a.tr_mask + a.tr_woff[seg];
        const uint8_t *tc = a.tr_chars + a.tr_off[seg];
        const int64_t nw = (L + 31) >> 5;
        const uint64_t c0 = cw[0];
        const uint32_t m0 = cm[0];
        const int first = L < 32 ? (int)L : 32;
        __syncthreads();
        if (tid == 0) s_best = 0x7fffffff;
        __syncthreads();
        const int64_t nq = LT - L + 1;
        for (int64_t qb = 0; qb < nq; qb += nthr) {
            const int64_t q = qb + tid;
            if (q < nq) {
                const int64_t idx = q >> 5;
                const uint32_t o = (uint32_t)(q & 31);
                // byte equality == equal 2-bit codes AND equal validity AND equal raw bytes where invalid
                bool ok = ((window64(tw[idx], tw[idx + 1], o) ^ c0) & keep_bases(first)) == 0 &&
                          ((window32(tm[idx], tm[idx + 1], o) ^ m0) & keep_bits(first)) == 0;
                bool any_invalid = (m0 & keep_bits(first)) != 0;
                for (int64_t j = 1; ok && j < nw; j++) {
                    const int rem = (L - 32 * j) < 32 ? (int)(L - 32 * j) : 32;
                    const uint32_t mj = cm[j];
                    ok = ((window64(tw[idx + j], tw[idx + j + 1], o) ^ cw[j]) & keep_bases(rem)) == 0 &&
                         ((window32(tm[idx + j], tm[idx + j + 1], o) ^ mj) & keep_bits(rem)) == 0;
                    any_invalid |= (mj & keep_bits(rem)) != 0;
                }
                if (ok && any_invalid) {
                    for (int64_t i = 0; ok && i < L; i++) ok = tc[q + i] == cc[i];
                }
                if (ok) atomicMin(&s_best, (int)q);
            }
            if (__syncthreads_or(s_best != 0x7fffffff)) break;  // positions are visited in ascending blocks
        }
        __syncthreads();
        if (tid == 0) publish(s_best == 0x7fffffff ? -1 : s_best);
    }
}

// ------------------------------------------------------------------------------------------
// infix edit distance of every contig against its truth (upstream calc_levenshtein,
// lib/BreakageScorer.cpp:41-55: edlib, EDLIB_MODE_HW, k = -1, distance only): the smallest
// Levenshtein distance between the whole contig and any substring of the truth.
//
// Myers' bit-vector recurrence in Hyyro's block form, one WARP per contig: lane b owns pattern
// block b (64 contig bases: Pv/Mv and one match mask per letter in registers) and works on truth
// column t - b at step t, so the horizontal delta of block b-1 and the truth letter arrive by
// shuffle from the lane above (a systolic array).  Contigs longer than 32 blocks are processed in
// chunks of 32 blocks; the last lane's deltas of a chunk go through a per-warp byte buffer.
// Exact substrings (known from k_startpos) have distance 0 and are skipped; everything else has
// distance >= 1, so a scan stops as soon as it sees 1.
// ------------------------------------------------------------------------------------------

struct LevArgs {
    const int32_t *order;   // [C] contig ids, longest first
    int32_t *work_counter;  // zero on entry
    const int64_t *ctg_off;
    const int64_t *ctg_woff;
    const uint64_t *ctg_words;
    const uint32_t *ctg_mask;
    const uint8_t *ctg_chars;
    const int32_t *ctg_seg;
    const int64_t *tr_off;
    const int64_t *tr_woff;
    const uint64_t *tr_words;
    const uint32_t *tr_mask;
    const uint8_t *tr_chars;
    const int32_t *exact;   // [C] position of the contig in its truth or -1
    int64_t n_contigs;
    int8_t *hbuf;           // [warps of the grid][hbuf_stride] chunk-boundary deltas (NULL when no contig exceeds 32 blocks)
    int64_t hbuf_stride;
    int32_t *lev;           // [C]
};

// match mask of one 64-base pattern block for a truth letter: tc < 4 is a packed ACGT code, else
// 4 + the raw byte of a non-ACGT truth letter (byte equality, as edlib compares symbols)
__device__ __forceinline__ uint64_t lev_eq(const uint64_t (&peq)[4], uint64_t inv, const uint8_t *blk_chars, int blk_len, int tc) {
    if (tc < 4) {  // selects, not an indexed load: the masks stay in registers
        const uint64_t lo = (tc & 1) ? peq[1] : peq[0], hi = (tc & 1) ? peq[3] : peq[2];
        return (tc & 2) ? hi : lo;
    }
    uint64_t m = 0;  // rare: compare the raw bytes of the block's non-ACGT positions
    for (int i = 0; i < blk_len; i++)
        if (((inv >> i) & 1) && blk_chars[i] == (uint8_t)(tc - 4)) m |= 1ull << i;
    return m;
}

// Cheap exact answers before the full scan.  lev[c] = 0 for exact substrings and empty contigs,
// the contig length against an empty truth, and 1 when the contig lies on some diagonal of the truth
// with exactly one substituted base (found through its first or last 32 bases as seeds and an
// XOR/popcount of the packed words): a contig that is not an exact substring cannot do better than
// 1.  Everything else is left at -1 for k_lev_infix.
__global__ void __launch_bounds__(128) k_lev_bound(LevArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < a.n_contigs; c += warps) {
        const int64_t m = a.ctg_off[c + 1] - a.ctg_off[c];
        const int64_t seg = a.ctg_seg[c];
        const int64_t n = a.tr_off[seg + 1] - a.tr_off[seg];
        int32_t res = -1;
        if (a.exact[c] >= 0 || m == 0) res = 0;
        else if (n == 0) res = (int32_t)m;
        if (res >= 0 || m < 32 || m > n) {
            if (lane == 0) a.lev[c] = res;
            continue;
        }
        const uint64_t *cw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *cm = a.ctg_mask + a.ctg_woff[c];
        const uint64_t *tw = a.tr_words + a.tr_woff[seg];
        const uint32_t *tm = a.tr_mask + a.tr_woff[seg];
        const int64_t toff = m - 32;  // the last 32 bases of the contig
        const bool head_ok = cm[0] == 0;
        const bool tail_ok = window32(cm[toff >> 5], cm[(toff >> 5) + 1], (uint32_t)(toff & 31)) == 0;
        const uint64_t head = cw[0];
        const uint64_t tail = window64(cw[toff >> 5], cw[(toff >> 5) + 1], (uint32_t)(toff & 31));
        const int64_t nw = (m + 31) >> 5;
        bool found = false;
        for (int64_t qb = 0; qb < n - 31 && !found; qb += 32) {
            const int64_t q = qb + lane;
            int64_t start = -1;  // contig start on the truth suggested by a seed hit at q
            if (q < n - 31 && window32(tm[q >> 5], tm[(q >> 5) + 1], (uint32_t)(q & 31)) == 0) {
                const uint64_t sq = window64(tw[q >> 5], tw[(q >> 5) + 1], (uint32_t)(q & 31));
                if (head_ok && sq == head && q + m <= n) start = q;
                else if (tail_ok && sq == tail && q - toff >= 0) start = q - toff;
            }
            unsigned hits = __ballot_sync(FULL_MASK, start >= 0);
            while (hits && !found) {
                const int src = __ffs((int)hits) - 1;
                hits &= hits - 1;
                const int64_t st = __shfl_sync(FULL_MASK, start, src);
                const int64_t idx = st >> 5;
                const uint32_t o = (uint32_t)(st & 31);
                int mism = 0;
                for (int64_t j0 = 0; j0 < nw && mism <= 1; j0 += 32) {
                    const int64_t j = j0 + lane;
                    int mine = 0;
                    if (j < nw) {
                        const int rem = (m - 32 * j) < 32 ? (int)(m - 32 * j) : 32;
                        const uint64_t d = (window64(tw[idx + j], tw[idx + j + 1], o) ^ cw[j]) & keep_bases(rem);
                        mine = __popcll((d | (d >> 1)) & 0x5555555555555555ull);
                        // a non-ACGT byte on either side: this diagonal is left to the full scan
                        if ((window32(tm[idx + j], tm[idx + j + 1], o) | cm[j]) & keep_bits(rem)) mine = 2;
                    }
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) mine += __shfl_xor_sync(FULL_MASK, mine, sft);
                    mism += mine;
                }
                if (mism == 1) found = true;  // (0 cannot happen: the contig is not an exact substring)
            }
        }
        if (lane == 0) a.lev[c] = found ? 1 : -1;
    }
}

__global__ void __launch_bounds__(128) k_lev_infix(LevArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t gwarp = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    int8_t *hbuf = a.hbuf ? a.hbuf + gwarp * a.hbuf_stride : nullptr;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(a.work_counter, 1);
        item = __shfl_sync(FULL_MASK, item, 0);
        if (item >= a.n_contigs) break;
        const int64_t c = a.order[item];
        const int64_t m = a.ctg_off[c + 1] - a.ctg_off[c];
        const int64_t seg = a.ctg_seg[c];
        const int64_t n = a.tr_off[seg + 1] - a.tr_off[seg];
        if (a.lev[c] >= 0) continue;  // settled by k_lev_bound
        const uint64_t *cw = a.ctg_words + a.ctg_woff[c];
        const uint32_t *cm = a.ctg_mask + a.ctg_woff[c];
        const uint8_t *cc = a.ctg_chars + a.ctg_off[c];
        const uint64_t *tw = a.tr_words + a.tr_woff[seg];
        const uint32_t *tm = a.tr_mask + a.tr_woff[seg];
        const uint8_t *tc = a.tr_chars + a.tr_off[seg];
        const int64_t n_blocks = (m + 63) >> 6;
        const int64_t n_chunks = (n_blocks + 31) >> 5;
        int32_t best = (int32_t)m;  // column 0: the contig against the empty substring
        bool done = false;
        for (int64_t g = 0; g < n_chunks && !done; g++) {
            const int64_t blk = g * 32 + lane;
            const bool owns = blk < n_blocks;
            const int nb = (int)(n_blocks - g * 32 < 32 ? n_blocks - g * 32 : 32);  // active lanes of this chunk
            const bool last_chunk = g == n_chunks - 1;
            const bool scorer = last_chunk && lane == nb - 1;  // owns the contig's last row
            // match masks of my 64 pattern bases: row i of the block at bit i
            uint64_t peq[4] = {0, 0, 0, 0}, inv = 0;
            int blk_len = 0;
            if (owns) {
                blk_len = (int)(m - blk * 64 < 64 ? m - blk * 64 : 64);
                for (int h = 0; h < 2; h++) {
                    const uint64_t w = cw[2 * blk + h];
                    const uint32_t mk = cm[2 * blk + h];
                    for (int i = 0; i < 32; i++) {
                        const int row = 32 * h + i;
                        if (row >= blk_len) break;
                        const uint64_t bit = 1ull << row;
                        const int code = (int)((w >> (62 - 2 * i)) & 3);
                        if ((mk >> (31 - i)) & 1) inv |= bit;
                        else {
                            peq[0] |= code == 0 ? bit : 0ull;
                            peq[1] |= code == 1 ? bit : 0ull;
                            peq[2] |= code == 2 ? bit : 0ull;
                            peq[3] |= code == 3 ? bit : 0ull;
                        }
                    }
                }
            }
            const int out_bit = scorer ? (int)((m - 1) & 63) : 63;
            uint64_t Pv = ~0ull, Mv = 0;
            int32_t score = (int32_t)m;  // meaningful on the scorer lane
            int hout_prev = 0, tc_prev = 0;
            const int64_t steps = n + nb - 1;
            for (int64_t t = 0; t < steps; t++) {
                // the truth letter and the horizontal delta travel down the lanes, one lane per step
                int tcin = __shfl_up_sync(FULL_MASK, tc_prev, 1);
                int hin = __shfl_up_sync(FULL_MASK, hout_prev, 1);
                if (lane == 0) {
                    tcin = 0;
                    hin = 0;
                    if (t < n) {
                        const bool bad = (tm[t >> 5] >> (31 - (t & 31))) & 1;
                        tcin = bad ? 4 + (int)tc[t] : (int)((tw[t >> 5] >> (62 - 2 * (t & 31))) & 3);
                        if (g > 0) hin = hbuf[t];
                    }
                }
                const int64_t col = t - lane;
                int hout = 0;
                if (owns && col >= 0 && col < n) {
                    uint64_t Eq = lev_eq(peq, inv, cc + blk * 64, blk_len, tcin);
                    const uint64_t Xv = Eq | Mv;
                    if (hin < 0) Eq |= 1ull;
                    const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                    uint64_t Ph = Mv | ~(Xh | Pv);
                    uint64_t Mh = Pv & Xh;
                    hout = (int)((Ph >> out_bit) & 1) - (int)((Mh >> out_bit) & 1);
                    Ph <<= 1;
                    Mh <<= 1;
                    if (hin < 0) Mh |= 1ull;
                    else if (hin > 0) Ph |= 1ull;
                    Pv = Mh | ~(Xv | Ph);
                    Mv = Ph & Xv;
                    if (scorer) {
                        score += hout;
                        if (score < best) best = score;
                    } else if (!last_chunk && lane == 31) {
                        hbuf[col] = (int8_t)hout;  // read back as lane 0's input in the next chunk
                    }
                }
                hout_prev = hout;
                tc_prev = tcin;
                // the contig is not an exact substring, so 1 is the floor: stop once it is reached
                if ((t & 63) == 63 && last_chunk) {
                    const int b = __shfl_sync(FULL_MASK, best, nb - 1);
                    if (b <= 1) break;
                }
            }
            if (last_chunk) best = __shfl_sync(FULL_MASK, best, nb - 1);
            __syncwarp();
        }
        if (lane == 0) a.lev[c] = best;
    }
}

}  // namespace bs
