// Descriptors of packed sequences and reads, 2-bit packing kernels and the per-segment read index.
#pragma once
#include "bs_device.cuh"

namespace bs {

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

// one thread per packed word of a SeqSet: the word's 32 bytes come in as nine aligned 32-bit loads realigned with
// byte funnel shifts, four bytes are converted per step inside a register (pack4 below); the pad words and the
// positions past the end of a string pack as code 0 with their mask bits set
__device__ __forceinline__ uint32_t pack4(uint32_t x, uint32_t &diff);
__global__ void k_pack_seqs(SeqSet s) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total_chars = s.n > 0 ? s.off[s.n] : 0;
    __shared__ int64_t s_first;  // string owning the block's first word of the current round
    for (int64_t wb = (int64_t)blockIdx.x * blockDim.x; wb < s.total_words; wb += stride) {
        // string owning word wi: largest i with woff[i] <= wi.  One binary search per block and round (its first word),
        // then every thread walks forward from there: a block's 256 words span a handful of strings
        __syncthreads();
        if (threadIdx.x == 0) {
            int64_t lo = 0, hi = s.n - 1;
            while (lo < hi) {
                int64_t mid = (lo + hi + 1) >> 1;
                if (s.woff[mid] <= wb) lo = mid; else hi = mid - 1;
            }
            s_first = lo;
        }
        __syncthreads();
        const int64_t wi = wb + threadIdx.x;
        if (wi >= s.total_words) continue;
        int64_t lo = s_first;
        while (lo + 1 < s.n && s.woff[lo + 1] <= wi) lo++;
        const int64_t c0 = s.off[lo];
        const int64_t L = s.off[lo + 1] - c0;
        const int64_t b0 = (wi - s.woff[lo]) * 32;
        const int nvalid = L - b0 >= 32 ? 32 : (L - b0 > 0 ? (int)(L - b0) : 0);  // bytes of the string in this word
        uint64_t w = 0;
        uint32_t m = ~0u;
        if (nvalid > 0) {
            const uint8_t *p = s.chars + c0 + b0;
            const uint32_t sh = (uint32_t)((uintptr_t)p & 3);
            const uint8_t *q = p - sh;                       // 4-byte aligned
            const int64_t q_off = (c0 + b0) - (int64_t)sh;   // its offset in the char buffer (may be < 0 for the first word)
            uint32_t a[9];
#pragma unroll
            for (int k = 0; k < 9; k++) {
                const int64_t o = q_off + 4 * k;
                a[k] = 0x41414141u;
                if (4 * k - (int)sh < nvalid && 4 * k + 4 - (int)sh > 0) {  // the aligned word overlaps the string's bytes
                    if (o >= 0 && o + 4 <= total_chars) {
                        a[k] = *reinterpret_cast<const uint32_t *>(q + 4 * k);
                    } else {  // sticks out of the buffer: the bytes that exist, one by one
                        uint32_t x = 0;
                        for (int b = 3; b >= 0; b--) x = (x << 8) | ((o + b >= 0 && o + b < total_chars) ? (uint32_t)q[4 * k + b] : 0x41u);
                        a[k] = x;
                    }
                }
            }
            uint32_t bad = 0;  // bit i: byte i of the word is outside ACGT
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t x = __funnelshift_r(a[k], a[k + 1], 8 * sh);  // bytes 4k .. 4k+3 of the word
                uint32_t diff = 0;
                const uint32_t code = pack4(x, diff) >> 24;  // c0 c1 c2 c3, first base in the most significant bits
                w = (w << 8) | code;
                // diff holds a non-zero BYTE per invalid base: one bit per base, first base highest
                const uint32_t nz = (diff | (diff >> 4)) & 0x0f0f0f0fu;
                const uint32_t nz2 = (nz | (nz >> 2)) & 0x03030303u;
                const uint32_t nz1 = (nz2 | (nz2 >> 1)) & 0x01010101u;  // bit 0 of byte i = base 4k+i invalid
                const uint32_t four = ((nz1 & 1u) << 3) | (((nz1 >> 8) & 1u) << 2) | (((nz1 >> 16) & 1u) << 1) | ((nz1 >> 24) & 1u);
                bad = (bad << 4) | four;
            }
            m = bad | ~keep_bits(nvalid);
            // code 0 wherever the mask is set: every mask bit spread to its base's two bits
            uint64_t mm = m;
            mm = (mm | (mm << 16)) & 0x0000ffff0000ffffull;
            mm = (mm | (mm << 8)) & 0x00ff00ff00ff00ffull;
            mm = (mm | (mm << 4)) & 0x0f0f0f0f0f0f0f0full;
            mm = (mm | (mm << 2)) & 0x3333333333333333ull;
            mm = (mm | (mm << 1)) & 0x5555555555555555ull;
            w &= ~(mm | (mm << 1));
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
__global__ void __launch_bounds__(PACK_THREADS, 5) k_pack_reads_uniform(ReadSet r, ReadIndex ix) {
    __shared__ uint32_t s_code_all[PACK_THREADS / 32][PACK_WARP_CELLS];
    __shared__ uint32_t s_bad_all[PACK_THREADS / 32][PACK_WARP_CELLS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ uint32_t s_seg_all[PACK_THREADS / 32][4];  // the warp's current segment: seed length, table mask, table offset
    uint32_t *s_code = s_code_all[warp], *s_bad = s_bad_all[warp], *s_seg = s_seg_all[warp];
    const int64_t n_words = r.n * r.W;
    const int64_t total_bytes = r.n * (int64_t)r.uniform_len;
    const uint32_t base15 = (uint32_t)((uintptr_t)r.chars & 15);
    const int L = r.uniform_len;
    const uint32_t W = (uint32_t)r.W;
    // x / W for the small x met below (x * W < 2^32) by one multiply
    const uint32_t w_magic = W > 1 ? (uint32_t)((0x100000000ull + W - 1) / W) : 0u;
    auto div_w = [&](uint32_t x) { return W > 1 ? __umulhi(x, w_magic) : x; };
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
    uint64_t *wp = r.words + k0;             // the tile's first output word
    uint32_t left = (uint32_t)(k_end - k0);  // words this warp still has to write
    // (read, word-in-read, first source byte) of the tile's first word: one 64-bit division here,
    // then increments
    uint32_t n0 = (uint32_t)(k0 / W);  // read ids fit 32 bits (the host refuses more reads per call)
    uint32_t j0 = (uint32_t)(k0 - (int64_t)n0 * W);
    int64_t lo0 = (int64_t)n0 * L + (int64_t)j0 * 32;
    const uint32_t adv = 32u % W, adv_n = 32u / W;
    const int32_t adv_bytes = (int32_t)adv_n * L + (int32_t)adv * 32, wrap_bytes = L - (int32_t)W * 32;

    // the segment of the tile's first read, its end and its index geometry, kept in registers
    int seg = 0;
    uint32_t seg_next = 0xffffffffu;  // first read of the next segment
    auto load_segment = [&]() {
        seg_next = seg + 1 < ix.n_seg ? (uint32_t)ix.seg_read_start[seg + 1] : 0xffffffffu;
        __syncwarp();
        if (lane == 0) {
            const int64_t off = ix.tab_off[seg];
            s_seg[0] = (uint32_t)ix.seed_len[seg];
            s_seg[1] = (uint32_t)ix.tab_mask[seg];
            s_seg[2] = (uint32_t)off;
            s_seg[3] = (uint32_t)(off >> 32);
        }
        __syncwarp();
    };
    if (ix.head) {
        seg = segment_of_read(ix, n0);
        load_segment();
    }

    // geometry of a tile and the (up to three) 16-byte cells this lane stages for it
    struct Tile { int64_t lo16; int n_cells; int ntile; int mis; bool inside; };
    auto tile_of = [&](uint32_t words_left, int64_t lo, uint32_t j) {
        Tile t;
        t.ntile = (int)(words_left < 32u ? words_left : 32u);
        const uint32_t xl = j + (uint32_t)t.ntile - 1;
        const uint32_t dnl = div_w(xl);
        // bytes from the tile's first source byte to one past its last
        const int32_t span = (int32_t)dnl * L + ((int32_t)(xl - dnl * W) - (int32_t)j) * 32 + 32;
        int64_t hi = lo + span;
        if (hi > total_bytes) hi = total_bytes;
        t.mis = (int)((base15 + ((uint32_t)lo & 15u)) & 15u);
        t.lo16 = lo - t.mis;  // 16-byte aligned address, may be < 0
        t.n_cells = (int)((hi - t.lo16 + 15) >> 4);
        t.inside = t.lo16 >= 0 && t.lo16 + 16 * (int64_t)t.n_cells <= total_bytes;
        return t;
    };
    auto load_cell = [&](const Tile &t, int ci) {
        uint4 v = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
        if (ci < t.n_cells) {
            const int64_t cb = t.lo16 + 16 * (int64_t)ci;
            if (t.inside || (cb >= 0 && cb + 16 <= total_bytes)) {
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

    Tile t = tile_of(left, lo0, j0);
    uint4 c0 = load_cell(t, lane), c1 = load_cell(t, lane + 32), c2 = c0;
    if (t.n_cells > 64) c2 = load_cell(t, lane + 64);
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
            if (t.n_cells > 64 && lane + 64 < PACK_WARP_CELLS) {  // a tile rarely reaches into a 65th cell
                diff = 0;
                s_code[lane + 64] = pack16(c2.x, c2.y, c2.z, c2.w, diff);
                s_bad[lane + 64] = diff;
            }
        }
        __syncwarp();
        // ---- issue the next tile's loads before cutting this tile's words ----
        const Tile cur = t;
        uint64_t *const wp_cur = wp;
        const uint32_t n_cur = n0, j_cur = j0;
        wp += 32;
        left = left > 32u ? left - 32u : 0u;
        n0 += adv_n;
        j0 += adv;
        lo0 += adv_bytes;
        if (j0 >= W) { j0 -= W; n0++; lo0 += wrap_bytes; }
        const bool more = left != 0u;
        if (more) {
            t = tile_of(left, lo0, j0);
            c0 = load_cell(t, lane);
            c1 = load_cell(t, lane + 32);
            if (t.n_cells > 64) c2 = load_cell(t, lane + 64);
        }
        if (lane < cur.ntile) {
            const uint32_t x = j_cur + (uint32_t)lane;
            const uint32_t dn = div_w(x);
            const uint32_t n = n_cur + dn;
            const int j = (int)(x - dn * W);
            const int a = cur.mis + (int)dn * L + (j - (int)j_cur) * 32;  // byte offset inside the staged span (< 1100)
            const int ci = a >> 4;
            const uint32_t sh = 2u * (uint32_t)(a & 15);
            const int rem = L - 32 * j;  // bases of this word that belong to the read (may exceed 32)
            // cells past the staged span are only touched by bases beyond the read: clamp the index
            const int i1 = ci + 1 < cur.n_cells ? ci + 1 : cur.n_cells - 1, i2 = ci + 2 < cur.n_cells ? ci + 2 : cur.n_cells - 1;
            const uint32_t w0 = s_code[ci], w1 = s_code[i1], w2 = s_code[i2];
            const uint32_t o_hi = sh ? ((w0 << sh) | (w1 >> (32u - sh))) : w0;
            const uint32_t o_lo = sh ? ((w1 << sh) | (w2 >> (32u - sh))) : w1;
            const uint64_t word = (((uint64_t)o_hi << 32) | o_lo) & keep_bases(rem);
            wp_cur[lane] = word;
            // cells overlapping this word's own bytes [a, a + min(rem, 32))
            const int last_cell = (a + (rem < 32 ? rem : 32) - 1) >> 4;
            uint32_t bad = s_bad[ci];
            if (last_cell >= ci + 1) bad |= s_bad[i1];
            if (last_cell >= ci + 2) bad |= s_bad[i2];
            if (bad) atomicOr(reinterpret_cast<unsigned *>(r.flags) + (n >> 2), 1u << (8 * (int)(n & 3)));
            if (j == 0 && ix.head) {
                if (n < seg_next && !bad) {  // the common case: same segment as the tile's first read
                    const uint64_t seed = word & keep_bases((int)s_seg[0]);
                    const uint32_t h = seed_hash(seed) & s_seg[1];
                    const int64_t slot = (int64_t)(((uint64_t)s_seg[3] << 32) | s_seg[2]) + h;
                    ix.next[n] = make_uint2(atomicExch(&ix.head[slot], n + 1u), seed_tag(seed));
                } else {
                    int sg = seg;
                    while (sg + 1 < ix.n_seg && (int64_t)n >= ix.seg_read_start[sg + 1]) sg++;
                    index_insert(ix, sg, n, word, L, bad != 0);
                }
            }
        }
        if (!more) break;
        if (ix.head && n0 >= seg_next) {
            while (seg + 1 < ix.n_seg && (int64_t)n0 >= ix.seg_read_start[seg + 1]) seg++;
            load_segment();
        }
    }
}

// ---- the same packing with the ASCII staged by bulk asynchronous copies (sm_90+/sm_100 TMA unit) ------------
// A block owns a contiguous run of TILES of R whole reads (R * L is a multiple of 16, so every tile starts at
// the same 16-byte phase `mis` of the buffer).  One thread arms an mbarrier with the tile's byte count and issues
// ONE cp.async.bulk (global -> shared, SASS UBLKCP) per tile; two stages, so the copy of tile t+1 is in flight
// while tile t is converted: no load instruction, address arithmetic or register of the SIMD path is spent on the
// input stream.  Phase A turns 16-byte cells into 32 bits of 2-bit codes + a validity word in shared memory
// (LDS.128 -> pack16), phase B cuts every output word out of two or three cells with funnel shifts, writes it
// coalesced, ORs the per-read flag in shared memory (written once per read, no global atomic, no memset) and
// inserts the read into its segment's index.  The index insertion's atomicExch result is consumed one tile
// later (the old kernel stalled on it: 54 % of its stall samples, profiles/r02a).  Code / flag buffers are
// double buffered as well: one block barrier per tile.
// Cells that are not completely inside the read buffer (the first cell of the buffer when it is not 16-byte
// aligned, the last one when the buffer does not end on a cell boundary) are never touched by the bulk copy:
// phase A reads their existing bytes one by one.

#ifndef BS_CPU_EMUL
__device__ __forceinline__ uint32_t bs_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bs_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bs_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bs_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(bs_smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(bs_smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "BS_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra BS_MBAR_DONE;\n"
        "bra BS_MBAR_WAIT;\n"
        "BS_MBAR_DONE:\n"
        "}\n" ::"r"(bs_smem_addr(bar)),
        "r"(parity)
        : "memory");
}
#else  // CPU emulation (tests): the issuing thread copies at once; the block barrier of the tile loop orders it before its readers
inline void mbar_init(uint64_t *, uint32_t) {}
inline void mbar_init_fence() {}
inline void mbar_arrive_expect_tx(uint64_t *, uint32_t) {}
inline void mbar_arrive(uint64_t *) {}
inline void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *) { std::memcpy(dst, src, bytes); }
inline void mbar_wait(uint64_t *, uint32_t) {}
#endif

constexpr int PACKB_THREADS = 256;
constexpr int PACKB_STAGES = 3;              // bulk copies in flight per block: tiles t+1 and t+2 while tile t is converted
#ifndef BS_PACKB_TILE_KB
#define BS_PACKB_TILE_KB 16
#endif
#ifndef BS_PACKB_BLOCKS
#define BS_PACKB_BLOCKS 3
#endif
constexpr int PACKB_TILE_BYTES = BS_PACKB_TILE_KB * 1024;  // target ASCII bytes per tile
constexpr int PACKB_MAX_STAGE = 20 * 1024;   // longest reads the staged kernel takes: 16 of them must fit a stage

// reads per tile for reads of L bytes: a multiple of 4 whose byte count is a multiple of 16, at most one per thread
BS_HD int packb_tile_reads(int L) {
    int g = L & 15;  // gcd(L, 16)
    g = g == 0 ? 16 : (g & -g);
    int q = 16 / g;
    if (q < 4) q = 4;
    int R = (PACKB_TILE_BYTES / L) / q * q;
    if (R > PACKB_THREADS / 2) R = PACKB_THREADS / 2 / q * q;  // a cutting thread per read; at least as many converting threads
    return R < q ? q : R;
}
BS_HD int packb_stage_bytes(int L, int R) { return ((R * L + 15 + 16) / 16 * 16 + 127) / 128 * 128; }
BS_HD size_t packb_smem_bytes(int L, int R) {
    const size_t cells = (size_t)packb_stage_bytes(L, R) / 16;
    return PACKB_STAGES * (size_t)packb_stage_bytes(L, R) + 2 * cells * 8 + 64;
}

// all NW words of one read out of its run of cells (cell = 2-bit codes of 16 bases + validity word): the cells are
// loaded first, all in flight together, then cut with funnel shifts.  `bad` = any byte outside ACGT in a cell the read
// overlaps (conservative: flagged reads are verified on the text, still exact); seed_bad = the same for the first word.
template <int NW>
__device__ __forceinline__ void packb_cut(const uint2 *cell, int ci, uint32_t sh, int last_cell, int L, uint64_t *wout, uint32_t &bad,
                                          uint32_t &seed_bad, uint64_t &word0) {
    uint2 x[2 * NW + 1];
#pragma unroll
    for (int k = 0; k <= 2 * NW; k++) x[k] = cell[ci + k < last_cell ? ci + k : last_cell];  // (cells past the read's last one are only touched by bases beyond the read)
    bad = 0;
#pragma unroll
    for (int k = 0; k <= 2 * NW; k++) bad |= x[k].y;
    seed_bad = x[0].y | x[1].y | x[2].y;
#pragma unroll
    for (int j = 0; j < NW; j++) {
        const uint32_t o_hi = __funnelshift_l(x[2 * j + 1].x, x[2 * j].x, sh), o_lo = __funnelshift_l(x[2 * j + 2].x, x[2 * j + 1].x, sh);
        const uint64_t word = (((uint64_t)o_hi << 32) | o_lo) & keep_bases(L - 32 * j);
        wout[j] = word;
        if (j == 0) word0 = word;
    }
}
__device__ __forceinline__ void packb_cut_any(const uint2 *cell, int ci, uint32_t sh, int last_cell, int L, int W, uint64_t *wout, uint32_t &bad,
                                              uint32_t &seed_bad, uint64_t &word0) {
    uint2 x0 = cell[ci];
    bad = x0.y;
    for (int j = 0; j < W; j++) {
        const int i1 = ci + 2 * j + 1 < last_cell ? ci + 2 * j + 1 : last_cell, i2 = ci + 2 * j + 2 < last_cell ? ci + 2 * j + 2 : last_cell;
        const uint2 x1 = cell[i1], x2 = cell[i2];
        const uint32_t o_hi = __funnelshift_l(x1.x, x0.x, sh), o_lo = __funnelshift_l(x2.x, x1.x, sh);
        const uint64_t word = (((uint64_t)o_hi << 32) | o_lo) & keep_bases(L - 32 * j);
        wout[j] = word;
        bad |= x1.y | x2.y;
        if (j == 0) { word0 = word; seed_bad = bad; }
        x0 = x2;
    }
}

__global__ void __launch_bounds__(PACKB_THREADS, BS_PACKB_BLOCKS) k_pack_reads_bulk(ReadSet r, ReadIndex ix, int R) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int L = r.uniform_len;
    const int W = r.W;
    const int SB = packb_stage_bytes(L, R);
    const int cells_max = SB / 16;
    unsigned char *sm = bs_dyn_smem();
    unsigned char *s_stage = sm;                          sm += PACKB_STAGES * (size_t)SB;
    uint2 *s_cell = (uint2 *)sm;                          sm += 2 * (size_t)cells_max * 8;  // (2-bit codes of 16 bases, validity) per cell, two tiles
    uint64_t *s_bar = (uint64_t *)(((uintptr_t)sm + 7) & ~(uintptr_t)7);  // one mbarrier per stage
    __shared__ int s_tileseg[2];  // segment of the first read of the tile being cut

    const int64_t N = r.n;
    const int64_t total_bytes = N * (int64_t)L;
    const int mis = (int)((uintptr_t)r.chars & 15);
    const int64_t tiles_total = (N + R - 1) / R;
    const int64_t per_block = (tiles_total + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per_block;
    const int64_t t1 = t0 + per_block < tiles_total ? t0 + per_block : tiles_total;
    if (t0 >= t1) return;

    // cells [c_lo, c_hi) of tile t are completely inside the buffer: the bulk copy's share
    auto tile_geom = [&](int64_t t, int &nr, int &ncell, int &c_lo, int &c_hi) {
        const int64_t n0 = t * R;
        nr = (int)(N - n0 < R ? N - n0 : R);
        ncell = (mis + nr * L + 15) >> 4;
        const int64_t g0 = n0 * L - mis;  // buffer offset of the tile's cell 0 (may be < 0 for tile 0)
        c_lo = g0 < 0 ? 1 : 0;
        c_hi = ncell;
        if (g0 + 16 * (int64_t)ncell > total_bytes) c_hi = ncell - 1;
        if (c_hi < c_lo) c_hi = c_lo;
    };
    auto issue = [&](int64_t t) {  // thread 0 only
        int nr, ncell, c_lo, c_hi;
        tile_geom(t, nr, ncell, c_lo, c_hi);
        const int st = (int)((t - t0) % PACKB_STAGES);
        const uint32_t bytes = 16u * (uint32_t)(c_hi - c_lo);
        if (bytes) {
            mbar_arrive_expect_tx(&s_bar[st], bytes);
            bulk_copy_g2s(s_stage + (size_t)st * SB + 16 * c_lo, r.chars + (t * R * (int64_t)L - mis + 16 * c_lo), bytes, &s_bar[st]);
        } else {
            mbar_arrive(&s_bar[st]);
        }
    };

    // warp roles: the last ceil(R / 32) warps CUT (a thread per read of the tile converted one iteration earlier), the
    // others CONVERT the current tile's cells: the two phases of consecutive tiles overlap, one block barrier per tile
    const int n_cut_threads = (R + 31) / 32 * 32;
    const int n_conv = nthr - n_cut_threads;  // (the host launches with enough threads: n_conv >= 32)
    const bool converter = tid < n_conv;
    int seg = 0;  // thread 0: segment of the current tile's first read, advanced tile by tile
    if (tid == 0) {
        for (int i = 0; i < PACKB_STAGES; i++) mbar_init(&s_bar[i], 1);
        mbar_init_fence();
        for (int i = 0; i < PACKB_STAGES && t0 + i < t1; i++) issue(t0 + i);
        if (ix.head) seg = segment_of_read(ix, t0 * R);
    }
    __syncthreads();

    // index insertion in flight: the bucket's previous head comes back from the atomic one tile later
    uint32_t pend_n = 0xffffffffu, pend_old = 0, pend_tag = 0;
    for (int64_t t = t0; t <= t1; t++) {
        if (converter && t < t1) {
            const int st = (int)((t - t0) % PACKB_STAGES);
            const uint32_t parity = (uint32_t)(((t - t0) / PACKB_STAGES) & 1);
            int nr, ncell, c_lo, c_hi;
            tile_geom(t, nr, ncell, c_lo, c_hi);
            const int64_t n0 = t * R;
            const unsigned char *stage = s_stage + (size_t)st * SB;
            uint2 *cell = s_cell + (size_t)((t - t0) & 1) * cells_max;
            mbar_wait(&s_bar[st], parity);
            // ---- 16-byte cells -> 2-bit codes + validity ----
            for (int ci = tid; ci < ncell; ci += n_conv) {
                uint4 v;
                if (ci >= c_lo && ci < c_hi) {
                    v = *reinterpret_cast<const uint4 *>(stage + 16 * ci);
                } else {  // a cell that sticks out of the buffer: its existing bytes one by one, 'A' for the rest
                    uint32_t xs[4];
                    const int64_t cb = n0 * L - mis + 16 * (int64_t)ci;
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
                uint32_t diff = 0;
                const uint32_t code = pack16(v.x, v.y, v.z, v.w, diff);
                cell[ci] = make_uint2(code, diff);
            }
            if (tid == 0 && ix.head) {
                while (seg + 1 < ix.n_seg && n0 >= ix.seg_read_start[seg + 1]) seg++;
                s_tileseg[(t - t0) & 1] = seg;
            }
        } else if (!converter && t > t0) {
            // ---- the previous tile: a thread cuts all W words of ONE read out of its run of cells ----
            const int64_t tp = t - 1;
            const int64_t n0 = tp * R;
            const int nr = (int)(N - n0 < R ? N - n0 : R);
            const uint2 *cell = s_cell + (size_t)((tp - t0) & 1) * cells_max;
            const int nl = tid - n_conv;
            if (nl < nr) {
                const uint32_t n = (uint32_t)n0 + (uint32_t)nl;
                const int a0 = mis + nl * L;  // byte offset of the read inside the staged span
                const int ci = a0 >> 4;
                const uint32_t sh = 2u * (uint32_t)(a0 & 15);
                const int last_cell = (a0 + L - 1) >> 4;
                uint64_t *wout = r.words + (int64_t)n * W;
                uint32_t bad = 0, seed_bad = 0;
                uint64_t word0 = 0;
                switch (W) {
                    case 1: packb_cut<1>(cell, ci, sh, last_cell, L, wout, bad, seed_bad, word0); break;
                    case 2: packb_cut<2>(cell, ci, sh, last_cell, L, wout, bad, seed_bad, word0); break;
                    case 3: packb_cut<3>(cell, ci, sh, last_cell, L, wout, bad, seed_bad, word0); break;
                    case 4: packb_cut<4>(cell, ci, sh, last_cell, L, wout, bad, seed_bad, word0); break;
                    case 5: packb_cut<5>(cell, ci, sh, last_cell, L, wout, bad, seed_bad, word0); break;
                    case 6: packb_cut<6>(cell, ci, sh, last_cell, L, wout, bad, seed_bad, word0); break;
                    case 7: packb_cut<7>(cell, ci, sh, last_cell, L, wout, bad, seed_bad, word0); break;
                    case 8: packb_cut<8>(cell, ci, sh, last_cell, L, wout, bad, seed_bad, word0); break;
                    default: packb_cut_any(cell, ci, sh, last_cell, L, W, wout, bad, seed_bad, word0); break;
                }
                r.flags[n] = bad ? 1 : 0;
                if (ix.head) {
                    if (pend_n != 0xffffffffu) ix.next[pend_n] = make_uint2(pend_old, pend_tag);
                    int sg = s_tileseg[(tp - t0) & 1];
                    while (sg + 1 < ix.n_seg && (int64_t)n >= ix.seg_read_start[sg + 1]) sg++;
                    if (!seed_bad) {
                        const uint64_t seed = word0 & keep_bases(ix.seed_len[sg]);
                        const uint32_t h = seed_hash(seed) & (uint32_t)ix.tab_mask[sg];
                        pend_old = atomicExch(&ix.head[ix.tab_off[sg] + h], n + 1u);
                        pend_tag = seed_tag(seed);
                    } else {
                        pend_old = atomicExch(&ix.odd_head[sg], n + 1u);
                        pend_tag = 0u;
                    }
                    pend_n = n;
                }
            }
        }
        __syncthreads();  // cells of tile t complete (its stage is free); tile t-1 is cut (its cells are free)
        if (tid == 0 && t + PACKB_STAGES < t1) issue(t + PACKB_STAGES);
    }
    if (pend_n != 0xffffffffu) ix.next[pend_n] = make_uint2(pend_old, pend_tag);
}

}  // namespace bs
