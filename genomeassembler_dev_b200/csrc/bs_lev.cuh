// Infix edit distance of every contig against its truth (the reference's edlib call).
#pragma once
#include "bs_startpos.cuh"

namespace bs {

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
