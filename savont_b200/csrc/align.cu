// align.cu -- K3: banded, both-ends-free unit-cost edit distance (NM), bit-parallel.
//
// Stands in for the minimap2 call of reference src/alignment.rs:1848-1863 (NM of read vs ASV) with
// the north-star definition (SURVEY.md 8c "S2"):
//   rows i = 0..n (ASV), columns j = 0..m (oriented read), cells |i-j| <= B,
//   D[0][j] = D[i][0] = 0, unit costs, NM = min over last-row / last-column cells,
//   tie -> largest i+j, then the last-row cell.
//
// One THREAD per (read, ASV, strand) pair; the W = 2B+1 band cells of a column live in NW 32-bit
// registers as Myers/Hyyro delta vectors (Pv, Mv) aligned to the diagonal band: bit b of column j is
// row j-B+b.  Per column: Eq from the two bit-planes of the ASV window (funnel shifts) and the
// read base (broadcast masks), D0 = (((X&Pv)+Pv)^Pv)|X with one IADD3.X carry chain,
// Ph/Mh, then the diagonal re-alignment X = D0>>1.  Rows <= 0 are virtual cells with D = 0
// (forced matches, only while j <= B); rows > n and bits >= W are garbage below the band that never
// feeds back (the only back-edge, bit W -> W-1 through D0>>1, is masked in Mv).
// The score of the band's top cell is tracked from bit 0 of D0; last-row cells are read with
// popcounts of the vertical deltas, last-column cells by a walk over the final column.
// No tensor cores: this is integer DP, ~14 INT32 ops per 32 cells.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "align_addchain.cuh"

namespace sb {

struct AlignArgs {
    const uint64_t *read_offsets;
    const uint2 *read_planes;
    const uint2 *asv_planes;
    const uint32_t *asv_plane_off;
    const uint32_t *asv_len;
    const uint32_t *pair_read;
    const uint32_t *pair_asv;
    const uint8_t *pair_strand;
    const uint32_t *n_pairs_dev;   // device-side pair count (may be null -> n_pairs_max)
    const uint32_t *err;           // device error word (may be null): non-zero -> an upstream pool overflowed, do nothing
    uint32_t n_pairs_max;
    int32_t *nm, *asv_end, *read_end;
    // traceback mode
    uint32_t *hist;                // checkpoints: [(block*2*NW + w) * hist_stride + t], block = 32 columns; Pv words then Mv words
    uint32_t hist_stride;
    int32_t *asv_start, *read_start;
    uint32_t *cigar;               // [pair * cigar_stride + ...] filled from the back
    uint32_t cigar_stride;
    uint32_t *cigar_cnt;
    int band, padl;
};

// s = a + b (+ carry chain across NW words) in ONE asm statement so nothing can clobber CC: align_addchain.cuh (generated)
// 32 bases [32t, 32t+32) of the oriented read as (low plane, high plane)
__device__ __forceinline__ uint2 read_word(const uint2 *rp, uint32_t m, uint32_t t, bool rev) {
    if (!rev) return rp[t];
    const int s0 = (int)m - 32 * (int)t - 32;          // forward index of the LAST base of this word
    uint32_t lo, hi;
    if (s0 >= 0) {
        const uint32_t ws = (uint32_t)s0 >> 5, sh = (uint32_t)s0 & 31u;
        const uint2 a = rp[ws], b = rp[ws + 1];
        lo = __funnelshift_r(a.x, b.x, sh);
        hi = __funnelshift_r(a.y, b.y, sh);
    } else {
        const uint2 a = rp[0];
        lo = a.x << (uint32_t)(-s0);
        hi = a.y << (uint32_t)(-s0);
    }
    return make_uint2(~__brev(lo), ~__brev(hi));        // reverse + complement (3 - code)
}

// TAIL: the band has exactly W = 32 * NW + 1 cells (B a multiple of 16: 2B + 1 = 129 for the default band 64).  Bits
// 0 .. 32NW-1 live in NW words; of the last cell (bit 32NW, row j + B) only the diagonal delta D0 is ever needed -- its vertical
// delta Mv is always 0 (masked after the diagonal shift), its Pv / Ph / Mh feed nothing but themselves, and no score is read at
// or below it -- and D0 = Eq | carry-out of the word chain.  So the bottom cell costs one shift, one carry read-back and one OR
// per column instead of a whole extra word (ten logic ops): 4 words instead of 5 for band 64, 2 instead of 3 for band 32.
template <int NW, bool TB, bool TAIL>
struct Band {
    uint32_t Pv[NW], Mv[NW];
    uint32_t Alo[NW + 1], Ahi[NW + 1];   // ASV window planes, word aligned; view = funnel shift by sh
    uint32_t vm2[TAIL ? 1 : NW];         // bits < W-1 (mask for Mv after the diagonal shift); all ones with TAIL
    uint32_t aw;                         // next ASV plane word to load
    const uint2 *ap;
    int T;                               // D at the band's top cell, up to the columns folded so far
    uint32_t tacc;                       // bit 0 of D0 of the columns since the last fold, newest on top; unfilled bits are ones
    int bd, bi, bj;                      // best end cell so far
    int dn;                              // D(n, j) of the last column whose last-row cell was evaluated
    // D at the band's top cell grows by 1 - (bit 0 of D0) per column
    __device__ __forceinline__ int top() const { return T + __popc(~tacc); }
    __device__ __forceinline__ void fold() { T += __popc(~tacc); tacc = 0xffffffffu; }
};

// one DP column.  TOP: rows <= 0 may be inside the window (j <= B).  LAST: j may be >= n-B, the last-row cell (n, j) is
// scored: 1 = from the band's top score and the vertical deltas (popcounts over the column), 2 = from the previous
// column's D(n, j-1) and the horizontal delta bits of row n (needs a LAST column right before this one).
// et (TAIL): match bit of the bottom cell in the sign bit.
template <int NW, bool TB, bool TOP, int LAST, bool TAIL, bool BTB>
__device__ __forceinline__ void dp_column(Band<NW, BTB, TAIL> &s, uint32_t (&TM)[NW], uint32_t clo, uint32_t chi, uint32_t sh, uint32_t et,
                                          int j, int n, int B, uint32_t *hist_col, uint32_t hist_stride) {
    uint32_t Nr[NW + 1], X[NW], t[NW], sum[NW], D0[NW], Ph[NW], Mh[NW];
    // MISmatch bits on the word-aligned window first (2 LOP3 per raw word), ONE funnel shift per band word afterwards;
    // the complement is folded into the LOP3 that forms X = Eq | Mv
#pragma unroll
    for (int w = 0; w <= NW; ++w) Nr[w] = (s.Alo[w] ^ clo) | (s.Ahi[w] ^ chi);
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const uint32_t Ns = __funnelshift_r(Nr[w], Nr[w + 1], sh);
        X[w] = TOP ? (~Ns | s.Mv[w] | TM[w]) : (~Ns | s.Mv[w]);
        t[w] = X[w] & s.Pv[w];
    }
    uint32_t d0t = 0u;
    if (TAIL) {
        add_chain_co<NW>(t, s.Pv, sum, et >> 31, d0t);      // D0 of the bottom cell = match | carry out (only bit 0 is used)
    } else {
        add_chain<NW>(t, s.Pv, sum);
    }
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        D0[w] = (sum[w] ^ s.Pv[w]) | X[w];
        // Ph = Mv | ~(D0 | Pv) as ONE LOP3 on (Mv, D0, Pv): left to itself the compiler rebuilds D0 | Pv from
        // (sum, Pv, X) with a second LOP3 per word
        asm("lop3.b32 %0, %1, %2, %3, 0xF1;" : "=r"(Ph[w]) : "r"(s.Mv[w]), "r"(D0[w]), "r"(s.Pv[w]));
        Mh[w] = s.Pv[w] & D0[w];
    }
    // collects bit 0 of D0 (one funnel shift; a mask + add per column otherwise): folded by popcount once per run of <= 32 columns
    s.tacc = __funnelshift_r(s.tacc, D0[0], 1);
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const uint32_t Xs = (w + 1 < NW) ? __funnelshift_r(D0[w], D0[w + 1], 1) : (TAIL ? __funnelshift_r(D0[w], d0t, 1) : (D0[w] >> 1));
        s.Pv[w] = Mh[w] | ~(Xs | Ph[w]);
        s.Mv[w] = TAIL ? (Ph[w] & Xs) : (Ph[w] & Xs & s.vm2[w]);
    }
    if (TB) {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            hist_col[(size_t)w * hist_stride] = D0[w];
            hist_col[(size_t)(NW + w) * hist_stride] = s.Pv[w];
        }
        if (TAIL) hist_col[(size_t)(2 * NW) * hist_stride] = d0t;       // D0 of the bottom cell (bit 0)
    }
    if (TOP) {
#pragma unroll
        for (int w = 0; w < NW; ++w) TM[w] = (w + 1 < NW) ? __funnelshift_r(TM[w], TM[w + 1], 1) : (TM[w] >> 1);
    }
    if (LAST == 1) {
        if (j + B >= n) {                                   // row n is inside the window: bit b = n - j + B
            const int b = n - j + B;
            int d = s.top();
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int rem = b - 32 * w;                 // number of delta bits of this word below b
                const uint32_t msk = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
                d += __popc(s.Pv[w] & msk) - __popc(s.Mv[w] & msk);
            }
            s.dn = d;
            if (d < s.bd || (d == s.bd && n + j > s.bi + s.bj)) { s.bd = d; s.bi = n; s.bj = j; }
        }
    }
    if (LAST == 2) {                                        // j > n - B here: D(n, j) = D(n, j-1) + Ph[b] - Mh[b]
        const int b = n - j + B;
        const int bw = b >> 5;
        uint32_t ph = Ph[0], mh = Mh[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) if (bw == w) { ph = Ph[w]; mh = Mh[w]; }
        const int d = s.dn + (int)((ph >> (b & 31)) & 1u) - (int)((mh >> (b & 31)) & 1u);
        s.dn = d;
        if (d <= s.bd) { s.bd = d; s.bi = n; s.bj = j; }     // n + j grows with j: a tie goes to the later column
    }
}

template <int NW, bool TB, bool TAIL>
__device__ __forceinline__ void align_pair(const AlignArgs &a, uint32_t pi, uint32_t tslot) {
    const uint32_t r = a.pair_read[pi], asv = a.pair_asv[pi];
    const bool rev = a.pair_strand[pi] != 0;
    const uint64_t roff = a.read_offsets[r];
    const int m = (int)(a.read_offsets[r + 1] - roff);
    const int n = (int)a.asv_len[asv];
    const int B = a.band, W = 2 * B + 1;
    const uint2 *rp = a.read_planes + plane_word_off(roff, r);
    const int jmax = m < n + B ? m : n + B;

    Band<NW, TB, TAIL> s;
    s.ap = a.asv_planes + a.asv_plane_off[asv];
    s.T = 0; s.tacc = 0xffffffffu; s.bd = 1 << 29; s.bi = -1; s.bj = -1;
    uint32_t TM[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        s.Pv[w] = 0u; s.Mv[w] = 0u;
        if (!TAIL) {
            const int remW = (W - 1) - 32 * w;              // bits of this word below W-1
            s.vm2[w] = remW >= 32 ? 0xffffffffu : (remW <= 0 ? 0u : ((1u << remW) - 1u));
        }
        const int remT = B - 32 * w;                        // column 1: rows <= 0 are bits 0..B-1
        TM[w] = remT >= 32 ? 0xffffffffu : (remT <= 0 ? 0u : ((1u << remT) - 1u));
    }
    if (TAIL) s.vm2[0] = 0xffffffffu;
    // The planes are left-padded by padl = B + 32 zero bits, so the window of column j starts at padded
    // bit 32 + (j - 1): its offset inside the register window equals the read's bit index (j-1) & 31,
    // and both the ASV window and the read word advance at the same 32-column boundaries.
    {
#pragma unroll
        for (int w = 0; w <= NW; ++w) { const uint2 v = s.ap[1 + w]; s.Alo[w] = v.x; s.Ahi[w] = v.y; }
        s.aw = NW + 2;
    }
    // degenerate: no columns
    if (jmax <= 0) {
        if (TB) { a.asv_start[pi] = 0; a.read_start[pi] = 0; a.cigar_cnt[pi] = 0; }
        // empty read: D[i][0] = 0 for i <= B; the tie rule picks the largest i
        a.nm[pi] = 0;
        if (a.asv_end) { a.asv_end[pi] = n <= B ? n : B; a.read_end[pi] = 0; }
        return;
    }

    uint32_t cw_lo = 0, cw_hi = 0;
    int j = 1;
    uint32_t *hist = TB ? (a.hist + tslot) : nullptr;
    const size_t hist_col_stride = (size_t)2 * NW * a.hist_stride;

    // runs columns j..jend of one phase; inner runs never cross a 32-column boundary, so the word
    // reload / register rotation is a real (rare) branch and the column body stays straight-line
    auto run_phase = [&](auto top_tag, auto last_tag, const int jend) {
        constexpr bool TOP = decltype(top_tag)::value;
        constexpr int LAST = decltype(last_tag)::value;
        while (j <= jend) {
            uint32_t bit = (uint32_t)(j - 1) & 31u;
            if (bit == 0u) {
                const uint2 v = read_word(rp, (uint32_t)m, (uint32_t)(j - 1) >> 5, rev);
                cw_lo = v.x; cw_hi = v.y;
                if (j > 1) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) { s.Alo[w] = s.Alo[w + 1]; s.Ahi[w] = s.Ahi[w + 1]; }
                    const uint2 nv = s.ap[s.aw++];
                    s.Alo[NW] = nv.x; s.Ahi[NW] = nv.y;
                }
                if (TB) {                                   // checkpoint: the vertical deltas in front of column j = 32 * block + 1
                    uint32_t *ck = hist + (size_t)((uint32_t)(j - 1) >> 5) * hist_col_stride;
#pragma unroll
                    for (int w = 0; w < NW; ++w) { ck[(size_t)w * a.hist_stride] = s.Pv[w]; ck[(size_t)(NW + w) * a.hist_stride] = s.Mv[w]; }
                }
            }
            // columns j..stop share the read word; the loop runs on the bit index alone.  cl / ch hold the bit-reversed
            // read word shifted so that the current column's bit is the sign bit: the broadcast mask is one arithmetic
            // shift and the advance is an add (IMAD on the FMA pipe), instead of a variable shift pair per plane.  (IMAD.HI as the
            // sign broadcast was measured: it costs an ALU-pipe slot as well on B200, tools/pipe_bench.cu, so no gain.)
            // TAIL: the bottom cell of column j is window bit 32NW + bit, i.e. bit `bit` of the last raw word -- the same bit
            // index as the read base -- so its match bits for the whole run are one word, walked like the read word.
            const int stop = min(jend, j + 31 - (int)bit);
            const uint32_t bit_end = bit + (uint32_t)(stop - j);
            const int jbase = j - (int)bit;
            uint32_t cl = __brev(cw_lo) << bit, ch = __brev(cw_hi) << bit;
            uint32_t el = TAIL ? (__brev(~((s.Alo[NW] ^ cw_lo) | (s.Ahi[NW] ^ cw_hi))) << bit) : 0u;
            s.fold();
            auto column = [&](const uint32_t sh) {
                const uint32_t clo = (uint32_t)((int32_t)cl >> 31);
                const uint32_t chi = (uint32_t)((int32_t)ch >> 31);
                const uint32_t et = el;
                cl += cl; ch += ch;
                if (TAIL) el += el;
                dp_column<NW, false, TOP, LAST, TAIL>(s, TM, clo, chi, sh, et, jbase + (int)sh, n, B, nullptr, 0u);
            };
            if (!TB && !TOP && LAST == 0 && bit == 0u && bit_end == 31u) {
                // a whole 32-column block of the interior: fixed trip count, the loop counter amortised over 8 columns
#pragma unroll 8
                for (uint32_t sh = 0; sh < 32u; ++sh) column(sh);
            } else {
                for (; bit <= bit_end; ++bit) column(bit);
            }
            j = stop + 1;
        }
    };
    // phase 1: virtual rows (<= 0) inside the window; phase 2: interior; phase 3: last row inside the window -- its first
    // column scores the last-row cell from the column's vertical deltas, the following ones incrementally
    using L0 = std::integral_constant<int, 0>; using L1 = std::integral_constant<int, 1>; using L2 = std::integral_constant<int, 2>;
    run_phase(std::true_type{}, L1{}, B < jmax ? B : jmax);
    run_phase(std::false_type{}, L0{}, (n - B - 1) < jmax ? (n - B - 1) : jmax);
    if (j <= jmax) run_phase(std::false_type{}, L1{}, j);
    run_phase(std::false_type{}, L2{}, jmax);
    // row n may already be inside the window of column 0 (n <= B): cell (n, 0) = 0 is a last-row cell
    // (handled below together with the last column when m == 0; for m > 0 the oracle also considers it)
    if (n <= B) {
        const int d = 0;
        if (d < s.bd || (d == s.bd && n + 0 > s.bi + s.bj)) { s.bd = d; s.bi = n; s.bj = 0; }
    }
    // last-column cells (only if column m was reached)
    if (jmax == m) {
        int d = s.top();                                    // b = 0: row m - B
        const int bmin = B - m > 0 ? B - m : 0;
        const int bmax = (n - m + B) < 2 * B ? (n - m + B) : 2 * B;
        if (bmin == 0 && bmax >= 0) {
            const int i = m - B;
            if (d < s.bd || (d == s.bd && i + m > s.bi + s.bj)) { s.bd = d; s.bi = i; s.bj = m; }
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            uint32_t pv = s.Pv[w], mv = s.Mv[w];
            for (int q = 0; q < 32; ++q) {
                const int b = 32 * w + q + 1;               // delta bit (b-1) leads to band bit b
                if (b > bmax) break;
                d += (int)(pv & 1u) - (int)(mv & 1u);
                pv >>= 1; mv >>= 1;
                if (b >= bmin) {
                    const int i = m - B + b;
                    if (d < s.bd || (d == s.bd && i + m > s.bi + s.bj)) { s.bd = d; s.bi = i; s.bj = m; }
                }
            }
        }
    }
    a.nm[pi] = s.bd;
    if (a.asv_end) { a.asv_end[pi] = s.bi; a.read_end[pi] = s.bj; }

    if constexpr (TB) {
        // ---- traceback; preference diagonal > D (ASV) > I (read) ----
        // The forward pass kept only a checkpoint of the vertical deltas every 32 columns (2 NW words: 1.9 KB per 1.5 kb pair at band
        // 64; round 1 stored D0 / Pv' of EVERY column, 60 KB per pair, and streamed it through HBM twice).  The walk goes back block by
        // block: the block's columns up to the current one are recomputed from the checkpoint into a per-thread buffer in local memory
        // (32 columns x 2 NW words, L1 / L2 resident), then walked.  One more forward pass in total, no history traffic to HBM.
        constexpr int HW = 2 * NW + (TAIL ? 1 : 0);          // words per recomputed column: D0, Pv' (+ the bottom cell's D0)
        uint32_t hb[32 * HW];
        const uint2 *app = a.asv_planes + a.asv_plane_off[asv];
        int i = s.bi, jj = s.bj;
        uint32_t *cg = a.cigar + (size_t)tslot * a.cigar_stride;
        uint32_t pos = a.cigar_stride;           // fill from the back
        uint32_t cur_op = 3u, cur_len = 0u;
        uint2 av = make_uint2(0u, 0u), rv = make_uint2(0u, 0u); uint32_t avw = 0xffffffffu, rvw = 0xffffffffu;
        auto emit = [&](const uint32_t op, const uint32_t len) {
            if (op == cur_op) cur_len += len;
            else {
                if (cur_len && pos > 0) cg[--pos] = (cur_len << 4) | cur_op;
                cur_op = op; cur_len = len;
            }
        };
        // A cell whose bases match is left along the diagonal whatever the DP says (the preference order), so runs of matches are
        // taken 32 bases at a time from the bit planes alone: window bit 31 - t = base (i - 1 - t) of the ASV / oriented base
        // (jj - 1 - t) of the read.  Only the cells where the bases differ need the DP's deltas.
        auto skip_matches = [&]() {
            while (i > 0 && jj > 0) {
                const uint32_t lo = (uint32_t)(i - 1 + a.padl) - 31u;                    // padl >= 32: never negative
                const uint2 a0 = app[lo >> 5], a1 = app[(lo >> 5) + 1];
                const uint32_t wax = __funnelshift_r(a0.x, a1.x, lo & 31u), way = __funnelshift_r(a0.y, a1.y, lo & 31u);
                uint32_t wrx, wry;
                if (!rev) {
                    const int r0 = jj - 32;                                             // bits r0 .. jj - 1 of the forward read
                    if (r0 >= 0) {
                        const uint2 b0 = rp[(uint32_t)r0 >> 5], b1 = rp[((uint32_t)r0 >> 5) + 1];
                        wrx = __funnelshift_r(b0.x, b1.x, (uint32_t)r0 & 31u); wry = __funnelshift_r(b0.y, b1.y, (uint32_t)r0 & 31u);
                    } else { const uint2 b0 = rp[0]; wrx = b0.x << (uint32_t)(-r0); wry = b0.y << (uint32_t)(-r0); }
                } else {
                    const uint32_t f0 = (uint32_t)(m - jj);                             // oriented base jj - 1 - t = complement of forward base f0 + t
                    const uint2 b0 = rp[f0 >> 5], b1 = rp[(f0 >> 5) + 1];
                    wrx = ~__brev(__funnelshift_r(b0.x, b1.x, f0 & 31u)); wry = ~__brev(__funnelshift_r(b0.y, b1.y, f0 & 31u));
                }
                const uint32_t mism = (wax ^ wrx) | (way ^ wry);
                uint32_t run = (uint32_t)__clz((int)mism);                              // 32 when the whole window matches
                run = min(run, min((uint32_t)i, (uint32_t)jj));
                if (run == 0u) break;
                i -= (int)run; jj -= (int)run;
                emit(0u, run);
                if (run < 32u) break;                                                   // stopped at a mismatch (or at a sequence start)
            }
        };
        // Outer loop = one block, inner loop = the walk inside it: the lanes of a warp leave the inner loop at different times and
        // reconverge behind it, so the recomputation (the expensive part, identical instructions for every lane) runs in lockstep.
        // It always covers the block's 32 columns -- those behind the current one, or behind the read's end, are computed but never read.
        // Blocks that the walk crosses inside a run of matches are never recomputed.
        while (true) {
            skip_matches();
            if (!(i > 0 && jj > 0)) break;
            const int k = (jj - 1) >> 5;
            {
                const uint32_t *ck = hist + (size_t)k * hist_col_stride;
#pragma unroll
                for (int w = 0; w < NW; ++w) { s.Pv[w] = ck[(size_t)w * a.hist_stride]; s.Mv[w] = ck[(size_t)(NW + w) * a.hist_stride]; }
#pragma unroll
                for (int w = 0; w <= NW; ++w) { const uint2 v = app[k + 1 + w]; s.Alo[w] = v.x; s.Ahi[w] = v.y; }
#pragma unroll
                for (int w = 0; w < NW; ++w) {               // the virtual-row mask of column 1 moved down by 32 k rows
                    const int remT = B - 32 * (w + k);
                    TM[w] = remT >= 32 ? 0xffffffffu : (remT <= 0 ? 0u : ((1u << remT) - 1u));
                }
                const uint2 cwv = read_word(rp, (uint32_t)m, (uint32_t)k, rev);
                const uint32_t elw = TAIL ? ~((s.Alo[NW] ^ cwv.x) | (s.Ahi[NW] ^ cwv.y)) : 0u;      // match bits of the bottom cells of the block's columns
                for (int sh = 0; sh < 32; ++sh) {
                    const uint32_t clo = 0u - ((cwv.x >> sh) & 1u), chi = 0u - ((cwv.y >> sh) & 1u);
                    const uint32_t et = (elw >> sh) << 31;
                    const int jc = 32 * k + 1 + sh;
                    if (jc <= B) dp_column<NW, true, true, 0, TAIL>(s, TM, clo, chi, (uint32_t)sh, et, jc, n, B, hb + sh * HW, 1u);
                    else dp_column<NW, true, false, 0, TAIL>(s, TM, clo, chi, (uint32_t)sh, et, jc, n, B, hb + sh * HW, 1u);
                }
            }
            while (i > 0 && jj > 32 * k) {
                const int c = jj - 1;
                const int b = i - jj + B;
                const uint32_t *hc = hb + (c & 31) * HW;
                const uint32_t d0 = (TAIL && (b >> 5) == NW) ? (hc[2 * NW] & 1u) : ((hc[b >> 5] >> (b & 31)) & 1u);
                // bases: ASV row i -> base i-1 at padded bit i-1+padl; read column jj -> oriented base jj-1
                const uint32_t abit = (uint32_t)(i - 1 + a.padl);
                if ((abit >> 5) != avw) { avw = abit >> 5; av = app[avw]; }
                const uint32_t ac = ((av.x >> (abit & 31u)) & 1u) | (((av.y >> (abit & 31u)) & 1u) << 1);
                const uint32_t rb = rev ? (uint32_t)(m - jj) : (uint32_t)(jj - 1);
                if ((rb >> 5) != rvw) { rvw = rb >> 5; rv = rp[rvw]; }
                uint32_t rc = ((rv.x >> (rb & 31u)) & 1u) | (((rv.y >> (rb & 31u)) & 1u) << 1);
                if (rev) rc = 3u - rc;
                uint32_t op;
                if (ac == rc || d0 == 0u) { op = 0u; --i; --jj; }
                else {
                    uint32_t pvb = 0u;
                    if (b >= 1) pvb = (hc[NW + ((b - 1) >> 5)] >> ((b - 1) & 31)) & 1u;
                    if (pvb) { op = 2u; --i; } else { op = 1u; --jj; }
                }
                emit(op, 1u);
                skip_matches();                                  // the cells behind an edit mostly match again: the run may leave the block
            }
        }
        if (cur_len && pos > 0) cg[--pos] = (cur_len << 4) | cur_op;
        a.asv_start[pi] = i; a.read_start[pi] = jj;
        a.cigar_cnt[pi] = a.cigar_stride - pos;
    }
}

template <int NW, bool TAIL>
__global__ void __launch_bounds__(128) align_nm_kernel(AlignArgs a) {
    if (a.err && *a.err) return;
    const uint32_t n_pairs = a.n_pairs_dev ? min(*a.n_pairs_dev, a.n_pairs_max) : a.n_pairs_max;
    for (uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x; pi < n_pairs; pi += gridDim.x * blockDim.x)
        align_pair<NW, false, TAIL>(a, pi, 0);
}
template <int NW, bool TAIL>
__global__ void __launch_bounds__(128) align_tb_kernel(AlignArgs a) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_pairs_max) return;
    align_pair<NW, true, TAIL>(a, t, t);
}

static int pick_nw(int band) {
    const int need = (2 * band + 1 + 31) / 32;
    const int avail[] = {1, 2, 3, 5, 7, 9, 13};
    for (int v : avail) if (v >= need) return v;
    return -1;
}
// bands with 2B + 1 = 32q + 1 cells: q words + the tail bit (0 = no such variant); SAVONT_B200_DP_TAIL=0 forces the
// whole-word kernels (tests compare the two)
static int pick_nw_tail(int band) {
    static const bool off = [] { const char *e = getenv("SAVONT_B200_DP_TAIL"); return e && atoi(e) == 0; }();
    if (off || band < 16 || band % 16 != 0) return 0;
    const int q = band / 16;
    const int avail[] = {1, 2, 3, 4, 6, 8, 12};
    for (int v : avail) if (v == q) return v;
    return 0;
}
int align_supported_band_max() { return (13 * 32 - 1) / 2; }
int align_words_for_band(int band) { return pick_nw(band); }
// vector bits the NM kernel computes per column for this band (the roofline's denominator): 32q + 1 with the tail variant
int align_bits_for_band(int band) { const int q = pick_nw_tail(band); return q ? 32 * q + 1 : 32 * pick_nw(band); }

static AlignArgs make_args(const SeqBatch &in, const uint2 *read_planes, const AsvIndex &ix, const uint32_t *pair_read,
                           const uint32_t *pair_asv, const uint8_t *pair_strand, const AlignOut &out, const DeviceParams &p) {
    AlignArgs a{};
    a.read_offsets = in.offsets; a.read_planes = read_planes;
    a.asv_planes = ix.planes; a.asv_plane_off = ix.plane_off; a.asv_len = ix.len;
    a.pair_read = pair_read; a.pair_asv = pair_asv; a.pair_strand = pair_strand;
    a.nm = out.nm; a.asv_end = out.asv_end; a.read_end = out.read_end;
    a.band = p.band; a.padl = p.asv_padl;
    return a;
}

void launch_align_nm(const SeqBatch &in, const uint2 *read_planes, const AsvIndex &ix, const uint32_t *pair_read,
                     const uint32_t *pair_asv, const uint8_t *pair_strand, const uint32_t *n_pairs_dev, uint32_t n_pairs_max,
                     const uint32_t *err, const AlignOut &out, const DeviceParams &p, int sm_count, cudaStream_t st) {
    if (n_pairs_max == 0) return;
    AlignArgs a = make_args(in, read_planes, ix, pair_read, pair_asv, pair_strand, out, p);
    a.n_pairs_dev = n_pairs_dev; a.n_pairs_max = n_pairs_max; a.err = err;
    uint32_t grid = (n_pairs_max + 127u) / 128u;
    // CTAs per SM: 8 x 128 threads x 64 registers fill an SM's register file; SAVONT_B200_DP_CTAS = 4..6 leaves room for the
    // sketch / filter kernels of the next part of the pipeline (api.cu map_resident)
    static const uint32_t per_sm = [] { const char *e = getenv("SAVONT_B200_DP_CTAS"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= 32) ? (uint32_t)v : 8u; }();
    const uint32_t cap = (uint32_t)sm_count * per_sm;
    if (grid > cap) grid = cap;
    switch (pick_nw_tail(p.band)) {
        case 1: align_nm_kernel<1, true><<<grid, 128, 0, st>>>(a); return;
        case 2: align_nm_kernel<2, true><<<grid, 128, 0, st>>>(a); return;
        case 3: align_nm_kernel<3, true><<<grid, 128, 0, st>>>(a); return;
        case 4: align_nm_kernel<4, true><<<grid, 128, 0, st>>>(a); return;
        case 6: align_nm_kernel<6, true><<<grid, 128, 0, st>>>(a); return;
        case 8: align_nm_kernel<8, true><<<grid, 128, 0, st>>>(a); return;
        case 12: align_nm_kernel<12, true><<<grid, 128, 0, st>>>(a); return;
        default: break;
    }
    switch (pick_nw(p.band)) {
        case 1: align_nm_kernel<1, false><<<grid, 128, 0, st>>>(a); break;
        case 2: align_nm_kernel<2, false><<<grid, 128, 0, st>>>(a); break;
        case 3: align_nm_kernel<3, false><<<grid, 128, 0, st>>>(a); break;
        case 5: align_nm_kernel<5, false><<<grid, 128, 0, st>>>(a); break;
        case 7: align_nm_kernel<7, false><<<grid, 128, 0, st>>>(a); break;
        case 9: align_nm_kernel<9, false><<<grid, 128, 0, st>>>(a); break;
        case 13: align_nm_kernel<13, false><<<grid, 128, 0, st>>>(a); break;
        default: break;
    }
}

size_t align_tb_hist_words(uint32_t n_pairs, uint32_t hist_cols, const DeviceParams &p) {
    const int nw = pick_nw_tail(p.band) ? pick_nw_tail(p.band) : pick_nw(p.band);
    return (size_t)n_pairs * ((hist_cols + 31u) / 32u + 1u) * 2u * (size_t)(nw > 0 ? nw : 1);      // one checkpoint per 32 columns
}

void launch_align_tb(const SeqBatch &in, const uint2 *read_planes, const AsvIndex &ix, const uint32_t *pair_read,
                     const uint32_t *pair_asv, const uint8_t *pair_strand, uint32_t n_pairs,
                     uint32_t *hist, const AlignOut &out, int32_t *asv_start, int32_t *read_start,
                     uint32_t *cigar, uint32_t cigar_stride, uint32_t *cigar_cnt, const DeviceParams &p, cudaStream_t st) {
    if (n_pairs == 0) return;
    AlignArgs a = make_args(in, read_planes, ix, pair_read, pair_asv, pair_strand, out, p);
    a.n_pairs_dev = nullptr; a.n_pairs_max = n_pairs;
    a.hist = hist; a.hist_stride = n_pairs;
    a.asv_start = asv_start; a.read_start = read_start;
    a.cigar = cigar; a.cigar_stride = cigar_stride; a.cigar_cnt = cigar_cnt;
    const uint32_t grid = (n_pairs + 127u) / 128u;
    switch (pick_nw_tail(p.band)) {                      // bands of 32 q + 1 cells: q words + the tail bit, as in the NM pass
        case 1: align_tb_kernel<1, true><<<grid, 128, 0, st>>>(a); return;
        case 2: align_tb_kernel<2, true><<<grid, 128, 0, st>>>(a); return;
        case 3: align_tb_kernel<3, true><<<grid, 128, 0, st>>>(a); return;
        case 4: align_tb_kernel<4, true><<<grid, 128, 0, st>>>(a); return;
        case 6: align_tb_kernel<6, true><<<grid, 128, 0, st>>>(a); return;
        case 8: align_tb_kernel<8, true><<<grid, 128, 0, st>>>(a); return;
        case 12: align_tb_kernel<12, true><<<grid, 128, 0, st>>>(a); return;
        default: break;
    }
    switch (pick_nw(p.band)) {
        case 1: align_tb_kernel<1, false><<<grid, 128, 0, st>>>(a); break;
        case 2: align_tb_kernel<2, false><<<grid, 128, 0, st>>>(a); break;
        case 3: align_tb_kernel<3, false><<<grid, 128, 0, st>>>(a); break;
        case 5: align_tb_kernel<5, false><<<grid, 128, 0, st>>>(a); break;
        case 7: align_tb_kernel<7, false><<<grid, 128, 0, st>>>(a); break;
        case 9: align_tb_kernel<9, false><<<grid, 128, 0, st>>>(a); break;
        case 13: align_tb_kernel<13, false><<<grid, 128, 0, st>>>(a); break;
        default: break;
    }
}

}  // namespace sb
