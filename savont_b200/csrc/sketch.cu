// sketch.cu -- K1: 2-bit packing + open-syncmer / SNPmer sketch of a batch of sequences.
//
// Stands in for seeding::get_twin_read_syncmer (reference src/seeding.rs:317-658), one warp per
// sequence:
//   phase A  coalesced byte loads -> BYTE_TO_SEQ codes (src/types.rs:92-101) -> warp ballots give
//            the two bit planes (stored for the DP) and a 16-bases-per-word packed copy in shared
//            memory; all-equal-quality check (src/seeding.rs:372-380).
//   phase B  the k-mer start positions are split into 32 contiguous chunks, one per lane, and every
//            lane ROLLS its forward / reverse-complement k-mer registers over its chunk exactly as
//            the reference does (:419-443), after a (c-1)-position warm-up that refills the
//            open-syncmer window.  Per position: splitmer-canonical choice (:426-434, tie ->
//            reverse); the s-mer hash is replaced by its RANK among all 4^s s-mers, looked up in a
//            shared-memory table indexed by the k-mer's last s bases (only the ORDER of the hashes is
//            ever used, :526-543); the window of c ranks lives in registers (loop unrolled by c) and
//            the strict-minimum-at-the-middle test is a min tree; SNPmer membership (:508-521) goes
//            through a two-hash Bloom filter in shared memory before the exact global table.
//            Minimizers are buffered per lane in shared memory and written out lane-major, which is
//            position order.
//   phase C  SNPmer de-duplication by splitmer multiplicity (:550-568), position sort by rank
//            counting, and pool allocation.
// The reference's s-mer warm-up quirk (SURVEY.md 8a note 1) is reproduced: for k-mer start
// positions q < s-1 the "s-mer" is bases q..s-2 followed by bases k-1..k-1+q.
#include <type_traits>

#include "common.cuh"
#include "sketch_dev.cuh"

namespace sb {

// phase A shared by the sketch kernel and pack_planes_kernel: 128 bases per warp iteration.
// Every lane loads one aligned 32-bit word of bases (and of qualities), the neighbour's word comes by
// shuffle and a funnel shift removes the misalignment of the sequence start; the four bytes go through
// the BYTE_TO_SEQ table in shared memory (src/types.rs:92-101).  The 4-bit plane nibbles of 8 lanes are
// OR-reduced into one 32-bit plane word, the 8-bit packed codes of 4 lanes into one 16-base word.
// Returns true iff qualities are present and all equal (src/seeding.rs:372-380).
__device__ __forceinline__ bool pack_sequence(const SeqBatch &in, uint64_t off, uint32_t L, uint2 *planes, uint32_t *packed,
                                              const uint8_t *lut, int lane) {
    const uint32_t nblk = (L + 127u) >> 7;
    const uint32_t nw = (L + 31u) >> 5;             // plane words; word nw is the zero guard word of the DP look-ahead
    const uint8_t *bp = in.bases + off;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(bp) & 3u);
    const uint32_t *bw = reinterpret_cast<const uint32_t *>(bp - mis);
    const uint8_t *qp = in.quals ? in.quals + off : nullptr;
    const uint32_t qmis = qp ? (uint32_t)(reinterpret_cast<uintptr_t>(qp) & 3u) : 0u;
    const uint32_t *qw = qp ? reinterpret_cast<const uint32_t *>(qp - qmis) : nullptr;
    const uint32_t q0x4 = qp ? (uint32_t)qp[0] * 0x01010101u : 0u;
    uint32_t qdiff = 0;
    // software pipeline: the words of block blk+1 are requested before block blk is processed
    auto load_b = [&](uint32_t w) -> uint32_t { return (w * 4u < L + 4u) ? __ldg(bw + w) : 0u; };
    auto load_q = [&](uint32_t w) -> uint32_t { return (qw && w * 4u < L + 4u) ? __ldg(qw + w) : 0u; };
    uint32_t xn = load_b(lane), yn = load_q(lane);
    uint32_t xe = (lane == 31 && mis) ? load_b(32u) : 0u, ye = (lane == 31 && qmis) ? load_q(32u) : 0u;
    for (uint32_t blk = 0; blk < nblk; ++blk) {
        const uint32_t w = blk * 32u + lane;
        const uint32_t g = w * 4u;                         // first base handled by this lane
        const uint32_t x0 = xn, y0 = yn, xe0 = xe, ye0 = ye;
        if (blk + 1u < nblk) {
            xn = load_b(w + 32u); yn = load_q(w + 32u);
            if (lane == 31) { xe = mis ? load_b(w + 33u) : 0u; ye = qmis ? load_q(w + 33u) : 0u; }
        }
        uint32_t x1 = __shfl_down_sync(0xffffffffu, x0, 1);
        if (lane == 31) x1 = xe0;
        const uint32_t x = __funnelshift_r(x0, x1, 8u * mis);
        const uint32_t nvalid = (g < L) ? min(4u, L - g) : 0u;
        uint32_t c0 = lut[x & 0xffu], c1 = lut[(x >> 8) & 0xffu], c2 = lut[(x >> 16) & 0xffu], c3 = lut[x >> 24];
        c0 = nvalid > 0u ? c0 : 0u; c1 = nvalid > 1u ? c1 : 0u; c2 = nvalid > 2u ? c2 : 0u; c3 = nvalid > 3u ? c3 : 0u;
        uint32_t lo = (c0 & 1u) | ((c1 & 1u) << 1) | ((c2 & 1u) << 2) | ((c3 & 1u) << 3);
        uint32_t hi = (c0 >> 1) | ((c1 >> 1) << 1) | ((c2 >> 1) << 2) | ((c3 >> 1) << 3);
        lo <<= 4u * (lane & 7u); hi <<= 4u * (lane & 7u);
        lo |= __shfl_xor_sync(0xffffffffu, lo, 1); hi |= __shfl_xor_sync(0xffffffffu, hi, 1);
        lo |= __shfl_xor_sync(0xffffffffu, lo, 2); hi |= __shfl_xor_sync(0xffffffffu, hi, 2);
        lo |= __shfl_xor_sync(0xffffffffu, lo, 4); hi |= __shfl_xor_sync(0xffffffffu, hi, 4);
        if ((lane & 7) == 0 && blk * 4u + (lane >> 3) <= nw) planes[blk * 4u + (lane >> 3)] = make_uint2(lo, hi);
        if (packed) {
            uint32_t v = ((c0 << 6) | (c1 << 4) | (c2 << 2) | c3) << (24u - 8u * (lane & 3u));
            v |= __shfl_xor_sync(0xffffffffu, v, 1);
            v |= __shfl_xor_sync(0xffffffffu, v, 2);
            if ((lane & 3) == 0) packed[blk * 8u + (lane >> 2)] = v;
        }
        if (qw) {
            uint32_t y1 = __shfl_down_sync(0xffffffffu, y0, 1);
            if (lane == 31) y1 = ye0;
            const uint32_t y = __funnelshift_r(y0, y1, 8u * qmis);
            const uint32_t vm = nvalid >= 4u ? 0xffffffffu : ((1u << (8u * nvalid)) - 1u);
            qdiff |= (y ^ q0x4) & vm;
        }
    }
    // guard words: one plane word read by the DP look-ahead, packed words read by window64
    if (lane == 0 && (nw & 3u) == 0u) planes[nw] = make_uint2(0u, 0u);
    if (packed && lane < 8) packed[nblk * 8u + lane] = 0u;
    return qp != nullptr && !__any_sync(0xffffffffu, qdiff != 0u);
}

// shared memory layout of the sketch kernel
struct SketchSmem {
    uint32_t table_bytes;    // s-mer rank table (0 when hashes are computed directly)
    uint32_t bloom_bytes;    // SNPmer Bloom filter (0 when the set is empty)
    uint32_t packed_words;   // per warp
    uint32_t ring_bytes;     // per warp: lane-private rank ring (runtime-c path only)
    uint32_t max_t;          // chunk length per lane (positions)
    uint32_t mini_cap;       // minimizers a lane can emit per chunk: max_t / spacing + 2
    __host__ __device__ size_t per_warp() const {
        const size_t b = (size_t)kSnpHitCap * 12u + 16u + ring_bytes + (size_t)mini_cap * 32u * 8u + (size_t)packed_words * 4u;
        return (b + 15u) & ~(size_t)15u;
    }
    __host__ __device__ size_t total() const { return 256u + (size_t)table_bytes + bloom_bytes + per_warp() * kSketchWarps; }
};

template <bool B> struct RankT { using type = uint64_t; };
template <> struct RankT<true> { using type = uint32_t; };

// K, C > 0: compile-time k-mer / window sizes (the default 17/11 path, window in registers);
// K == 0: runtime parameters (window in a lane-private shared-memory ring).
// USE_TABLE: s-mer hash ranks from the 4^s-entry shared-memory table instead of mm_hash64.
template <int K, int C, bool USE_TABLE>
__global__ void __launch_bounds__(kSketchWarps * 32)
sketch_kernel(SeqBatch in, SketchOut out, SnpSet set, const uint16_t *__restrict__ rank_table_g, DeviceParams p, SketchSmem lay) {
    using rank_t = typename RankT<USE_TABLE>::type;
    constexpr int CC = C > 0 ? C : 1;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t k = K ? (uint32_t)K : (uint32_t)p.k;
    const uint32_t c = K ? (uint32_t)C : (uint32_t)p.c;
    const uint32_t s = k - c + 1u;
    const uint64_t kmer_mask = (1ull << (2u * k)) - 1ull;
    const uint64_t km1_mask = (1ull << (2u * (k - 1u))) - 1ull;
    const uint64_t split_mask = ~(3ull << (k - 1u));
    const uint32_t smask = (s >= 16u) ? 0xffffffffu : ((1u << (2u * s)) - 1u);
    const uint32_t rc_shift = 2u * (k - 1u);

    uint8_t *s_lut = smem;                        // BYTE_TO_SEQ
    const uint16_t *s_rank = reinterpret_cast<const uint16_t *>(smem + 256);
    const uint32_t *s_bloom = reinterpret_cast<const uint32_t *>(smem + 256 + lay.table_bytes);
    {
        if (threadIdx.x < 256) s_lut[threadIdx.x] = (uint8_t)enc_base(threadIdx.x);
        uint32_t *dst = reinterpret_cast<uint32_t *>(smem + 256);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(rank_table_g);
        const uint32_t nt = lay.table_bytes / 4u, nb = lay.bloom_bytes / 4u;
        for (uint32_t i = threadIdx.x; i < nt; i += blockDim.x) dst[i] = __ldg(src + i);
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) dst[nt + i] = __ldg(set.bloom + i);
        __syncthreads();
    }
    uint8_t *wbase = smem + 256 + lay.table_bytes + lay.bloom_bytes + (size_t)warp * lay.per_warp();
    uint64_t *hits_k = reinterpret_cast<uint64_t *>(wbase);
    uint64_t *mbuf = hits_k + kSnpHitCap;                                           // [mini_cap][32]
    rank_t *ring = reinterpret_cast<rank_t *>(mbuf + (size_t)lay.mini_cap * 32u);   // [c][32] (runtime-c path)
    uint32_t *hits_p = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(ring) + lay.ring_bytes);
    uint32_t *hcnt = hits_p + kSnpHitCap;                                           // [0] hits
    uint32_t *packed = hcnt + 4;

    const uint32_t mid_idx = (c - 1u) >> 1;      // (k - s) / 2, src/seeding.rs:528
    const uint32_t half_k = k >> 1;
    // SNPmer hits beyond the shared-memory list spill to a per-warp global region (rare)
    uint64_t *spill_k = out.spill_kmer + (size_t)(blockIdx.x * kSketchWarps + warp) * out.spill_stride;
    uint32_t *spill_p = out.spill_pos + (size_t)(blockIdx.x * kSketchWarps + warp) * out.spill_stride;
    auto hit_k = [&](uint32_t h) -> uint64_t { return h < (uint32_t)kSnpHitCap ? hits_k[h] : spill_k[h - kSnpHitCap]; };
    auto hit_p = [&](uint32_t h) -> uint32_t { return h < (uint32_t)kSnpHitCap ? hits_p[h] : spill_p[h - kSnpHitCap]; };
    auto set_hit_p = [&](uint32_t h, uint32_t v) { if (h < (uint32_t)kSnpHitCap) hits_p[h] = v; else spill_p[h - kSnpHitCap] = v; };

    for (uint32_t seq = blockIdx.x * kSketchWarps + warp; seq < in.n; seq += gridDim.x * kSketchWarps) {
        const uint64_t off = in.offsets[seq];
        const uint32_t L = (uint32_t)(in.offsets[seq + 1] - off);
        uint2 *planes = out.planes + plane_word_off(off, seq);
        if (L < k) {                              // src/seeding.rs:339-341 -> None
            pack_sequence(in, off, L, planes, nullptr, s_lut, lane);   // planes still serve align_pairs
            if (lane == 0) {
                out.status[seq] = 4; out.flatq[seq] = 0; out.mini_cnt[seq] = 0;
                out.snp_start[seq] = 0; out.snp_cnt[seq] = 0;
            }
            continue;
        }
        __syncwarp();
        const bool flat = pack_sequence(in, off, L, planes, packed, s_lut, lane);
        if (lane == 0) hcnt[0] = 0u;
        __syncwarp();

        // exact SNPmer set probe + mid-base quality gate (src/seeding.rs:509-521) -> hit list
        auto resolve_snp = [&](const uint64_t canon, const uint32_t q) {
            uint32_t slot = hash_slot(canon, set.shift);
            while (true) {
                const uint64_t v = __ldg(set.keys + slot);
                if (v == canon) break;
                if (v == kEmpty64) return;
                slot = (slot + 1u) & set.mask;
            }
            int qv = 60;
            if (in.quals) qv = (int)in.quals[off + q + half_k] - 33;
            const bool gate = (qv > p.min_bq) || flat;                      // strict > (src/seeding.rs:517)
            const uint32_t o = atomicAdd(hcnt, 1u);
            const uint32_t pv = q | (gate ? 0x80000000u : 0u);
            if (o < (uint32_t)kSnpHitCap) { hits_k[o] = canon; hits_p[o] = pv; }
            else if (o - kSnpHitCap < out.spill_stride) { spill_k[o - kSnpHitCap] = canon; spill_p[o - kSnpHitCap] = pv; }
        };

        const uint32_t n_pos = L - k + 1u;
        uint64_t *mini_out = out.mini_kmer + mini_slot_off(off, seq, (uint32_t)p.mini_spacing);
        uint32_t *mini_pos_out = out.mini_pos ? out.mini_pos + mini_slot_off(off, seq, (uint32_t)p.mini_spacing) : nullptr;
        uint32_t n_mini = 0;

        // positions are processed in segments of at most 32 * max_t; inside a segment lane l owns
        // positions [p0, p1) and starts rolling c-1 positions earlier to refill the window
        const uint32_t seg_span = 32u * lay.max_t;
        const uint32_t n_seg = (n_pos + seg_span - 1u) / seg_span;
        const uint32_t seg_len = (n_pos + n_seg - 1u) / n_seg;
        for (uint32_t seg = 0; seg < n_seg; ++seg) {
            const uint32_t pstart = seg * seg_len;
            const uint32_t pend = min(n_pos, pstart + seg_len);
            const uint32_t T = (pend - pstart + 31u) >> 5;
            const uint32_t p0 = min(pstart + lane * T, pend);
            const uint32_t p1 = min(p0 + T, pend);
            const uint32_t qs = (p0 >= c - 1u) ? p0 - (c - 1u) : 0u;       // first position rolled by this lane
            const uint32_t ns = p1 - qs;                                    // steps of this lane
            const uint32_t NS = T + c - 1u;                                 // uniform loop bound
            // prime the registers with the k-1 bases before the first pushed base
            uint64_t F, Rv;
            {
                const uint64_t x = window64(packed, qs);
                F = x >> (64u - 2u * (k - 1u));
                Rv = revcomp_top(x, km1_mask) << 2;
            }
            uint32_t bidx = qs + k - 1u;                                    // next base to push
            uint32_t pw = K ? 0u : packed[bidx >> 4] << (2u * (bidx & 15u));
            uint32_t cnt = 0;                                               // minimizers buffered by this lane
            uint64_t cmask = 0;                                             // bit j: step j was Bloom-positive
            rank_t win[CC];
#pragma unroll
            for (int t = 0; t < CC; ++t) win[t] = 0;

            // one position.  `first` (warp-uniform): the step belongs to the first block of the segment, where
            // warm-up steps (q < p0), a not-yet-full window and (segment 0) the warm-up quirk can occur.
            // `e`: the base pushed by this step.  Returns whether the k-mer passed the SNPmer Bloom filter.
            auto step = [&](const bool first, const uint32_t j, const int jj, const uint32_t e) -> bool {
                const uint32_t q = qs + j;
                F = ((F << 2) | (uint64_t)e) & kmer_mask;
                Rv = (Rv >> 2) | ((uint64_t)(3u - e) << rc_shift);
                const bool fwd = (F & split_mask) < (Rv & split_mask);      // tie -> reverse (src/seeding.rs:429)
                const uint64_t canon = fwd ? F : Rv;
                // the s-mer ending at base q+k-1 is the tail of this k-mer ...
                uint32_t f = (uint32_t)F & smask;
                if (first && seg == 0u && q + 1u < s) {
                    // ... except for the reference's warm-up quirk (src/seeding.rs:392-397,413)
                    const uint32_t nt = q + 1u;
                    const uint32_t head = (uint32_t)(F >> (2u * (k - (s - nt))));
                    f = (head << (2u * nt)) | ((uint32_t)F & ((1u << (2u * nt)) - 1u));
                }
                rank_t rk;
                if (USE_TABLE) {
                    rk = (rank_t)s_rank[f];
                } else {
                    const uint32_t r = revcomp_smer(f, s);
                    rk = (rank_t)mm_hash64((uint64_t)(f < r ? f : r));
                }
                if (K) win[jj] = rk; else ring[(j % c) * 32u + lane] = rk;
                if (first && q < p0) return false;                          // warm-up step: roll and rank only
                // ---- SNPmer membership: Bloom filter now (branch-free), exact probe after the loop ----
                bool bloom_hit = false;
                if (set.keys != nullptr) {
                    const uint32_t h = bloom_hash32_sym((uint32_t)F, (uint32_t)Rv);   // strand symmetric key (shared with sketch_fast.cu)
                    const uint32_t bmask = (1u << (h & 31u)) | (1u << ((h >> 5) & 31u));
                    bloom_hit = (s_bloom[h >> set.bloom_shift] & bmask) == bmask;
                }
                // ---- open syncmer: strict minimum at the middle of the last c ranks ----
                if (!first || j + 1u >= c) {
                    bool is_mini;
                    if (K) {
                        const int mslot = (jj + 1 + ((CC - 1) >> 1)) % CC;
                        rank_t others = ~(rank_t)0;
#pragma unroll
                        for (int t = 0; t < CC; ++t) if (t != mslot) others = min(others, win[t]);
                        is_mini = win[mslot] < others || CC == 1;
                    } else {
                        const uint32_t mslot = (j + 1u + mid_idx) % c;
                        const rank_t mid = ring[mslot * 32u + lane];
                        rank_t others = ~(rank_t)0;
                        for (uint32_t t = 0; t < c; ++t) {
                            const rank_t v = ring[t * 32u + lane];
                            if (t != mslot) others = min(others, v);
                        }
                        is_mini = mid < others || c == 1u;
                    }
                    if (is_mini) {
                        mbuf[cnt * 32u + lane] = canon | (fwd ? kFwdBit : 0ull) | ((uint64_t)(q - p0) << 48);
                        ++cnt;
                    }
                }
                return bloom_hit;
            };

            if (K) {
                // blocks of CC steps: the CC bases of a block are fetched at once, the window slot of a step is static
                for (uint32_t jb = 0; jb < NS; jb += (uint32_t)CC) {
                    const uint32_t bb = extract32(packed, bidx, (uint32_t)CC);
                    bidx += (uint32_t)CC;
                    uint32_t cm = 0;
#pragma unroll
                    for (int jj = 0; jj < CC; ++jj) {
                        const uint32_t j = jb + (uint32_t)jj;
                        if (j < ns && step(jb == 0u, j, jj, (bb >> (2 * (CC - 1 - jj))) & 3u)) cm |= 1u << jj;
                    }
                    cmask |= (uint64_t)cm << jb;
                }
            } else {
                for (uint32_t j = 0; j < NS; ++j) {
                    if (j < ns) {
                        const uint32_t e = pw >> 30;
                        pw <<= 2;
                        ++bidx;
                        if ((bidx & 15u) == 0u) pw = packed[bidx >> 4];
                        if (step(j < max(c, s), j, 0, e)) cmask |= 1ull << j;   // quirk region (q < s-1) may exceed c steps
                    }
                }
            }
            // ---- write the buffered minimizers lane-major (= position order) ----
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t o0 = n_mini + incl - cnt;
            for (uint32_t i = 0; i < cnt; ++i) {
                const uint64_t v = mbuf[i * 32u + lane];
                mini_out[o0 + i] = v & ~(0x3full << 48);
                if (mini_pos_out) mini_pos_out[o0 + i] = p0 + (uint32_t)((v >> 48) & 0x3fu);
            }
            n_mini += total;
            // ---- exact probe of this lane's Bloom-positive positions (k-mer re-extracted from the packed read) ----
            while (cmask) {
                const uint32_t j = (uint32_t)__ffsll((long long)cmask) - 1u;
                cmask &= cmask - 1ull;
                const uint32_t q = qs + j;
                const uint64_t x = window64(packed, q);
                const uint64_t kf = x >> (64u - 2u * k);
                const uint64_t kr = revcomp_top(x, kmer_mask);
                resolve_snp(((kf & split_mask) < (kr & split_mask)) ? kf : kr, q);
            }
            __syncwarp();
        }


        // ---- SNPmer dedup: keep gated hits whose splitmer occurs exactly once among ALL hits ----
        uint32_t nh = hcnt[0];
        if (nh > (uint32_t)kSnpHitCap + out.spill_stride) nh = kSnpHitCap + out.spill_stride;   // cannot happen: stride > max_len
        uint32_t kept_total = 0;
        for (uint32_t base = 0; base < nh; base += 32u) {
            const uint32_t h = base + lane;
            bool keep = false;
            if (h < nh) {
                const uint64_t sp = hit_k(h) & split_mask;
                uint32_t dup = 0;
                for (uint32_t j = 0; j < nh; ++j) dup += ((hit_k(j) & split_mask) == sp) ? 1u : 0u;
                keep = (hit_p(h) >> 31) && dup == 1u;
            }
            const uint32_t b = __ballot_sync(0xffffffffu, keep);
            __syncwarp();
            if (h < nh) set_hit_p(h, (hit_p(h) & 0x7fffffffu) | (keep ? 0x80000000u : 0u));
            kept_total += __popc(b);
        }
        __syncwarp();
        uint32_t start = 0;
        if (lane == 0 && kept_total) start = atomicAdd(out.snp_pool_ctr, kept_total);
        start = __shfl_sync(0xffffffffu, start, 0);
        if (kept_total && start + kept_total > out.snp_pool_cap) {
            if (lane == 0) atomicOr(out.err, kErrSnpPool);
            kept_total = 0;
        }
        if (kept_total) {
            // hits were appended in arbitrary order: the output slot of a kept hit is its position rank
            for (uint32_t base = 0; base < nh; base += 32u) {
                const uint32_t h = base + lane;
                if (h < nh && (hit_p(h) >> 31)) {
                    const uint32_t ph = hit_p(h) & 0x7fffffffu;
                    uint32_t rank = 0;
                    for (uint32_t j = 0; j < nh; ++j) {
                        const uint32_t pj = hit_p(j);
                        rank += ((pj >> 31) && (pj & 0x7fffffffu) < ph) ? 1u : 0u;
                    }
                    out.snp_kmer[start + rank] = hit_k(h);
                    out.snp_pos[start + rank] = ph;
                }
            }
        }
        if (lane == 0) {
            out.status[seq] = 0;
            out.flatq[seq] = flat ? 1 : 0;
            out.mini_cnt[seq] = n_mini;
            out.snp_start[seq] = start;
            out.snp_cnt[seq] = kept_total;
        }
        __syncwarp();
    }
}

static SketchSmem sketch_layout(uint32_t max_len, bool use_table, const DeviceParams &p, const SnpSet &set) {
    SketchSmem lay{};
    const bool fixed = (p.k == 17 && p.c == 11 && use_table);
    lay.table_bytes = use_table ? (uint32_t)((((size_t)1 << (2 * p.s)) * 2u + 15u) & ~(size_t)15u) : 0u;
    lay.bloom_bytes = set.keys ? set.bloom_words * 4u : 0u;
    lay.packed_words = (max_len >> 4) + 20u;
    lay.ring_bytes = fixed ? 0u : (uint32_t)((size_t)p.c * 32u * (use_table ? 4u : 8u));
    // chunk length: at most 64 positions per lane, and small enough that a lane emits <= 12 minimizers
    uint32_t max_t = 10u * (uint32_t)p.mini_spacing;
    if (max_t > 65u - (uint32_t)p.c) max_t = 65u - (uint32_t)p.c;       // T + c - 1 <= 64: one Bloom-mask bit per step
    lay.max_t = max_t;
    // the longest chunk any lane can get for this batch
    const uint32_t n_pos = max_len >= (uint32_t)p.k ? max_len - (uint32_t)p.k + 1u : 1u;
    const uint32_t n_seg = (n_pos + 32u * max_t - 1u) / (32u * max_t);
    const uint32_t seg_len = (n_pos + n_seg - 1u) / n_seg;
    const uint32_t t_eff = (seg_len + 31u) / 32u;
    lay.mini_cap = t_eff / (uint32_t)p.mini_spacing + 2u;
    return lay;
}

// grid used by launch_sketch (the host sizes the per-warp spill region with it)
uint32_t sketch_grid(uint32_t n, uint32_t max_len, bool use_table, const DeviceParams &p, const SnpSet &set, int sm_count) {
    const SketchSmem lay = sketch_layout(max_len, use_table, p, set);
    int ctas_per_sm = (int)(220u * 1024u / (lay.total() + 1024u));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 4) ctas_per_sm = 4;          // 4 x 16 warps = 64 resident warps
    uint32_t grid = (uint32_t)(sm_count * ctas_per_sm);
    const uint32_t need = (n + kSketchWarps - 1) / kSketchWarps;
    if (grid > need) grid = need;
    return grid ? grid : 1u;
}

void launch_sketch(const SeqBatch &in, const SketchOut &out, const SnpSet &set, const uint16_t *rank_table,
                   const DeviceParams &p, int sm_count, cudaStream_t st) {
    if (in.n == 0) return;
    const bool use_table = rank_table != nullptr;
    const SketchSmem lay = sketch_layout(in.max_len, use_table, p, set);
    const size_t smem = lay.total();
    const uint32_t grid = sketch_grid(in.n, in.max_len, use_table, p, set, sm_count);
    if (p.k == 17 && p.c == 11 && use_table) {
        cudaFuncSetAttribute(sketch_kernel<17, 11, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        sketch_kernel<17, 11, true><<<grid, kSketchWarps * 32, smem, st>>>(in, out, set, rank_table, p, lay);
    } else if (use_table) {
        cudaFuncSetAttribute(sketch_kernel<0, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        sketch_kernel<0, 0, true><<<grid, kSketchWarps * 32, smem, st>>>(in, out, set, rank_table, p, lay);
    } else {
        cudaFuncSetAttribute(sketch_kernel<0, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        sketch_kernel<0, 0, false><<<grid, kSketchWarps * 32, smem, st>>>(in, out, set, rank_table, p, lay);
    }
}

// planes only (align_pairs entry point: reads that were never sketched)
__global__ void pack_planes_kernel(SeqBatch in, uint2 *planes_all) {
    __shared__ uint8_t s_lut[256];
    if (threadIdx.x < 256) s_lut[threadIdx.x] = (uint8_t)enc_base(threadIdx.x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t seq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; seq < in.n; seq += warps) {
        const uint64_t off = in.offsets[seq];
        const uint32_t L = (uint32_t)(in.offsets[seq + 1] - off);
        SeqBatch nq = in; nq.quals = nullptr;
        pack_sequence(nq, off, L, planes_all + plane_word_off(off, seq), nullptr, s_lut, lane);
    }
}
void launch_pack_planes(const SeqBatch &in, uint2 *planes, cudaStream_t st) {
    if (in.n == 0) return;
    uint32_t grid = (in.n + 7u) / 8u;
    if (grid > 148u * 8u) grid = 148u * 8u;
    pack_planes_kernel<<<grid, 256, 0, st>>>(in, planes);
}

// est_id (src/seeding.rs:801-817): 100 - 100 * mean(10^(-q/10)), a strictly sequential f64 sum in
// read order so that the result is bit-identical to the reference; one thread per read.  The
// table holds pow(10, -(b-33)/10) for every byte b, computed on the host with libm.
__global__ void est_id_kernel(SeqBatch in, const uint8_t *status, const uint8_t *flatq, const double *__restrict__ table,
                              double *est_id) {
    __shared__ double s_table[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_table[i] = table[i];
    __syncthreads();
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= in.n) return;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    if (in.quals == nullptr || status[r] != 0 || flatq[r]) { est_id[r] = nan; return; }
    const uint64_t off = in.offsets[r];
    const uint32_t L = (uint32_t)(in.offsets[r + 1] - off);
    const uint8_t *q = in.quals + off;
    double sum = 0.0;
    uint32_t i = 0;
    while (i < L && ((uintptr_t)(q + i) & 15u)) { sum = __dadd_rn(sum, s_table[q[i]]); ++i; }
    for (; i + 16u <= L; i += 16u) {
        const uint4 v = *reinterpret_cast<const uint4 *>(q + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            sum = __dadd_rn(sum, s_table[w[t] & 0xffu]);
            sum = __dadd_rn(sum, s_table[(w[t] >> 8) & 0xffu]);
            sum = __dadd_rn(sum, s_table[(w[t] >> 16) & 0xffu]);
            sum = __dadd_rn(sum, s_table[w[t] >> 24]);
        }
    }
    for (; i < L; ++i) sum = __dadd_rn(sum, s_table[q[i]]);
    // 100. - (sum / count * 100.)  -- no FMA contraction
    est_id[r] = __dsub_rn(100.0, __dmul_rn(__ddiv_rn(sum, (double)L), 100.0));
}
void launch_est_id(const SeqBatch &in, const uint8_t *status, const uint8_t *flatq, const double *phred_table, double *est_id,
                   cudaStream_t st) {
    if (in.n == 0) return;
    est_id_kernel<<<(in.n + 127u) / 128u, 128, 0, st>>>(in, status, flatq, phred_table, est_id);
}

// one 2-bit code per base (optional `codes` output of savont_b200_sketch)
__global__ void codes_kernel(const uint8_t *bases, uint8_t *codes, uint64_t total) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
        codes[i] = (uint8_t)enc_base(bases[i]);
}
void launch_codes(const uint8_t *bases, uint8_t *codes, uint64_t total, cudaStream_t st) {
    if (total == 0) return;
    codes_kernel<<<148 * 8, 256, 0, st>>>(bases, codes, total);
}

// compact CSR copies of the sketch (C-ABI savont_b200_sketch): one warp per sequence
__global__ void gather_sketch_kernel(SeqBatch in, SketchOut sk, const uint32_t *mini_off, const uint32_t *snp_off,
                                     uint32_t *mini_pos, uint64_t *mini_kmer, uint8_t *mini_fwd, uint32_t *snp_pos,
                                     uint64_t *snp_kmer, uint32_t mini_spacing) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t seq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; seq < in.n; seq += warps) {
        const uint64_t mo = mini_slot_off(in.offsets[seq], seq, mini_spacing);
        const uint32_t nm = sk.mini_cnt[seq], d0 = mini_off[seq];
        for (uint32_t t = lane; t < nm; t += 32u) {
            const uint64_t v = sk.mini_kmer[mo + t];
            mini_kmer[d0 + t] = v & ~kFwdBit;
            mini_fwd[d0 + t] = (uint8_t)(v >> 63);
            mini_pos[d0 + t] = sk.mini_pos[mo + t];
        }
        const uint32_t ns = sk.snp_cnt[seq], s0 = sk.snp_start[seq], e0 = snp_off[seq];
        for (uint32_t t = lane; t < ns; t += 32u) {
            snp_kmer[e0 + t] = sk.snp_kmer[s0 + t];
            snp_pos[e0 + t] = sk.snp_pos[s0 + t];
        }
    }
}
void launch_gather_sketch(const SeqBatch &in, const SketchOut &sk, const uint32_t *mini_off, const uint32_t *snp_off,
                          uint32_t *mini_pos, uint64_t *mini_kmer, uint8_t *mini_fwd, uint32_t *snp_pos, uint64_t *snp_kmer,
                          const DeviceParams &p, cudaStream_t st) {
    if (in.n == 0) return;
    uint32_t grid = (in.n + 7u) / 8u;
    if (grid > 148u * 8u) grid = 148u * 8u;
    gather_sketch_kernel<<<grid, 256, 0, st>>>(in, sk, mini_off, snp_off, mini_pos, mini_kmer, mini_fwd, snp_pos, snp_kmer,
                                               (uint32_t)p.mini_spacing);
}

}  // namespace sb
