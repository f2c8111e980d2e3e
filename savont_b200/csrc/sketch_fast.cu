// sketch_fast.cu -- K1 for the default parameters (k = 17, c = 11, s = 7): 2-bit packing + open-syncmer /
// SNPmer sketch, one warp per sequence, 32 warps per CTA, one persistent CTA per SM.
//
// Same contract as sketch_kernel (sketch.cu; reference seeding::get_twin_read_syncmer, src/seeding.rs:317-658)
// with two input encodings:
//   SRC_ASCII   ASCII bases (+ Phred+33 qualities): the FASTQ/FASTA side (src/kmer_comp.rs:39-66, :68-258)
//   SRC_TWIN    TwinRead.dna_seq as stored by the reference (bio-seq Seq<Dna> raw words, base i of a read in
//               bits 2(i%32).. of word i/32, first base in the LOW bits, src/types.rs:575-579); the read's SNPmers
//               are handed over by the caller (TwinRead.snpmer_kmers), so only the minimizers are recomputed.
//
// Differences from the generic kernel that make it ~2.5x cheaper per base:
//   phase A  16 bases per lane and iteration (one 128-bit load): an 8-byte LUT entry per base byte yields the
//            2-bit code and its two plane bits at once; packed words and ALIGNED plane words are staged in
//            shared memory and the misalignment of the sequence start is removed once per 32-base plane word.
//            The all-equal-quality scan (src/seeding.rs:372-380) stops at the first differing chunk.
//   phase B  steady-state loop without any boundary logic: per position only the two rolling registers, the
//            s-mer rank lookup, an 11-wide register window (min tree) and ONE Bloom probe keyed by F + revcomp(F)
//            (strand symmetric, so the canonical choice is not needed per position).  Decisions are collected
//            as bit masks; canonical k-mers are re-extracted from the packed read only for the ~1/11 selected
//            positions, balanced over the lanes and written coalesced.  The reference's s-mer warm-up quirk
//            (SURVEY.md 8a note 1: k-mer positions 0..9 untested, 10..15 decided on non-contiguous s-mers) is
//            patched in by a short warp-cooperative fix-up instead of living in the loop.
//   phase C  SNPmer hits arrive in position order (ballot compaction), so no sort; the splitmer multiplicity
//            (src/seeding.rs:550-568) is counted in a small shared-memory hash table.
// Anything that does not fit the shared-memory budgets (more SNPmer hits / Bloom positives than the per-warp
// lists hold) raises kErrSnpHits and the host re-runs the batch with the generic kernel.
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "sketch_dev.cuh"

namespace sb {

constexpr int kFastWarps = 32;
constexpr uint32_t kFastK = 17, kFastC = 11, kFastS = 7;
constexpr uint32_t kFastMaxT = 54;          // positions per lane and segment: 5 blocks of 11 steps <= 64 mask bits
constexpr uint32_t kFastListCap = 384;      // positions per warp and segment (minimizers <= 1728/6+2; Bloom positives)

struct FastSmem {
    uint32_t bloom_bytes;    // multiple of 16
    uint32_t packed_words;   // per warp
    uint32_t stage_bytes;    // per warp: aligned plane staging / position list / SNPmer dedup table (aliased)
    uint32_t hit_cap;        // per warp, power of two
    __host__ __device__ size_t per_warp() const { return (size_t)packed_words * 4u + stage_bytes + (size_t)hit_cap * 12u + 16u; }
    uint32_t warps;          // warps per CTA (32 unless long reads need more shared memory per warp)
    __host__ __device__ size_t total() const { return 32768u + 2048u + bloom_bytes + per_warp() * warps; }
};

// 32 -> 16: keep the even bits of x
__device__ __forceinline__ uint32_t compress_even(uint32_t x) {
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    x = (x | (x >> 8)) & 0x0000ffffu;
    return x;
}
// bio-seq word half (base t in bits 2t..2t+1) -> packed word (base 0 in the top bit pair)
__device__ __forceinline__ uint32_t bioseq_to_packed(uint32_t v) {
    const uint32_t r = __brev(v);
    return ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
}
__device__ __forceinline__ uint32_t revcomp16(uint32_t f) {       // 16 bases, first base in the top pair
    const uint32_t r = __brev(~f);
    return ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
}

template <int SRC, bool SNP>
__global__ void __launch_bounds__(kFastWarps * 32, 1)
sketch_fast_kernel(SeqBatch in, SketchOut out, SnpSet set, const uint16_t *__restrict__ rank_table_g, DeviceParams p, FastSmem lay) {
    constexpr uint32_t K = kFastK, C = kFastC, S = kFastS;
    constexpr uint64_t kmer_mask = (1ull << (2u * K)) - 1ull;
    constexpr uint64_t split_mask = ~(3ull << (K - 1u));
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;

    const uint16_t *s_rank = reinterpret_cast<const uint16_t *>(smem);                 // 4^7 x u16
    uint2 *s_lut = reinterpret_cast<uint2 *>(smem + 32768);                            // 256 x (code, plane bits)
    const uint32_t *s_bloom = reinterpret_cast<const uint32_t *>(smem + 32768 + 2048);
    {
        uint32_t *dst = reinterpret_cast<uint32_t *>(smem);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(rank_table_g);
        for (uint32_t i = threadIdx.x; i < 8192u; i += blockDim.x) dst[i] = __ldg(src + i);
        if (threadIdx.x < 256) {
            const uint32_t c = enc_base(threadIdx.x);
            s_lut[threadIdx.x] = make_uint2(c, (c & 1u) | ((c >> 1) << 16));
        }
        uint32_t *bd = reinterpret_cast<uint32_t *>(smem + 32768 + 2048);
        const uint32_t nb = lay.bloom_bytes / 4u;
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) bd[i] = (set.keys != nullptr && i < set.bloom_words) ? __ldg(set.bloom + i) : 0u;
        __syncthreads();
    }
    uint8_t *wbase = smem + 32768 + 2048 + lay.bloom_bytes + (size_t)warp * lay.per_warp();
    uint32_t *packed = reinterpret_cast<uint32_t *>(wbase);
    uint8_t *stage = wbase + (size_t)lay.packed_words * 4u;
    uint2 *astage = reinterpret_cast<uint2 *>(stage);                                  // aligned plane words (phase A)
    uint16_t *plist = reinterpret_cast<uint16_t *>(stage);                             // selected positions (phase B)
    unsigned long long *dtab = reinterpret_cast<unsigned long long *>(stage);          // splitmer -> count (phase C)
    unsigned long long *hits_k = reinterpret_cast<unsigned long long *>(stage + lay.stage_bytes);
    uint32_t *hits_p = reinterpret_cast<uint32_t *>(hits_k + lay.hit_cap);
    constexpr bool want_snp = (SRC == 0) && SNP;

    // sequences are handed out dynamically: the kernel makes progress at full efficiency with however many of its CTAs are
    // resident (it shares the SMs with the DP kernel of the previous part, api.cu map_resident)
    while (true) {
        uint32_t seq = 0;
        if (lane == 0) seq = atomicAdd(out.work_ctr, 1u);
        seq = __shfl_sync(0xffffffffu, seq, 0);
        if (seq >= in.n) break;
        const uint64_t off = in.offsets[seq];
        const uint32_t L = (uint32_t)(in.offsets[seq + 1] - off);
        uint2 *planes = out.planes + plane_word_off(off, seq);
        const uint32_t nw = (L + 31u) >> 5;                 // plane words; word nw is the zero guard word of the DP look-ahead
        uint32_t pb = 0;                                    // base index of the read's first base inside `packed`
        bool flat = false;
        __syncwarp();

        // ------------------------------------------------------------------ phase A
        if (SRC == 0) {
            const uint8_t *bp = in.bases + off;
            const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(bp) & 15u);
            pb = mis;
            const uint4 *b4 = reinterpret_cast<const uint4 *>(bp - mis);
            const uint32_t nchunk = (L + mis + 15u) >> 4;
            for (uint32_t c0 = 0; c0 < 2u * nw + 8u; c0 += 32u) {       // covers the zero guard words behind the read too
                const uint32_t ch = c0 + lane;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (ch < nchunk) v = __ldg(b4 + ch);
                const uint32_t xs[4] = {v.x, v.y, v.z, v.w};
                uint32_t pk = 0, pl = 0;
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const uint32_t byte = (xs[t >> 2] >> (8 * (t & 3))) & 0xffu;
                    const uint2 e = s_lut[byte];
                    pk = mad_lo(e.x, 1u << (30 - 2 * t), pk);                 // disjoint bit fields: OR as multiply-add
                    pl = mad_lo(e.y, 1u << t, pl);
                }
                // bases outside [0, L) (before the misaligned start, after the end) are zero
                const int first = (int)mis - (int)(ch * 16u);               // chunk positions < first are invalid
                const int last = (int)(L + mis) - (int)(ch * 16u);          // chunk positions >= last are invalid
                uint32_t keep = 0xffffu;
                if (first > 0) keep &= first >= 16 ? 0u : (0xffffu << first);
                if (last < 16) keep &= last <= 0 ? 0u : (0xffffu >> (16 - last));
                if (ch >= nchunk) keep = 0u;
                pl &= keep | (keep << 16);
                uint32_t keep2 = 0u;                                        // 2 bits per position, position 0 in the top pair
                if (keep == 0xffffu) keep2 = 0xffffffffu;
                else if (keep != 0u) {
                    const int lo_t = first > 0 ? first : 0, hi_t = last < 16 ? last : 16;
                    keep2 = (0xffffffffu >> (2 * lo_t)) & ~(hi_t >= 16 ? 0u : (0xffffffffu >> (2 * hi_t)));
                }
                pk &= keep2;
                if (ch < lay.packed_words) packed[ch] = pk;
                // two lanes -> one aligned 32-base plane word
                const uint32_t other = __shfl_xor_sync(0xffffffffu, pl, 1);
                if ((lane & 1) == 0 && (ch >> 1) <= nw + 1u) {
                    const uint32_t lo = (pl & 0xffffu) | (other << 16);
                    const uint32_t hi = (pl >> 16) | (other & 0xffff0000u);
                    astage[ch >> 1] = make_uint2(lo, hi);
                }
            }
            __syncwarp();
            // plane word w = aligned bits [32w + mis, 32w + mis + 32)
            for (uint32_t w = lane; w <= nw; w += 32u) {
                const uint2 a = astage[w], b = astage[w + 1];
                planes[w] = make_uint2(__funnelshift_r(a.x, b.x, mis), __funnelshift_r(a.y, b.y, mis));
            }
            // all-equal-quality scan, stops at the first chunk with a difference
            if (in.quals != nullptr) {
                const uint8_t *qp = in.quals + off;
                const uint32_t qmis = (uint32_t)(reinterpret_cast<uintptr_t>(qp) & 15u);
                const uint4 *q4 = reinterpret_cast<const uint4 *>(qp - qmis);
                const uint32_t q0x4 = (uint32_t)qp[0] * 0x01010101u;
                const uint32_t qchunk = (L + qmis + 15u) >> 4;
                bool differ = false;
                for (uint32_t c0 = 0; c0 < qchunk && !differ; c0 += 32u) {
                    const uint32_t ch = c0 + lane;
                    uint32_t d = 0;
                    if (ch < qchunk) {
                        const uint4 v = __ldg(q4 + ch);
                        const uint32_t xs[4] = {v.x ^ q0x4, v.y ^ q0x4, v.z ^ q0x4, v.w ^ q0x4};
                        const int first = (int)qmis - (int)(ch * 16u), last = (int)(L + qmis) - (int)(ch * 16u);
#pragma unroll
                        for (int t4 = 0; t4 < 4; ++t4) {
                            uint32_t m = 0xffffffffu;                                   // byte t valid iff first <= t < last
                            const int f = first - 4 * t4, l = last - 4 * t4;
                            if (f > 0) m &= f >= 4 ? 0u : (0xffffffffu << (8 * f));
                            if (l < 4) m &= l <= 0 ? 0u : (0xffffffffu >> (8 * (4 - l)));
                            d |= xs[t4] & m;
                        }
                    }
                    differ = __any_sync(0xffffffffu, d != 0u);
                }
                flat = !differ;
            }
            __syncwarp();
        } else {
            const uint64_t w0 = in.word_offsets[seq];
            const uint32_t nwords = (L + 31u) >> 5;
            for (uint32_t j0 = 0; j0 < nwords + 8u; j0 += 32u) {
                const uint32_t j = j0 + lane;
                uint64_t v = 0;
                if (j < nwords) {
                    v = __ldg(reinterpret_cast<const unsigned long long *>(in.dna_words) + w0 + j);
                    const uint32_t rem = L - 32u * j;                                   // bases of this word that exist
                    if (rem < 32u) v &= (1ull << (2u * rem)) - 1ull;
                }
                const uint32_t vl = (uint32_t)v, vh = (uint32_t)(v >> 32);
                if (2u * j + 1u < lay.packed_words) { packed[2u * j] = bioseq_to_packed(vl); packed[2u * j + 1u] = bioseq_to_packed(vh); }
                if (j <= nw) {
                    const uint32_t lo = compress_even(vl) | (compress_even(vh) << 16);
                    const uint32_t hi = compress_even(vl >> 1) | (compress_even(vh >> 1) << 16);
                    planes[j] = make_uint2(lo, hi);
                }
            }
            __syncwarp();
        }

        if (L < K) {                                        // src/seeding.rs:339-341 -> None
            if (lane == 0) {
                out.status[seq] = 4; out.flatq[seq] = 0; out.mini_cnt[seq] = 0;
                if (SRC == 0) { out.snp_start[seq] = 0; out.snp_cnt[seq] = 0; }
            }
            continue;
        }

        const uint32_t n_pos = L - K + 1u;
        uint64_t *mini_out = out.mini_kmer + mini_slot_off(off, seq, (uint32_t)p.mini_spacing);
        uint32_t *mini_pos_out = out.mini_pos ? out.mini_pos + mini_slot_off(off, seq, (uint32_t)p.mini_spacing) : nullptr;
        uint32_t n_mini = 0, nh = 0;
        bool overflow = false;

        const uint32_t seg_span = 32u * kFastMaxT;
        const uint32_t n_seg = (n_pos + seg_span - 1u) / seg_span;
        const uint32_t seg_len = (n_pos + n_seg - 1u) / n_seg;
        for (uint32_t seg = 0; seg < n_seg; ++seg) {
            const uint32_t pstart = seg * seg_len;
            const uint32_t pend = min(n_pos, pstart + seg_len);
            const uint32_t T = (pend - pstart + 31u) >> 5;
            const uint32_t p0 = min(pstart + lane * T, pend);
            const uint32_t p1 = min(p0 + T, pend);
            const uint32_t ns = p1 - p0;

            // ---- registers after bases p0..p0+15; ranks of the s-mers starting at p0..p0+9 ----
            const uint64_t x0 = window64(packed, pb + p0);
            uint32_t Flo = (uint32_t)(x0 >> 32);
            uint32_t Rlo, Rhi;
            { const uint32_t rc = revcomp16(Flo); Rlo = rc << 2; Rhi = rc >> 30; }
            uint32_t win[C];
#pragma unroll
            for (int t = 0; t < (int)C - 1; ++t) win[t] = s_rank[(uint32_t)(x0 >> (64 - 2 * (t + (int)S))) & 0x3fffu];
            win[C - 1] = 0u;
            // sliding minima for the open-syncmer test: m3[i] = min(rank i..i+2), m5[i] = min(rank i..i+4), both indexed modulo C like
            // `win`.  The k-mer whose middle s-mer is q is a syncmer iff rank q < min(m5[q-5], m5[q+1]): per position two 3-input
            // minima extend m3 / m5 and one 2-input minimum + compare decide, instead of a 10-input minimum over the window.
            uint32_t m3[C], m5[C];
#pragma unroll
            for (int t = 0; t < (int)C; ++t) { m3[t] = 0u; m5[t] = 0u; }
#pragma unroll
            for (int t = 0; t + 2 < (int)C - 1; ++t) m3[t] = min(min(win[t], win[t + 1]), win[t + 2]);
#pragma unroll
            for (int t = 0; t + 4 < (int)C - 1; ++t) m5[t] = min(min(m3[t], win[t + 3]), win[t + 4]);

            uint64_t cm = 0, mm = 0;
            uint32_t bidx = pb + p0 + K - 1u;
            // one block = C consecutive positions (the register window rotates with the unrolled step index).  The last block of
            // a chunk is usually partial (T = 47 positions = 4 blocks + 3 steps for 1.5 kb reads): it leaves the unrolled body early
            // instead of computing up to C - 1 positions nobody reads
            auto block = [&](auto partial_tag, const uint32_t b, const uint32_t steps) {
                constexpr bool PARTIAL = decltype(partial_tag)::value;
                // (the word index is clamped to the warp's packed words: compute-sanitizer saw this load leave the CTA's shared memory once
                // per launch in the release build -- its result is never used there, the tests are bit-exact -- and the clamp costs one
                // instruction per 11 positions)
                const uint32_t bwi = min(bidx >> 4, lay.packed_words - 2u);
                const uint32_t bb = (__funnelshift_l(packed[bwi + 1u], packed[bwi], (bidx & 15u) * 2u) >> (32u - 2u * C)) << (32u - 2u * C);      // the block's bases, first one on top
                bidx += C;
                uint32_t cmb = 0, mmb = 0;
#pragma unroll
                for (int jj = 0; jj < (int)C; ++jj) {
                    if (PARTIAL && (uint32_t)jj >= steps) break;
                    const uint32_t e = (bb >> (30 - 2 * jj)) & 3u;
                    Flo = mad_lo(Flo, 4u, e);                                  // (Flo << 2) | e
                    Rlo = __funnelshift_r(Rlo, Rhi, 2);
                    Rhi = e ^ 3u;
                    win[(jj + (int)C - 1) % (int)C] = s_rank[Flo & 0x3fffu];
                    if (want_snp) {
                        // Bloom probe on low(F) + low(revcomp(F)) (identical for both strands); the funnel shifts take
                        // their shift amount modulo 32, so the two bit indices need no masking
                        const uint32_t h = bloom_hash32_sym(Flo, Rlo);
                        const uint32_t w = s_bloom[h >> set.bloom_shift];
                        const uint32_t hit = __funnelshift_r(w, 0u, h) & __funnelshift_r(w, 0u, h >> 5) & 1u;
                        cmb = mad_lo(hit, 1u << jj, cmb);
                    }
                    // open syncmer: strict minimum at the middle of the 11 ranks (src/seeding.rs:526-543); tau = index of the newest rank
                    constexpr int CI = (int)C;
                    static_assert(CI == 11, "the sliding minima below are laid out for c = 11 (windows of 5 on either side of the middle)");
                    const int tau = (jj + CI - 1) % CI;
                    m3[(tau + CI - 2) % CI] = min(min(win[(tau + CI - 2) % CI], win[(tau + CI - 1) % CI]), win[tau]);
                    m5[(tau + CI - 4) % CI] = min(min(m3[(tau + CI - 4) % CI], win[(tau + CI - 1) % CI]), win[tau]);
                    const uint32_t others = min(m5[(tau + CI - 10) % CI], m5[(tau + CI - 4) % CI]);
                    if (win[(tau + CI - 5) % CI] < others) mmb |= 1u << jj;
                }
                cm |= (uint64_t)cmb << (C * b);
                mm |= (uint64_t)mmb << (C * b);
            };
            const uint32_t full_blocks = T / C, rest = T - full_blocks * C;
            for (uint32_t b = 0; b < full_blocks; ++b) block(std::false_type{}, b, C);
            if (rest) block(std::true_type{}, full_blocks, rest);
            const uint64_t valid = ns >= 64u ? ~0ull : ((1ull << ns) - 1ull);
            mm &= valid; cm &= valid;

            if (seg == 0u) {
                // reference quirk: the window is full from k-mer position 10 on, and the s-mers pushed for k-mer
                // positions q < 6 are bases q..5 followed by bases 16..16+q (src/seeding.rs:392-397, :413)
                uint32_t rk = 0xffffu;
                if (lane < 16) {
                    const uint32_t x = 10u + lane;
                    if (lane < 6) {
                        if (lane < n_pos) {
                            const uint32_t head = extract32(packed, pb + lane, 6u - lane);
                            const uint32_t tail = extract32(packed, pb + 16u, lane + 1u);
                            rk = s_rank[(head << (2u * (lane + 1u))) | tail];
                        }
                    } else if (x + S <= L) {
                        rk = s_rank[extract32(packed, pb + x, S)];
                    }
                }
                uint32_t others = 0xffffffffu, mid = 0;
#pragma unroll
                for (int t = 0; t < (int)C; ++t) {
                    const uint32_t v = __shfl_sync(0xffffffffu, rk, (lane + t) & 31);
                    if (t == ((int)C - 1) / 2) mid = v; else others = min(others, v);
                }
                const bool dec = lane < 6 && (10u + lane) < n_pos && mid < others;
                const uint32_t fix = __ballot_sync(0xffffffffu, dec);
#pragma unroll
                for (int d = 0; d < 6; ++d) {
                    const uint32_t q = 10u + d;
                    if (q >= p0 && q < p1) {
                        const uint64_t bit = 1ull << (q - p0);
                        mm = (mm & ~bit) | (((fix >> d) & 1u) ? bit : 0ull);
                    }
                }
                if (p0 < 10u) mm &= ~((1ull << min(10u - p0, 63u)) - 1ull);
            }

            // ---- minimizers: positions -> list (lane-major = position order), then balanced re-extraction ----
            {
                const uint32_t cnt = (uint32_t)__popcll(mm);
                uint32_t incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
                const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                uint32_t o = incl - cnt;
                uint64_t m2 = mm;
                while (m2) {
                    const uint32_t j = (uint32_t)__ffsll((long long)m2) - 1u;
                    m2 &= m2 - 1ull;
                    if (o < kFastListCap) plist[o] = (uint16_t)(p0 + j - pstart);
                    ++o;
                }
                __syncwarp();
                for (uint32_t i = lane; i < total && i < kFastListCap; i += 32u) {
                    const uint32_t q = pstart + plist[i];
                    const uint64_t x = window64(packed, pb + q);
                    const uint64_t kf = x >> (64u - 2u * K);
                    const uint64_t kr = revcomp_top(x, kmer_mask);
                    const bool fwd = (kf & split_mask) < (kr & split_mask);            // tie -> reverse (src/seeding.rs:429)
                    mini_out[n_mini + i] = (fwd ? kf : kr) | (fwd ? kFwdBit : 0ull);
                    if (mini_pos_out) mini_pos_out[n_mini + i] = q;
                }
                if (total > kFastListCap) overflow = true;       // impossible (two open syncmers are >= 6 apart); defensive
                n_mini += total;
                __syncwarp();
            }
            // ---- Bloom positives: exact probe + mid-base quality gate (src/seeding.rs:509-521), position order ----
            if (want_snp) {
                const uint32_t cnt = (uint32_t)__popcll(cm);
                uint32_t incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
                uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                uint32_t o = incl - cnt;
                uint64_t m2 = cm;
                while (m2) {
                    const uint32_t j = (uint32_t)__ffsll((long long)m2) - 1u;
                    m2 &= m2 - 1ull;
                    if (o < kFastListCap) plist[o] = (uint16_t)(p0 + j - pstart);
                    ++o;
                }
                if (total > kFastListCap) { overflow = true; total = kFastListCap; }
                __syncwarp();
                for (uint32_t i0 = 0; i0 < total; i0 += 32u) {
                    const uint32_t i = i0 + lane;
                    bool found = false, gate = false;
                    uint64_t canon = 0; uint32_t q = 0;
                    if (i < total) {
                        q = pstart + plist[i];
                        const uint64_t x = window64(packed, pb + q);
                        const uint64_t kf = x >> (64u - 2u * K);
                        const uint64_t kr = revcomp_top(x, kmer_mask);
                        canon = ((kf & split_mask) < (kr & split_mask)) ? kf : kr;
                        uint32_t slot = hash_slot(canon, set.shift);
                        while (true) {
                            const uint64_t v = __ldg(set.keys + slot);
                            if (v == canon) { found = true; break; }
                            if (v == kEmpty64) break;
                            slot = (slot + 1u) & set.mask;
                        }
                        if (found) {
                            int qv = 60;
                            if (in.quals) qv = (int)in.quals[off + q + (K >> 1)] - 33;
                            gate = (qv > p.min_bq) || flat;                             // strict > (src/seeding.rs:517)
                        }
                    }
                    const uint32_t bal = __ballot_sync(0xffffffffu, found);
                    const uint32_t slot_o = nh + (uint32_t)__popc(bal & lt_mask);
                    if (found && slot_o < lay.hit_cap) { hits_k[slot_o] = canon; hits_p[slot_o] = q | (gate ? 0x80000000u : 0u); }
                    nh += (uint32_t)__popc(bal);
                }
                __syncwarp();
            }
        }

        // ------------------------------------------------------------------ phase C (SRC_ASCII only)
        uint32_t start = 0, kept_total = 0;
        if (SRC == 0) {
            if (nh > lay.hit_cap) { overflow = true; nh = lay.hit_cap; }
            if (nh) {
                // splitmer multiplicities over ALL hits, gated or not (src/seeding.rs:522-524)
                const uint32_t tcap = 2u * lay.hit_cap, tmask = tcap - 1u;
                for (uint32_t i = lane; i < tcap; i += 32u) dtab[i] = 0ull;
                __syncwarp();
                for (uint32_t h0 = 0; h0 < nh; h0 += 32u) {
                    const uint32_t h = h0 + lane;
                    if (h < nh) {
                        const uint64_t sp = hits_k[h] & split_mask;
                        // k = 17: the splitmer has exactly 32 significant bits
                        const uint32_t sp32 = ((uint32_t)sp & 0xffffu) | ((uint32_t)(sp >> 18) << 16);
                        const unsigned long long tag = ((unsigned long long)sp32 << 32) | 0x80000000ull;   // never zero
                        uint32_t slot = (sp32 * 0x9E3779B1u) >> 16 & tmask;
                        while (true) {
                            unsigned long long old = dtab[slot];
                            if (old == 0ull) old = atomicCAS(&dtab[slot], 0ull, tag);
                            if (old == 0ull || (old & 0xffffffff80000000ull) == tag) { atomicAdd(&dtab[slot], 1ull); break; }
                            slot = (slot + 1u) & tmask;
                        }
                    }
                }
                __syncwarp();
                for (uint32_t h0 = 0; h0 < nh; h0 += 32u) {
                    const uint32_t h = h0 + lane;
                    bool keep = false;
                    if (h < nh) {
                        const uint64_t sp = hits_k[h] & split_mask;
                        const uint32_t sp32 = ((uint32_t)sp & 0xffffu) | ((uint32_t)(sp >> 18) << 16);
                        const unsigned long long tag = ((unsigned long long)sp32 << 32) | 0x80000000ull;
                        uint32_t slot = (sp32 * 0x9E3779B1u) >> 16 & tmask;
                        unsigned long long v;
                        while (((v = dtab[slot]) & 0xffffffff80000000ull) != tag) slot = (slot + 1u) & tmask;
                        keep = (hits_p[h] >> 31) && (uint32_t)(v & 0x7fffffffull) == 1u;
                        hits_p[h] = (hits_p[h] & 0x7fffffffu) | (keep ? 0x80000000u : 0u);
                    }
                    kept_total += (uint32_t)__popc(__ballot_sync(0xffffffffu, keep));
                }
                __syncwarp();
            }
            if (lane == 0 && kept_total) start = atomicAdd(out.snp_pool_ctr, kept_total);
            start = __shfl_sync(0xffffffffu, start, 0);
            if (kept_total && start + kept_total > out.snp_pool_cap) {
                if (lane == 0) atomicOr(out.err, kErrSnpPool);
                kept_total = 0;
            }
            if (kept_total) {
                uint32_t o = start;
                for (uint32_t h0 = 0; h0 < nh; h0 += 32u) {
                    const uint32_t h = h0 + lane;
                    const bool keep = h < nh && (hits_p[h] >> 31);
                    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
                    if (keep) {
                        const uint32_t d = o + (uint32_t)__popc(bal & lt_mask);
                        out.snp_kmer[d] = hits_k[h];
                        out.snp_pos[d] = hits_p[h] & 0x7fffffffu;
                    }
                    o += (uint32_t)__popc(bal);
                }
            }
        }
        if (overflow && lane == 0) atomicOr(out.err, kErrSnpHits);
        if (lane == 0) {
            out.status[seq] = 0;
            out.flatq[seq] = flat ? 1 : 0;
            out.mini_cnt[seq] = n_mini;
            if (SRC == 0) { out.snp_start[seq] = start; out.snp_cnt[seq] = kept_total; }
        }
        __syncwarp();
    }
}

static uint32_t pow2_at_least(uint32_t x) { uint32_t v = 1; while (v < x) v <<= 1; return v; }

static FastSmem fast_layout(uint32_t max_len, const SnpSet &set) {
    FastSmem lay{};
    lay.bloom_bytes = set.keys ? ((set.bloom_words * 4u + 15u) & ~15u) : 16u;
    lay.packed_words = ((((max_len + 15u + 15u) >> 4) + 8u) + 3u) & ~3u;      // keeps the per-warp regions 16-byte aligned
    lay.hit_cap = pow2_at_least(max_len / 24u + 1u);
    if (lay.hit_cap < 64u) lay.hit_cap = 64u;
    uint32_t stage = (((max_len + 15u + 31u) >> 5) + 3u) * 8u;     // aligned plane words + look-ahead
    if (stage < kFastListCap * 2u) stage = kFastListCap * 2u;
    if (stage < lay.hit_cap * 16u) stage = lay.hit_cap * 16u;
    lay.stage_bytes = (stage + 15u) & ~15u;
    // one persistent CTA per SM: as many warps as the per-warp regions allow in 225 KB (32 for 16S reads, fewer for
    // multi-kb operon reads)
    static const uint32_t want = [] { const char *e = getenv("SAVONT_B200_SKETCH_WARPS"); const int v = e ? atoi(e) : 0;
                                      return (v == 8 || v == 16 || v == 24 || v == 32) ? (uint32_t)v : 16u; }();
    lay.warps = want;
    while (lay.warps > 8u && lay.total() > 225u * 1024u / (kFastWarps / lay.warps)) lay.warps -= 8u;
    return lay;
}

// the fast kernel serves k = 17, c = 11 with the s-mer rank table when its shared-memory layout fits one SM
bool sketch_fast_supported(uint32_t max_len, const DeviceParams &p, bool use_table, const SnpSet &set) {
    if (!(p.k == (int)kFastK && p.c == (int)kFastC && use_table)) return false;
    if (max_len >= 60000u) return false;                          // u16 position lists are per segment; keep clear of any limit
    return fast_layout(max_len, set).total() <= 225u * 1024u;
}

void launch_sketch_fast(const SeqBatch &in, const SketchOut &out, const SnpSet &set, const uint16_t *rank_table,
                        const DeviceParams &p, int sm_count, cudaStream_t st) {
    if (in.n == 0) return;
    const FastSmem lay = fast_layout(in.max_len, set);
    const size_t smem = lay.total();
    // as many CTAs per SM as 32 warps and the shared memory allow (two 16-warp CTAs by default: when the DP kernel holds half of
    // an SM's registers one of them still fits)
    uint32_t per_sm = kFastWarps / lay.warps;
    while (per_sm > 1u && per_sm * smem > 225u * 1024u) --per_sm;
    uint32_t grid = (uint32_t)sm_count * per_sm;
    const uint32_t need = (in.n + lay.warps - 1) / lay.warps;
    if (grid > need) grid = need;
    static const bool trace = getenv("SAVONT_B200_TRACE") != nullptr;
    if (trace) fprintf(stderr, "[trace] sketch_fast: n %u max_len %u warps %u packed_words %u stage %u hit_cap %u bloom %u per_warp %zu smem %zu grid %u\n", in.n, in.max_len, lay.warps, lay.packed_words, lay.stage_bytes, lay.hit_cap, lay.bloom_bytes, lay.per_warp(), smem, grid);
    if (in.dna_words != nullptr) {
        cudaFuncSetAttribute(sketch_fast_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        sketch_fast_kernel<1, false><<<grid, lay.warps * 32, smem, st>>>(in, out, set, rank_table, p, lay);
    } else if (set.keys != nullptr) {
        cudaFuncSetAttribute(sketch_fast_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        sketch_fast_kernel<0, true><<<grid, lay.warps * 32, smem, st>>>(in, out, set, rank_table, p, lay);
    } else {
        cudaFuncSetAttribute(sketch_fast_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        sketch_fast_kernel<0, false><<<grid, lay.warps * 32, smem, st>>>(in, out, set, rank_table, p, lay);
    }
}

// TwinRead.dna_seq words -> one code byte (0..3) per base, for the generic sketch kernel (k, c other than 17/11) and
// for savont_b200_align_pairs-style consumers; one thread per base.
__global__ void unpack_twin_kernel(SeqBatch in, uint8_t *codes) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t seq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; seq < in.n; seq += warps) {
        const uint64_t off = in.offsets[seq];
        const uint32_t L = (uint32_t)(in.offsets[seq + 1] - off);
        const unsigned long long *wp = reinterpret_cast<const unsigned long long *>(in.dna_words) + in.word_offsets[seq];
        for (uint32_t i = lane; i < L; i += 32u) codes[off + i] = (uint8_t)((wp[i >> 5] >> (2u * (i & 31u))) & 3ull);
    }
}
void launch_unpack_twin(const SeqBatch &in, uint8_t *codes, cudaStream_t st) {
    if (in.n == 0) return;
    uint32_t grid = (in.n + 7u) / 8u;
    if (grid > 148u * 8u) grid = 148u * 8u;
    unpack_twin_kernel<<<grid, 256, 0, st>>>(in, codes);
}

// Kmer48 (6 bytes little endian, src/types.rs:104-134) -> u64
__global__ void widen_kmer48_kernel(const uint8_t *src, uint64_t *dst, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t *s = src + 6u * i;
        uint64_t v = 0;
#pragma unroll
        for (int b = 0; b < 6; ++b) v |= (uint64_t)s[b] << (8 * b);
        dst[i] = v;
    }
}
void launch_widen_kmer48(const uint8_t *src, uint64_t *dst, uint64_t n, cudaStream_t st) {
    if (n == 0) return;
    uint32_t grid = (uint32_t)((n + 255u) / 256u);
    if (grid > 148u * 8u) grid = 148u * 8u;
    widen_kmer48_kernel<<<grid, 256, 0, st>>>(src, dst, n);
}

}  // namespace sb
