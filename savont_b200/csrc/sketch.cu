// sketch.cu -- K1: 2-bit packing + open-syncmer / SNPmer sketch of a batch of sequences.
//
// Stands in for seeding::get_twin_read_syncmer (reference src/seeding.rs:317-658), one warp per
// sequence, position-parallel (lane = k-mer start position inside a 32-position tile) instead of
// the reference's rolling registers:
//   phase A  coalesced byte loads -> BYTE_TO_SEQ codes (src/types.rs:92-101) -> warp ballots give
//            the two bit planes (stored for the DP) and a 16-bases-per-word packed copy in shared
//            memory; all-equal-quality check (src/seeding.rs:372-380).
//   phase B  per tile: s-mer hash rank at every base index (a 4^s-entry rank table of
//            mm_hash64(canonical s-mer) staged in shared memory replaces the 64-bit hash: only the
//            ORDER of the hashes is ever used, src/seeding.rs:526-543), forward/reverse k-mer
//            extraction, splitmer-canonical choice (:426-434, tie -> reverse), SNPmer set probe +
//            mid-base quality gate (:508-521), open-syncmer test (:527-543); ordered compaction by
//            ballot/popc.
//   phase C  SNPmer de-duplication by splitmer multiplicity (:550-568) and pool allocation.
// The reference's s-mer warm-up quirk (SURVEY.md 8a note 1) is reproduced: the s-mers ending at
// base indices k-1 .. k+s-3 are formed from bases 0..s-2 followed by bases k-1.. .
#include "common.cuh"

namespace sb {

__device__ __forceinline__ uint32_t enc_base(uint32_t b) {
    const uint32_t u = b & 0xDFu;
    uint32_t code = (b <= 3u) ? b : 0u;
    code = (u == 'C') ? 1u : code;
    code = (u == 'G') ? 2u : code;
    code = (u == 'T' || u == 'U') ? 3u : code;
    return code;
}

// 64 bits of the packed sequence starting at base a (base a in the top bit pair)
__device__ __forceinline__ uint64_t window64(const uint32_t *packed, uint32_t a) {
    const uint32_t w = a >> 4, sh = (a & 15u) * 2u;
    const uint32_t w0 = packed[w], w1 = packed[w + 1], w2 = packed[w + 2];
    const uint32_t hi = __funnelshift_l(w1, w0, sh);
    const uint32_t lo = __funnelshift_l(w2, w1, sh);
    return ((uint64_t)hi << 32) | lo;
}
// len (1..16) bases starting at base a, right-aligned
__device__ __forceinline__ uint32_t extract32(const uint32_t *packed, uint32_t a, uint32_t len) {
    const uint32_t w = a >> 4, sh = (a & 15u) * 2u;
    return __funnelshift_l(packed[w + 1], packed[w], sh) >> (32u - 2u * len);
}
__device__ __forceinline__ uint64_t swap_pairs64(uint64_t r) {
    return ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
}
// reverse complement of the k-mer held in the TOP 2k bits of x; result right-aligned
__device__ __forceinline__ uint64_t revcomp_top(uint64_t x, uint64_t kmer_mask) {
    const uint64_t z = ~x;
    const uint64_t r = ((uint64_t)__brev((uint32_t)z) << 32) | (uint64_t)__brev((uint32_t)(z >> 32));
    return swap_pairs64(r) & kmer_mask;
}
__device__ __forceinline__ uint32_t revcomp_smer(uint32_t f, uint32_t s) {
    const uint32_t z = ~(f << (32u - 2u * s));
    const uint32_t r = __brev(z);
    const uint32_t sw = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    return (s == 16u) ? sw : (sw & ((1u << (2u * s)) - 1u));
}

// phase A shared by the sketch kernel and pack_planes_kernel
__device__ __forceinline__ bool pack_sequence(const SeqBatch &in, uint64_t off, uint32_t L, uint2 *planes, uint32_t *packed,
                                              int lane) {
    const uint32_t nblk = (L + 31u) >> 5;
    bool flat = true;
    const uint32_t q0 = in.quals ? in.quals[off] : 0u;
    for (uint32_t blk = 0; blk < nblk; ++blk) {
        const uint32_t i = blk * 32u + lane;
        const uint32_t code = (i < L) ? enc_base(in.bases[off + i]) : 0u;
        const uint32_t lo = __ballot_sync(0xffffffffu, code & 1u);
        const uint32_t hi = __ballot_sync(0xffffffffu, code & 2u);
        if (lane == 0) planes[blk] = make_uint2(lo, hi);
        if (packed) {
            uint32_t v = code << (30u - 2u * (lane & 15u));
            v |= __shfl_xor_sync(0xffffffffu, v, 1);
            v |= __shfl_xor_sync(0xffffffffu, v, 2);
            v |= __shfl_xor_sync(0xffffffffu, v, 4);
            v |= __shfl_xor_sync(0xffffffffu, v, 8);
            if ((lane & 15) == 0) packed[blk * 2u + (lane >> 4)] = v;
        }
        if (in.quals) {
            const uint32_t q = (i < L) ? in.quals[off + i] : q0;
            flat = flat && __all_sync(0xffffffffu, q == q0);
        }
    }
    if (lane == 0) planes[nblk] = make_uint2(0u, 0u);   // guard word read by the DP look-ahead
    if (packed && lane < 4) packed[nblk * 2u + lane] = 0u;
    return flat && in.quals != nullptr;
}

// shared memory layout of the sketch kernel
struct SketchSmem {
    uint32_t table_bytes;    // s-mer rank table (0 when hashes are computed directly)
    uint32_t bloom_bytes;    // SNPmer Bloom filter (0 when the set is empty)
    uint32_t packed_words;   // per warp
    uint32_t rank_bytes;     // per warp: (32 + c) ranks, u32 or u64
    __host__ __device__ size_t per_warp() const {
        const size_t b = (size_t)kSnpHitCap * 12u + rank_bytes + (size_t)packed_words * 4u;
        return (b + 15u) & ~(size_t)15u;
    }
    __host__ __device__ size_t total() const { return (size_t)table_bytes + bloom_bytes + per_warp() * kSketchWarps; }
};

template <bool B> struct RankT { using type = uint64_t; };
template <> struct RankT<true> { using type = uint32_t; };

// K, C > 0: compile-time k-mer / window sizes (the default 17/11 path); K == 0: runtime parameters.
// USE_TABLE: s-mer hash ranks from the 4^s-entry shared-memory table instead of mm_hash64.
template <int K, int C, bool USE_TABLE>
__global__ void __launch_bounds__(kSketchWarps * 32)
sketch_kernel(SeqBatch in, SketchOut out, SnpSet set, const uint16_t *__restrict__ rank_table_g, DeviceParams p, SketchSmem lay) {
    using rank_t = typename RankT<USE_TABLE>::type;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t k = K ? (uint32_t)K : (uint32_t)p.k;
    const uint32_t c = K ? (uint32_t)C : (uint32_t)p.c;
    const uint32_t s = k - c + 1u;
    const uint64_t kmer_mask = (1ull << (2u * k)) - 1ull;
    const uint64_t split_mask = ~(3ull << (k - 1u));
    const uint32_t smask = (s >= 16u) ? 0xffffffffu : ((1u << (2u * s)) - 1u);

    const uint16_t *s_rank = reinterpret_cast<const uint16_t *>(smem);
    const uint32_t *s_bloom = reinterpret_cast<const uint32_t *>(smem + lay.table_bytes);
    {
        uint32_t *dst = reinterpret_cast<uint32_t *>(smem);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(rank_table_g);
        const uint32_t nt = lay.table_bytes / 4u, nb = lay.bloom_bytes / 4u;
        for (uint32_t i = threadIdx.x; i < nt; i += blockDim.x) dst[i] = __ldg(src + i);
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) dst[nt + i] = __ldg(set.bloom + i);
        __syncthreads();
    }
    uint8_t *wbase = smem + lay.table_bytes + lay.bloom_bytes + (size_t)warp * lay.per_warp();
    uint64_t *hits_k = reinterpret_cast<uint64_t *>(wbase);
    rank_t *rbuf = reinterpret_cast<rank_t *>(wbase + (size_t)kSnpHitCap * 8u);
    uint32_t *hits_p = reinterpret_cast<uint32_t *>(wbase + (size_t)kSnpHitCap * 8u + lay.rank_bytes);
    uint32_t *packed = hits_p + kSnpHitCap;

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
            pack_sequence(in, off, L, planes, nullptr, lane);   // planes still serve align_pairs
            if (lane == 0) {
                out.status[seq] = 4; out.flatq[seq] = 0; out.mini_cnt[seq] = 0;
                out.snp_start[seq] = 0; out.snp_cnt[seq] = 0;
            }
            continue;
        }
        __syncwarp();
        const bool flat = pack_sequence(in, off, L, planes, packed, lane);
        __syncwarp();

        const uint32_t n_pos = L - k + 1u;
        uint64_t *mini_out = out.mini_kmer + mini_slot_off(off, seq, (uint32_t)p.mini_spacing);
        uint32_t *mini_pos_out = out.mini_pos ? out.mini_pos + mini_slot_off(off, seq, (uint32_t)p.mini_spacing) : nullptr;
        uint32_t n_mini = 0, n_hits = 0;

        // rbuf[e] holds the s-mer hash rank of base index  base_p + (k - c) + e,  e in [0, 31 + c)
        for (uint32_t base_p = 0; base_p < n_pos; base_p += 32u) {
            const uint32_t pos = base_p + lane;
            const bool valid = pos < n_pos;
            // carry the c-1 trailing ranks of the previous tile to the front
            rank_t carry = 0;
            if ((uint32_t)lane < c - 1u) carry = rbuf[32u + lane];
            __syncwarp();
            if ((uint32_t)lane < c - 1u) rbuf[lane] = carry;

            uint64_t canon = 0; bool fwd = false;
            bool is_snp = false, gate = false;
            if (valid) {
                const uint64_t x = window64(packed, pos);
                const uint64_t kf = x >> (64u - 2u * k);
                const uint64_t kr = revcomp_top(x, kmer_mask);
                fwd = (kf & split_mask) < (kr & split_mask);       // tie -> reverse (src/seeding.rs:429)
                canon = fwd ? kf : kr;
                // the s-mer ending at base pos+k-1 is the tail of this k-mer ...
                uint32_t f = (uint32_t)kf & smask;
                if (pos + 1u < s) {
                    // ... except for the reference's warm-up quirk (src/seeding.rs:392-397,413): the s-mer
                    // register was primed with bases 0..s-2 only, so the first s-1 s-mers are
                    // bases pos..s-2 (head of this k-mer) followed by bases k-1..k-1+pos (its tail)
                    const uint32_t nt = pos + 1u;
                    const uint32_t head = (uint32_t)(kf >> (2u * (k - (s - nt))));
                    f = (head << (2u * nt)) | ((uint32_t)kf & ((1u << (2u * nt)) - 1u));
                }
                rank_t rk;
                if (USE_TABLE) {
                    rk = (rank_t)s_rank[f];
                } else {
                    const uint32_t r = revcomp_smer(f, s);
                    rk = (rank_t)mm_hash64((uint64_t)(f < r ? f : r));
                }
                rbuf[c - 1u + lane] = rk;
                if (set.keys != nullptr) {
                    const uint64_t prod = canon * kFibMul;
                    const uint32_t bi = (uint32_t)(prod >> set.bloom_shift);
                    if ((s_bloom[bi >> 5] >> (bi & 31u)) & 1u) {
                        uint32_t slot = (uint32_t)(prod >> set.shift);
                        while (true) {
                            const uint64_t v = __ldg(set.keys + slot);
                            if (v == canon) { is_snp = true; break; }
                            if (v == kEmpty64) break;
                            slot = (slot + 1u) & set.mask;
                        }
                        if (is_snp) {
                            int q = 60;
                            if (in.quals) q = (int)in.quals[off + pos + half_k] - 33;
                            gate = (q > p.min_bq) || flat;             // strict > (src/seeding.rs:517)
                        }
                    }
                }
            }
            __syncwarp();
            bool is_mini = false;
            if (valid && pos + 1u >= c) {                              // the window of c hashes is full
                const rank_t mid = rbuf[lane + mid_idx];
                rank_t others = ~(rank_t)0;
                if (K) {
#pragma unroll
                    for (uint32_t j = 0; j < (uint32_t)(C ? C : 1); ++j)
                        if (j != mid_idx) { const rank_t v = rbuf[lane + j]; others = v < others ? v : others; }
                } else {
                    for (uint32_t j = 0; j < c; ++j)
                        if (j != mid_idx) { const rank_t v = rbuf[lane + j]; others = v < others ? v : others; }
                }
                is_mini = mid < others || (c == 1u);                   // strict minimum at the middle offset
            }
            const uint32_t mb = __ballot_sync(0xffffffffu, is_mini);
            if (is_mini) {
                const uint32_t o = n_mini + __popc(mb & lt_mask);
                mini_out[o] = canon | (fwd ? kFwdBit : 0ull);
                if (mini_pos_out) mini_pos_out[o] = pos;
            }
            n_mini += __popc(mb);
            const uint32_t hb = __ballot_sync(0xffffffffu, is_snp);
            if (is_snp) {
                const uint32_t o = n_hits + __popc(hb & lt_mask);
                if (o < (uint32_t)kSnpHitCap) { hits_k[o] = canon; hits_p[o] = pos | (gate ? 0x80000000u : 0u); }
                else { spill_k[o - kSnpHitCap] = canon; spill_p[o - kSnpHitCap] = pos | (gate ? 0x80000000u : 0u); }
            }
            n_hits += __popc(hb);
        }
        __syncwarp();

        // ---- SNPmer dedup: keep gated hits whose splitmer occurs exactly once among ALL hits ----
        const uint32_t nh = n_hits;
        uint32_t kept_total = 0;
        for (uint32_t base = 0; base < nh; base += 32u) {
            const uint32_t h = base + lane;
            bool keep = false;
            if (h < nh) {
                const uint64_t sp = hit_k(h) & split_mask;
                uint32_t cnt = 0;
                for (uint32_t j = 0; j < nh; ++j) cnt += ((hit_k(j) & split_mask) == sp) ? 1u : 0u;
                keep = (hit_p(h) >> 31) && cnt == 1u;
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
            uint32_t idx = 0;
            for (uint32_t base = 0; base < nh; base += 32u) {
                const uint32_t h = base + lane;
                const bool keep = (h < nh) && (hit_p(h) >> 31);
                const uint32_t b = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const uint32_t o = start + idx + __popc(b & lt_mask);
                    out.snp_kmer[o] = hit_k(h);
                    out.snp_pos[o] = hit_p(h) & 0x7fffffffu;
                }
                idx += __popc(b);
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
    lay.table_bytes = use_table ? (uint32_t)((((size_t)1 << (2 * p.s)) * 2u + 15u) & ~(size_t)15u) : 0u;
    lay.bloom_bytes = set.keys ? set.bloom_words * 4u : 0u;
    lay.packed_words = (max_len >> 4) + 6u;
    lay.rank_bytes = (uint32_t)(((32u + (uint32_t)p.c) * (use_table ? 4u : 8u) + 15u) & ~15u);
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
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t seq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; seq < in.n; seq += warps) {
        const uint64_t off = in.offsets[seq];
        const uint32_t L = (uint32_t)(in.offsets[seq + 1] - off);
        SeqBatch nq = in; nq.quals = nullptr;
        pack_sequence(nq, off, L, planes_all + plane_word_off(off, seq), nullptr, lane);
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
