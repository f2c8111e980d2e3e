// kmercount.cu -- K5: Stage 1 split-k-mer counting, the integer part of seq_parse::read_to_split_kmers
// (reference src/seq_parse.rs:12-83, :315-497, bloom_filter_size == 0) over seeding::split_kmer_mid
// (src/seeding.rs:975-1068).
//
// The reference counts every k-mer of every read in sharded FxHashMap<u64, [u32; 2]> and afterwards keeps the entries
// seen on both strands more than twice (src/seq_parse.rs:30-44).  1 M ONT 16S reads hold 1.48 G k-mer occurrences of
// 48 M distinct k-mers.  Two algorithms, both exact:
//   partitioned (k <= 21, the default)   histogram of the occurrences over 8192 hash partitions -> scatter of 32-bit
//       entries into the partitions (two levels with a shared-memory sort per tile, so that global writes are coalesced
//       runs; one level of direct stores for k = 21) -> one CTA per partition counts in a shared-memory table and emits
//       the survivors.  Sequential traffic only: 32 ms per 1 M reads (hist 5.8, scatter 18.5, count 7.9).
//   single table (k = 23, or when a partition holds more distinct k-mers than the spill table takes)   one pass, one
//       open-addressing table in HBM (16-byte slots), per occurrence one key probe + one atomicAdd in the same sector:
//       83 ms per 1 M reads.  Measured: extraction alone 5.5 ms, + probes 42.7 ms, + count atomics 81.7 ms -- scattered
//       32-byte sector operations run at ~40 G/s on B200 whatever their kind (load, reduction, store), far below what the
//       issue slots (7 % busy) or DRAM (8 %) could take; 4 probes in flight per lane change nothing.  That is why the
//       partitioned algorithm exists.
// Reads whose id ends in "rc" are reverse-complemented first exactly like utils::reverse_complement does it
// (src/utils.rs:51-65: anything but ACGT/acgt becomes N, which encodes as A).
#include "common.cuh"
#include "sketch_dev.cuh"

namespace sb {

constexpr int kKcWarps = 16;
constexpr int kKcUnroll = 4;
constexpr uint64_t kKcMul2 = 0xD6E8FEB86659FD93ull;

struct KcArgs {
    SeqBatch in;
    const uint8_t *rc;        // per read, may be null
    int k, min_bq;
    KmerSlot *table; uint32_t tshift, tmask;        // slot = (kmer * kKcMul2) >> tshift
    uint32_t *ctr;                                  // [0] slots in use (= distinct k-mers), [1] error bits
    uint32_t fill_limit;
    uint32_t packed_words;                          // per warp
};

__global__ void __launch_bounds__(kKcWarps * 32, 2)
kmer_count_kernel(KcArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *s_lut = smem;                           // [0..255] forward codes, [256..511] codes of the reverse complement
    uint32_t *packed = reinterpret_cast<uint32_t *>(smem + 512) + (size_t)warp * a.packed_words;
    for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) {
        const uint32_t c = enc_base(i);
        const uint32_t u = i & 0xDFu;
        const bool acgt = (u == 'A' || u == 'C' || u == 'G' || u == 'T');
        s_lut[i] = (uint8_t)c;
        s_lut[256 + i] = (uint8_t)(acgt ? 3u - c : 0u);         // reverse_complement maps everything else to N (= A)
    }
    __syncthreads();
    const uint32_t K = (uint32_t)a.k;
    const uint64_t kmer_mask = (K == 32u) ? ~0ull : ((1ull << (2u * K)) - 1ull);
    const uint64_t split_mask = ~(3ull << (K - 1u));

    for (uint32_t seq = blockIdx.x * kKcWarps + warp; seq < a.in.n; seq += gridDim.x * kKcWarps) {
        const uint64_t off = a.in.offsets[seq];
        const uint32_t L = (uint32_t)(a.in.offsets[seq + 1] - off);
        if (L < K) continue;                                                  // src/seeding.rs:982-984
        const bool rc = a.rc != nullptr && a.rc[seq] != 0;
        const uint8_t *bp = a.in.bases + off;
        const uint8_t *qp = a.in.quals ? a.in.quals + off : nullptr;
        __syncwarp();
        // ---- 2-bit packing in processing orientation: 16 bases per word, first base in the top pair ----
        const uint32_t nwords = (L + 15u) >> 4;
        for (uint32_t w = lane; w < nwords + 3u; w += 32u) {
            uint32_t pk = 0;
            if (w < nwords) {
#pragma unroll 4
                for (uint32_t t = 0; t < 16u; ++t) {
                    const uint32_t i = 16u * w + t;
                    uint32_t c = 0;
                    if (i < L) c = rc ? s_lut[256u + bp[L - 1u - i]] : s_lut[bp[i]];
                    pk |= c << (30u - 2u * t);
                }
            }
            packed[w] = pk;
        }
        // ---- quality skip is active only when the qualities are not all equal (src/seeding.rs:1004-1018) ----
        bool use_q = false;
        if (qp != nullptr) {
            const uint8_t q0 = qp[0];
            for (uint32_t i0 = 0; i0 < L && !use_q; i0 += 128u) {
                bool d = false;
#pragma unroll
                for (uint32_t u = 0; u < 4u; ++u) { const uint32_t i = i0 + 32u * u + lane; if (i < L && qp[i] != q0) d = true; }
                use_q = __any_sync(0xffffffffu, d);
            }
        }
        __syncwarp();
        // ---- kKcUnroll k-mers per lane and step: all key probes of a step are issued before the first one is
        //      consumed (the kernel is bound by the latency of these random L2 / DRAM accesses, not by issue slots) ----
        const uint32_t n_pos = L - K + 1u;
        for (uint32_t p0 = 0; p0 < n_pos; p0 += 32u * kKcUnroll) {
            unsigned long long canon[kKcUnroll], found[kKcUnroll];
            uint32_t slot[kKcUnroll];
            bool valid[kKcUnroll], fwd[kKcUnroll];
            int q[kKcUnroll];
#pragma unroll
            for (int u = 0; u < kKcUnroll; ++u) {                             // middle-base qualities first (global loads)
                const uint32_t p = p0 + 32u * u + lane;
                q[u] = 1000;
                if (use_q && p < n_pos) { const uint32_t mid = p + (K >> 1); q[u] = (int)qp[rc ? L - 1u - mid : mid] - 33; }
            }
#pragma unroll
            for (int u = 0; u < kKcUnroll; ++u) {
                const uint32_t p = p0 + 32u * u + lane;
                valid[u] = false; fwd[u] = false; canon[u] = 0; slot[u] = 0;
                if (p < n_pos) {
                    const uint64_t x = window64(packed, p);
                    const uint64_t kf = x >> (64u - 2u * K);
                    const uint64_t kr = revcomp_top(x, kmer_mask);
                    const uint64_t sf = kf & split_mask, sr = kr & split_mask;
                    fwd[u] = sf < sr;                                         // :1053
                    canon[u] = fwd[u] ? kf : kr;
                    valid[u] = (sf != sr) && (q[u] >= a.min_bq);              // palindromic splitmer :1044-1046; quality :1012
                    slot[u] = (uint32_t)((canon[u] * kKcMul2) >> a.tshift);
                }
            }
#pragma unroll
            for (int u = 0; u < kKcUnroll; ++u) found[u] = valid[u] ? __ldcg(&a.table[slot[u]].key) : 0ull;
#pragma unroll
            for (int u = 0; u < kKcUnroll; ++u) {
                if (!valid[u]) continue;
                uint32_t sl = slot[u];
                unsigned long long v = found[u];
                uint32_t *cnt = nullptr;
                for (uint32_t probes = 0;; ++probes) {
                    if (v == canon[u]) { cnt = fwd[u] ? &a.table[sl].c1 : &a.table[sl].c0; break; }
                    if (v == kEmpty64) {
                        const unsigned long long old = atomicCAS(&a.table[sl].key, (unsigned long long)kEmpty64, canon[u]);
                        if (old == kEmpty64) {
                            if (atomicAdd(a.ctr, 1u) >= a.fill_limit) atomicOr(a.ctr + 1, 1u);  // the host regrows and re-runs
                            cnt = fwd[u] ? &a.table[sl].c1 : &a.table[sl].c0; break;
                        }
                        if (old == canon[u]) { cnt = fwd[u] ? &a.table[sl].c1 : &a.table[sl].c0; break; }
                    }
                    if (probes > 4096u) { atomicOr(a.ctr + 1, 1u); break; }
                    sl = (sl + 1u) & a.tmask;
                    v = __ldcg(&a.table[sl].key);
                }
                if (cnt != nullptr) atomicAdd(cnt, 1u);                       // src/seq_parse.rs:455-460
            }
        }
    }
}

// =================================================================================================
// Partitioned counting (k <= 21): sequential traffic instead of one scattered sector operation per occurrence.
//   1. kmer_hist_kernel     every occurrence -> histogram over P = 2^pbits partitions (top bits of h = kmer * M mod 2^2k,
//                           M odd: a bijection on the k-mer space, so partition + remainder identify the k-mer)
//   2. host                 exclusive scan of the histogram = partition offsets
//   3. kmer_scatter_kernel  every occurrence again -> 32-bit entry (remainder | flag << 31) appended to its partition;
//                           a CTA reserves the space of a whole tile of reads per partition with ONE global atomic
//   4. kmer_count_part_kernel  one CTA per partition: the partition's k-mers are counted in a shared-memory table
//                           (16 K slots of 12 B), in R rounds over a slice of the remainders when the partition is large;
//                           survivors of the reference's filter are reconstructed (h * M^-1) and appended to the output.
//                           k-mers that find the shared table full are counted in the global table instead (spill).
// =================================================================================================
constexpr int kKpTileReads = 4;              // reads per warp and tile in the scatter kernel
constexpr uint32_t kKpSlots = 16384;         // shared-memory table slots of the counting kernel
constexpr uint32_t kKpFillLimit = 11000;     // distinct k-mers per round before spilling

struct KpArgs {
    KcArgs kc;                               // reads, k, min_bq, the global (spill) table, counters
    uint64_t mul, inv, kmask;                // h = (kmer * mul) & kmask; kmer = (h * inv) & kmask
    int pbits, rem_bits;                     // partition = h >> rem_bits; entry = (h & rem_mask) | flag << 31
    uint32_t *hist;                          // [P] occurrences per partition
    uint32_t *cursor;                        // [P] next free entry of the partition, relative to part_off (scatter)
    uint32_t *entries;                       // all partitions back to back
    const unsigned long long *part_off;      // [P + 1]
    uint32_t *const *dst_ptr;                // sharded counting: [P] where this rank's entries of a partition go (a peer's HBM
                                             // through NVLink, or local); null = entries + part_off
    int single_strand;
    uint64_t *out_kmers; uint32_t *out_counts; uint32_t out_cap;
    uint32_t *out_n;                         // survivors written
    unsigned long long *n_distinct;          // distinct k-mers counted in shared memory
    uint32_t *queue;                         // partition work queue
    uint32_t packed_words;
    uint32_t smem_fill_limit;               // distinct k-mers per round before the round goes to the global table
};

// one warp: pack `seq` into `packed` (processing orientation) and report whether the quality skip is active
__device__ __forceinline__ bool kp_pack_read(const KcArgs &a, const uint8_t *s_lut, uint32_t *packed, uint32_t seq, int lane,
                                             uint32_t &L, bool &rc, const uint8_t *&qp) {
    const uint64_t off = a.in.offsets[seq];
    L = (uint32_t)(a.in.offsets[seq + 1] - off);
    rc = a.rc != nullptr && a.rc[seq] != 0;
    const uint8_t *bp = a.in.bases + off;
    qp = a.in.quals ? a.in.quals + off : nullptr;
    const uint32_t nwords = (L + 15u) >> 4;
    for (uint32_t w = lane; w < nwords + 3u; w += 32u) {
        uint32_t pk = 0;
        if (w < nwords) {
#pragma unroll 4
            for (uint32_t t = 0; t < 16u; ++t) {
                const uint32_t i = 16u * w + t;
                uint32_t c = 0;
                if (i < L) c = rc ? s_lut[256u + bp[L - 1u - i]] : s_lut[bp[i]];
                pk |= c << (30u - 2u * t);
            }
        }
        packed[w] = pk;
    }
    bool use_q = false;
    if (qp != nullptr && L > 0u) {
        const uint8_t q0 = qp[0];
        for (uint32_t i0 = 0; i0 < L && !use_q; i0 += 128u) {
            bool d = false;
#pragma unroll
            for (uint32_t u = 0; u < 4u; ++u) { const uint32_t i = i0 + 32u * u + lane; if (i < L && qp[i] != q0) d = true; }
            use_q = __any_sync(0xffffffffu, d);
        }
    }
    __syncwarp();
    return use_q;
}

// the occurrence at position p of a packed read: false = skipped (palindromic splitmer / low middle-base quality)
__device__ __forceinline__ bool kp_occurrence(const uint32_t *packed, uint32_t p, uint32_t K, uint64_t kmer_mask, uint64_t split_mask,
                                              bool use_q, const uint8_t *qp, bool rc, uint32_t L, int min_bq, uint64_t &canon, bool &fwd) {
    const uint64_t x = window64(packed, p);
    const uint64_t kf = x >> (64u - 2u * K);
    const uint64_t kr = revcomp_top(x, kmer_mask);
    const uint64_t sf = kf & split_mask, sr = kr & split_mask;
    if (sf == sr) return false;
    if (use_q) {
        const uint32_t mid = p + (K >> 1);
        if ((int)qp[rc ? L - 1u - mid : mid] - 33 < min_bq) return false;
    }
    fwd = sf < sr;
    canon = fwd ? kf : kr;
    return true;
}

__device__ __forceinline__ void kp_lut(uint8_t *s_lut) {
    for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) {
        const uint32_t c = enc_base(i);
        const uint32_t u = i & 0xDFu;
        const bool acgt = (u == 'A' || u == 'C' || u == 'G' || u == 'T');
        s_lut[i] = (uint8_t)c;
        s_lut[256 + i] = (uint8_t)(acgt ? 3u - c : 0u);
    }
}

// count one occurrence in the global table (spill path of the partitioned kernels; same protocol as kmer_count_kernel)
__device__ __forceinline__ void kc_global_count(const KcArgs &a, unsigned long long canon, bool fwd) {
    uint32_t sl = (uint32_t)((canon * kKcMul2) >> a.tshift);
    for (uint32_t probes = 0;; ++probes) {
        const unsigned long long v = __ldcg(&a.table[sl].key);
        bool mine = v == canon;
        if (!mine && v == kEmpty64) {
            const unsigned long long old = atomicCAS(&a.table[sl].key, (unsigned long long)kEmpty64, canon);
            if (old == kEmpty64) { if (atomicAdd(a.ctr, 1u) >= a.fill_limit) atomicOr(a.ctr + 1, 1u); mine = true; }
            else mine = old == canon;
        }
        if (mine) { atomicAdd(fwd ? &a.table[sl].c1 : &a.table[sl].c0, 1u); return; }
        if (probes > 4096u) { atomicOr(a.ctr + 1, 1u); return; }
        sl = (sl + 1u) & a.tmask;
    }
}

__global__ void __launch_bounds__(kKcWarps * 32)
kmer_hist_kernel(KpArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *s_lut = smem;
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem + 512);
    const uint32_t P = 1u << a.pbits;
    uint32_t *packed = s_hist + P + (size_t)warp * a.packed_words;
    kp_lut(s_lut);
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    const uint32_t K = (uint32_t)a.kc.k;
    const uint64_t kmer_mask = (1ull << (2u * K)) - 1ull, split_mask = ~(3ull << (K - 1u));
    for (uint32_t seq = blockIdx.x * kKcWarps + warp; seq < a.kc.in.n; seq += gridDim.x * kKcWarps) {
        uint32_t L; bool rc; const uint8_t *qp;
        __syncwarp();
        const bool use_q = kp_pack_read(a.kc, s_lut, packed, seq, lane, L, rc, qp);
        if (L < K) continue;
        for (uint32_t p = lane; p < L - K + 1u; p += 32u) {
            uint64_t canon; bool fwd;
            if (!kp_occurrence(packed, p, K, kmer_mask, split_mask, use_q, qp, rc, L, a.kc.min_bq, canon, fwd)) continue;
            const uint64_t h = (canon * a.mul) & a.kmask;
            atomicAdd(&s_hist[(uint32_t)(h >> a.rem_bits)], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) if (s_hist[i]) atomicAdd(&a.hist[i], s_hist[i]);
}

__global__ void __launch_bounds__(kKcWarps * 32)
kmer_scatter_kernel(KpArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *s_lut = smem;
    const uint32_t P = 1u << a.pbits;
    uint32_t *s_base = reinterpret_cast<uint32_t *>(smem + 512);
    uint32_t *s_hist = s_base + P;
    uint32_t *packed = s_hist + P + (size_t)warp * a.packed_words * kKpTileReads;
    kp_lut(s_lut);
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    const uint32_t K = (uint32_t)a.kc.k;
    const uint64_t kmer_mask = (1ull << (2u * K)) - 1ull, split_mask = ~(3ull << (K - 1u));
    const uint64_t rem_mask = (1ull << a.rem_bits) - 1ull;
    const uint32_t tile = kKcWarps * kKpTileReads;
    const uint32_t n_tiles = (a.kc.in.n + tile - 1u) / tile;
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        uint32_t Ls[kKpTileReads]; bool rcs[kKpTileReads], uqs[kKpTileReads]; const uint8_t *qps[kKpTileReads];
        // phase A: pack this warp's reads, count the tile's occurrences per partition
#pragma unroll
        for (int j = 0; j < kKpTileReads; ++j) {
            const uint32_t seq = t * tile + (uint32_t)warp * kKpTileReads + (uint32_t)j;
            Ls[j] = 0; rcs[j] = false; uqs[j] = false; qps[j] = nullptr;
            if (seq >= a.kc.in.n) continue;
            uint32_t *pk = packed + (size_t)j * a.packed_words;
            uqs[j] = kp_pack_read(a.kc, s_lut, pk, seq, lane, Ls[j], rcs[j], qps[j]);
            if (Ls[j] < K) { Ls[j] = 0; continue; }
            for (uint32_t p = lane; p < Ls[j] - K + 1u; p += 32u) {
                uint64_t canon; bool fwd;
                if (!kp_occurrence(pk, p, K, kmer_mask, split_mask, uqs[j], qps[j], rcs[j], Ls[j], a.kc.min_bq, canon, fwd)) continue;
                const uint64_t h = (canon * a.mul) & a.kmask;
                atomicAdd(&s_hist[(uint32_t)(h >> a.rem_bits)], 1u);
            }
        }
        __syncthreads();
        // one global reservation per non-empty partition
        for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
            const uint32_t c = s_hist[i];
            if (c) s_base[i] = atomicAdd(&a.cursor[i], c);
            s_hist[i] = 0u;
        }
        __syncthreads();
        // phase B: the same occurrences again, now written to their reserved places
#pragma unroll
        for (int j = 0; j < kKpTileReads; ++j) {
            if (Ls[j] == 0u) continue;
            const uint32_t *pk = packed + (size_t)j * a.packed_words;
            for (uint32_t p = lane; p < Ls[j] - K + 1u; p += 32u) {
                uint64_t canon; bool fwd;
                if (!kp_occurrence(pk, p, K, kmer_mask, split_mask, uqs[j], qps[j], rcs[j], Ls[j], a.kc.min_bq, canon, fwd)) continue;
                const uint64_t h = (canon * a.mul) & a.kmask;
                const uint32_t bin = (uint32_t)(h >> a.rem_bits);
                const uint32_t r = atomicAdd(&s_hist[bin], 1u);
                a.entries[__ldg(a.part_off + bin) + s_base[bin] + r] = (uint32_t)(h & rem_mask) | (fwd ? 0x80000000u : 0u);
            }
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) s_hist[i] = 0u;
        __syncthreads();
    }
}

// ---- two-level scatter (2k - 7 + 1 <= 32): both levels sort a tile by bin in shared memory first, so that the global
// writes are runs of consecutive entries (coalesced) and a tile costs one global atomic per bin.  The direct scatter above
// issues one 4-byte store per occurrence to 8192 different places: scattered sector operations run at ~40 G/s on B200
// whatever their kind (load, reduction, store), which made it the slowest step. ----
constexpr int kKp1Bits = 7;                  // level 1: 128 bins (the top bits of h)
constexpr uint32_t kKp1TileCap = 24576;      // entries per level-1 tile (16 reads of up to 1536 k-mers)
constexpr uint32_t kKp2Tile = 12288;         // entries per level-2 tile

struct Kp2Args {
    KpArgs p;
    uint32_t *entries1;                      // level-1 output: (h & rem1_mask) | flag << 31, 128 bins back to back
    uint32_t *cursor1;                       // [128] relative
    const uint32_t *tile_start;              // [129] first level-2 tile of every level-1 bin
    uint32_t n_tiles2, tile_reads;
};

__global__ void __launch_bounds__(kKcWarps * 32)
kmer_scatter1_kernel(Kp2Args b) {
    const KpArgs &a = b.p;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr uint32_t P1 = 1u << kKp1Bits;
    uint8_t *s_lut = smem;
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem + 512);           // counts, then fill cursors
    uint32_t *s_off = s_hist + P1;                                         // [P1 + 1]
    uint32_t *s_gbase = s_off + P1 + 1;
    uint32_t *sorted = s_gbase + P1;
    uint32_t *packed = sorted + kKp1TileCap + (size_t)warp * a.packed_words;
    kp_lut(s_lut);
    for (uint32_t i = threadIdx.x; i < P1; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    const uint32_t K = (uint32_t)a.kc.k;
    const uint64_t kmer_mask = (1ull << (2u * K)) - 1ull, split_mask = ~(3ull << (K - 1u));
    const int rem1_bits = 2 * a.kc.k - kKp1Bits;
    const uint64_t rem1_mask = (1ull << rem1_bits) - 1ull;
    const int sub_bits = a.pbits - kKp1Bits;
    const uint32_t tile = b.tile_reads;
    const uint32_t n_tiles = (a.kc.in.n + tile - 1u) / tile;
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint32_t seq = t * tile + (uint32_t)warp;
        uint32_t L = 0; bool rc = false, use_q = false; const uint8_t *qp = nullptr;
        const bool have = (uint32_t)warp < tile && seq < a.kc.in.n;
        if (have) {
            use_q = kp_pack_read(a.kc, s_lut, packed, seq, lane, L, rc, qp);
            if (L < K) L = 0;
        }
        if (L) {
            for (uint32_t p = lane; p < L - K + 1u; p += 32u) {
                uint64_t canon; bool fwd;
                if (!kp_occurrence(packed, p, K, kmer_mask, split_mask, use_q, qp, rc, L, a.kc.min_bq, canon, fwd)) continue;
                const uint64_t h = (canon * a.mul) & a.kmask;
                atomicAdd(&s_hist[(uint32_t)(h >> rem1_bits)], 1u);
            }
        }
        __syncthreads();
        if (warp == 0) {                                   // exclusive scan of 128 counts + one global reservation per bin
            uint32_t c[4], sum = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) { c[j] = s_hist[4 * lane + j]; sum += c[j]; }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
            uint32_t run = incl - sum;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t bin = 4u * lane + j;
                s_off[bin] = run; run += c[j];
                s_gbase[bin] = c[j] ? atomicAdd(&b.cursor1[bin], c[j]) : 0u;
                s_hist[bin] = 0u;
            }
            if (lane == 31) s_off[P1] = run;
        }
        __syncthreads();
        if (L) {
            for (uint32_t p = lane; p < L - K + 1u; p += 32u) {
                uint64_t canon; bool fwd;
                if (!kp_occurrence(packed, p, K, kmer_mask, split_mask, use_q, qp, rc, L, a.kc.min_bq, canon, fwd)) continue;
                const uint64_t h = (canon * a.mul) & a.kmask;
                const uint32_t bin = (uint32_t)(h >> rem1_bits);
                const uint32_t pos = s_off[bin] + atomicAdd(&s_hist[bin], 1u);
                sorted[pos] = (uint32_t)(h & rem1_mask) | (fwd ? 0x80000000u : 0u);
            }
        }
        __syncthreads();
        for (uint32_t bin = warp; bin < P1; bin += kKcWarps) {             // coalesced runs
            const uint32_t s0 = s_off[bin], cnt = s_off[bin + 1] - s0;
            uint32_t *dst = b.entries1 + __ldg(a.part_off + ((size_t)bin << sub_bits)) + s_gbase[bin];
            for (uint32_t i = lane; i < cnt; i += 32u) dst[i] = sorted[s0 + i];
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < P1; i += blockDim.x) s_hist[i] = 0u;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(512)
kmer_scatter2_kernel(Kp2Args b) {
    const KpArgs &a = b.p;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr uint32_t P1 = 1u << kKp1Bits;
    const int sub_bits = a.pbits - kKp1Bits;
    const uint32_t P2 = 1u << sub_bits;                                    // <= 64
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem);
    uint32_t *s_off = s_hist + 64;
    uint32_t *s_gbase = s_off + 65;
    uint32_t *sorted = s_gbase + 64;
    const uint32_t rem_mask = (uint32_t)((1ull << a.rem_bits) - 1ull);
    for (uint32_t t = blockIdx.x; t < b.n_tiles2; t += gridDim.x) {
        uint32_t lo = 0, hi = P1;                                          // level-1 bin of this tile
        while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(b.tile_start + mid) <= t) lo = mid; else hi = mid; }
        const uint32_t bin1 = lo;
        const unsigned long long b0 = __ldg(a.part_off + ((size_t)bin1 << sub_bits));
        const unsigned long long nb = __ldg(a.part_off + ((size_t)(bin1 + 1) << sub_bits)) - b0;
        const unsigned long long start = (unsigned long long)(t - __ldg(b.tile_start + bin1)) * kKp2Tile;
        const uint32_t cnt = (uint32_t)min((unsigned long long)kKp2Tile, nb - start);
        const uint32_t *src = b.entries1 + b0 + start;
        if (threadIdx.x < 64) s_hist[threadIdx.x] = 0u;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x)
            atomicAdd(&s_hist[((__ldg(src + i) & 0x7fffffffu) >> a.rem_bits) & (P2 - 1u)], 1u);
        __syncthreads();
        if (warp == 0) {
            uint32_t c[2], sum = 0;
#pragma unroll
            for (int j = 0; j < 2; ++j) { c[j] = s_hist[2 * lane + j]; sum += c[j]; }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
            uint32_t run = incl - sum;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t s = 2u * lane + j;
                s_off[s] = run; run += c[j];
                s_gbase[s] = (c[j] && s < P2) ? atomicAdd(&a.cursor[((size_t)bin1 << sub_bits) + s], c[j]) : 0u;
                s_hist[s] = 0u;
            }
            if (lane == 31) s_off[64] = run;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
            const uint32_t e = __ldg(src + i);
            const uint32_t s = ((e & 0x7fffffffu) >> a.rem_bits) & (P2 - 1u);
            sorted[s_off[s] + atomicAdd(&s_hist[s], 1u)] = (e & rem_mask) | (e & 0x80000000u);
        }
        __syncthreads();
        for (uint32_t s = warp; s < P2; s += 16u) {
            const uint32_t s0 = s_off[s], c = s_off[s + 1] - s0;
            // the final partition: local, or (sharded counting) the owning GPU's buffer -- the runs are written straight into
            // peer memory, so the exchange step of the multi-GPU algorithm is this store
            const size_t part = ((size_t)bin1 << sub_bits) + s;
            uint32_t *dst = (a.dst_ptr ? a.dst_ptr[part] : a.entries + __ldg(a.part_off + part)) + s_gbase[s];
            for (uint32_t i = lane; i < c; i += 32u) dst[i] = sorted[s0 + i];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024, 1)
kmer_count_part_kernel(KpArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *s_key = reinterpret_cast<uint32_t *>(smem);                 // remainder or 0xffffffff
    uint32_t *s_cnt = s_key + kKpSlots;                                   // [slot][2]
    __shared__ uint32_t s_fill, s_part, s_over;
    const uint32_t P = 1u << a.pbits;
    const int lane = threadIdx.x & 31;
    unsigned long long distinct = 0;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_part = atomicAdd(a.queue, 1u);
        __syncthreads();
        const uint32_t part = s_part;
        if (part >= P) break;
        const unsigned long long e0 = a.part_off[part], n = a.part_off[part + 1] - e0;
        if (n == 0) continue;
        // rounds: a slice of the remainders per round so that the distinct k-mers of a round fit the table (assumes >= 32
        // occurrences per distinct k-mer; whatever does not fit spills into the global table)
        uint32_t R = 1;
        while (R < 64u && (unsigned long long)R * kKpFillLimit * 32ull < n) R <<= 1;
        for (uint32_t r = 0; r < R; ++r) {
            for (uint32_t i = threadIdx.x; i < kKpSlots; i += blockDim.x) { s_key[i] = 0xffffffffu; s_cnt[2 * i] = 0u; s_cnt[2 * i + 1] = 0u; }
            if (threadIdx.x == 0) { s_fill = 0u; s_over = 0u; }
            __syncthreads();
            volatile uint32_t *vkey = s_key;
            // four independent entry loads per thread in flight (the kernel is bound by the latency of these streaming loads)
            for (unsigned long long i = threadIdx.x; i < n; i += 4ull * blockDim.x) {
                if (*(volatile uint32_t *)&s_over) break;
                uint32_t ev[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned long long j = i + (unsigned long long)u * blockDim.x;
                    ev[u] = j < n ? __ldg(a.entries + e0 + j) : 0xffffffffu;       // no entry has all remainder bits set
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t e = ev[u];
                    const uint32_t key = e & 0x7fffffffu, flag = e >> 31;
                    if (e == 0xffffffffu || (key & (R - 1u)) != r) continue;
                    uint32_t slot = (key * 0x9E3779B1u) >> 18;                // 14 bits
                    for (uint32_t probes = 0; probes < kKpSlots; ++probes) {
                        const uint32_t v = vkey[slot];
                        bool mine = v == key;
                        if (!mine && v == 0xffffffffu) {
                            const uint32_t old = atomicCAS(&s_key[slot], 0xffffffffu, key);
                            if (old == 0xffffffffu) { if (atomicAdd(&s_fill, 1u) >= a.smem_fill_limit) s_over = 1u; mine = true; }
                            else mine = old == key;
                        }
                        if (mine) { atomicAdd(&s_cnt[2u * slot + flag], 1u); break; }
                        slot = (slot + 1u) & (kKpSlots - 1u);
                    }
                }
            }
            __syncthreads();
            if (s_over) {
                // more distinct k-mers than the table holds: this round is counted in the global table instead (exact: the
                // shared-memory counts of the round are dropped as a whole)
                for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) {
                    const uint32_t e = __ldg(a.entries + e0 + i);
                    const uint32_t key = e & 0x7fffffffu;
                    if ((key & (R - 1u)) != r) continue;
                    const unsigned long long h = ((unsigned long long)part << a.rem_bits) | key;
                    kc_global_count(a.kc, (h * a.inv) & a.kmask, (e >> 31) != 0u);
                }
                __syncthreads();
                continue;
            }
            // survivors of the reference's filter (src/seq_parse.rs:30-44)
            for (uint32_t i0 = 0; i0 < kKpSlots; i0 += blockDim.x) {
                const uint32_t i = i0 + threadIdx.x;
                const uint32_t key = s_key[i], c0 = s_cnt[2 * i], c1 = s_cnt[2 * i + 1];
                const bool used = key != 0xffffffffu;
                const bool keep = used && (a.single_strand ? (c0 > 2u) : (c0 > 0u && c1 > 0u && c0 + c1 > 2u));
                if (used) ++distinct;
                const uint32_t bal = __ballot_sync(0xffffffffu, keep);
                if (bal) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(a.out_n, (uint32_t)__popc(bal));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (keep) {
                        const uint32_t o = base + (uint32_t)__popc(bal & ((1u << lane) - 1u));
                        if (o < a.out_cap) {
                            const unsigned long long h = ((unsigned long long)part << a.rem_bits) | key;
                            a.out_kmers[o] = (h * a.inv) & a.kmask; a.out_counts[2u * o] = c0; a.out_counts[2u * o + 1u] = c1;
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) distinct += __shfl_xor_sync(0xffffffffu, distinct, d);
    if (lane == 0 && distinct) atomicAdd(a.n_distinct, distinct);
}

static KpArgs kp_args(const KmerPartPlan &pl, const SeqBatch &in, const uint8_t *rc, int k, int min_bq) {
    KpArgs a{};
    a.kc.in = in; a.kc.rc = rc; a.kc.k = k; a.kc.min_bq = min_bq;
    a.kc.table = pl.table; a.kc.tshift = pl.tshift; a.kc.tmask = pl.tmask; a.kc.ctr = pl.ctr; a.kc.fill_limit = pl.fill_limit;
    a.mul = pl.mul; a.inv = pl.inv; a.kmask = pl.kmask; a.pbits = pl.pbits; a.rem_bits = pl.rem_bits;
    a.hist = pl.hist; a.cursor = pl.cursor; a.entries = pl.entries; a.part_off = pl.part_off; a.dst_ptr = pl.dst_ptr;
    a.single_strand = pl.single_strand; a.out_kmers = pl.out_kmers; a.out_counts = pl.out_counts; a.out_cap = pl.out_cap;
    a.out_n = pl.out_n; a.n_distinct = pl.n_distinct; a.queue = pl.queue;
    a.packed_words = ((in.max_len + 15u) >> 4) + 4u;
    a.smem_fill_limit = pl.smem_fill_limit ? pl.smem_fill_limit : kKpFillLimit;
    return a;
}

bool kmer_part_two_level(int k, uint32_t max_len);
// 32-bit entries need 2k - 13 + 1 <= 32; the scatter kernels keep whole reads in shared memory: up to 24576 bases with
// the two-level scatter (k <= 19), up to 9000 with the one-level scatter (k = 21)
bool kmer_part_supported(int k, uint32_t max_len) {
    if (!(2 * k - kKmerPartBits + 1 <= 32 && 2 * k > kKmerPartBits)) return false;
    return kmer_part_two_level(k, max_len) || max_len <= 9000u;
}

void launch_kmer_hist(const KmerPartPlan &pl, const SeqBatch &in, const uint8_t *rc, int k, int min_bq, int sm_count, cudaStream_t st) {
    if (in.n == 0) return;
    KpArgs a = kp_args(pl, in, rc, k, min_bq);
    const size_t smem = 512u + (4u << a.pbits) + (size_t)a.packed_words * 4u * kKcWarps;
    int ctas = (int)(200u * 1024u / (smem + 1024u));
    if (ctas < 1) ctas = 1;
    if (ctas > 4) ctas = 4;
    uint32_t grid = (uint32_t)(sm_count * ctas);
    const uint32_t need = (in.n + kKcWarps - 1) / kKcWarps;
    if (grid > need) grid = need;
    cudaFuncSetAttribute(kmer_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kmer_hist_kernel<<<grid, kKcWarps * 32, smem, st>>>(a);
}

void launch_kmer_scatter(const KmerPartPlan &pl, const SeqBatch &in, const uint8_t *rc, int k, int min_bq, int sm_count, cudaStream_t st) {
    if (in.n == 0) return;
    KpArgs a = kp_args(pl, in, rc, k, min_bq);
    const size_t smem = 512u + (8u << a.pbits) + (size_t)a.packed_words * 4u * kKcWarps * kKpTileReads;
    int ctas = (int)(220u * 1024u / (smem + 1024u));
    if (ctas < 1) ctas = 1;
    if (ctas > 4) ctas = 4;
    uint32_t grid = (uint32_t)(sm_count * ctas);
    const uint32_t tile = kKcWarps * kKpTileReads;
    const uint32_t need = (in.n + tile - 1) / tile;
    if (grid > need) grid = need;
    cudaFuncSetAttribute(kmer_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kmer_scatter_kernel<<<grid, kKcWarps * 32, smem, st>>>(a);
}

bool kmer_part_two_level(int k, uint32_t max_len) {
    return 2 * k - kKp1Bits + 1 <= 32 && max_len >= 1u && max_len <= 24000u && kKmerPartBits - kKp1Bits <= 6;
}

void launch_kmer_scatter1(const KmerPartPlan &pl, uint32_t *entries1, uint32_t *cursor1, const SeqBatch &in, const uint8_t *rc, int k,
                          int min_bq, int sm_count, cudaStream_t st) {
    if (in.n == 0) return;
    Kp2Args b{};
    b.p = kp_args(pl, in, rc, k, min_bq);
    b.entries1 = entries1; b.cursor1 = cursor1;
    b.tile_reads = kKp1TileCap / (in.max_len ? in.max_len : 1u);
    if (b.tile_reads > (uint32_t)kKcWarps) b.tile_reads = kKcWarps;
    const size_t smem = 512u + (size_t)(3u * (1u << kKp1Bits) + 1u) * 4u + (size_t)kKp1TileCap * 4u + (size_t)b.p.packed_words * 4u * kKcWarps;
    int ctas = (int)(220u * 1024u / (smem + 1024u));
    if (ctas < 1) ctas = 1;
    uint32_t grid = (uint32_t)(sm_count * ctas);
    const uint32_t need = (in.n + b.tile_reads - 1) / b.tile_reads;
    if (grid > need) grid = need;
    cudaFuncSetAttribute(kmer_scatter1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kmer_scatter1_kernel<<<grid, kKcWarps * 32, smem, st>>>(b);
}

uint32_t kmer_scatter2_tile() { return kKp2Tile; }
int kmer_scatter1_bits() { return kKp1Bits; }

void launch_kmer_scatter2(const KmerPartPlan &pl, const uint32_t *entries1, const uint32_t *tile_start, uint32_t n_tiles2, int k,
                          int sm_count, cudaStream_t st) {
    if (n_tiles2 == 0) return;
    Kp2Args b{};
    b.p = kp_args(pl, SeqBatch{}, nullptr, k, 0);
    b.entries1 = const_cast<uint32_t *>(entries1); b.tile_start = tile_start; b.n_tiles2 = n_tiles2;
    const size_t smem = (size_t)(64 + 65 + 64) * 4u + (size_t)kKp2Tile * 4u;
    uint32_t grid = (uint32_t)sm_count * 4u;
    if (grid > n_tiles2) grid = n_tiles2;
    cudaFuncSetAttribute(kmer_scatter2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kmer_scatter2_kernel<<<grid, 512, smem, st>>>(b);
}

void launch_kmer_count_part(const KmerPartPlan &pl, int sm_count, cudaStream_t st) {
    KpArgs a = kp_args(pl, SeqBatch{}, nullptr, 17, 0);
    const size_t smem = (size_t)kKpSlots * 12u;
    cudaFuncSetAttribute(kmer_count_part_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kmer_count_part_kernel<<<sm_count, 1024, smem, st>>>(a);
}


__global__ void kmer_table_clear_kernel(KmerSlot *table, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        KmerSlot s; s.key = kEmpty64; s.c0 = 0u; s.c1 = 0u;
        table[i] = s;
    }
}

// survivors of the reference's filter (src/seq_parse.rs:30-44) -> out (unordered, like the reference's hash-map order)
__global__ void kmer_compact_kernel(const KmerSlot *table, uint64_t n_slots, int single_strand, uint64_t *out_kmers,
                                    uint32_t *out_counts, uint32_t out_cap, uint32_t *out_n) {
    const int lane = threadIdx.x & 31;
    for (uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; i0 < n_slots; i0 += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = i0 + lane;
        bool keep = false;
        KmerSlot s; s.key = kEmpty64; s.c0 = 0; s.c1 = 0;
        if (i < n_slots) {
            s = table[i];
            if (s.key != kEmpty64) keep = single_strand ? (s.c0 > 2u) : (s.c0 > 0u && s.c1 > 0u && s.c0 + s.c1 > 2u);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (bal == 0u) continue;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(out_n, (uint32_t)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) {
            const uint32_t o = base + (uint32_t)__popc(bal & ((1u << lane) - 1u));
            if (o < out_cap) { out_kmers[o] = s.key; out_counts[2u * o] = s.c0; out_counts[2u * o + 1u] = s.c1; }
        }
    }
}

void launch_kmer_count(const SeqBatch &in, const uint8_t *rc, int k, int min_bq, KmerSlot *table, uint32_t tshift, uint32_t tmask,
                       uint32_t *ctr, uint32_t fill_limit, int sm_count, cudaStream_t st) {
    if (in.n == 0) return;
    KcArgs a{};
    a.in = in; a.rc = rc; a.k = k; a.min_bq = min_bq;
    a.table = table; a.tshift = tshift; a.tmask = tmask; a.ctr = ctr; a.fill_limit = fill_limit;
    a.packed_words = ((in.max_len + 15u) >> 4) + 4u;
    const size_t smem = 512u + (size_t)a.packed_words * 4u * kKcWarps;
    int ctas = (int)(200u * 1024u / (smem + 1024u));
    if (ctas < 1) ctas = 1;
    if (ctas > 4) ctas = 4;                                      // 64 warps per SM
    uint32_t grid = (uint32_t)(sm_count * ctas);
    const uint32_t need = (in.n + kKcWarps - 1) / kKcWarps;
    if (grid > need) grid = need;
    cudaFuncSetAttribute(kmer_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kmer_count_kernel<<<grid, kKcWarps * 32, smem, st>>>(a);
}

void launch_kmer_table_clear(KmerSlot *table, uint64_t n, cudaStream_t st) {
    if (n == 0) return;
    kmer_table_clear_kernel<<<148 * 8, 256, 0, st>>>(table, n);
}

void launch_kmer_compact(const KmerSlot *table, uint64_t n_slots, int single_strand, uint64_t *out_kmers, uint32_t *out_counts,
                         uint32_t out_cap, uint32_t *out_n, cudaStream_t st) {
    if (n_slots == 0) return;
    kmer_compact_kernel<<<148 * 8, 256, 0, st>>>(table, n_slots, single_strand, out_kmers, out_counts, out_cap, out_n);
}

}  // namespace sb
