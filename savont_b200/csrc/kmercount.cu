// kmercount.cu -- K5: Stage 1 split-k-mer counting, the integer part of seq_parse::read_to_split_kmers
// (reference src/seq_parse.rs:12-83, :315-497, bloom_filter_size == 0) over seeding::split_kmer_mid
// (src/seeding.rs:975-1068).
//
// The reference counts every k-mer of every read in sharded FxHashMap<u64, [u32; 2]> and afterwards keeps the entries
// seen on both strands more than twice (src/seq_parse.rs:30-44).  Here ONE pass over the reads counts every occurrence
// in one open-addressing table in HBM (16-byte slots: key, c0, c1; two slots per 32-byte sector), sized from the number
// of occurrences (load <= 0.5, regrown x4 and re-run otherwise): per occurrence one key probe and one atomicAdd into the
// same sector.  A 16S read set is extremely redundant -- 1 M ONT reads hold 1.48 G occurrences of 48 M distinct k-mers,
// two thirds of the occurrences belong to < 1 M true k-mers -- so the hot sectors stay in L2 and only the error k-mers
// pay a DRAM round trip.  (A first version with a Bloom filter in front of the table and a second counting pass touched
// two structures per occurrence and took twice as long.)  Measured on 1 M reads: extraction alone 5.5 ms, + key probes
// 42.7 ms, + count atomics 81.7 ms: the kernel is bound by the rate of scattered 32-byte L2 / DRAM operations
// (~40 G/s for loads as well as for reductions; 41.7 GB of DRAM reads = one 64-byte burst per error k-mer occurrence),
// not by issue slots or memory-level parallelism (4 probes in flight per lane change nothing).  Partitioning the k-mers
// by hash first, so that every partition is counted in shared memory, is the round-2 design.
// A final scan applies the reference's filter and appends the survivors to the output.
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
