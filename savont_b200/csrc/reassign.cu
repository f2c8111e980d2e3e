// reassign.cu -- Stage 3 reclustering: reads (or consensus lists) against the clusters' consensus SNPmers.
//
// Stands in for the inner loops of asv_cluster::reassign_reads_to_best_cluster (reference src/asv_cluster.rs:1052-1098) and
// compare_consensus_snpmers (:968-994).  The reference probes, per (read, cluster) pair, a per-cluster hash map with every
// marker of the read; here ONE inverted index over all clusters (splitmer -> bucket of (cluster, middle base), the layout of
// the ASV SNPmer index) is probed once per marker and the (matches, mismatches) of all clusters are accumulated in a dense
// shared-memory row: CTA per read, warp per marker, lanes over the bucket.
#include "common.cuh"

namespace sb {

constexpr int kCmpThreads = 128;

__global__ void __launch_bounds__(kCmpThreads) snpmer_compare_kernel(SnpCompareArgs a) {
    extern __shared__ uint32_t acc[];                 // [n_cons] matches (low 16 bits) | mismatches (high 16 bits)
    __shared__ unsigned long long s_best[kCmpThreads / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t split_mask = ~(3ull << (a.k - 1));
    for (uint32_t r = blockIdx.x; r < a.n_reads; r += gridDim.x) {
        for (uint32_t c = tid; c < a.n_cons; c += kCmpThreads) acc[c] = 0u;
        __syncthreads();
        const uint64_t m0 = a.read_offsets[r], m1 = a.read_offsets[r + 1];
        for (uint64_t i = m0 + warp; i < m1; i += kCmpThreads / 32) {
            const uint64_t kmer = __ldg(a.read_kmers + i);
            const uint64_t key = kmer & split_mask;
            const uint32_t allele = (uint32_t)(kmer >> (a.k - 1)) & 3u;
            if (a.idx_keys == nullptr) continue;
            uint32_t slot = hash_slot(key, a.idx_shift);
            uint2 range = make_uint2(0u, 0u);
            while (true) {                                         // warp-uniform probe
                const uint64_t kk = __ldg(a.idx_keys + slot);
                if (kk == key) { range = __ldg(a.idx_vals + slot); break; }
                if (kk == kEmpty64) break;
                slot = (slot + 1u) & a.idx_mask;
            }
            for (uint32_t e = lane; e < range.y; e += 32u) {
                const uint32_t v = __ldg(a.idx_entries + range.x + e);
                atomicAdd(&acc[v >> 2], (v & 3u) == allele ? 1u : 0x10000u);
            }
        }
        __syncthreads();
        // first cluster with the fewest mismatches, then the most matches (:1087-1091)
        unsigned long long best = ~0ull;
        uint32_t *row = a.matrix ? a.matrix + (size_t)(r - a.first_row) * a.n_cons : nullptr;
        for (uint32_t c = tid; c < a.n_cons; c += kCmpThreads) {
            const uint32_t v = acc[c];
            if (row) row[c] = v;
            const unsigned long long key = ((unsigned long long)(v >> 16) << 48) | ((unsigned long long)(0xffffu - (v & 0xffffu)) << 32) | c;
            best = key < best ? key : best;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = ((unsigned long long)__shfl_xor_sync(0xffffffffu, (uint32_t)(best >> 32), d) << 32) |
                                         __shfl_xor_sync(0xffffffffu, (uint32_t)best, d);
            best = o < best ? o : best;
        }
        if (lane == 0) s_best[warp] = best;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kCmpThreads / 32; ++w) best = s_best[w] < best ? s_best[w] : best;
            if (a.n_cons == 0) { a.best_cons[r] = 0xffffffffu; a.best_mismatches[r] = 0; a.best_matches[r] = 0; }
            else {
                a.best_cons[r] = (uint32_t)best; a.best_mismatches[r] = (uint32_t)(best >> 48);
                a.best_matches[r] = 0xffffu - (uint32_t)((best >> 32) & 0xffffu);
            }
        }
        __syncthreads();
    }
}

int launch_snpmer_compare(const SnpCompareArgs &a, uint32_t first, uint32_t n, int sm_count, cudaStream_t st) {
    if (n == 0) return 0;
    const size_t smem = (size_t)(a.n_cons ? a.n_cons : 1) * 4u;
    if (smem > 200u * 1024u) return 1;
    cudaFuncSetAttribute(snpmer_compare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SnpCompareArgs b = a;
    // reads [first, first + n): shift the per-read pointers, the matrix rows are relative to first_row
    b.read_offsets += first; b.best_cons += first; b.best_mismatches += first; b.best_matches += first;
    b.n_reads = n; b.first_row = 0;
    int per_sm = (int)((200u * 1024u) / (smem + 1024u));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16) per_sm = 16;
    uint32_t grid = (uint32_t)(sm_count * per_sm);
    if (grid > n) grid = n;
    snpmer_compare_kernel<<<grid, kCmpThreads, smem, st>>>(b);
    return 0;
}

}  // namespace sb
