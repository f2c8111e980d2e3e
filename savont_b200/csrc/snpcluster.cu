// snpcluster.cu -- Stage 3 greedy SNPmer clustering inside the k-mer clusters.
//
// Stands in for the per-cluster loop of asv_cluster::cluster_reads_by_snpmers (reference src/asv_cluster.rs:561-700): the reads of a
// k-mer cluster are visited in order; a read joins the best representative it is compatible with (no SNPmer mismatch, at least
// one match), else it becomes a representative itself.
//   <= iterative_above representatives (find_best_representative_indexed, :467-510, blockmers off): the candidate with the most
//      matches, then the smallest cluster, then the smallest read id (the sort key (-matches, size, id) of :496-501);
//   above (find_best_representative_iterative, :413-464): the reference asks rayon for ANY compatible representative, which is
//      not deterministic; like the oracle this takes the FIRST one in creation order (what find_first would return).
// The k-mer clusters are independent (the reference runs them in a par_iter, :593): one CTA per cluster, largest first, the
// reads of a cluster strictly in sequence.  Per read the CTA puts the read's SNPmers into a shared-memory map (splitmer -> k-mer)
// and every thread scores one representative against it by walking that representative's own SNPmer list (the lists are the
// input CSR: nothing is copied, and the hot ones stay in L1 / L2); a block-wide lexicographic reduction picks the winner.  This is
// the reference's inverted index turned around: the index answers "which representatives share this SNPmer", the walk answers the
// same question per representative, and with the splitmers of a read being distinct (src/seeding.rs:550-568 keeps only SNPmers
// whose splitmer occurs once) both count exactly the same (matches, mismatches).
#include "common.cuh"

namespace sb {

constexpr int kSnpClusterThreads = 256;
constexpr uint32_t kSnpClusterMap = 1024;          // slots of the query map: reads with up to 512 SNPmers

struct SnpClusterCand { uint32_t m, size, rid, r; };        // m == 0: none

__device__ __forceinline__ bool snpcluster_better(const SnpClusterCand &a, const SnpClusterCand &b, bool iterative) {
    if (a.m == 0u) return false;
    if (b.m == 0u) return true;
    if (iterative) return a.r < b.r;
    if (a.m != b.m) return a.m > b.m;
    if (a.size != b.size) return a.size < b.size;
    return a.rid < b.rid;
}

__global__ void __launch_bounds__(kSnpClusterThreads)
snp_cluster_kernel(SnpClusterArgs a) {
    __shared__ unsigned long long qmap[kSnpClusterMap];
    __shared__ SnpClusterCand warp_best[kSnpClusterThreads / 32];
    __shared__ uint32_t s_next, s_nreps;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t mask = ~(3ull << (a.k - 1));
    while (true) {
        if (tid == 0) s_next = atomicAdd(a.work_ctr, 1u);
        __syncthreads();
        const uint32_t oi = s_next;
        if (oi >= a.n_clusters) break;
        const uint32_t ci = a.order[oi];
        const uint64_t c0 = a.cluster_offsets[ci], c1 = a.cluster_offsets[ci + 1];
        uint32_t *reps = a.reps + c0, *rep_size = a.rep_size + c0;
        uint32_t n_reps = 0;
        for (uint64_t pos = c0; pos < c1; ++pos) {
            const uint32_t read = a.cluster_reads[pos];
            const uint64_t q0 = a.snp_offsets[read], q1 = a.snp_offsets[read + 1];
            const uint32_t nq = (uint32_t)(q1 - q0);
            // ---- the read's SNPmers -> map (splitmer -> k-mer) ----
            for (uint32_t i = tid; i < kSnpClusterMap; i += kSnpClusterThreads) qmap[i] = kEmpty64;
            __syncthreads();
            if (nq > kSnpClusterMap / 2u) { if (tid == 0) atomicOr(a.err, 1u); }
            else {
                for (uint32_t i = tid; i < nq; i += kSnpClusterThreads) {
                    const unsigned long long km = a.snp_kmers[q0 + i];
                    const unsigned long long sp = km & mask;
                    uint32_t slot = (uint32_t)((sp * kFibMul) >> 54) & (kSnpClusterMap - 1u);
                    while (true) {
                        const unsigned long long old = atomicCAS(&qmap[slot], (unsigned long long)kEmpty64, km);
                        if (old == kEmpty64) break;
                        if ((old & mask) == sp) { atomicOr(a.err, 2u); break; }      // the splitmers of a read must be distinct
                        slot = (slot + 1u) & (kSnpClusterMap - 1u);
                    }
                }
            }
            __syncthreads();
            // ---- every thread scores representatives r = tid, tid + 256, ... ----
            const bool iterative = n_reps > a.iterative_above;
            SnpClusterCand best{0u, 0u, 0u, 0u};
            if (nq > 0u) {
                for (uint32_t r = tid; r < n_reps; r += kSnpClusterThreads) {
                    const uint32_t rid = reps[r];
                    const uint64_t e0 = a.snp_offsets[rid], e1 = a.snp_offsets[rid + 1];
                    uint32_t m = 0, mm = 0;
                    for (uint64_t e = e0; e < e1 && mm == 0u; ++e) {
                        const unsigned long long km = __ldg(a.snp_kmers + e);
                        const unsigned long long sp = km & mask;
                        uint32_t slot = (uint32_t)((sp * kFibMul) >> 54) & (kSnpClusterMap - 1u);
                        while (true) {
                            const unsigned long long v = qmap[slot];
                            if (v == kEmpty64) break;
                            if ((v & mask) == sp) { if (v == km) ++m; else ++mm; break; }
                            slot = (slot + 1u) & (kSnpClusterMap - 1u);
                        }
                    }
                    if (mm == 0u && m > 0u) {
                        const SnpClusterCand cnd{m, rep_size[r], rid, r};
                        if (snpcluster_better(cnd, best, iterative)) best = cnd;
                    }
                }
            }
            // ---- block-wide winner ----
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                SnpClusterCand o;
                o.m = __shfl_xor_sync(0xffffffffu, best.m, d); o.size = __shfl_xor_sync(0xffffffffu, best.size, d);
                o.rid = __shfl_xor_sync(0xffffffffu, best.rid, d); o.r = __shfl_xor_sync(0xffffffffu, best.r, d);
                if (snpcluster_better(o, best, iterative)) best = o;
            }
            if (lane == 0) warp_best[warp] = best;
            __syncthreads();
            if (tid == 0) {
                SnpClusterCand w = warp_best[0];
                for (int i = 1; i < kSnpClusterThreads / 32; ++i) if (snpcluster_better(warp_best[i], w, iterative)) w = warp_best[i];
                if (w.m > 0u) { a.rep_of[read] = w.rid; rep_size[w.r] += 1u; }                      // src/asv_cluster.rs:386-394
                else { reps[n_reps] = read; rep_size[n_reps] = 1u; a.rep_of[read] = read; s_nreps = n_reps + 1u; }   // :397-410
                if (w.m > 0u) s_nreps = n_reps;
            }
            __syncthreads();
            n_reps = s_nreps;
        }
        if (tid == 0) atomicAdd(a.total_reps, (unsigned long long)n_reps);
        __syncthreads();
    }
}

void launch_snp_cluster(const SnpClusterArgs &a, int sm_count, cudaStream_t st) {
    if (a.n_clusters == 0) return;
    uint32_t grid = (uint32_t)sm_count * 4u;
    if (grid > a.n_clusters) grid = a.n_clusters;
    snp_cluster_kernel<<<grid, kSnpClusterThreads, 0, st>>>(a);
}

}  // namespace sb
