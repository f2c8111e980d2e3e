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
// and every WARP scores one representative against it, the lanes taking that representative's own SNPmers (the lists are the
// input CSR: nothing is copied, and the hot ones stay in L1 / L2); every thread then picks the winner of the 8 per-warp candidates
// by the same lexicographic rule.  The lists and sizes of the first 64 representatives live in shared memory, the read's map is double
// buffered and emptied slot by slot by the threads that filled it, the next read's list is loaded one read ahead: two block barriers
// and no global-memory round trip on the critical path per read.  This is
// the reference's inverted index turned around: the index answers "which representatives share this SNPmer", the walk answers the
// same question per representative, and with the splitmers of a read being distinct (src/seeding.rs:550-568 keeps only SNPmers
// whose splitmer occurs once) both count exactly the same (matches, mismatches).
#include <algorithm>

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

constexpr uint32_t kSnpCacheReps = 64;             // representatives whose SNPmer list and size live in shared memory
constexpr uint32_t kSnpCacheLen = 32;              // ... if the list has at most this many SNPmers

__global__ void __launch_bounds__(kSnpClusterThreads)
snp_cluster_kernel(SnpClusterArgs a) {
    __shared__ unsigned long long qmap[2][kSnpClusterMap];                 // the read's SNPmers, double buffered (even / odd reads)
    __shared__ unsigned long long rep_km[kSnpCacheReps][kSnpCacheLen];
    __shared__ uint32_t rep_n[kSnpCacheReps], rep_sz[kSnpCacheReps], rep_rid[kSnpCacheReps];
    __shared__ SnpClusterCand warp_best[kSnpClusterThreads / 32];
    __shared__ uint32_t s_next;
    constexpr int kWarps = kSnpClusterThreads / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t mask = ~(3ull << (a.k - 1));
    for (uint32_t i = tid; i < 2u * kSnpClusterMap; i += kSnpClusterThreads) (&qmap[0][0])[i] = kEmpty64;
    while (true) {
        __syncthreads();
        if (tid == 0) s_next = atomicAdd(a.work_ctr, 1u);
        __syncthreads();
        const uint32_t oi = s_next;
        if (oi >= a.n_clusters) break;
        const uint32_t ci = a.order[oi];
        const uint64_t c0 = a.cluster_offsets[ci], c1 = a.cluster_offsets[ci + 1];
        uint32_t *reps = a.reps + c0, *rep_size = a.rep_size + c0;
        uint32_t n_reps = 0;                                     // every thread keeps the same copy (the winner is computed by all)
        // the SNPmers of the read about to be processed, one or two per thread, loaded one read ahead
        uint32_t q0 = 0, nq = 0;
        unsigned long long km[2] = {0ull, 0ull};
        auto load_read = [&](const uint64_t pos) {
            q0 = a.poff[pos]; nq = a.poff[pos + 1] - q0;
#pragma unroll
            for (int u = 0; u < 2; ++u) { const uint32_t i = tid + u * kSnpClusterThreads; if (i < nq && nq <= kSnpClusterMap / 2u) km[u] = a.pkmers[q0 + i]; }
        };
        if (c0 < c1) load_read(c0);
        long long tprof[5] = {0, 0, 0, 0, 0};
        for (uint64_t pos = c0; pos < c1; ++pos) {
            const long long tp0 = a.prof ? clock64() : 0;
            unsigned long long (&qm)[kSnpClusterMap] = qmap[pos & 1];
            const uint32_t read = a.cluster_reads[pos];
            const uint32_t nq_cur = nq;
            const unsigned long long km_cur[2] = {km[0], km[1]};
            // ---- 1. the read's SNPmers -> map (splitmer -> k-mer); every inserting thread remembers its slots and empties them again
            //         after the scoring, so a map is never cleared as a whole ----
            uint32_t my_slot[2] = {0xffffffffu, 0xffffffffu};
            if (nq_cur > kSnpClusterMap / 2u) { if (tid == 0) atomicOr(a.err, 1u); }
            else {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const uint32_t i = tid + u * kSnpClusterThreads;
                    if (i >= nq_cur) break;
                    const unsigned long long sp = km_cur[u] & mask;
                    uint32_t slot = (uint32_t)((sp * kFibMul) >> 54) & (kSnpClusterMap - 1u);
                    while (true) {
                        const unsigned long long old = atomicCAS(&qm[slot], (unsigned long long)kEmpty64, km_cur[u]);
                        if (old == kEmpty64) { my_slot[u] = slot; break; }
                        if ((old & mask) == sp) { atomicOr(a.err, 2u); break; }      // the splitmers of a read must be distinct
                        slot = (slot + 1u) & (kSnpClusterMap - 1u);
                    }
                }
            }
            // the next read's list is requested now and consumed in the next iteration (the chain is latency bound); further ahead
            // the lists are pulled into L2
            if (pos + 1 < c1) load_read(pos + 1);
            if (tid == 0) asm volatile("prefetch.global.L2 [%0];" :: "l"(a.pkmers + q0 + 512));
            if (tid == 32) asm volatile("prefetch.global.L2 [%0];" :: "l"(a.poff + pos + 64));
            if (tid == 64) asm volatile("prefetch.global.L2 [%0];" :: "l"(a.cluster_reads + pos + 64));
            const long long tp1 = a.prof ? clock64() : 0;
            __syncthreads();
            const long long tp2 = a.prof ? clock64() : 0;
            // ---- 2. one WARP per representative (r = warp, warp + 8, ...): the lanes take the representative's SNPmers, from shared
            //         memory for the first 64 representatives, from the cluster-major lists otherwise ----
            const bool iterative = n_reps > a.iterative_above;
            SnpClusterCand best{0u, 0u, 0u, 0u};
            if (nq_cur > 0u && nq_cur <= kSnpClusterMap / 2u) {
                for (uint32_t r = warp; r < n_reps; r += kWarps) {
                    const bool cached = r < kSnpCacheReps && rep_n[r] != 0xffffffffu;
                    uint32_t e0 = 0, e1 = 0, rpos = 0;
                    if (cached) e1 = rep_n[r];
                    else { rpos = reps[r]; e0 = a.poff[rpos]; e1 = a.poff[rpos + 1]; }
                    uint32_t m = 0, mm = 0;
                    for (uint32_t eb = e0; eb < e1 && mm == 0u; eb += 32u) {
                        const uint32_t e = eb + lane;
                        bool hit = false, eq = false;
                        if (e < e1) {
                            const unsigned long long k2 = cached ? rep_km[r][e] : a.pkmers[e];
                            const unsigned long long sp = k2 & mask;
                            uint32_t slot = (uint32_t)((sp * kFibMul) >> 54) & (kSnpClusterMap - 1u);
                            while (true) {
                                const unsigned long long v = qm[slot];
                                if (v == kEmpty64) break;
                                if ((v & mask) == sp) { hit = true; eq = v == k2; break; }
                                slot = (slot + 1u) & (kSnpClusterMap - 1u);
                            }
                        }
                        m += (uint32_t)__popc(__ballot_sync(0xffffffffu, hit && eq));
                        mm += (uint32_t)__popc(__ballot_sync(0xffffffffu, hit && !eq));
                    }
                    if (mm == 0u && m > 0u) {
                        const SnpClusterCand cnd{m, r < kSnpCacheReps ? rep_sz[r] : rep_size[r], r < kSnpCacheReps ? rep_rid[r] : a.cluster_reads[reps[r]], r};
                        if (snpcluster_better(cnd, best, iterative)) best = cnd;
                    }
                }
            }
            if (lane == 0) warp_best[warp] = best;
            const long long tp3 = a.prof ? clock64() : 0;
            __syncthreads();
            const long long tp4 = a.prof ? clock64() : 0;
            // ---- 3. the winner, computed by every thread from the 8 per-warp candidates; bookkeeping; the map is emptied.  No barrier
            //         behind this: the next read uses the other map, and everything written here is read after its first barrier ----
            SnpClusterCand w = warp_best[0];
#pragma unroll
            for (int i = 1; i < kWarps; ++i) { const SnpClusterCand o = warp_best[i]; if (snpcluster_better(o, w, iterative)) w = o; }
            if (w.m > 0u) {                                                                            // src/asv_cluster.rs:386-394
                if (tid == 0) { a.rep_of[read] = w.rid; if (w.r < kSnpCacheReps) rep_sz[w.r] += 1u; else rep_size[w.r] += 1u; }
            } else {                                                                                   // :397-410
                if (tid == 0) {
                    reps[n_reps] = (uint32_t)pos; a.rep_of[read] = read;
                    if (n_reps < kSnpCacheReps) { rep_sz[n_reps] = 1u; rep_rid[n_reps] = read; rep_n[n_reps] = nq_cur <= kSnpCacheLen ? nq_cur : 0xffffffffu; }
                    else rep_size[n_reps] = 1u;
                }
                if (n_reps < kSnpCacheReps && nq_cur <= kSnpCacheLen && (uint32_t)tid < nq_cur) rep_km[n_reps][tid] = km_cur[0];
                ++n_reps;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) if (my_slot[u] != 0xffffffffu) qm[my_slot[u]] = kEmpty64;
            if (a.prof) { const long long tp5 = clock64(); tprof[0] += tp1 - tp0; tprof[1] += tp2 - tp1; tprof[2] += tp3 - tp2; tprof[3] += tp4 - tp3; tprof[4] += tp5 - tp4; }
        }
        if (a.prof && tid == 0 && oi == 0) { for (int i = 0; i < 5; ++i) a.prof[i] = (unsigned long long)tprof[i]; a.prof[5] = c1 - c0; a.prof[6] = n_reps; }
        if (tid == 0) atomicAdd(a.total_reps, (unsigned long long)n_reps);
    }
}

// cluster-major copy of the SNPmer lists: position p of the concatenated clusters gets the list of read cluster_reads[p]
__global__ void snp_cluster_count_kernel(SnpClusterArgs a, uint32_t n_memb, uint32_t *cnt) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_memb; p += gridDim.x * blockDim.x) {
        const uint32_t read = a.cluster_reads[p];
        cnt[p] = (uint32_t)(a.snp_offsets[read + 1] - a.snp_offsets[read]);
    }
}
__global__ void snp_cluster_gather_kernel(SnpClusterArgs a, uint32_t n_memb, unsigned long long *dst) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n_memb; p += warps) {
        const uint32_t read = a.cluster_reads[p];
        const uint64_t s0 = a.snp_offsets[read];
        const uint32_t n = a.poff[p + 1] - a.poff[p], d0 = a.poff[p];
        for (uint32_t i = lane; i < n; i += 32u) dst[d0 + i] = a.snp_kmers[s0 + i];
    }
}
void launch_snp_cluster_count(const SnpClusterArgs &a, uint32_t n_memb, uint32_t *cnt, cudaStream_t st) {
    if (n_memb == 0) return;
    snp_cluster_count_kernel<<<std::min<uint32_t>((n_memb + 255u) / 256u, 148u * 8u), 256, 0, st>>>(a, n_memb, cnt);
}
void launch_snp_cluster_gather(const SnpClusterArgs &a, uint32_t n_memb, unsigned long long *dst, cudaStream_t st) {
    if (n_memb == 0) return;
    snp_cluster_gather_kernel<<<std::min<uint32_t>((n_memb + 7u) / 8u, 148u * 16u), 256, 0, st>>>(a, n_memb, dst);
}

void launch_snp_cluster(const SnpClusterArgs &a, int sm_count, cudaStream_t st) {
    if (a.n_clusters == 0) return;
    uint32_t grid = (uint32_t)sm_count * 4u;
    if (grid > a.n_clusters) grid = a.n_clusters;
    snp_cluster_kernel<<<grid, kSnpClusterThreads, 0, st>>>(a);
}

}  // namespace sb
