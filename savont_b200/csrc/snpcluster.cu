// snpcluster.cu -- Stage 3 greedy SNPmer clustering inside the k-mer clusters.
//
// Stands in for the per-cluster loop of asv_cluster::cluster_reads_by_snpmers (reference src/asv_cluster.rs:561-700): the reads of a
// k-mer cluster are visited in order; a read joins the best representative it is compatible with (no SNPmer mismatch, at least
// one match), else it becomes a representative itself.
//   <= iterative_above representatives (find_best_representative_indexed, :467-510, blockmers off): the candidate with the most
//      matches, then the smallest cluster, then the smallest read id (the sort key (-matches, size, id) of :496-501);
//   above (find_best_representative_iterative, :413-464): the reference asks rayon for ANY compatible representative, which is
//      not deterministic; like the oracle this takes the FIRST one in creation order (what find_first would return).
//
// The loop looks serial, but only two things in it depend on the order: WHICH reads become representatives, and the cluster sizes
// that break ties.  Whether read i is compatible with representative r, and with how many matches, depends on the two lists alone.
// So the work is cut differently from the reference (round 2, third version; the first two walked the reads of a cluster one by
// one inside one CTA and were latency bound at 4 M reads/s):
//
//   rounds   (snp_round_kernel, the whole GPU, all clusters at once)  Round j knows the j-th representative of every cluster that
//            still has one (position rp in the cluster-major order).  Its SNPmers go into a shared-memory map (splitmer -> k-mer);
//            every read BEHIND rp is scored against it, 8 lanes per read, and the read's running state is updated: best match
//            count so far, the representatives that reach it (up to 6, in creation order), the first compatible representative.
//            The first read behind rp that still has no compatible representative is the cluster's next representative
//            (atomicMin) -- exactly the read the sequential loop would have promoted, because everything before it has found a
//            home among representatives 0..j and nothing created later can change that.  A read without SNPmers is always a
//            representative that nobody can join; those are marked up front and need no round.
//   resolve  (snp_resolve_kernel, one warp per cluster, 32 reads at a time, in order)  With the representatives known, a read's
//            winner is fixed unless several representatives tie on the match count; then the cluster sizes AT THAT READ decide.
//            The warp applies the size increments of the untied reads in parallel, in order segments between tied reads, and
//            resolves a tied read from the live sizes.  A read with more than 6 tied representatives is re-scored against all
//            representatives of its cluster by the warp (slow path, exact).
//
// The reference's inverted index answers "which representatives share this SNPmer"; the rounds answer the same question per
// representative, and with the splitmers of a read being distinct (src/seeding.rs:550-568 keeps only SNPmers whose splitmer occurs
// once; the gather kernel checks it) both count exactly the same (matches, mismatches).
#include <algorithm>

#include "common.cuh"

namespace sb {

constexpr int kSnpThreads = 256;
constexpr uint32_t kSnpMapMax = 1024;              // slots of a SNPmer map: lists of up to 512 SNPmers
constexpr uint32_t kSnpInf = 0xffffffffu;
constexpr uint32_t kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t snp_slot(const unsigned long long sp, const uint32_t smask) { return (uint32_t)((sp * kFibMul) >> 44) & smask; }
__device__ __forceinline__ uint32_t snp_map_slots(const uint32_t n) { uint32_t s = 64u; while (s < 2u * n) s <<= 1; return s; }

// per-tile: state reset, reads without SNPmers are representatives, the first read with SNPmers is the cluster's first "real" one
__global__ void __launch_bounds__(kSnpThreads) snp_init_kernel(SnpClusterArgs a) {
    __shared__ uint32_t s_min;
    const int tid = threadIdx.x;
    const uint32_t t = blockIdx.x, cl = a.tile_cl[t], p0 = a.tile_p0[t];
    const uint32_t c1 = (uint32_t)a.cluster_offsets[cl + 1], p1 = min(p0 + kSnpTile, c1);
    if (tid == 0) s_min = kSnpInf;
    __syncthreads();
    uint32_t mn = kSnpInf;
    for (uint32_t p = p0 + tid; p < p1; p += kSnpThreads) {
        const uint32_t nq = a.poff[p + 1] - a.poff[p];
        a.st_mt[p] = nq == 0u ? (1u << 30) : 0u; a.st_first[p] = kSnpInf;      // bits 30-31 of st_mt: 1 = representative without SNPmers, 2 = with
        if (nq > kSnpMapMax / 2u) atomicOr(a.err, 1u);
        else if (nq > 0u && mn == kSnpInf) mn = p;
    }
    mn = __reduce_min_sync(kFull, mn);
    if ((tid & 31) == 0 && mn != kSnpInf) atomicMin(&s_min, mn);
    __syncthreads();
    if (tid == 0 && s_min != kSnpInf) atomicMin(&a.next3[cl], s_min);
}

// round j: every read behind the cluster's j-th (non-empty) representative against that representative
__global__ void __launch_bounds__(kSnpThreads) snp_round_kernel(SnpClusterArgs a, const uint32_t j) {
    __shared__ unsigned long long map[kSnpMapMax];
    __shared__ uint32_t s_min;
    const int tid = threadIdx.x;
    const uint32_t t = blockIdx.x, cl = a.tile_cl[t], p0 = a.tile_p0[t], nc = a.n_clusters;
    const uint32_t c0 = (uint32_t)a.cluster_offsets[cl], c1 = (uint32_t)a.cluster_offsets[cl + 1], p1 = min(p0 + kSnpTile, c1);
    const uint32_t rp = a.next3[(j % 3u) * nc + cl];
    if (p0 == c0 && tid == 0) {                                    // the cluster's first tile keeps the three-deep ring of "next representative"
        a.next3[((j + 2u) % 3u) * nc + cl] = kSnpInf;
        if (rp != kSnpInf) { atomicOr(&a.st_mt[rp], 2u << 30); a.reps[c0 + j] = rp; atomicMax(a.last_round, (int)j); }     // the cluster's j-th representative with SNPmers
    }
    if (rp == kSnpInf || p1 <= rp + 1u) return;
    const unsigned long long mask = ~(3ull << (a.k - 1));
    const uint32_t r0 = a.poff[rp], nr = a.poff[rp + 1] - r0;
    if (nr > kSnpMapMax / 2u) return;                              // flagged by snp_init_kernel: the call fails
    const uint32_t slots = snp_map_slots(nr), smask = slots - 1u;
    for (uint32_t i = tid; i < slots; i += kSnpThreads) map[i] = kEmpty64;
    if (tid == 0) s_min = kSnpInf;
    __syncthreads();
    for (uint32_t i = tid; i < nr; i += kSnpThreads) {
        const unsigned long long k2 = a.pkmers[r0 + i], sp = k2 & mask;
        uint32_t s = snp_slot(sp, smask);
        while (true) {
            const unsigned long long old = atomicCAS(&map[s], (unsigned long long)kEmpty64, k2);
            if (old == kEmpty64 || (old & mask) == sp) break;       // distinct splitmers: checked by the gather kernel
            s = (s + 1u) & smask;
        }
    }
    __syncthreads();
    const uint32_t g = (uint32_t)tid >> 3, gl = (uint32_t)tid & 7u;
    uint32_t mn = kSnpInf;
    auto probe = [&](const unsigned long long k2) -> uint32_t {
        const unsigned long long sp = k2 & mask;
        uint32_t s = snp_slot(sp, smask);
        while (true) {
            const unsigned long long v = map[s];
            if (v == kEmpty64) return 0u;
            if ((v & mask) == sp) return v == k2 ? 1u : 0x10000u;
            s = (s + 1u) & smask;
        }
    };
    for (uint32_t base = max(p0, rp + 1u); base < p1; base += kSnpThreads / 8) {
        const uint32_t p = base + g;
        const bool valid = p < p1;
        uint32_t q0 = 0, nq = 0;
        if (valid) { q0 = a.poff[p]; nq = a.poff[p + 1] - q0; }
        // up to 32 SNPmers per read in one go (4 independent loads per lane), the rest in a loop
        unsigned long long kk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const uint32_t i = gl + 8u * u; kk[u] = i < nq ? a.pkmers[q0 + i] : kEmpty64; }
        uint32_t acc = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) if (kk[u] != kEmpty64) acc += probe(kk[u]);
        for (uint32_t i = gl + 32u; i < nq; i += 8u) acc += probe(a.pkmers[q0 + i]);
        acc += __shfl_xor_sync(kFull, acc, 1); acc += __shfl_xor_sync(kFull, acc, 2); acc += __shfl_xor_sync(kFull, acc, 4);
        if (valid && gl == 0u && nq > 0u) {
            const uint32_t m = acc & 0xffffu, mm = acc >> 16, mt = a.st_mt[p];
            uint32_t bm = mt & 0xffffu, nt = (mt >> 16) & 0xffu;
            if (mm == 0u && m > 0u) {
                if (bm == 0u) a.st_first[p] = j;
                if (m > bm) { bm = m; nt = 1u; a.st_tie[p] = j; }
                else if (m == bm) { if (nt < kSnpTies) a.st_tie[(size_t)nt * a.n_memb + p] = j; nt = min(nt + 1u, kSnpTies + 1u); }
                a.st_mt[p] = bm | (nt << 16) | (mt & 0xc0000000u);
            } else if (bm == 0u && mn == kSnpInf) mn = p;          // still homeless; a thread's positions increase
        }
    }
    mn = __reduce_min_sync(kFull, mn);
    if ((tid & 31) == 0 && mn != kSnpInf) atomicMin(&s_min, mn);
    __syncthreads();
    if (tid == 0 && s_min != kSnpInf) atomicMin(&a.next3[((j + 1u) % 3u) * nc + cl], s_min);
}

// the order-dependent rest: one warp per cluster, 32 reads per step.  Representatives are named by their ordinal among the cluster's
// representatives with SNPmers (= the round that scored them); sizes and read ids of the first kSnpSizeCache live in shared memory.
constexpr int kSnpResolveWarps = 4;
constexpr uint32_t kSnpSizeCache = 512;
constexpr uint32_t kSnpAhead = 4;                  // steps of read state in flight per warp (cp.async groups)
constexpr uint32_t kSnpResolveWarpBytes = kSnpMapMax * 8u + 2u * kSnpSizeCache * 4u + kSnpAhead * 5u * 32u * 4u;
__global__ void __launch_bounds__(kSnpResolveWarps * 32) snp_resolve_kernel(SnpClusterArgs a) {
    extern __shared__ __align__(16) uint8_t resolve_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *wb = resolve_smem + (size_t)warp * kSnpResolveWarpBytes;
    unsigned long long *qm = reinterpret_cast<unsigned long long *>(wb);                      // the slow path's read map
    uint32_t *ssz = reinterpret_cast<uint32_t *>(wb + kSnpMapMax * 8u), *srid = ssz + kSnpSizeCache;
    uint32_t *ring = srid + kSnpSizeCache;                                                     // [kSnpAhead][5 fields][32 lanes]
    const unsigned long long mask = ~(3ull << (a.k - 1));
    const uint32_t lt = (1u << lane) - 1u;
    while (true) {
        uint32_t oi = 0;
        if (lane == 0) oi = atomicAdd(a.work_ctr, 1u);
        oi = __shfl_sync(kFull, oi, 0);
        if (oi >= a.n_clusters) break;
        const uint32_t ci = a.order[oi];
        const uint32_t c0 = (uint32_t)a.cluster_offsets[ci], c1 = (uint32_t)a.cluster_offsets[ci + 1];
        const uint32_t *reps = a.reps + c0;                 // positions of the representatives with SNPmers, by ordinal (written by the rounds)
        uint32_t *gsize = a.rep_size + c0;                  // sizes beyond the shared-memory cache, by ordinal
        auto size_get = [&](const uint32_t w) -> uint32_t { return w < kSnpSizeCache ? ssz[w] : __ldcg(&gsize[w]); };
        auto size_inc = [&](const uint32_t w) { if (w < kSnpSizeCache) atomicAdd(&ssz[w], 1u); else atomicAdd(&gsize[w], 1u); };
        auto rid_get = [&](const uint32_t w) -> uint32_t { return w < kSnpSizeCache ? srid[w] : a.cluster_reads[reps[w]]; };
        uint32_t n_reps = 0, n_ne = 0;
        // The state of the reads is requested kSnpAhead steps ahead with cp.async into a shared-memory ring, one commit group per step.
        // (Register prefetching does not work here: a consumer waits on a scoreboard that the loads issued since share, so the loop ran
        // at the speed of the memory latency whatever the lead -- measured with 1 and 3 steps of lead.)  A lane reads back only what it
        // copied itself, so the ring needs no barrier.
        uint32_t step = 0;
        auto issue = [&](const uint32_t stage, const uint32_t pb) {
            const uint32_t p = pb + (uint32_t)lane;
            uint32_t *dst = ring + stage * 160u + (uint32_t)lane;
            if (p < c1) {
                cp_async4(dst, a.st_mt + p); cp_async4(dst + 32, a.st_first + p); cp_async4(dst + 64, a.cluster_reads + p);
                cp_async4(dst + 96, a.st_tie + p); cp_async4(dst + 128, a.st_tie + (size_t)a.n_memb + p);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (uint32_t i = 0; i < kSnpAhead; ++i) issue(i, c0 + 32u * i);
        long long tpa = 0, tpb = 0, tpc = 0, tpn3 = 0;
        for (uint32_t base = c0; base < c1; base += 32u) {
            const long long t0 = a.prof ? clock64() : 0;
            const uint32_t p = base + lane;
            const bool valid = p < c1;
            asm volatile("cp.async.wait_group %0;" ::"n"(kSnpAhead - 1) : "memory");
            const uint32_t stage = step % kSnpAhead;
            const uint32_t *src = ring + stage * 160u + (uint32_t)lane;
            uint32_t mt = 0, first = kSnpInf, read = 0, tie0 = kSnpInf, tie1 = kSnpInf;
            if (valid) { mt = src[0]; first = src[32]; read = src[64]; tie0 = src[96]; tie1 = src[128]; }
            const uint32_t isr = mt >> 30;
            issue(stage, base + 32u * kSnpAhead);
            ++step;
            const uint32_t rep_mask = __ballot_sync(kFull, valid && isr != 0u), ne_mask = __ballot_sync(kFull, valid && isr == 2u);
            const uint32_t my_nreps = n_reps + (uint32_t)__popc(rep_mask & lt);          // representatives created before this read
            const uint32_t my_ne = n_ne + (uint32_t)__popc(ne_mask & lt);                // ... of those, with SNPmers
            const uint32_t nt = (mt >> 16) & 0xffu;
            bool need = false;
            uint32_t winner = kSnpInf;
            if (valid) {
                if (isr != 0u) {                                                                               // src/asv_cluster.rs:397-410
                    a.rep_of[read] = read;
                    if (isr == 2u) { if (my_ne < kSnpSizeCache) { ssz[my_ne] = 1u; srid[my_ne] = read; } else gsize[my_ne] = 1u; }
                }
                else if ((mt & 0xffffu) == 0u) atomicOr(a.err, 4u);                                            // cannot happen: a homeless read is a representative
                else if (my_nreps > a.iterative_above) winner = first;
                else if (nt == 1u) winner = tie0;
                else need = true;
            }
            __syncwarp();
            // a two-way tie whose sizes are 32 or more apart cannot be turned by the (at most 31) reads of this step before it
            if (need && nt == 2u) {
                const uint32_t sz0 = size_get(tie0), sz1 = size_get(tie1);
                if (sz0 + 32u <= sz1) { winner = tie0; need = false; }
                else if (sz1 + 32u <= sz0) { winner = tie1; need = false; }
            }
            uint32_t need_mask = __ballot_sync(kFull, need);
            uint32_t lo = 0;
            const long long t1 = a.prof ? clock64() : 0;
            if (a.prof) tpn3 += __popc(__ballot_sync(kFull, need && nt >= 3u));
            while (true) {
                const uint32_t L = need_mask ? (uint32_t)__ffs(need_mask) - 1u : 32u;
                if (winner != kSnpInf && (uint32_t)lane >= lo && (uint32_t)lane < L) { size_inc(winner); a.rep_of[read] = rid_get(winner); }   // :386-394
                if (L == 32u) break;
                __syncwarp();
                const uint32_t ntL = __shfl_sync(kFull, nt, L);
                if (a.prof && (uint32_t)lane == L) atomicAdd(a.work_ctr + (ntL <= kSnpTies ? 4 : 5), 1u);       // statistics (SAVONT_B200_SNPCLUSTER_PROF)
                if (ntL <= kSnpTies) {
                    if ((uint32_t)lane == L) {
                        uint32_t bw = kSnpInf, bs = 0, br = 0;
                        for (uint32_t i = 0; i < nt; ++i) {
                            const uint32_t w = i == 0u ? tie0 : i == 1u ? tie1 : a.st_tie[(size_t)i * a.n_memb + p];
                            const uint32_t sz = size_get(w), rid = rid_get(w);
                            if (bw == kSnpInf || sz < bs || (sz == bs && rid < br)) { bw = w; bs = sz; br = rid; }
                        }
                        size_inc(bw); a.rep_of[read] = br;
                    }
                } else {
                    // more tied representatives than the state holds: the read against every representative before it, by the whole warp
                    const uint32_t pL = base + L, neL = __shfl_sync(kFull, my_ne, L), readL = __shfl_sync(kFull, read, L);
                    const uint32_t q0 = a.poff[pL], nq = min(a.poff[pL + 1] - a.poff[pL], kSnpMapMax / 2u);       // longer lists: the call fails (snp_init_kernel)
                    const uint32_t slots = snp_map_slots(nq), smask = slots - 1u;
                    for (uint32_t i = lane; i < slots; i += 32u) qm[i] = kEmpty64;
                    __syncwarp();
                    for (uint32_t i = lane; i < nq; i += 32u) {
                        const unsigned long long k2 = a.pkmers[q0 + i], sp = k2 & mask;
                        uint32_t s = snp_slot(sp, smask);
                        while (true) {
                            const unsigned long long old = atomicCAS(&qm[s], (unsigned long long)kEmpty64, k2);
                            if (old == kEmpty64 || (old & mask) == sp) break;
                            s = (s + 1u) & smask;
                        }
                    }
                    __syncwarp();
                    uint32_t bw = kSnpInf, bm = 0, bs = 0, br = 0;
                    for (uint32_t r = 0; r < neL; ++r) {
                        const uint32_t rpos = reps[r];
                        const uint32_t e0 = a.poff[rpos], e1 = a.poff[rpos + 1];
                        uint32_t m = 0, mm = 0;
                        for (uint32_t eb = e0; eb < e1 && mm == 0u; eb += 32u) {
                            const uint32_t e = eb + lane;
                            bool hit = false, eq = false;
                            if (e < e1) {
                                const unsigned long long k2 = a.pkmers[e], sp = k2 & mask;
                                uint32_t s = snp_slot(sp, smask);
                                while (true) {
                                    const unsigned long long v = qm[s];
                                    if (v == kEmpty64) break;
                                    if ((v & mask) == sp) { hit = true; eq = v == k2; break; }
                                    s = (s + 1u) & smask;
                                }
                            }
                            m += (uint32_t)__popc(__ballot_sync(kFull, hit && eq));
                            mm += (uint32_t)__popc(__ballot_sync(kFull, hit && !eq));
                        }
                        if (mm == 0u && m > 0u) {
                            const uint32_t sz = size_get(r), rid = rid_get(r);
                            if (bw == kSnpInf || m > bm || (m == bm && (sz < bs || (sz == bs && rid < br)))) { bw = r; bm = m; bs = sz; br = rid; }
                        }
                    }
                    if (bw == kSnpInf) { if (lane == 0) atomicOr(a.err, 4u); }
                    else if (lane == 0) { size_inc(bw); a.rep_of[readL] = br; }
                }
                __syncwarp();
                need_mask &= need_mask - 1u; lo = L + 1u;
            }
            __syncwarp();
            n_reps += (uint32_t)__popc(rep_mask); n_ne += (uint32_t)__popc(ne_mask);
            const long long t2 = a.prof ? clock64() : 0;
            tpa += t1 - t0; tpb += t2 - t1;
        }
        if (a.prof && oi == 0 && lane == 0) { unsigned long long *pr = reinterpret_cast<unsigned long long *>(a.work_ctr + 8); pr[0] = tpa; pr[1] = tpb; pr[2] = tpn3; pr[3] = c1 - c0; }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (lane == 0) atomicAdd(a.total_reps, (unsigned long long)n_reps);
    }
}

// cluster-major copy of the SNPmer lists: position p of the concatenated clusters gets the list of read cluster_reads[p]
__global__ void snp_cluster_count_kernel(SnpClusterArgs a, uint32_t n_memb, uint32_t *cnt) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_memb; p += gridDim.x * blockDim.x) {
        const uint32_t read = a.cluster_reads[p];
        cnt[p] = (uint32_t)(a.snp_offsets[read + 1] - a.snp_offsets[read]);
    }
}
// ... and the check that the splitmers of a list are distinct (err bit 2)
__global__ void snp_cluster_gather_kernel(SnpClusterArgs a, uint32_t n_memb, unsigned long long *dst) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const unsigned long long mask = ~(3ull << (a.k - 1));
    for (uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n_memb; p += warps) {
        const uint32_t read = a.cluster_reads[p];
        const uint64_t s0 = a.snp_offsets[read];
        const uint32_t n = a.poff[p + 1] - a.poff[p], d0 = a.poff[p];
        bool dup = false;
        for (uint32_t ib = 0; ib < n; ib += 32u) {
            const uint32_t i = ib + lane;
            unsigned long long sp = 0xffffffffffffff00ull | (unsigned)lane;                  // no splitmer looks like this
            if (i < n) { const unsigned long long k2 = a.snp_kmers[s0 + i]; dst[d0 + i] = k2; sp = k2 & mask; }
            __syncwarp();
            if (__match_any_sync(kFull, sp) != (1u << lane)) dup = true;
            if (ib > 0u && i < n) for (uint32_t e = 0; e < ib; ++e) if ((dst[d0 + e] & mask) == sp) dup = true;      // lists beyond 32: rare
        }
        if (dup) atomicOr(a.err, 2u);
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

void launch_snp_init(const SnpClusterArgs &a, cudaStream_t st) {
    if (a.n_tiles) snp_init_kernel<<<a.n_tiles, kSnpThreads, 0, st>>>(a);
}
void launch_snp_round(const SnpClusterArgs &a, uint32_t j, cudaStream_t st) {
    if (a.n_tiles) snp_round_kernel<<<a.n_tiles, kSnpThreads, 0, st>>>(a, j);
}
void launch_snp_resolve(const SnpClusterArgs &a, int sm_count, cudaStream_t st) {
    if (a.n_clusters == 0) return;
    const uint32_t want = (a.n_clusters + kSnpResolveWarps - 1u) / kSnpResolveWarps;
    const size_t smem = (size_t)kSnpResolveWarps * kSnpResolveWarpBytes;
    cudaFuncSetAttribute(snp_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    snp_resolve_kernel<<<std::min<uint32_t>(want, (uint32_t)sm_count * 3u), kSnpResolveWarps * 32, smem, st>>>(a);
}

}  // namespace sb
