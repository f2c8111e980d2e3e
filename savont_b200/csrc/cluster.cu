// cluster.cu -- Stage 2 greedy k-mer clustering: the device side of savont_b200_lsh_signatures and
// savont_b200_cluster_reads_by_kmers.
//
// Stands in for TwinRead::compute_lsh_signatures (reference src/types.rs:719-747) and for the per-read work of
// asv_cluster::cluster_reads_by_kmers (src/asv_cluster.rs:72-190): bucket query of the 20 LSH tables, candidate window
// (all representatives with the maximum number of bucket hits plus the first 10 by (hits, id) descending) and the
// containment count of every window member.  The greedy order itself is resolved on the host (api_cluster.cu) from these
// per-read results, block by block.
//
//   lsh_sig_kernel      warp per read: 20 x (FxHasher64 of every minimizer, the 3 smallest (hash, index))
//   sort_dedup_kernel   CTA per read: bitonic sort of the minimizers in shared memory -> sorted distinct list
//   rep_query_kernel    warp per read: lanes 0..19 own one table's bucket (equal range of the read's signature in the table
//                       sorted by (signature, representative)); two 20-way merges by representative id: hits histogram,
//                       then window membership + containment counts
//   pair_stats_kernel   warp per (read, earlier read of the same block): bucket hits and containment count
#include "common.cuh"

namespace sb {

__device__ __forceinline__ uint64_t fx_write(uint64_t h, uint64_t w) { return (((h << 5) | (h >> 59)) ^ w) * 0x517cc1b727220a95ull; }
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    return ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src) << 32) | __shfl_sync(0xffffffffu, (uint32_t)v, src);
}
__device__ __forceinline__ uint64_t shfl_xor64(uint64_t v, int m) {
    return ((uint64_t)__shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m) << 32) | __shfl_xor_sync(0xffffffffu, (uint32_t)v, m);
}

// ---------------- LSH signatures (src/types.rs:719-747) ----------------
__global__ void lsh_sig_kernel(const uint64_t *__restrict__ kmers, const uint64_t *__restrict__ offsets, uint32_t n_reads,
                               uint64_t *__restrict__ sig, uint8_t *__restrict__ valid) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (r >= n_reads) return;
    const uint64_t off = offsets[r];
    const uint32_t n = (uint32_t)(offsets[r + 1] - off);
    const bool ok = n >= (uint32_t)kLshBucket;
    if (lane == 0) valid[r] = ok ? 1 : 0;
    if (!ok) { if (lane < (uint32_t)kLshTables) sig[(uint64_t)r * kLshTables + lane] = 0; return; }
    const uint64_t *km = kmers + off;
    for (int t = 0; t < kLshTables; ++t) {
        const uint64_t seed = fx_write(0, (uint64_t)t);
        // the lane's three smallest (hash, index) of its strided share; sort_by_key is stable, so equal hashes (equal
        // k-mers) keep their input order
        uint64_t h0 = ~0ull, h1 = ~0ull, h2 = ~0ull; uint32_t i0 = ~0u, i1 = ~0u, i2 = ~0u;
        for (uint32_t i = lane; i < n; i += 32u) {
            const uint64_t h = fx_write(seed, __ldg(km + i));
            if (h < h2 || (h == h2 && i < i2)) {
                h2 = h; i2 = i;
                if (h2 < h1 || (h2 == h1 && i2 < i1)) { uint64_t th = h1; h1 = h2; h2 = th; uint32_t ti = i1; i1 = i2; i2 = ti; }
                if (h1 < h0 || (h1 == h0 && i1 < i0)) { uint64_t th = h0; h0 = h1; h1 = th; uint32_t ti = i0; i0 = i1; i1 = ti; }
            }
        }
        uint64_t sg = 0;
#pragma unroll
        for (int round = 0; round < kLshBucket; ++round) {
            uint64_t bh = h0; uint32_t bi = i0;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const uint64_t oh = shfl_xor64(bh, d); const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, d);
                if (oh < bh || (oh == bh && oi < bi)) { bh = oh; bi = oi; }
            }
            sg ^= __ldg(km + bi) * (uint64_t)(round + 1);
            if (i0 == bi) { h0 = h1; i0 = i1; h1 = h2; i1 = i2; h2 = ~0ull; i2 = ~0u; }     // the winner pops its head
        }
        if (lane == 0) sig[(uint64_t)r * kLshTables + t] = sg;
    }
}
void launch_lsh_sig(const uint64_t *kmers, const uint64_t *offsets, uint32_t n_reads, uint64_t *sig, uint8_t *valid, cudaStream_t st) {
    if (n_reads) lsh_sig_kernel<<<(n_reads + 7u) / 8u, 256, 0, st>>>(kmers, offsets, n_reads, sig, valid);
}

// signatures of a few reads (the representatives a block created), for the host's bucket tables
__global__ void gather_sig_kernel(const uint64_t *__restrict__ sig, const uint32_t *__restrict__ ids, uint32_t n, uint64_t *__restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n * (uint32_t)kLshTables) out[t] = sig[(uint64_t)ids[t / kLshTables] * kLshTables + t % kLshTables];
}
void launch_gather_sig(const uint64_t *sig, const uint32_t *ids, uint32_t n, uint64_t *out, cudaStream_t st) {
    if (n) gather_sig_kernel<<<(n * kLshTables + 255u) / 256u, 256, 0, st>>>(sig, ids, n, out);
}

// ---------------- sorted distinct minimizers per read ----------------
// reads with more than kSortMax minimizers are left to the host (nd[r] = UINT32_MAX marks them)
constexpr uint32_t kSortMax = 2048;
__global__ void __launch_bounds__(128) sort_dedup_kernel(const uint64_t *__restrict__ kmers, const uint64_t *__restrict__ offsets,
                                                         uint32_t n_reads, uint64_t *__restrict__ dk, uint32_t *__restrict__ nd) {
    __shared__ uint64_t s[kSortMax];
    __shared__ uint32_t s_cnt[4];
    const uint32_t tid = threadIdx.x;
    for (uint32_t r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const uint64_t off = offsets[r];
        const uint32_t n = (uint32_t)(offsets[r + 1] - off);
        if (n > kSortMax) { if (tid == 0) nd[r] = 0xffffffffu; continue; }
        uint32_t m = 32; while (m < n) m <<= 1;
        for (uint32_t i = tid; i < m; i += 128u) s[i] = i < n ? kmers[off + i] : ~0ull;
        __syncthreads();
        for (uint32_t k = 2; k <= m; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = tid; i < m; i += 128u) {
                    const uint32_t p = i ^ j;
                    if (p > i) {
                        const uint64_t a = s[i], b = s[p];
                        const bool up = (i & k) == 0;
                        if ((a > b) == up) { s[i] = b; s[p] = a; }
                    }
                }
                __syncthreads();
            }
        // compaction of the distinct values, in order: per-warp ballot ranks + running base
        uint32_t base = 0;
        for (uint32_t i0 = 0; i0 < m; i0 += 128u) {
            const uint32_t i = i0 + tid;
            const bool keep = i < n && (i == 0 || s[i] != s[i - 1]);
            const uint32_t bal = __ballot_sync(0xffffffffu, keep);
            if ((tid & 31u) == 0) s_cnt[tid >> 5] = __popc(bal);
            __syncthreads();
            uint32_t before = base;
            for (uint32_t w = 0; w < (tid >> 5); ++w) before += s_cnt[w];
            if (keep) dk[off + before + __popc(bal & ((1u << (tid & 31u)) - 1u))] = s[i];
            base += s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
            __syncthreads();
        }
        if (tid == 0) nd[r] = base;
        __syncthreads();
    }
}
void launch_sort_dedup(const uint64_t *kmers, const uint64_t *offsets, uint32_t n_reads, uint64_t *dk, uint32_t *nd, int sm_count,
                       cudaStream_t st) {
    if (n_reads == 0) return;
    uint32_t grid = (uint32_t)sm_count * 12u;
    if (grid > n_reads) grid = n_reads;
    sort_dedup_kernel<<<grid, 128, 0, st>>>(kmers, offsets, n_reads, dk, nd);
}
uint32_t cluster_sort_max() { return kSortMax; }

// |distinct(a) n distinct(b)|: the lanes share the elements of a, binary search in b
__device__ __forceinline__ uint32_t intersect_count(const uint64_t *__restrict__ a, uint32_t na, const uint64_t *__restrict__ b, uint32_t nb,
                                                    uint32_t lane) {
    uint32_t cnt = 0;
    for (uint32_t i0 = 0; i0 < na; i0 += 32u) {
        const uint32_t i = i0 + lane;
        bool hit = false;
        if (i < na) {
            const uint64_t key = __ldg(a + i);
            uint32_t lo = 0, hi = nb;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(b + mid) < key) lo = mid + 1; else hi = mid; }
            hit = lo < nb && __ldg(b + lo) == key;
        }
        cnt += __popc(__ballot_sync(0xffffffffu, hit));
    }
    return cnt;
}

// the same with b staged in the warp's shared-memory buffer first (one round of coalesced loads instead of a dependent global
// load per binary-search step); falls back to the global search when b does not fit
constexpr uint32_t kIsectBuf = 512;
__device__ __forceinline__ uint32_t intersect_count_staged(const uint64_t *__restrict__ a, uint32_t na, const uint64_t *__restrict__ b, uint32_t nb,
                                                           uint64_t *buf, uint32_t lane) {
    if (nb > kIsectBuf) return intersect_count(a, na, b, nb, lane);
    __syncwarp();
    for (uint32_t i = lane; i < nb; i += 32u) buf[i] = __ldg(b + i);
    __syncwarp();
    uint32_t cnt = 0;
    for (uint32_t i0 = 0; i0 < na; i0 += 32u) {
        const uint32_t i = i0 + lane;
        bool hit = false;
        if (i < na) {
            const uint64_t key = __ldg(a + i);
            uint32_t lo = 0, hi = nb;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (buf[mid] < key) lo = mid + 1; else hi = mid; }
            hit = lo < nb && buf[lo] == key;
        }
        cnt += __popc(__ballot_sync(0xffffffffu, hit));
    }
    return cnt;
}

// ---------------- bucket query + candidate window + containment counts ----------------
__device__ __forceinline__ void equal_range_u64(const uint64_t *__restrict__ v, uint32_t t0, uint32_t t1, uint64_t key, uint32_t &lo, uint32_t &hi) {
    uint32_t a = t0, b = t1;
    while (a < b) { const uint32_t mid = (a + b) >> 1; if (__ldg(v + mid) < key) a = mid + 1; else b = mid; }
    lo = a; b = t1;
    while (a < b) { const uint32_t mid = (a + b) >> 1; if (__ldg(v + mid) <= key) a = mid + 1; else b = mid; }
    hi = a;
}
__global__ void rep_query_kernel(ClusterQuery q) {
    __shared__ uint64_t s_isect[4][kIsectBuf];
    uint64_t *isect_buf = s_isect[threadIdx.x >> 5];
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (w >= q.n_block) return;
    const uint32_t r = q.first + w;
    ClusterResult res{};
    res.best_rep = 0xffffffffu;
    const uint64_t off = q.offsets[r];
    const uint32_t ndr = q.nd[r];
    // equal ranges of the read's signature in table `lane` of the main and of the delta level; the lane's bucket is the main
    // range followed by the delta range (ascending representative ids)
    uint32_t lo = 0, hi = 0, lo2 = 0, hi2 = 0;
    if (lane < (uint32_t)kLshTables && q.valid[r]) {
        const uint64_t sg = q.sig[(uint64_t)r * kLshTables + lane];
        equal_range_u64(q.table_sig, q.table_off[lane], q.table_off[lane + 1], sg, lo, hi);
        equal_range_u64(q.delta_sig, q.delta_off[lane], q.delta_off[lane + 1], sg, lo2, hi2);
    }
    // element x of the lane's bucket (main range, then delta range); the lane keeps the next element loaded ahead of the one
    // the merge looks at, so the load is off the REDUX -> ballot -> advance chain
    const uint32_t n1 = hi - lo, n_all = n1 + (hi2 - lo2);
    auto fetch = [&](uint32_t x) -> uint32_t {
        return x < n1 ? __ldg(q.table_rep + lo + x) : (x < n_all ? __ldg(q.delta_rep + lo2 + (x - n1)) : 0xffffffffu);
    };
    res.n_entries = __reduce_add_sync(0xffffffffu, (hi - lo) + (hi2 - lo2));
    // pass 1: histogram of the hit counts (lane h holds the number of representatives with h hits)
    uint32_t cnt_h = 0;
    {
        uint32_t x = 0, cur = fetch(0), nxt = fetch(1);
        while (true) {
            const uint32_t m = __reduce_min_sync(0xffffffffu, cur);
            if (m == 0xffffffffu) break;
            const uint32_t h = __popc(__ballot_sync(0xffffffffu, cur == m));
            if (lane == h) cnt_h++;
            if (cur == m) { ++x; cur = nxt; nxt = fetch(x + 1); }
        }
    }
    const uint32_t nz = __ballot_sync(0xffffffffu, cnt_h != 0);
    if (nz != 0) {
        const uint32_t max_hits = 31u - __clz(nz);
        // cum_h = number of representatives with >= h hits (suffix sum over the lanes)
        uint32_t cum = cnt_h;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_down_sync(0xffffffffu, cum, d); if (lane + d < 32u) cum += y; }
        const uint32_t total = __shfl_sync(0xffffffffu, cum, 1);
        // h10 = the level at which the first kTopN of the (hits, id)-descending order end (0: fewer than kTopN candidates)
        const uint32_t ge = __ballot_sync(0xffffffffu, cum >= (uint32_t)kClusterTopN && lane >= 1u);
        const uint32_t h10 = ge ? 31u - __clz(ge) : 0u;
        const uint32_t cum_above_h10 = h10 < 31u ? __shfl_sync(0xffffffffu, cum, h10 + 1) : 0u;
        const uint32_t cnt_h10 = __shfl_sync(0xffffffffu, cnt_h, h10);
        const uint32_t need10 = h10 ? (uint32_t)kClusterTopN - cum_above_h10 : 0u;
        // pass 2
        uint32_t seen = 0;
        uint32_t b_rep = 0xffffffffu, b_hits = 0, b_cnt = 0, b_den = 1;
        uint32_t x = 0, cur = fetch(0), nxt = fetch(1);
        while (true) {
            const uint32_t m = __reduce_min_sync(0xffffffffu, cur);
            if (m == 0xffffffffu) break;
            const uint32_t h = __popc(__ballot_sync(0xffffffffu, cur == m));
            if (h < h10) {                                   // below the first kClusterTopN and (max_hits >= h10) not in the max-hits group
                if (cur == m) { ++x; cur = nxt; nxt = fetch(x + 1); }
                continue;
            }
            const uint32_t seen_before = __shfl_sync(0xffffffffu, seen, h);
            const uint32_t cnt_lvl = __shfl_sync(0xffffffffu, cnt_h, h);
            const uint32_t cum_above = h < 31u ? __shfl_sync(0xffffffffu, cum, h + 1) : 0u;
            if (lane == h) seen++;
            const bool in_top = h10 == 0u || h > h10 || (h == h10 && seen_before >= cnt_h10 - need10);
            if (in_top || h == max_hits) {
                const uint64_t moff = q.offsets[m];
                const uint32_t tot_m = (uint32_t)(q.offsets[m + 1] - moff);
                const uint32_t c = intersect_count_staged(q.dk + off, ndr, q.dk + moff, q.nd[m], isect_buf, lane);
                const uint32_t den = max(ndr, tot_m);
                if (in_top) {
                    const uint32_t pos = cum_above + (cnt_lvl - 1u - seen_before);
                    if (pos < (uint32_t)kClusterTopN) { res.top_rep[pos] = m; res.top_hits[pos] = (uint8_t)h; res.top_cnt[pos] = c; res.top_tot[pos] = tot_m; }
                }
                // best of the whole window: larger count / den, then more hits, then larger id (= first in the reference's order)
                if (c > 0u) {
                    const uint64_t lhs = (uint64_t)c * b_den, rhs = (uint64_t)b_cnt * den;
                    if (lhs > rhs || (lhs == rhs && (h > b_hits || (h == b_hits && (b_rep == 0xffffffffu || m > b_rep))))) {
                        b_rep = m; b_hits = h; b_cnt = c; b_den = den;
                    }
                }
            }
            if (cur == m) { ++x; cur = nxt; nxt = fetch(x + 1); }
        }
        res.n_cand = total; res.max_hits = max_hits; res.n_max = __shfl_sync(0xffffffffu, cnt_h, max_hits);
        res.best_rep = b_rep; res.best_hits = b_hits; res.best_cnt = b_cnt; res.best_den = b_den;
    }
    // hits histogram for the host-side merge with the representatives created inside the block
    const uint32_t my = cnt_h;
    for (int h = 1; h <= kLshTables; ++h) { const uint32_t v = __shfl_sync(0xffffffffu, my, h); if (lane == 0) res.cnt_h[h] = v; }
    if (lane == 0) q.out[w] = res;
}
void launch_rep_query(const ClusterQuery &q, cudaStream_t st) {
    if (q.n_block) rep_query_kernel<<<(q.n_block + 3u) / 4u, 128, 0, st>>>(q);
}

// ---------------- (read, in-block potential representative) pairs ----------------
__global__ void pair_stats_kernel(ClusterQuery q, const uint32_t *__restrict__ pair_i, const uint32_t *__restrict__ pair_j, uint32_t n_pairs,
                                  uint8_t *__restrict__ hits_out, uint32_t *__restrict__ cnt_out) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (w >= n_pairs) return;
    const uint32_t i = pair_i[w], j = pair_j[w];
    bool eq = false;
    if (lane < (uint32_t)kLshTables && q.valid[i] && q.valid[j])
        eq = q.sig[(uint64_t)i * kLshTables + lane] == q.sig[(uint64_t)j * kLshTables + lane];
    const uint32_t h = __popc(__ballot_sync(0xffffffffu, eq));
    uint32_t c = 0;
    if (h) c = intersect_count(q.dk + q.offsets[i], q.nd[i], q.dk + q.offsets[j], q.nd[j], lane);
    if (lane == 0) { hits_out[w] = (uint8_t)h; cnt_out[w] = c; }
}
// Only the pairs that share a bucket matter to the host: they are appended to `out` (any order) as
// (read index in the block, index into p0, hits, containment count).
__global__ void block_pair_stats_kernel(ClusterQuery q, const uint32_t *__restrict__ pair_start, const uint32_t *__restrict__ p0, uint32_t n_pairs,
                                        uint4 *__restrict__ out, uint32_t *__restrict__ out_n) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (w >= n_pairs) return;
    uint32_t a = 0, b = q.n_block;                       // li = the last read with pair_start[li] <= w
    while (a < b) { const uint32_t mid = (a + b) >> 1; if (__ldg(pair_start + mid + 1) <= w) a = mid + 1; else b = mid; }
    const uint32_t pidx = w - __ldg(pair_start + a);
    const uint32_t i = q.first + a, j = __ldg(p0 + pidx);
    bool eq = false;
    if (lane < (uint32_t)kLshTables && q.valid[i] && q.valid[j])
        eq = q.sig[(uint64_t)i * kLshTables + lane] == q.sig[(uint64_t)j * kLshTables + lane];
    const uint32_t h = __popc(__ballot_sync(0xffffffffu, eq));
    if (h == 0u) return;
    // (staging j's list in shared memory as rep_query_kernel does was measured here too: slower, 22.6 -> 29.5 ms per 1M reads)
    const uint32_t c = intersect_count(q.dk + q.offsets[i], q.nd[i], q.dk + q.offsets[j], q.nd[j], lane);
    if (lane == 0) out[atomicAdd(out_n, 1u)] = make_uint4(a, pidx, h, c);
}
void launch_block_pair_stats(const ClusterQuery &q, const uint32_t *pair_start, const uint32_t *p0, uint32_t n_pairs, uint4 *out,
                             uint32_t *out_n, cudaStream_t st) {
    if (n_pairs) block_pair_stats_kernel<<<(n_pairs + 7u) / 8u, 256, 0, st>>>(q, pair_start, p0, n_pairs, out, out_n);
}
void launch_pair_stats(const ClusterQuery &q, const uint32_t *pair_i, const uint32_t *pair_j, uint32_t n_pairs, uint8_t *hits, uint32_t *cnt,
                       cudaStream_t st) {
    if (n_pairs) pair_stats_kernel<<<(n_pairs + 7u) / 8u, 256, 0, st>>>(q, pair_i, pair_j, n_pairs, hits, cnt);
}

}  // namespace sb
