// assign.cu -- K4: best-NM ties -> equivalence class + per-ASV counters, plus the small scan /
// gather kernels that turn per-read slots into CSR outputs.
//
// Stands in for reference src/alignment.rs:1865-1925 (Stage 7) and :2151-2171 (Stage 7b):
// best_nm = min NM over the kept candidates, class = sorted ASV ids with NM == best_nm (the pair
// chunk is already in ascending ASV order), unambig / ambig / leq10nm counters, total assigned /
// filtered, and the per-(sample, ASV) singleton-class matrix that the multi-GPU layer all-reduces.
#include "common.cuh"

namespace sb {

__global__ void assign_kernel(FilterOut fo, PairPool pool, const int32_t *__restrict__ nm, const uint32_t *__restrict__ file_idx,
                              uint32_t n_reads, uint32_t n_asv, int n_samples, int strand_mode,
                              const uint32_t *__restrict__ asv_len, uint32_t cells_per_row, unsigned long long *dp_cells,
                              uint8_t *status, int32_t *best_nm, uint32_t *class_cnt, uint32_t *class_slot,
                              unsigned long long *counters) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool assigned = false, filtered = false;
    unsigned long long cells = 0;                 // banded DP cells of this read's pairs: L_asv * (2B+1) each
    const bool skip = *fo.err != 0u;              // a pool overflowed upstream: the host re-runs the batch
    if (skip) { if (r < n_reads) class_cnt[r] = 0; }
    else if (r < n_reads) {
        const uint8_t st = fo.cand_status[r];
        if (st != 0) {
            status[r] = st; best_nm[r] = 0x7fffffff; class_cnt[r] = 0;
            filtered = (st < 4);                  // 4: no TwinRead (too short); 5-7: dropped by the read filter (a4)
        } else {
            const uint32_t start = fo.pair_start[r], cnt = fo.pair_cnt[r];
            const uint32_t per = (strand_mode == 1) ? 2u : 1u;
            const uint32_t ncand = cnt / per;
            for (uint32_t c = 0; c < cnt; ++c) cells += (unsigned long long)asv_len[pool.asv[start + c]] * cells_per_row;
            int best = 0x7fffffff;
            for (uint32_t c = 0; c < ncand; ++c) {
                int v = nm[start + c * per];
                if (per == 2u) { const int v2 = nm[start + c * per + 1]; v = v2 < v ? v2 : v; }
                best = v < best ? v : best;
            }
            uint32_t kcls = 0;
            for (uint32_t c = 0; c < ncand; ++c) {
                int v = nm[start + c * per];
                if (per == 2u) { const int v2 = nm[start + c * per + 1]; v = v2 < v ? v2 : v; }
                if (v == best) class_slot[start + kcls++] = pool.asv[start + c * per];
            }
            class_cnt[r] = kcls;
            if (kcls == 0) { status[r] = 3; best_nm[r] = 0x7fffffff; filtered = true; }
            else {
                status[r] = 0; best_nm[r] = best; assigned = true;
                unsigned long long *unambig = counters, *ambig = counters + n_asv, *leq10 = counters + 2ull * n_asv;
                unsigned long long *single = counters + 3ull * n_asv + 2ull;
                if (kcls == 1u) {
                    const uint32_t a = class_slot[start];
                    atomicAdd(unambig + a, 1ull);
                    const uint32_t s = file_idx ? file_idx[r] : 0u;
                    if (s < (uint32_t)n_samples) atomicAdd(single + (size_t)s * n_asv + a, 1ull);
                } else {
                    for (uint32_t t = 0; t < kcls; ++t) atomicAdd(ambig + class_slot[start + t], 1ull);
                }
                if (best <= 10)
                    for (uint32_t t = 0; t < kcls; ++t) atomicAdd(leq10 + class_slot[start + t], 1ull);
            }
        }
    }
    const uint32_t na = __popc(__ballot_sync(0xffffffffu, assigned));
    const uint32_t nf = __popc(__ballot_sync(0xffffffffu, filtered));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, d);
    if (lane == 0) {
        if (na) atomicAdd(counters + 3ull * n_asv, (unsigned long long)na);
        if (nf) atomicAdd(counters + 3ull * n_asv + 1ull, (unsigned long long)nf);
        if (cells) atomicAdd(dp_cells, cells);
    }
}

void launch_assign(const FilterOut &fo, const PairPool &pool, const int32_t *nm, const uint32_t *file_idx, uint32_t n_reads,
                   uint32_t n_asv, int n_samples, int strand_mode, const uint32_t *asv_len, uint32_t cells_per_row,
                   unsigned long long *dp_cells, uint8_t *status, int32_t *best_nm, uint32_t *class_cnt, uint32_t *class_slot,
                   unsigned long long *counters, cudaStream_t st) {
    if (n_reads == 0) return;
    assign_kernel<<<(n_reads + 255u) / 256u, 256, 0, st>>>(fo, pool, nm, file_idx, n_reads, n_asv, n_samples, strand_mode, asv_len,
                                                         cells_per_row, dp_cells, status, best_nm, class_cnt, class_slot, counters);
}

// ---------------- exclusive scan (u32), three small kernels ----------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;                       // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t warp_sums[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t ws = lane < kScanThreads / 32 ? warp_sums[lane] : 0u;
        uint32_t wi = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += y; }
        if (lane < kScanThreads / 32) warp_sums[lane] = wi - ws;
        if (lane == kScanThreads / 32 - 1 && total) *total = wi;
    }
    __syncthreads();
    const uint32_t res = incl - v + warp_sums[warp];
    __syncthreads();
    return res;
}

__global__ void scan_tile_sums(const uint32_t *counts, uint32_t n, uint32_t *tile_sums) {
    __shared__ uint32_t total;
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) if (base + i < n) s += counts[base + i];
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void scan_tiles(uint32_t *tile_sums, uint32_t n_tiles, uint32_t *grand_total) {
    __shared__ uint32_t total;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_tiles; base += kScanThreads) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_tiles ? tile_sums[i] : 0u;
        const uint32_t ex = block_exclusive_scan(v, &total);
        if (i < n_tiles) tile_sums[i] = ex + carry;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}
__global__ void scan_finish(const uint32_t *counts, uint32_t n, const uint32_t *tile_sums, uint32_t *offsets) {
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems], s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) { v[i] = (base + i < n) ? counts[base + i] : 0u; s += v[i]; }
    uint32_t ex = block_exclusive_scan(s, nullptr) + tile_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) { if (base + i < n) offsets[base + i] = ex; ex += v[i]; }
}
// offsets[n] receives the grand total.  tmp needs ceil(n / 2048) + 1 words.
void launch_scan_u32(const uint32_t *counts, uint32_t *offsets, uint32_t n, uint32_t *tmp, cudaStream_t st) {
    if (n == 0) { cudaMemsetAsync(offsets, 0, sizeof(uint32_t), st); return; }
    const uint32_t n_tiles = (n + kScanTile - 1) / kScanTile;
    scan_tile_sums<<<n_tiles, kScanThreads, 0, st>>>(counts, n, tmp);
    scan_tiles<<<1, kScanThreads, 0, st>>>(tmp, n_tiles, offsets + n);
    scan_finish<<<n_tiles, kScanThreads, 0, st>>>(counts, n, tmp, offsets);
}

__global__ void gather_u32_kernel(const uint32_t *src, const uint32_t *src_start, const uint32_t *cnt, const uint32_t *dst_off,
                                  uint32_t *dst, uint32_t n) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t c = cnt[r], s = src_start[r], d = dst_off[r];
    for (uint32_t t = 0; t < c; ++t) dst[d + t] = src[s + t];
}
void launch_gather_u32(const uint32_t *src, const uint32_t *src_start, const uint32_t *cnt, const uint32_t *dst_off,
                       uint32_t *dst, uint32_t n, cudaStream_t st) {
    if (n == 0) return;
    gather_u32_kernel<<<(n + 255u) / 256u, 256, 0, st>>>(src, src_start, cnt, dst_off, dst, n);
}

// Best hit of every assigned read as an explicit (read, ASV, strand) pair for the traceback kernel: the first member of the
// class (lowest ASV index; the reference's stable sort by NM keeps candidate order, src/alignment.rs:1865-1867) and, when
// both strands were aligned, the strand with the smaller NM (tie -> forward).  flag[r] = 1 iff the read has a best hit.
__global__ void best_pair_kernel(FilterOut fo, PairPool pool, const int32_t *__restrict__ nm, const uint8_t *__restrict__ status,
                                 const int32_t *__restrict__ best_nm, uint32_t n_reads, int strand_mode, uint32_t *flag,
                                 uint32_t *best_asv, uint8_t *best_strand) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    uint32_t a = 0xffffffffu; uint8_t s = 0xffu;
    if (status[r] == 0) {
        const uint32_t start = fo.pair_start[r], cnt = fo.pair_cnt[r];
        const uint32_t per = (strand_mode == 1) ? 2u : 1u;
        const int best = best_nm[r];
        for (uint32_t c = 0; c < cnt / per; ++c) {
            const uint32_t i0 = start + c * per;
            int v = nm[i0]; uint32_t pick = i0;
            if (per == 2u) { const int v2 = nm[i0 + 1]; if (v2 < v) { v = v2; pick = i0 + 1; } }
            if (v == best) { a = pool.asv[pick]; s = pool.strand[pick]; break; }
        }
    }
    flag[r] = a != 0xffffffffu ? 1u : 0u;
    best_asv[r] = a; best_strand[r] = s;
}
__global__ void best_pair_scatter_kernel(const uint32_t *__restrict__ flag, const uint32_t *__restrict__ off,
                                         const uint32_t *__restrict__ best_asv, const uint8_t *__restrict__ best_strand,
                                         uint32_t n_reads, uint32_t *p_read, uint32_t *p_asv, uint8_t *p_strand) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads || !flag[r]) return;
    const uint32_t d = off[r];
    p_read[d] = r; p_asv[d] = best_asv[r]; p_strand[d] = best_strand[r];
}
// off needs n_reads + 1 words (off[n_reads] = number of pairs), tmp as launch_scan_u32
void launch_best_pairs(const FilterOut &fo, const PairPool &pool, const int32_t *nm, const uint8_t *status, const int32_t *best_nm,
                       uint32_t n_reads, int strand_mode, uint32_t *flag, uint32_t *off, uint32_t *tmp, uint32_t *best_asv,
                       uint8_t *best_strand, uint32_t *p_read, uint32_t *p_asv, uint8_t *p_strand, cudaStream_t st) {
    if (n_reads == 0) { cudaMemsetAsync(off, 0, sizeof(uint32_t), st); return; }
    const uint32_t g = (n_reads + 255u) / 256u;
    best_pair_kernel<<<g, 256, 0, st>>>(fo, pool, nm, status, best_nm, n_reads, strand_mode, flag, best_asv, best_strand);
    launch_scan_u32(flag, off, n_reads, tmp, st);
    best_pair_scatter_kernel<<<g, 256, 0, st>>>(flag, off, best_asv, best_strand, n_reads, p_read, p_asv, p_strand);
}

// CIGAR slots (filled from the back by the traceback kernel) -> one compact run per pair; a warp per pair
__global__ void cigar_compact_kernel(const uint32_t *__restrict__ slots, uint32_t stride, const uint32_t *__restrict__ cnt,
                                     const uint32_t *__restrict__ off, uint32_t n_pairs, uint32_t *__restrict__ dst) {
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= n_pairs) return;
    const uint32_t c = cnt[t];
    const uint32_t *src = slots + (size_t)t * stride + (stride - c);
    uint32_t *d = dst + off[t];
    for (uint32_t i = lane; i < c; i += 32) d[i] = src[i];
}
void launch_cigar_compact(const uint32_t *slots, uint32_t stride, const uint32_t *cnt, const uint32_t *off, uint32_t n_pairs,
                          uint32_t *dst, cudaStream_t st) {
    if (n_pairs == 0) return;
    cigar_compact_kernel<<<(n_pairs + 7u) / 8u, 256, 0, st>>>(slots, stride, cnt, off, n_pairs, dst);
}

// gather for find_candidates: per read copy asv / mini / strand of its pair chunk
__global__ void gather_cand_kernel(PairPool pool, const uint32_t *pair_start, const uint32_t *pair_cnt, const uint32_t *low_mm,
                                   const uint32_t *dst_off, uint32_t *asv, uint32_t *mm, uint32_t *mini, uint8_t *strand,
                                   uint32_t n) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t c = pair_cnt[r], s = pair_start[r], d = dst_off[r], lm = low_mm[r];
    for (uint32_t t = 0; t < c; ++t) {
        asv[d + t] = pool.asv[s + t]; mm[d + t] = lm; mini[d + t] = pool.mini[s + t]; strand[d + t] = pool.strand[s + t];
    }
}
void launch_gather_cand(const PairPool &pool, const uint32_t *pair_start, const uint32_t *pair_cnt, const uint32_t *low_mm,
                        const uint32_t *dst_off, uint32_t *asv, uint32_t *mm, uint32_t *mini, uint8_t *strand, uint32_t n,
                        cudaStream_t st) {
    if (n == 0) return;
    gather_cand_kernel<<<(n + 255u) / 256u, 256, 0, st>>>(pool, pair_start, pair_cnt, low_mm, dst_off, asv, mm, mini, strand, n);
}

}  // namespace sb
