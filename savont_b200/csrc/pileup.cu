// pileup.cu -- Stage 4 pileups: the device side of savont_b200_consensus_pileups.
//
// Stands in for alignment::generate_consensus_pileups (reference src/alignment.rs:415-641): every read of a cluster is
// homopolymer-compressed (utils::homopolymer_compress_with_quality, src/utils.rs:135-184), aligned to the cluster's
// compressed consensus (K3 + traceback, align.cu) and its CIGAR is walked into one column per consensus position
// (:524-569); the median run length of the aligned bases becomes the column's ref_hp_length (:589-622).
//
// Columns are built without atomics and in the reference's push order (read order, then CIGAR order):
//   count    warp per read: CIGAR walk -> cnt[read row][position] (0..2 entries)
//   rank     thread per consensus position: running sum of cnt over the reads of the consensus -> excl (u16), column total
//   (scan of the column totals -> pos_offsets)
//   emit     warp per read: second CIGAR walk, entry index = pos_offsets[position] + excl + (insertion after a base ? 1 : 0)
//   median   warp per position: 256-bin histogram of the run lengths of the Base entries in shared memory
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "sketch_dev.cuh"

namespace sb {

// ---------------- homopolymer compression ----------------
// position i starts a run iff its code differs from the previous base's or it lies a multiple of 255 behind the start of
// its homopolymer stretch (the reference caps the u8 run length: `current_length < 255`, src/utils.rs:152)
__device__ __forceinline__ uint32_t hpc_starts(const uint8_t *seq, uint32_t len, uint32_t i0, int lane, int &carry_start,
                                               uint32_t &code_out) {
    const uint32_t i = i0 + (uint32_t)lane;
    const bool in = i < len;
    const uint32_t code = in ? enc_base(seq[i]) : 0xffu;
    code_out = code;
    uint32_t prev = __shfl_up_sync(0xffffffffu, code, 1);
    if (lane == 0) prev = (i0 == 0) ? 0xfeu : enc_base(seq[i0 - 1]);
    const bool nat = in && code != prev;
    int s = nat ? (int)i : -1;                       // last natural start <= i (max scan), carried across chunks
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, s, d); if (lane >= d) s = max(s, y); }
    s = max(s, carry_start);
    carry_start = __shfl_sync(0xffffffffu, s, 31);
    const bool start = in && (nat || ((i - (uint32_t)s) % 255u) == 0u);
    return __ballot_sync(0xffffffffu, start);
}

__global__ void hpc_count_kernel(SeqBatch in, int do_hpc, uint32_t *cnt) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= in.n) return;
    const uint64_t off = in.offsets[r];
    const uint32_t len = (uint32_t)(in.offsets[r + 1] - off);
    uint32_t total = len;
    if (do_hpc) {
        total = 0;
        int carry = -1; uint32_t code;
        for (uint32_t i0 = 0; i0 < len; i0 += 32) total += __popc(hpc_starts(in.bases + off, len, i0, lane, carry, code));
    }
    if (lane == 0) cnt[r] = total;
}

// out_off[r] = first compressed base of read r; hc ASCII (upper-case ACGT decoded from the 2-bit code, as the reference
// decodes TwinRead.dna_seq), hq = minimum quality of the run, hl = run length
__global__ void hpc_write_kernel(SeqBatch in, int do_hpc, const uint32_t *out_off, uint8_t *hc, uint8_t *hq, uint8_t *hl) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= in.n) return;
    const uint64_t off = in.offsets[r];
    const uint32_t len = (uint32_t)(in.offsets[r + 1] - off);
    const uint8_t *seq = in.bases + off;
    const uint8_t *q = in.quals ? in.quals + off : nullptr;
    uint32_t o = out_off[r];
    int carry = -1;
    for (uint32_t i0 = 0; i0 < len; i0 += 32) {
        uint32_t code;
        uint32_t starts = do_hpc ? hpc_starts(seq, len, i0, lane, carry, code) : 0u;
        const uint32_t i = i0 + (uint32_t)lane;
        if (!do_hpc) { code = i < len ? enc_base(seq[i]) : 0xffu; starts = __ballot_sync(0xffffffffu, i < len); }
        if ((starts >> lane) & 1u) {
            const uint32_t dst = o + __popc(starts & ((1u << lane) - 1u));
            uint32_t run = 1, mq = q ? q[i] : 33u;
            if (do_hpc)
                for (uint32_t j = i + 1; j < len && run < 255u && enc_base(seq[j]) == code; ++j) { run++; if (q) mq = min(mq, (uint32_t)q[j]); }
            hc[dst] = (uint8_t)("ACGT"[code]); hq[dst] = (uint8_t)mq; hl[dst] = (uint8_t)run;
        }
        o += __popc(starts);
    }
}

void launch_hpc_count(const SeqBatch &in, int do_hpc, uint32_t *cnt, cudaStream_t st) {
    if (in.n == 0) return;
    hpc_count_kernel<<<(in.n + 7u) / 8u, 256, 0, st>>>(in, do_hpc, cnt);
}
void launch_hpc_write(const SeqBatch &in, int do_hpc, const uint32_t *out_off, uint8_t *hc, uint8_t *hq, uint8_t *hl, cudaStream_t st) {
    if (in.n == 0) return;
    hpc_write_kernel<<<(in.n + 7u) / 8u, 256, 0, st>>>(in, do_hpc, out_off, hc, hq, hl);
}

// ---------------- strand choice: both strands were aligned, the smaller NM wins, tie -> forward ----------------
__global__ void pick_strand_kernel(const int32_t *__restrict__ nm2, const uint32_t *__restrict__ pr2, const uint32_t *__restrict__ pa2,
                                   uint32_t n, uint32_t *pr, uint32_t *pa, uint8_t *ps) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    pr[t] = pr2[2 * t]; pa[t] = pa2[2 * t];
    ps[t] = nm2[2 * t + 1] < nm2[2 * t] ? 1 : 0;
}
void launch_pick_strand(const int32_t *nm2, const uint32_t *pr2, const uint32_t *pa2, uint32_t n, uint32_t *pr, uint32_t *pa,
                        uint8_t *ps, cudaStream_t st) {
    if (n == 0) return;
    pick_strand_kernel<<<(n + 255u) / 256u, 256, 0, st>>>(nm2, pr2, pa2, n, pr, pa, ps);
}

// ---------------- CIGAR walk (src/alignment.rs:524-569) ----------------
// Warp per aligned read: the ops are taken in order (warp-uniform), the lanes share the positions of a match / deletion run.
// EMIT = false: cnt[row + position]++ per entry; EMIT = true: write the entries.
template <bool EMIT>
__global__ void pileup_walk_kernel(PileupArgs a) {
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (t >= a.n_aln) return;
    const uint32_t r = a.aln_read[t];
    const uint32_t ci = a.aln_cons[t];
    const uint32_t clen = a.cons_len[ci];
    const uint64_t row = a.row_off[t];
    const uint32_t coff = a.cons_pos_off[ci];
    const bool rev = a.strand[t] != 0;
    const uint64_t hoff = a.hpc_offsets[r];
    const uint32_t hn = (uint32_t)(a.hpc_offsets[r + 1] - hoff);
    const uint32_t qs = (uint32_t)a.read_start[t], mlen = (uint32_t)(a.read_end[t] - a.read_start[t]);
    uint32_t ref_pos = (uint32_t)a.cons_start[t], query_pos = 0;
    const uint32_t c0 = a.cigar_off[t], c1 = a.cigar_off[t + 1];
    // base / quality / run length of position x of the read in alignment orientation (:497-510)
    auto fetch = [&](uint32_t x, uint8_t &b, uint8_t &q, uint8_t &h) {
        const uint64_t src = hoff + (rev ? (uint64_t)(hn - 1u - x) : (uint64_t)x);
        const uint32_t code = enc_base(a.hc[src]);
        b = (uint8_t)("ACGT"[rev ? 3u - code : code]); q = a.hq[src]; h = a.hl[src];
    };
    for (uint32_t k = c0; k < c1; ++k) {
        const uint32_t ln = a.cigar[k] >> 4, op = a.cigar[k] & 15u;
        if (op == 0u) {
            for (uint32_t j = lane; j < ln; j += 32u) {
                if (ref_pos + j < clen && query_pos + j < mlen) {
                    if (!EMIT) a.cnt[row + ref_pos + j] += 1;
                    else {
                        PileupEntry e{}; e.read = r + a.read_base; e.kind = 0; e.n = 1;
                        fetch(qs + query_pos + j, e.base[0], e.qual[0], e.hp[0]);
                        a.entries[(uint64_t)a.pos_off[coff + ref_pos + j] + a.excl[row + ref_pos + j]] = e;
                    }
                }
            }
            ref_pos += ln; query_pos += ln;
        } else if (op == 1u) {
            if (lane == 0u && ref_pos > 0 && ref_pos - 1 < clen && query_pos + ln <= mlen) {
                if (!EMIT) a.cnt[row + ref_pos - 1] += 1;
                else {
                    PileupEntry e{}; e.read = r + a.read_base; e.kind = 2; e.n = (uint8_t)min(ln, 2u);
                    for (uint32_t j = 0; j < e.n; ++j) fetch(qs + query_pos + j, e.base[j], e.qual[j], e.hp[j]);
                    // the read's base / deletion entry of this column (if any) precedes its insertion entry
                    a.entries[(uint64_t)a.pos_off[coff + ref_pos - 1] + a.excl[row + ref_pos - 1] + (k > c0 ? 1u : 0u)] = e;
                }
            }
            query_pos += ln;
        } else {
            for (uint32_t j = lane; j < ln; j += 32u) {
                if (ref_pos + j < clen) {
                    if (!EMIT) a.cnt[row + ref_pos + j] += 1;
                    else {
                        PileupEntry e{}; e.read = r + a.read_base; e.kind = 1; e.n = 0;
                        a.entries[(uint64_t)a.pos_off[coff + ref_pos + j] + a.excl[row + ref_pos + j]] = e;
                    }
                }
            }
            ref_pos += ln;
        }
        if (!EMIT) __syncwarp();      // the insertion count lands on a cell another lane has just written
    }
}

// thread per (consensus, position): rank of every read's entries inside the column, column total
__global__ void pileup_rank_kernel(PileupArgs a) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;      // global position
    if (g >= a.n_pos) return;
    const uint32_t ci = a.pos_cons[g];
    const uint32_t p = g - a.cons_pos_off[ci];
    uint32_t run = 0;
    for (uint32_t t = a.cons_aln_off[ci]; t < a.cons_aln_off[ci + 1]; ++t) {
        const uint64_t at = a.row_off[t] + p;
        a.excl[at] = (uint16_t)run;
        run += a.cnt[at];
    }
    a.col_total[g] = run;
}

// warp per position: median run length of the Base entries (:589-622); no base -> 1
__global__ void pileup_median_kernel(PileupArgs a, uint8_t *ref_hp) {
    __shared__ uint32_t hist_all[8][256];
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (g >= a.n_pos) return;
    uint32_t *hist = hist_all[warp];
    for (int v = lane; v < 256; v += 32) hist[v] = 0;
    __syncwarp();
    const uint32_t e0 = a.pos_off[g], e1 = a.pos_off[g + 1];
    uint32_t nb = 0;
    for (uint32_t e = e0 + lane; e < e1; e += 32) {
        const PileupEntry &x = a.entries[e];
        if (x.kind == 0) { atomicAdd(&hist[x.hp[0]], 1u); nb++; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) nb += __shfl_xor_sync(0xffffffffu, nb, d);
    __syncwarp();
    if (lane == 0) {
        uint8_t out = 1;
        if (nb) {
            const uint32_t mid = nb / 2;
            int lo = -1, hi = -1; uint32_t acc = 0;
            for (int v = 0; v < 256; ++v) {
                acc += hist[v];
                if (lo < 0 && mid >= 1 && acc >= mid) lo = v;
                if (acc >= mid + 1) { hi = v; break; }
            }
            out = (nb % 2 == 0) ? (uint8_t)(((uint32_t)lo + (uint32_t)hi) / 2u) : (uint8_t)hi;
        }
        ref_hp[g] = out;
    }
}

// Device -> page-locked host copy done by a few CTAs instead of the copy engine: the entry stream of a chunk (hundreds of MB)
// must not sit in the DMA queue in front of the small result copies the next chunk's host logic waits for.  Warps store
// 128-byte aligned runs of the destination (dst is only 4-byte aligned).
__global__ void copy_to_host_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, uint64_t n_words, uint32_t pad) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_words + pad; g += 4 * stride) {
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const uint64_t w = g + u * stride; v[u] = (w >= pad && w < n_words + pad) ? src[w - pad] : 0u; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { const uint64_t w = g + u * stride; if (w >= pad && w < n_words + pad) dst[w - pad] = v[u]; }
    }
}
void launch_copy_to_host(const void *src_dev, void *dst_host, uint64_t bytes, cudaStream_t st) {   // bytes % 4 == 0, both 4-byte aligned
    if (bytes == 0) return;
    const uint32_t pad = (uint32_t)((reinterpret_cast<uintptr_t>(dst_host) & 127u) >> 2);
    static const int ctas = [] { const char *e = getenv("SAVONT_B200_COPY_CTAS"); return e ? std::max(1, atoi(e)) : 16; }();   // tuning: 4 / 8 / 16 / 32 CTAs -> 40.5 / 39.8 / 37.6 / 39.7 ms per call; more CTAs finish the copy sooner but slow the next chunk's kernels
    copy_to_host_kernel<<<ctas, 512, 0, st>>>(reinterpret_cast<const uint32_t *>(src_dev), reinterpret_cast<uint32_t *>(dst_host), bytes / 4, pad);
}

void launch_pileup_count(const PileupArgs &a, cudaStream_t st) {
    if (a.n_aln) pileup_walk_kernel<false><<<(a.n_aln + 7u) / 8u, 256, 0, st>>>(a);
}
void launch_pileup_rank(const PileupArgs &a, cudaStream_t st) {
    if (a.n_pos) pileup_rank_kernel<<<(a.n_pos + 255u) / 256u, 256, 0, st>>>(a);
}
void launch_pileup_emit(const PileupArgs &a, cudaStream_t st) {
    if (a.n_aln) pileup_walk_kernel<true><<<(a.n_aln + 7u) / 8u, 256, 0, st>>>(a);
}
void launch_pileup_median(const PileupArgs &a, uint8_t *ref_hp, cudaStream_t st) {
    if (a.n_pos) pileup_median_kernel<<<(a.n_pos + 7u) / 8u, 256, 0, st>>>(a, ref_hp);
}

}  // namespace sb
