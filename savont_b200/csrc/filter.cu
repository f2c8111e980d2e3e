// filter.cu -- K2: candidate-ASV filter, one warp per read.
//
// Stands in for asv_cluster::find_compatible_candidates (reference src/asv_cluster.rs:356-383) and
// the minimizer-overlap / ratio / minimum-mismatch selection of src/alignment.rs:1786-1843:
//   1. the read's minimizer k-mers go into a shared-memory open-addressing set (set semantics of
//      the reference's FxHashSet<Kmer48>, src/alignment.rs:1788); the forward flags of duplicate
//      k-mers are OR-ed so the result does not depend on insertion order;
//   2. every read SNPmer probes the replicated splitmer -> bucket index; the warp walks each bucket
//      cooperatively (coalesced) and bumps a packed u16 mismatch counter per ASV in shared memory
//      plus a two-level candidate bitmap (the reference's FxHashMap<asv,(matches,mismatches)>;
//      `matches` is never used downstream);
//   3. candidates are visited in ascending ASV order; the warp streams the ASV's distinct
//      minimizers (coalesced 64-bit loads, L2 resident) through the read's set: m = |R n A| and the
//      same/opposite-orientation vote; the f64 tests use the reference's operation order
//      (src/alignment.rs:1805-1811, :1829);
//   4. candidates with ratio <= threshold and the minimum SNPmer mismatch count are emitted, in
//      ASV order, into a chunk of the pair pool (one atomic bump per read).
#include <cstdlib>

#include "common.cuh"

namespace sb {

// the reference's two f64 tests on integer counts (src/alignment.rs:1805-1811, :1829): looked up in the host-built threshold tables
// when the arguments are in range, evaluated in f64 in the reference's operation order otherwise
__device__ __forceinline__ bool frac_too_low(const DeviceParams &p, uint32_t m, uint32_t mn) {
    if (mn < p.thresh_n) return m < (uint32_t)__ldg(p.frac_min + mn);
    return (double)m / (double)mn < p.min_frac;
}
__device__ __forceinline__ bool ratio_passes(const DeviceParams &p, uint32_t mm, uint32_t m) {
    if (m < p.thresh_n) return mm <= (uint32_t)__ldg(p.mm_max + m);
    return (double)mm / (double)m / (double)p.c <= p.ratio_thresh;
}

// slot of a k-mer in the per-read minimizer set (shared memory, capacity 2^bits): a 32-bit multiplicative hash -- the set is
// private to the kernel, so insert and probe only have to agree with each other
__device__ __forceinline__ uint32_t rset_slot(uint64_t key, uint32_t shift32) {
    return (((uint32_t)key ^ ((uint32_t)(key >> 32) * 0x85EBCA6Bu)) * 0x9E3779B1u) >> shift32;
}

struct FilterSmem {
    uint32_t rcap;      // read set capacity (power of two)
    uint32_t acc_words; // ceil(A/2)
    uint32_t bm_words;  // ceil(A/32)
    uint32_t sum_words; // ceil(bm_words/32)
    uint32_t same_words; // low-polymorphism mode: ceil(A/2) same-orientation counters, else 0
    __host__ __device__ size_t per_warp() const {
        size_t b = (size_t)rcap * 8u + ((size_t)acc_words + 2u * bm_words + sum_words + same_words) * 4u;
        return (b + 15u) & ~(size_t)15u;
    }
};

// A candidate ASV's distinct minimizers streamed through the read's set: m = shared ones, same = of those, with the read's orientation
// flag.  128 entries per step, the four loads of a lane issued before the first probe (the ASV lists sit in L2: one latency per
// step instead of one per 32 entries); the counts stay in the lanes and are summed once.
__device__ __forceinline__ void count_shared_minimizers(const uint64_t *__restrict__ list, const uint32_t nA, const unsigned long long *rset,
                                                        const uint32_t rshift, const uint32_t rmask, const int lane, uint32_t &m, uint32_t &same) {
    uint32_t acc = 0;                                        // found in the low half, same orientation in the high half
    for (uint32_t t0 = 0; t0 < nA; t0 += 128u) {
        uint64_t vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const uint32_t t = t0 + 32u * u + lane; vv[u] = t < nA ? __ldg(list + t) : kEmpty64; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t v = vv[u];
            if (v == kEmpty64) continue;
            const uint64_t key = v & ~kFwdBit;
            uint32_t slot = rset_slot(key, rshift);
            while (true) {
                const uint64_t r = rset[slot];
                if (r == kEmpty64) break;
                if ((r & ~kFwdBit) == key) { acc += 1u + ((((r ^ v) >> 63) == 0ull) ? 0x10000u : 0u); break; }
                slot = (slot + 1u) & rmask;
            }
        }
    }
    acc = __reduce_add_sync(0xffffffffu, acc);
    m = acc & 0xffffu; same = acc >> 16;
}

// What a warp needs to know about a read before it touches the read's lists.  The kernels fetch it for the NEXT read while the
// current one is processed, and pull that read's minimizer / SNPmer lists towards L2 half-way through (the sketches were written
// by the previous kernel and come from HBM: the filter is bound by such latencies, profiles/r02s_filter_kernel_full.txt).
struct ReadMeta { uint64_t off; uint32_t status, nmini, ns, s0; };
template <bool SNP>
__device__ __forceinline__ ReadMeta load_read_meta(const SeqBatch &in, const SketchOut &sk, const uint32_t seq) {
    ReadMeta m;
    m.status = sk.status[seq]; m.off = in.offsets[seq]; m.nmini = sk.mini_cnt[seq];
    m.ns = SNP ? sk.snp_cnt[seq] : 0u; m.s0 = SNP ? sk.snp_start[seq] : 0u;
    return m;
}
template <bool SNP>
__device__ __forceinline__ void prefetch_read_lists(const SketchOut &sk, const DeviceParams &p, const ReadMeta &m, const uint32_t seq, const int lane) {
    if (m.status != 0u) return;
    if (lane < 16) {
        const char *a = reinterpret_cast<const char *>(sk.mini_kmer + mini_slot_off(m.off, seq, (uint32_t)p.mini_spacing));
        if ((uint32_t)lane * 128u < m.nmini * 8u + 127u) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + lane * 128));
    } else if (SNP) {
        const char *b = reinterpret_cast<const char *>(sk.snp_kmer + m.s0);
        if ((uint32_t)(lane - 16) * 128u < m.ns * 8u + 127u) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + (lane - 16) * 128));
    }
}

// LOWPOLY = candidate_mode 1 (refine_asv_depths_with_minimap2, reference src/alignment.rs:1527-1719): no SNPmer index;
// the read's distinct minimizers probe the minimizer inverted index, `acc` counts the shared minimizers per ASV and
// `same` those with equal orientation flags; every ASV passing the containment test is kept (0 SNPmer mismatches).
template <bool LOWPOLY>
__global__ void __launch_bounds__(kFilterWarps * 32)
filter_kernel(SeqBatch in, SketchOut sk, AsvIndex ix, PairPool pool, FilterOut out, DeviceParams p, FilterSmem lay, int redo_only) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint8_t *wbase = smem + (size_t)warp * lay.per_warp();
    unsigned long long *rset = reinterpret_cast<unsigned long long *>(wbase);
    uint32_t *acc = reinterpret_cast<uint32_t *>(wbase + (size_t)lay.rcap * 8u);
    uint32_t *bm1 = acc + lay.acc_words;
    uint32_t *bm2 = bm1 + lay.bm_words;
    uint32_t *sum1 = bm2 + lay.bm_words;
    uint32_t *same = sum1 + lay.sum_words;
    const uint32_t rmask = lay.rcap - 1u;
    const uint32_t rshift = 32u - (uint32_t)__popc(rmask);
    const uint32_t clear_words = lay.acc_words + 2u * lay.bm_words + lay.sum_words + lay.same_words;
    const int mid_shift = p.k - 1;
    if (*out.err & kErrSnpPool) return;          // the sketch ran out of pool space: the host grows it and re-runs the batch

    uint32_t *work = out.work_ctr + (redo_only ? 1 : 0);
    uint32_t seq_n = 0;
    if (lane == 0) seq_n = atomicAdd(work, 1u);
    seq_n = __shfl_sync(0xffffffffu, seq_n, 0);
    ReadMeta meta_n{};
    if (seq_n < in.n) meta_n = load_read_meta<!LOWPOLY>(in, sk, seq_n);
    while (true) {                                                     // reads are handed out dynamically (see sketch_fast.cu)
        const uint32_t seq = seq_n;
        const ReadMeta me = meta_n;
        if (seq >= in.n) break;
        if (lane == 0) seq_n = atomicAdd(work, 1u);
        seq_n = __shfl_sync(0xffffffffu, seq_n, 0);
        if (seq_n < in.n) meta_n = load_read_meta<!LOWPOLY>(in, sk, seq_n);      // in flight while this read is processed
        if (redo_only && out.cand_status[seq] != 0xFE) continue;       // second launch after filter_compact_kernel (kCandRedo)
        if (me.status != 0u) {
            if (lane == 0) { out.cand_status[seq] = (uint8_t)me.status; out.pair_start[seq] = 0; out.pair_cnt[seq] = 0; out.low_mm[seq] = 0; }
            continue;
        }
        // ---- clear per-read state ----
        for (uint32_t i = lane; i < lay.rcap; i += 32u) rset[i] = kEmpty64;
        for (uint32_t i = lane; i < clear_words; i += 32u) acc[i] = 0u;
        __syncwarp();

        // ---- 1. read minimizer set ----
        const uint64_t off = me.off;
        const uint64_t mo = mini_slot_off(off, seq, (uint32_t)p.mini_spacing);
        const uint32_t nmini = me.nmini;
        uint32_t nR = 0;
        for (uint32_t t0 = 0; t0 < nmini; t0 += 128u) {
            // four independent loads per lane in flight before the (dependent) inserts
            unsigned long long vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t t = t0 + 32u * u + lane;
                vv[u] = (t < nmini) ? sk.mini_kmer[mo + t] : kEmpty64;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned long long v = vv[u];
                bool inserted = false;
                if (v != kEmpty64) {
                    const unsigned long long key = v & ~kFwdBit;
                    uint32_t slot = rset_slot(key, rshift);
                    while (true) {
                        const unsigned long long old = atomicCAS(&rset[slot], (unsigned long long)kEmpty64, v);
                        if (old == kEmpty64) { inserted = true; break; }
                        if ((old & ~kFwdBit) == key) {
                            if (v & kFwdBit) atomicOr(&rset[slot], (unsigned long long)kFwdBit);
                            break;
                        }
                        slot = (slot + 1u) & rmask;
                    }
                }
                nR += __popc(__ballot_sync(0xffffffffu, inserted));
            }
        }
        __syncwarp();
        if (!redo_only && seq_n < in.n) prefetch_read_lists<!LOWPOLY>(sk, p, meta_n, seq_n, lane);

        uint32_t lowest = 0xffffffffu, n_low = 0;
        bool any_scored = false;
        if (LOWPOLY) {
            // ---- 2'. every distinct read minimizer probes the minimizer inverted index (flattened bucket walk) ----
            for (uint32_t s0 = 0; s0 < lay.rcap; s0 += 32u) {
                const unsigned long long rv = rset[s0 + lane];
                uint32_t bstart = 0, blen = 0, rf = 0;
                if (rv != kEmpty64 && ix.mx_keys != nullptr) {
                    const uint64_t key = rv & ~kFwdBit;
                    rf = (uint32_t)(rv >> 63);
                    uint32_t slot = hash_slot(key, ix.mx_shift);
                    while (true) {
                        const uint64_t kk = __ldg(ix.mx_keys + slot);
                        if (kk == key) { const uint2 v = __ldg(ix.mx_vals + slot); bstart = v.x; blen = v.y; break; }
                        if (kk == kEmpty64) break;
                        slot = (slot + 1u) & ix.mx_mask;
                    }
                }
                uint32_t incl = blen;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
                const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                const uint32_t packed_owner = (bstart - (incl - blen));
                for (uint32_t f0 = 0; f0 < total; f0 += 32u) {
                    const uint32_t f = f0 + lane;
                    uint32_t lo = 0;
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        const uint32_t probe = __shfl_sync(0xffffffffu, incl, (int)(lo + step - 1u));
                        if (probe <= f) lo += (uint32_t)step;
                    }
                    const uint32_t own = lo < 32u ? lo : 31u;
                    const uint32_t ebase = __shfl_sync(0xffffffffu, packed_owner, (int)own);
                    const uint32_t of = __shfl_sync(0xffffffffu, rf, (int)own);
                    if (f < total) {
                        const uint32_t ent = __ldg(ix.mx_entries + (uint32_t)(ebase + f));
                        const uint32_t a = ent >> 1;
                        atomicAdd(&acc[a >> 1], 1u << (16u * (a & 1u)));
                        if ((ent & 1u) == of) atomicAdd(&same[a >> 1], 1u << (16u * (a & 1u)));
                        const uint32_t w = a >> 5;
                        const uint32_t old = atomicOr(&bm1[w], 1u << (a & 31u));
                        if (old == 0u) atomicOr(&sum1[w >> 5], 1u << (w & 31u));
                    }
                }
            }
            __syncwarp();
            // ---- 3'. containment test per candidate, ascending ASV order; all survivors are kept ----
            for (uint32_t w = lane; w < lay.bm_words; w += 32u) {
                uint32_t bits = bm1[w], keep = 0;
                while (bits) {
                    const uint32_t bit = (uint32_t)(__ffs(bits) - 1);
                    bits &= bits - 1u;
                    const uint32_t a = w * 32u + bit;
                    const uint32_t sh = 16u * (a & 1u);
                    const uint32_t m = (acc[a >> 1] >> sh) & 0xffffu;
                    const uint32_t sm = (same[a >> 1] >> sh) & 0xffffu;
                    const uint32_t nA = __ldg(ix.mset_off + a + 1) - __ldg(ix.mset_off + a);
                    const uint32_t mn = nR < nA ? nR : nA;
                    if (m == 0u || frac_too_low(p, m, mn)) continue;                   // src/alignment.rs:1801-1808 analogue
                    keep |= 1u << bit;
                    const uint32_t stash = (m & 0x7fffu) | (((m - sm) > sm) ? 0x8000u : 0u);
                    atomicAnd(&acc[a >> 1], ~(0xffffu << sh));
                    atomicOr(&acc[a >> 1], stash << sh);
                }
                bm2[w] = keep;
                n_low += __popc(keep);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) n_low += __shfl_xor_sync(0xffffffffu, n_low, d);
            any_scored = n_low > 0u;
            lowest = 0u;
            __syncwarp();
        } else {
        // ---- 2. SNPmer buckets -> per-ASV mismatch counters + candidate bitmap ----
        // 32 SNPmers per round: every lane probes the splitmer table for its SNPmer, then the (SNPmer, bucket
        // entry) pairs of the round are FLATTENED over the lanes (prefix sum of the bucket lengths + a 5-step
        // search for the owner), so that all entry loads of a round are independent and in flight together.
        const uint32_t ns = me.ns, s0 = me.s0;
        for (uint32_t t0 = 0; t0 < ns; t0 += 32u) {
            const uint32_t t = t0 + lane;
            uint32_t bstart = 0, blen = 0, mid = 0;
            if (t < ns && ix.idx_keys != nullptr) {
                const uint64_t q = sk.snp_kmer[s0 + t];
                const uint64_t sp = q & p.split_mask;
                mid = (uint32_t)(q >> mid_shift) & 3u;
                uint32_t slot = hash_slot(sp, ix.idx_shift);
                while (true) {
                    const uint64_t key = __ldg(ix.idx_keys + slot);
                    if (key == sp) { const uint2 v = __ldg(ix.idx_vals + slot); bstart = v.x; blen = v.y; break; }
                    if (key == kEmpty64) break;
                    slot = (slot + 1u) & ix.idx_mask;
                }
            }
            // short buckets (the common case: a site is variable in one group of related ASVs): every lane walks the bucket of
            // its own SNPmer, four entry loads in flight; long buckets (conserved sites) are flattened over the lanes below
            if (__reduce_max_sync(0xffffffffu, blen) <= 24u) {
                for (uint32_t i0 = 0; i0 < blen; i0 += 4u) {
                    uint32_t ent[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) ent[u] = (i0 + u < blen) ? __ldg(ix.idx_entries + bstart + i0 + u) : 0xffffffffu;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (ent[u] == 0xffffffffu) continue;
                        const uint32_t a = ent[u] >> 2;
                        if ((ent[u] & 3u) != mid) atomicAdd(&acc[a >> 1], 1u << (16u * (a & 1u)));
                        const uint32_t w = a >> 5;
                        const uint32_t old = atomicOr(&bm1[w], 1u << (a & 31u));
                        if (old == 0u) atomicOr(&sum1[w >> 5], 1u << (w & 31u));
                    }
                }
                continue;
            }
            uint32_t incl = blen;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t packed_owner = (bstart - (incl - blen));       // entry index = packed_owner(owner) + flat index
            for (uint32_t f0 = 0; f0 < total; f0 += 32u) {
                const uint32_t f = f0 + lane;
                // owner = first lane whose inclusive prefix exceeds f
                uint32_t lo = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const uint32_t probe = __shfl_sync(0xffffffffu, incl, (int)(lo + step - 1u));
                    if (probe <= f) lo += (uint32_t)step;
                }
                const uint32_t own = lo < 32u ? lo : 31u;
                const uint32_t ebase = __shfl_sync(0xffffffffu, packed_owner, (int)own);
                const uint32_t md = __shfl_sync(0xffffffffu, mid, (int)own);
                if (f < total) {
                    const uint32_t ent = __ldg(ix.idx_entries + (uint32_t)(ebase + f));   // 32-bit wrap-around sum
                    const uint32_t a = ent >> 2;
                    if ((ent & 3u) != md) atomicAdd(&acc[a >> 1], 1u << (16u * (a & 1u)));
                    const uint32_t w = a >> 5;
                    const uint32_t old = atomicOr(&bm1[w], 1u << (a & 31u));
                    if (old == 0u) atomicOr(&sum1[w >> 5], 1u << (w & 31u));
                }
            }
        }
        __syncwarp();

        // ---- 3. evaluate candidates, lowest SNPmer mismatch count first ----
        // Only candidates with the MINIMUM mismatch count among those that pass the minimizer / ratio
        // tests are kept (src/alignment.rs:1829-1843), so the groups are visited in ascending
        // mismatch order and the walk stops at the first group with a passing candidate; every
        // group is visited (and any_scored is exact) only when nothing passes.
        // The candidate bitmap is walked word by word (the summary bitmap names the non-empty words; related ASVs sit in one or two of
        // them), lane b looking at ASV 32w + b: one shared-memory read per lane and word, no per-lane bit loops.
        auto word_mm = [&](const uint32_t w, const uint32_t bits) -> uint32_t {          // this lane's candidate of word w: its mismatches
            const uint32_t a = w * 32u + (uint32_t)lane;
            return ((bits >> lane) & 1u) ? ((acc[a >> 1] >> (16u * (a & 1u))) & 0xffffu) : 0xffffffffu;
        };
        uint32_t cur = 0;
        while (true) {
            uint32_t round_min = 0xffffffffu;                                             // smallest mismatch count >= cur
            for (uint32_t sw = 0; sw < lay.sum_words; ++sw) {
                uint32_t sbits = sum1[sw];
                while (sbits) {
                    const uint32_t w = sw * 32u + (uint32_t)(__ffs(sbits) - 1);
                    sbits &= sbits - 1u;
                    const uint32_t mm = word_mm(w, bm1[w]);
                    if (mm >= cur && mm < round_min) round_min = mm;
                }
            }
            round_min = __reduce_min_sync(0xffffffffu, round_min);
            if (round_min == 0xffffffffu) break;
            // visit the candidates of this group in ascending ASV order
            for (uint32_t sw = 0; sw < lay.sum_words; ++sw) {
                uint32_t sbits = sum1[sw];
                while (sbits) {
                    const uint32_t w = sw * 32u + (uint32_t)(__ffs(sbits) - 1);
                    sbits &= sbits - 1u;
                    uint32_t grp = __ballot_sync(0xffffffffu, word_mm(w, bm1[w]) == round_min);
                    while (grp) {
                        const uint32_t bit = (uint32_t)(__ffs(grp) - 1);
                        grp &= grp - 1u;
                        const uint32_t a = w * 32u + bit;
                        const uint32_t mm = round_min;
                        const uint32_t m0 = __ldg(ix.mset_off + a);
                        const uint32_t nA = __ldg(ix.mset_off + a + 1) - m0;
                        uint32_t m = 0, same = 0;
                        count_shared_minimizers(ix.mset + m0, nA, rset, rshift, rmask, lane, m, same);
                        if (m == 0u) continue;                                        // src/alignment.rs:1801-1803
                        const uint32_t mn = nR < nA ? nR : nA;
                        if (frac_too_low(p, m, mn)) continue;                         // :1805-1808
                        any_scored = true;
                        if (ratio_passes(p, mm, m)) {                                 // :1811, :1829-1833
                            ++n_low;
                            if (lane == 0) {
                                bm2[w] |= 1u << bit;
                                // the mismatch count of a kept ASV is `lowest`: reuse its counter for (m | strand<<15)
                                const uint32_t opp = m - same;
                                const uint32_t stash = (m & 0x7fffu) | ((opp > same) ? 0x8000u : 0u);
                                const uint32_t sh = 16u * (a & 1u);
                                acc[a >> 1] = (acc[a >> 1] & ~(0xffffu << sh)) | (stash << sh);
                            }
                            __syncwarp();          // the next candidate's counter may share this 32-bit word (warp-uniform branch)
                        }
                    }
                }
            }
            if (n_low) { lowest = round_min; break; }
            cur = round_min + 1u;
        }
        __syncwarp();

        }

        // ---- 4. emit the kept candidates ----
        uint8_t status = 0;
        uint32_t total = 0, start = 0;
        if (!any_scored) status = 1;            // SAVONT_B200_READ_NO_CANDIDATE
        else if (n_low == 0u) status = 2;       // SAVONT_B200_READ_RATIO_FILTERED
        else {
            const uint32_t per = (p.strand_mode == 1) ? 2u : 1u;
            total = n_low * per;
            if (lane == 0) start = atomicAdd(pool.counter, total);
            start = __shfl_sync(0xffffffffu, start, 0);
            if ((uint64_t)start + total > pool.cap) {
                if (lane == 0) atomicOr(out.err, kErrPairPool);
                total = 0;
            } else {
                uint32_t running = 0;
                for (uint32_t w0 = 0; w0 < lay.bm_words; w0 += 32u) {
                    const uint32_t w = w0 + lane;
                    uint32_t bits = (w < lay.bm_words) ? bm2[w] : 0u;
                    const uint32_t pc = __popc(bits);
                    uint32_t incl = pc;                               // inclusive warp scan
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += y;
                    }
                    uint32_t o = start + (running + incl - pc) * per;
                    while (bits) {
                        const uint32_t bit = (uint32_t)(__ffs(bits) - 1);
                        bits &= bits - 1u;
                        const uint32_t a = w * 32u + bit;
                        const uint32_t stash = (acc[a >> 1] >> (16u * (a & 1u))) & 0xffffu;
                        if (per == 1u) {
                            pool.read[o] = seq; pool.asv[o] = a; pool.mini[o] = stash & 0x7fffu; pool.strand[o] = (uint8_t)(stash >> 15);
                            o += 1u;
                        } else {
                            pool.read[o] = seq; pool.asv[o] = a; pool.mini[o] = stash & 0x7fffu; pool.strand[o] = 0;
                            pool.read[o + 1] = seq; pool.asv[o + 1] = a; pool.mini[o + 1] = stash & 0x7fffu; pool.strand[o + 1] = 1;
                            o += 2u;
                        }
                    }
                    running += __shfl_sync(0xffffffffu, incl, 31);
                }
            }
        }
        if (lane == 0) {
            out.cand_status[seq] = status;
            out.pair_start[seq] = start;
            out.pair_cnt[seq] = total;
            out.low_mm[seq] = (status == 0) ? lowest : 0u;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Compact variant for large ASV sets (candidate_mode 0): the dense kernel above keeps one 16-bit mismatch counter per
// ASV and two bitmaps per warp in shared memory -- 10.6 KB per warp at 5000 ASVs, which cuts the occupancy to 18 % and
// makes every read clear and scan structures sized by A.  A read only ever meets the ASVs of its SNPmer buckets (tens),
// so this kernel counts mismatches in a small per-warp hash table (asv -> mismatches).  A read that meets more distinct
// ASVs than the table holds is marked kCandRedo and handled afterwards by the dense kernel (redo_only = 1).
// Same results: candidates are evaluated group by group in ascending mismatch order, the kept ones are emitted in
// ascending ASV order.
// ------------------------------------------------------------------------------------------------
constexpr uint8_t kCandRedo = 0xFE;
constexpr uint32_t kCompactEmpty = 0xffffffffu;

struct CompactSmem {
    uint32_t rcap;       // read set capacity (power of two)
    uint32_t ccap;       // candidate table capacity (power of two)
    uint32_t climit;     // distinct candidates before the read is handed to the dense kernel
    __host__ __device__ size_t per_warp() const { return ((size_t)rcap * 8u + (size_t)ccap * 16u + 15u) & ~(size_t)15u; }
};

__global__ void __launch_bounds__(kFilterWarps * 32)
filter_compact_kernel(SeqBatch in, SketchOut sk, AsvIndex ix, PairPool pool, FilterOut out, DeviceParams p, CompactSmem lay) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint8_t *wbase = smem + (size_t)warp * lay.per_warp();
    unsigned long long *rset = reinterpret_cast<unsigned long long *>(wbase);
    uint32_t *ckey = reinterpret_cast<uint32_t *>(wbase + (size_t)lay.rcap * 8u);   // asv id or kCompactEmpty
    uint32_t *cval = ckey + lay.ccap;                                               // SNPmer mismatches
    uint32_t *clist = cval + lay.ccap;                                              // occupied slots, then kept slots
    uint32_t *cstash = clist + lay.ccap;                                            // (m | strand << 15) of kept candidates
    const uint32_t rmask = lay.rcap - 1u, cmask = lay.ccap - 1u;
    const uint32_t rshift = 32u - (uint32_t)__popc(rmask);
    const int mid_shift = p.k - 1;
    if (*out.err & kErrSnpPool) return;

    while (true) {
        uint32_t seq = 0;
        if (lane == 0) seq = atomicAdd(out.work_ctr, 1u);
        seq = __shfl_sync(0xffffffffu, seq, 0);
        if (seq >= in.n) break;
        if (sk.status[seq] != 0) {
            if (lane == 0) { out.cand_status[seq] = sk.status[seq]; out.pair_start[seq] = 0; out.pair_cnt[seq] = 0; out.low_mm[seq] = 0; }
            continue;
        }
        for (uint32_t i = lane; i < lay.rcap; i += 32u) rset[i] = kEmpty64;
        for (uint32_t i = lane; i < lay.ccap; i += 32u) { ckey[i] = kCompactEmpty; cval[i] = 0u; }
        __syncwarp();

        // ---- 1. read minimizer set (as in the dense kernel) ----
        const uint64_t off = in.offsets[seq];
        const uint64_t mo = mini_slot_off(off, seq, (uint32_t)p.mini_spacing);
        const uint32_t nmini = sk.mini_cnt[seq];
        uint32_t nR = 0;
        for (uint32_t t0 = 0; t0 < nmini; t0 += 128u) {
            unsigned long long vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t t = t0 + 32u * u + lane;
                vv[u] = (t < nmini) ? sk.mini_kmer[mo + t] : kEmpty64;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned long long v = vv[u];
                bool inserted = false;
                if (v != kEmpty64) {
                    const unsigned long long key = v & ~kFwdBit;
                    uint32_t slot = rset_slot(key, rshift);
                    while (true) {
                        const unsigned long long old = atomicCAS(&rset[slot], (unsigned long long)kEmpty64, v);
                        if (old == kEmpty64) { inserted = true; break; }
                        if ((old & ~kFwdBit) == key) {
                            if (v & kFwdBit) atomicOr(&rset[slot], (unsigned long long)kFwdBit);
                            break;
                        }
                        slot = (slot + 1u) & rmask;
                    }
                }
                nR += __popc(__ballot_sync(0xffffffffu, inserted));
            }
        }
        __syncwarp();

        // ---- 2. SNPmer buckets -> candidate table (asv -> mismatches) ----
        uint32_t n_distinct = 0;
        bool overflow = false;
        const uint32_t ns = sk.snp_cnt[seq], s0 = sk.snp_start[seq];
        for (uint32_t t0 = 0; t0 < ns && !overflow; t0 += 32u) {
            const uint32_t t = t0 + lane;
            uint32_t bstart = 0, blen = 0, mid = 0;
            if (t < ns && ix.idx_keys != nullptr) {
                const uint64_t q = sk.snp_kmer[s0 + t];
                const uint64_t sp = q & p.split_mask;
                mid = (uint32_t)(q >> mid_shift) & 3u;
                uint32_t slot = hash_slot(sp, ix.idx_shift);
                while (true) {
                    const uint64_t key = __ldg(ix.idx_keys + slot);
                    if (key == sp) { const uint2 v = __ldg(ix.idx_vals + slot); bstart = v.x; blen = v.y; break; }
                    if (key == kEmpty64) break;
                    slot = (slot + 1u) & ix.idx_mask;
                }
            }
            uint32_t incl = blen;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t packed_owner = (bstart - (incl - blen));
            for (uint32_t f0 = 0; f0 < total; f0 += 32u) {
                const uint32_t f = f0 + lane;
                uint32_t lo = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const uint32_t probe = __shfl_sync(0xffffffffu, incl, (int)(lo + step - 1u));
                    if (probe <= f) lo += (uint32_t)step;
                }
                const uint32_t own = lo < 32u ? lo : 31u;
                const uint32_t ebase = __shfl_sync(0xffffffffu, packed_owner, (int)own);
                const uint32_t md = __shfl_sync(0xffffffffu, mid, (int)own);
                bool fresh = false;
                if (f < total) {
                    const uint32_t ent = __ldg(ix.idx_entries + (uint32_t)(ebase + f));
                    const uint32_t a = ent >> 2;
                    uint32_t slot = (a * 0x9E3779B1u) >> (32u - (uint32_t)__popc(cmask));
                    for (uint32_t probes = 0; probes < lay.ccap; ++probes) {
                        const uint32_t v = *(volatile uint32_t *)&ckey[slot];
                        bool mine = v == a;
                        if (!mine && v == kCompactEmpty) {
                            const uint32_t old = atomicCAS(&ckey[slot], kCompactEmpty, a);
                            if (old == kCompactEmpty) { fresh = true; mine = true; }
                            else mine = old == a;
                        }
                        if (mine) { if ((ent & 3u) != md) atomicAdd(&cval[slot], 1u); break; }
                        slot = (slot + 1u) & cmask;
                    }
                }
                n_distinct += __popc(__ballot_sync(0xffffffffu, fresh));
                if (n_distinct > lay.climit) { overflow = true; break; }       // warp-uniform
            }
        }
        __syncwarp();
        if (overflow) {                                                         // too many distinct candidate ASVs: dense kernel
            if (lane == 0) { out.cand_status[seq] = kCandRedo; out.pair_start[seq] = 0; out.pair_cnt[seq] = 0; out.low_mm[seq] = 0; }
            continue;
        }
        // occupied slots -> clist (any order)
        uint32_t n_c = 0;
        for (uint32_t i0 = 0; i0 < lay.ccap; i0 += 32u) {
            const bool occ = ckey[i0 + lane] != kCompactEmpty;
            const uint32_t bal = __ballot_sync(0xffffffffu, occ);
            if (occ) clist[n_c + (uint32_t)__popc(bal & lt_mask)] = i0 + lane;
            n_c += (uint32_t)__popc(bal);
        }
        __syncwarp();

        // ---- 3. evaluate candidates, lowest SNPmer mismatch count first (src/alignment.rs:1829-1843) ----
        uint32_t lowest = 0xffffffffu, n_low = 0, cur = 0;
        bool any_scored = false;
        while (true) {
            uint32_t round_min = 0xffffffffu;
            for (uint32_t i = lane; i < n_c; i += 32u) {
                const uint32_t mm = cval[clist[i]];
                if (mm >= cur && mm < round_min) round_min = mm;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const uint32_t y = __shfl_xor_sync(0xffffffffu, round_min, d);
                round_min = y < round_min ? y : round_min;
            }
            if (round_min == 0xffffffffu) break;
            for (uint32_t ci = 0; ci < n_c; ++ci) {
                const uint32_t slot = clist[ci];
                const uint32_t mm = cval[slot];
                if (mm != round_min) continue;
                const uint32_t a = ckey[slot];
                const uint32_t m0 = __ldg(ix.mset_off + a);
                const uint32_t nA = __ldg(ix.mset_off + a + 1) - m0;
                uint32_t m = 0, same = 0;
                count_shared_minimizers(ix.mset + m0, nA, rset, rshift, rmask, lane, m, same);
                if (m == 0u) continue;                                        // src/alignment.rs:1801-1803
                const uint32_t mn = nR < nA ? nR : nA;
                if (frac_too_low(p, m, mn)) continue;                         // :1805-1808
                any_scored = true;
                if (ratio_passes(p, mm, m)) {                                 // :1811, :1829-1833
                    if (lane == 0) {
                        const uint32_t opp = m - same;
                        cstash[n_low] = (m & 0x7fffu) | ((opp > same) ? 0x8000u : 0u);
                        cval[slot] = 0x80000000u | n_low;                     // kept marker (mm of a kept ASV = lowest)
                    }
                    ++n_low;
                    __syncwarp();
                }
            }
            if (n_low) { lowest = round_min; break; }
            cur = round_min + 1u;
        }
        __syncwarp();

        // ---- 4. emit the kept candidates in ascending ASV order ----
        uint8_t status = 0;
        uint32_t total = 0, start = 0;
        if (!any_scored) status = 1;
        else if (n_low == 0u) status = 2;
        else {
            const uint32_t per = (p.strand_mode == 1) ? 2u : 1u;
            total = n_low * per;
            if (lane == 0) start = atomicAdd(pool.counter, total);
            start = __shfl_sync(0xffffffffu, start, 0);
            if ((uint64_t)start + total > pool.cap) {
                if (lane == 0) atomicOr(out.err, kErrPairPool);
                total = 0;
            } else {
                for (uint32_t i = lane; i < n_c; i += 32u) {
                    const uint32_t slot = clist[i];
                    const uint32_t v = cval[slot];
                    if (!(v & 0x80000000u)) continue;
                    const uint32_t a = ckey[slot];
                    uint32_t rank = 0;                                        // kept ASVs with a smaller id
                    for (uint32_t j = 0; j < n_c; ++j) {
                        const uint32_t sj = clist[j];
                        if ((cval[sj] & 0x80000000u) && ckey[sj] < a) ++rank;
                    }
                    const uint32_t stash = cstash[v & 0x7fffffffu];
                    const uint32_t o = start + rank * per;
                    if (per == 1u) {
                        pool.read[o] = seq; pool.asv[o] = a; pool.mini[o] = stash & 0x7fffu; pool.strand[o] = (uint8_t)(stash >> 15);
                    } else {
                        pool.read[o] = seq; pool.asv[o] = a; pool.mini[o] = stash & 0x7fffu; pool.strand[o] = 0;
                        pool.read[o + 1] = seq; pool.asv[o + 1] = a; pool.mini[o + 1] = stash & 0x7fffu; pool.strand[o + 1] = 1;
                    }
                }
            }
        }
        if (lane == 0) {
            out.cand_status[seq] = status;
            out.pair_start[seq] = start;
            out.pair_cnt[seq] = total;
            out.low_mm[seq] = (status == 0) ? lowest : 0u;
        }
        __syncwarp();
    }
}

static uint32_t next_pow2(uint32_t x) { uint32_t p = 1; while (p < x) p <<= 1; return p; }

static uint32_t env_u32(const char *name, uint32_t dflt) {
    const char *e = getenv(name);
    return (e && atoi(e) > 0) ? (uint32_t)atoi(e) : dflt;
}

static void launch_filter_dense(const SeqBatch &in, const SketchOut &sk, const AsvIndex &ix, const PairPool &pool, const FilterOut &out,
                                const DeviceParams &p, int sm_count, int redo_only, cudaStream_t st) {
    FilterSmem lay;
    const uint32_t max_mini = in.max_len / (uint32_t)p.mini_spacing + 2u;
    lay.rcap = next_pow2(2u * max_mini + 2u);       // load factor <= 0.5: a 256-slot set for 1.5 kb reads (48 instead of 40 warps per SM) measured 12 % slower
    if (lay.rcap < 64u) lay.rcap = 64u;
    lay.acc_words = (ix.n_asv + 1u) / 2u + 1u;
    lay.bm_words = (ix.n_asv + 31u) / 32u;
    if (lay.bm_words == 0u) lay.bm_words = 1u;
    lay.sum_words = (lay.bm_words + 31u) / 32u;
    const bool lowpoly = p.candidate_mode == 1;
    lay.same_words = lowpoly ? lay.acc_words : 0u;
    const size_t smem = lay.per_warp() * kFilterWarps;
    if (lowpoly) cudaFuncSetAttribute(filter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else cudaFuncSetAttribute(filter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int ctas_per_sm = (int)(220u * 1024u / (smem + 1024u));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 12) ctas_per_sm = 12;
    uint32_t grid = (uint32_t)(sm_count * ctas_per_sm);
    const uint32_t need = (in.n + kFilterWarps - 1) / kFilterWarps;
    if (grid > need) grid = need;
    if (lowpoly) filter_kernel<true><<<grid, kFilterWarps * 32, smem, st>>>(in, sk, ix, pool, out, p, lay, redo_only);
    else filter_kernel<false><<<grid, kFilterWarps * 32, smem, st>>>(in, sk, ix, pool, out, p, lay, redo_only);
}

// Large ASV sets go through the compact kernel first and the dense kernel picks up the reads it could not hold
// (SAVONT_B200_FILTER_COMPACT_MIN_ASV / _LIMIT override the switch-over point and the table size: tests).
int launch_filter(const SeqBatch &in, const SketchOut &sk, const AsvIndex &ix, const PairPool &pool, const FilterOut &out,
                  const DeviceParams &p, int sm_count, cudaStream_t st) {
    if (in.n == 0) return 0;
    static const uint32_t min_asv = env_u32("SAVONT_B200_FILTER_COMPACT_MIN_ASV", 1024u);
    static const uint32_t climit = env_u32("SAVONT_B200_FILTER_COMPACT_LIMIT", 48u);
    if (p.candidate_mode == 1 || ix.n_asv < min_asv) {
        launch_filter_dense(in, sk, ix, pool, out, p, sm_count, 0, st);
        return 1;
    }
    CompactSmem lay;
    const uint32_t max_mini = in.max_len / (uint32_t)p.mini_spacing + 2u;
    lay.rcap = next_pow2(2u * max_mini + 2u);       // load factor <= 0.5: a 256-slot set for 1.5 kb reads (48 instead of 40 warps per SM) measured 12 % slower
    if (lay.rcap < 64u) lay.rcap = 64u;
    lay.climit = climit;
    lay.ccap = next_pow2(2u * climit + 32u);        // one flattened step can add 32 entries before the limit is checked
    const size_t smem = lay.per_warp() * kFilterWarps;
    cudaFuncSetAttribute(filter_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int ctas_per_sm = (int)(220u * 1024u / (smem + 1024u));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 12) ctas_per_sm = 12;
    uint32_t grid = (uint32_t)(sm_count * ctas_per_sm);
    const uint32_t need = (in.n + kFilterWarps - 1) / kFilterWarps;
    if (grid > need) grid = need;
    filter_compact_kernel<<<grid, kFilterWarps * 32, smem, st>>>(in, sk, ix, pool, out, p, lay);
    launch_filter_dense(in, sk, ix, pool, out, p, sm_count, 1, st);
    return 2;
}

}  // namespace sb
