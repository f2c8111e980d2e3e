// prefilter.cu -- the tests a read must pass to become a TwinRead: the read-side wrapper
// kmer_comp::twin_reads_from_snpmers (reference src/kmer_comp.rs:68-258), one warp per read, run on the sketch
// of the read (K1) before the candidate filter (K2):
//   length filter                       len < min || len > max                               :117-120  -> status 5
//   get_twin_read_syncmer returned None                                                      :160-161  -> status 4 (from K1)
//   repetitive read                     #solid minimizers < len / c / 20                     :185-188  -> status 6
//       solid = minimizer occurrences whose k-mer occurs <= MAX_KMER_COUNT_IN_READ times in
//       the read and is not in the high-frequency set (USE_SOLID_KMERS = false)              :163-183
//   quality cutoff                      est_id is Some and < quality_value_cutoff            :248      -> status 7
// n_solid_mini / n_solid_snp are the sizes retain_mini_indices / retain_snpmer_indices leave in the POSITION vectors
// (src/types.rs:702-715; the k-mer vectors Stage 7 reads are not pruned).
#include <algorithm>

#include "common.cuh"

namespace sb {

constexpr int kPrefilterWarps = 4;

__device__ __forceinline__ bool hf_contains(const HfSet &hf, uint64_t key) {
    if (hf.keys == nullptr) return false;
    uint32_t slot = hash_slot(key, hf.shift);
    while (true) {
        const uint64_t v = __ldg(hf.keys + slot);
        if (v == key) return true;
        if (v == kEmpty64) return false;
        slot = (slot + 1u) & hf.mask;
    }
}

__global__ void __launch_bounds__(kPrefilterWarps * 32)
read_filter_kernel(SeqBatch in, SketchOut sk, HfSet hf, const double *est_id, ReadFilter f, DeviceParams p, uint32_t rcap,
                   uint32_t *n_solid_mini, uint32_t *n_solid_snp) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem + (size_t)warp * rcap * 12u);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(keys + rcap);
    const uint32_t rmask = rcap - 1u, rshift = 64u - (uint32_t)__popc(rmask);
    for (uint32_t seq = blockIdx.x * kPrefilterWarps + warp; seq < in.n; seq += gridDim.x * kPrefilterWarps) {
        const uint64_t off = in.offsets[seq];
        const uint32_t L = (uint32_t)(in.offsets[seq + 1] - off);
        uint32_t status = sk.status[seq], solid = 0, solid_snp = 0;
        if (L < f.min_len || (f.max_len != 0u && L > f.max_len)) status = 5u;
        else if (status == 0u) {
            for (uint32_t i = lane; i < rcap; i += 32u) { keys[i] = kEmpty64; cnt[i] = 0u; }
            __syncwarp();
            const uint64_t mo = mini_slot_off(off, seq, (uint32_t)p.mini_spacing);
            const uint32_t nmini = sk.mini_cnt[seq];
            for (uint32_t t = lane; t < nmini; t += 32u) {           // kmer_counter_map
                const unsigned long long key = sk.mini_kmer[mo + t] & ~kFwdBit;
                uint32_t slot = hash_slot(key, rshift);
                while (true) {
                    const unsigned long long old = atomicCAS(&keys[slot], (unsigned long long)kEmpty64, key);
                    if (old == kEmpty64 || old == key) { atomicAdd(&cnt[slot], 1u); break; }
                    slot = (slot + 1u) & rmask;
                }
            }
            __syncwarp();
            for (uint32_t i = lane; i < rcap; i += 32u) {            // every occurrence of a solid k-mer is a solid index
                const unsigned long long key = keys[i];
                if (key != kEmpty64 && cnt[i] <= f.max_count && !hf_contains(hf, key)) solid += cnt[i];
            }
            const uint32_t ns = sk.snp_cnt[seq], s0 = sk.snp_start[seq];
            for (uint32_t t = lane; t < ns; t += 32u) if (!hf_contains(hf, sk.snp_kmer[s0 + t])) solid_snp += 1u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                solid += __shfl_xor_sync(0xffffffffu, solid, d);
                solid_snp += __shfl_xor_sync(0xffffffffu, solid_snp, d);
            }
            if (solid < L / (uint32_t)p.c / 20u) { status = 6u; solid = 0; solid_snp = 0; }
            else if (f.use_cutoff && est_id != nullptr) {
                const double e = est_id[seq];
                if (e == e && e < f.cutoff) status = 7u;             // None (NaN) is never dropped
            }
            __syncwarp();
        }
        if (lane == 0) {
            sk.status[seq] = (uint8_t)status;
            if (n_solid_mini) n_solid_mini[seq] = (status == 0u || status == 7u) ? solid : 0u;
            if (n_solid_snp) n_solid_snp[seq] = (status == 0u || status == 7u) ? solid_snp : 0u;
        }
    }
}

static uint32_t next_pow2(uint32_t x) { uint32_t v = 1; while (v < x) v <<= 1; return v; }

void launch_read_filter(const SeqBatch &in, const SketchOut &sk, const HfSet &hf, const double *est_id, const ReadFilter &f,
                        const DeviceParams &p, uint32_t *n_solid_mini, uint32_t *n_solid_snp, int sm_count, cudaStream_t st) {
    if (in.n == 0) return;
    const uint32_t max_mini = in.max_len / (uint32_t)p.mini_spacing + 2u;
    uint32_t rcap = next_pow2(2u * max_mini + 2u);
    if (rcap < 64u) rcap = 64u;
    const size_t smem = (size_t)rcap * 12u * kPrefilterWarps;
    cudaFuncSetAttribute(read_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int ctas_per_sm = (int)(220u * 1024u / (smem + 1024u));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 12) ctas_per_sm = 12;
    uint32_t grid = (uint32_t)(sm_count * ctas_per_sm);
    const uint32_t need = (in.n + kPrefilterWarps - 1) / kPrefilterWarps;
    if (grid > need) grid = need;
    read_filter_kernel<<<grid, kPrefilterWarps * 32, smem, st>>>(in, sk, hf, est_id, f, p, rcap, n_solid_mini, n_solid_snp);
}

// TwinRead.qual_seq (seeding::get_twin_read_syncmer, reference src/seeding.rs:578-602): the minimum quality BYTE of every QUALITY_SEQ_BIN = 4
// bases (src/constants.rs:48), rounded by QualCompact3::try_from_bits (src/types.rs:437-459: 0..=34 -> 0, then one code per 3, 15 from 77 on
// = clamp((byte - 32) / 3, 0, 15)), packed like bio-seq stores a Seq of a 4-bit codec: bin i in bits 4(i % 16).. of the read's word i / 16.
// One warp per read, a lane per bin (the byte loads of a warp cover 128 consecutive bytes), 8 lanes -> one 32-bit half word.
__global__ void __launch_bounds__(256) qual_seq_kernel(const uint8_t *__restrict__ quals, const uint64_t *__restrict__ offsets,
                                                       const uint64_t *__restrict__ word_offsets, uint32_t n, uint32_t *__restrict__ half_words) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
        const uint64_t off = offsets[r];
        const uint32_t L = (uint32_t)(offsets[r + 1] - off), nb = (L + 3u) >> 2;
        const uint64_t h0 = 2ull * word_offsets[r];
        const uint32_t nh = (uint32_t)(2ull * (word_offsets[r + 1] - word_offsets[r]));
        for (uint32_t b0 = 0; b0 < nh * 8u; b0 += 32u) {
            const uint32_t b = b0 + (uint32_t)lane;
            uint32_t code = 0;
            if (b < nb) {
                uint32_t mn = 255u;
#pragma unroll
                for (uint32_t t = 0; t < 4u; ++t) { const uint32_t i = 4u * b + t; if (i < L) mn = min(mn, (uint32_t)__ldg(quals + off + i)); }
                code = mn < 32u ? 0u : min(15u, (mn - 32u) / 3u);
            }
            uint32_t v = code << (4u * ((uint32_t)lane & 7u));
            v |= __shfl_xor_sync(0xffffffffu, v, 1); v |= __shfl_xor_sync(0xffffffffu, v, 2); v |= __shfl_xor_sync(0xffffffffu, v, 4);
            if ((lane & 7) == 0 && (b >> 3) < nh) half_words[h0 + (b >> 3)] = v;
        }
    }
}
void launch_qual_seq(const uint8_t *quals, const uint64_t *offsets, const uint64_t *word_offsets, uint32_t n, uint64_t *words, int sm_count, cudaStream_t st) {
    if (n == 0) return;
    const uint32_t need = (n + 7u) / 8u;
    qual_seq_kernel<<<std::min<uint32_t>(need, (uint32_t)sm_count * 8u), 256, 0, st>>>(quals, offsets, word_offsets, n, reinterpret_cast<uint32_t *>(words));
}

}  // namespace sb
