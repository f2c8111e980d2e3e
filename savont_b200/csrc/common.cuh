// common.cuh -- shared device/host definitions for libsavont_b200 (sm_100a).
//
// Data layout in HBM (per batch of sequences; all offsets derive from the byte offsets of the
// concatenated input so no prefix sums are needed):
//   bases/quals  u8[total]                 raw ASCII / Phred+33 as handed over by the caller
//   planes       uint2[(off + 63*r) >> 5 ..] bit-planes of the 2-bit codes, bit i of word w = base
//                                           32*w+i; .x = low code bit, .y = high code bit
//   mini slots   u64[(off / 6) + 2*r ..]    minimizer k-mers (bit 63 = forward flag), <= len/6+1 each
//   snp pool     u64/u32 chunks            SNPmer k-mers+positions, chunk per read (atomic bump)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

constexpr uint64_t kEmpty64 = ~0ull;
constexpr uint64_t kFwdBit = 1ull << 63;
constexpr int kSnpHitCap = 64;         // per-sequence SNPmer hits held in shared memory (more spill to global)
constexpr int kSketchWarps = 16;       // warps per CTA in the sketch kernel
constexpr int kFilterWarps = 4;        // warps per CTA in the filter kernel

// error flag bits written by kernels into Context::d_err
enum : uint32_t {
    kErrSnpPool = 1u,      // SNPmer pool exhausted
    kErrSnpHits = 2u,      // > kSnpHitCap SNPmer hits in one sequence
    kErrPairPool = 4u,     // candidate pair pool exhausted
    kErrMiniSet = 8u,      // read minimizer set overflow (should be impossible)
};

struct DeviceParams {
    int k, c, s;            // s = k - c + 1
    int min_bq, band, strand_mode, candidate_mode;
    int asv_padl;           // left padding (bits) of the ASV DP planes
    int mini_spacing;       // minimum distance between two open syncmers = (c-1)/2 + 1
    uint64_t split_mask;    // ~(3 << (k-1))
    uint64_t kmer_mask;     // low 2k bits
    double ratio_thresh, min_frac;
    // the two f64 tests of the candidate filter (src/alignment.rs:1805-1811, :1829) as integer thresholds, tabulated on the host
    // with the very divisions the reference performs: frac_min[mn] = smallest m with m / mn >= min_frac (mn = min(|R|, |A|)),
    // mm_max[m] = largest SNPmer mismatch count with (mm / m) / c <= ratio_thresh; valid for arguments < thresh_n
    const uint16_t *frac_min, *mm_max;
    uint32_t thresh_n;
};
constexpr uint32_t kThreshTable = 4096;

// Fibonacci hashing shared by host-built and device-probed open-addressing tables: the slot is the
// TOP log2(cap) bits of key * 2^64/phi (shift = 64 - log2(cap), cap >= 2).
constexpr uint64_t kFibMul = 0x9E3779B97F4A7C15ull;
#ifdef __CUDACC__
// 4-byte asynchronous copy global -> shared (LDGSTS): no destination register, hence no scoreboard for a later consumer to wait on --
// the way to prefetch along a serial chain inside one thread (cp.async.commit_group / wait_group N order the copies)
__device__ __forceinline__ void cp_async4(uint32_t *smem_dst, const uint32_t *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
#endif
__host__ __device__ __forceinline__ uint32_t hash_slot(uint64_t key, uint32_t shift) { return (uint32_t)((key * kFibMul) >> shift); }
// 32-bit hash for the SNPmer Bloom filter, strand symmetric: the key is the sum of the low words of the k-mer and of its
// reverse complement (the sketch kernels probe it without choosing the canonical strand first).
// word = h >> (32 - log2(words)), bits h & 31 and (h >> 5) & 31
__host__ __device__ __forceinline__ uint32_t bloom_hash32_sym(uint32_t lo_fwd, uint32_t lo_rev) { return (lo_fwd + lo_rev) * 0x9E3779B1u; }

// src/seeding.rs:18-28
__host__ __device__ __forceinline__ uint64_t mm_hash64(uint64_t key) {
    key = (~key) + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}

// plane word offset of sequence r starting at byte offset `off` (>= ceil(len/32)+1 words each)
__host__ __device__ __forceinline__ uint64_t plane_word_off(uint64_t off, uint64_t r) { return (off + 63ull * r) >> 5; }
// minimizer slot offset (>= len/spacing + 2 slots each; two open syncmers are >= spacing apart)
__host__ __device__ __forceinline__ uint64_t mini_slot_off(uint64_t off, uint64_t r, uint32_t spacing) {
    return off / spacing + 2ull * r;
}

struct SeqBatch {            // device view of a batch of sequences
    const uint8_t *bases;
    const uint8_t *quals;    // may be null
    const uint64_t *offsets; // [n+1]
    uint32_t n;
    uint32_t max_len;
    // TwinRead input (savont_b200_map_twin_reads): bases == nullptr, the sequences are the reference's 2-bit
    // Seq<Dna> words (first base in the low bits), sequence r starting at word word_offsets[r]
    const uint64_t *dna_words;
    const uint64_t *word_offsets;
};

struct SketchOut {
    uint8_t *status;         // [n] 0 / 4
    uint8_t *flatq;          // [n] 1 if all qualities equal
    uint2 *planes;           // bit planes
    uint32_t *mini_cnt;      // [n]
    uint64_t *mini_kmer;     // slots
    uint32_t *mini_pos;      // slots or null
    uint32_t *snp_start;     // [n]
    uint32_t *snp_cnt;       // [n]
    uint64_t *snp_kmer;      // pool
    uint32_t *snp_pos;       // pool
    uint32_t *snp_pool_ctr;  // [1]
    uint32_t snp_pool_cap;
    uint64_t *spill_kmer;    // per-warp overflow of the shared-memory SNPmer hit list
    uint32_t *spill_pos;
    uint32_t spill_stride;   // entries per warp (>= max_len)
    uint32_t *err;           // [1] error bits
    uint32_t *work_ctr;      // [1] next sequence to sketch (dynamic distribution over however many CTAs are resident), zeroed per launch
};

struct SnpSet {              // open addressing, kEmpty64 = empty
    const uint64_t *keys;    // null when the set is empty
    uint32_t mask;           // cap-1
    uint32_t shift;          // 64 - log2(cap)
    const uint32_t *bloom;   // blocked two-bit Bloom filter keyed by bloom_hash32_sym (staged in shared memory)
    uint32_t bloom_shift;    // 32 - log2(bloom words)
    uint32_t bloom_words;
};

struct AsvIndex {
    // SNPmer inverted index: splitmer -> bucket of (asv << 2 | mid base)
    const uint64_t *idx_keys;
    const uint2 *idx_vals;       // (start, len)
    uint32_t idx_mask;
    uint32_t idx_shift;
    const uint32_t *idx_entries;
    // distinct minimizer sets per ASV (sorted; bit 63 = forward-any flag)
    const uint64_t *mset;
    const uint32_t *mset_off;    // [A+1]
    // minimizer inverted index (candidate_mode 1): k-mer -> bucket of (asv << 1 | forward-any flag)
    const uint64_t *mx_keys;
    const uint2 *mx_vals;        // (start, len)
    uint32_t mx_mask, mx_shift;
    const uint32_t *mx_entries;
    // DP planes (left-padded by asv_padl zero bits)
    const uint2 *planes;
    const uint32_t *plane_off;   // [A] word offsets
    const uint32_t *len;         // [A]
    uint32_t n_asv;
};

struct PairPool {
    uint32_t *read;      // [cap]
    uint32_t *asv;       // [cap]
    uint32_t *mini;      // [cap]
    uint8_t *strand;     // [cap]
    uint32_t *counter;   // [1]
    uint32_t cap;
};

struct FilterOut {
    uint8_t *cand_status;  // [n]
    uint32_t *pair_start;  // [n]
    uint32_t *pair_cnt;    // [n]
    uint32_t *low_mm;      // [n] SNPmer mismatches of the kept candidates
    uint32_t *err;
    uint32_t *work_ctr;    // [2] next read of the first / second filter launch (zeroed with the batch counters)
};

struct AlignOut {
    int32_t *nm;        // [pairs]
    int32_t *asv_end;   // [pairs] or null
    int32_t *read_end;  // [pairs] or null
};

// K5 (Stage 1 split-k-mer counting, kmercount.cu): one slot of the exact table
struct KmerSlot { unsigned long long key; uint32_t c0, c1; };     // key = kEmpty64 when free; c[flag], flag 1 = forward canonical
void launch_kmer_count(const SeqBatch &in, const uint8_t *rc, int k, int min_bq, KmerSlot *table, uint32_t tshift, uint32_t tmask,
                       uint32_t *ctr, uint32_t fill_limit, int sm_count, cudaStream_t st);
void launch_kmer_table_clear(KmerSlot *table, uint64_t n, cudaStream_t st);
// partitioned counting (k <= 21): histogram -> scatter of 32-bit entries -> per-partition shared-memory counting
constexpr int kKmerPartBits = 13;              // 8192 partitions
struct KmerPartPlan {
    uint64_t mul, inv, kmask; int pbits, rem_bits;
    uint32_t *hist; uint32_t *cursor; uint32_t *entries; const unsigned long long *part_off;
    KmerSlot *table; uint32_t tshift, tmask; uint32_t *ctr; uint32_t fill_limit;       // spill table
    int single_strand; uint64_t *out_kmers; uint32_t *out_counts; uint32_t out_cap; uint32_t *out_n;
    unsigned long long *n_distinct; uint32_t *queue;
    uint32_t smem_fill_limit;                // 0 = default
    uint32_t *const *dst_ptr;                // sharded counting (scatter level 2): per-partition destination, null = local
};
bool kmer_part_supported(int k, uint32_t max_len);
void launch_kmer_hist(const KmerPartPlan &pl, const SeqBatch &in, const uint8_t *rc, int k, int min_bq, int sm_count, cudaStream_t st);
void launch_kmer_scatter(const KmerPartPlan &pl, const SeqBatch &in, const uint8_t *rc, int k, int min_bq, int sm_count, cudaStream_t st);
void launch_kmer_count_part(const KmerPartPlan &pl, int sm_count, cudaStream_t st);
// two-level scatter with coalesced writes (2k <= 38): level 1 from the reads into 128 bins, level 2 into the partitions
bool kmer_part_two_level(int k, uint32_t max_len);
uint32_t kmer_scatter2_tile();
int kmer_scatter1_bits();
void launch_kmer_scatter1(const KmerPartPlan &pl, uint32_t *entries1, uint32_t *cursor1, const SeqBatch &in, const uint8_t *rc, int k,
                          int min_bq, int sm_count, cudaStream_t st);
void launch_kmer_scatter2(const KmerPartPlan &pl, const uint32_t *entries1, const uint32_t *tile_start, uint32_t n_tiles2, int k,
                          int sm_count, cudaStream_t st);
void launch_kmer_compact(const KmerSlot *table, uint64_t n_slots, int single_strand, uint64_t *out_kmers, uint32_t *out_counts,
                         uint32_t out_cap, uint32_t *out_n, cudaStream_t st);

// a4 (twin_reads_from_snpmers) read tests, prefilter.cu
struct HfSet { const uint64_t *keys; uint32_t mask, shift; };       // high-frequency k-mers, open addressing (hash_slot)
struct ReadFilter {
    uint32_t min_len, max_len;     // max_len 0 = no upper bound
    uint32_t max_count;            // MAX_KMER_COUNT_IN_READ
    uint32_t use_cutoff;
    double cutoff;
};

// launchers (defined in the .cu files)
void launch_read_filter(const SeqBatch &in, const SketchOut &sk, const HfSet &hf, const double *est_id, const ReadFilter &f,
                        const DeviceParams &p, uint32_t *n_solid_mini, uint32_t *n_solid_snp, int sm_count, cudaStream_t st);
void launch_sketch(const SeqBatch &in, const SketchOut &out, const SnpSet &set, const uint16_t *rank_table,
                   const DeviceParams &p, int sm_count, cudaStream_t st);
// k = 17, c = 11 fast path (sketch_fast.cu); in.dna_words != nullptr selects the TwinRead input encoding
bool sketch_fast_supported(uint32_t max_len, const DeviceParams &p, bool use_table, const SnpSet &set);
void launch_sketch_fast(const SeqBatch &in, const SketchOut &out, const SnpSet &set, const uint16_t *rank_table,
                        const DeviceParams &p, int sm_count, cudaStream_t st);
void launch_unpack_twin(const SeqBatch &in, uint8_t *codes, cudaStream_t st);
void launch_widen_kmer48(const uint8_t *src, uint64_t *dst, uint64_t n, cudaStream_t st);
uint32_t sketch_grid(uint32_t n, uint32_t max_len, bool use_table, const DeviceParams &p, const SnpSet &set, int sm_count);
void launch_qual_seq(const uint8_t *quals, const uint64_t *offsets, const uint64_t *word_offsets, uint32_t n, uint64_t *words, int sm_count, cudaStream_t st);
void launch_est_id(const SeqBatch &in, const uint8_t *status, const uint8_t *flatq, const double *phred_table,
                   double *est_id, cudaStream_t st);
void launch_codes(const uint8_t *bases, uint8_t *codes, uint64_t total, cudaStream_t st);
int launch_filter(const SeqBatch &in, const SketchOut &sk, const AsvIndex &ix, const PairPool &pool, const FilterOut &out,
                  const DeviceParams &p, int sm_count, cudaStream_t st);   // returns the number of kernels launched
void launch_align_nm(const SeqBatch &in, const uint2 *read_planes, const AsvIndex &ix, const uint32_t *pair_read,
                     const uint32_t *pair_asv, const uint8_t *pair_strand, const uint32_t *n_pairs_dev, uint32_t n_pairs_max,
                     const uint32_t *err, const AlignOut &out, const DeviceParams &p, int sm_count, cudaStream_t st);
// traceback variant for a sub-batch of n_pairs (all pointers pre-offset to the sub-batch): history workspace
// hist[(col * 2*NW + w) * n_pairs + t]; cigar slots of `cigar_stride` ops per pair filled from the back
// (the last cigar_cnt[t] entries of slot t are the ops in forward order).
void launch_align_tb(const SeqBatch &in, const uint2 *read_planes, const AsvIndex &ix, const uint32_t *pair_read,
                     const uint32_t *pair_asv, const uint8_t *pair_strand, uint32_t n_pairs,
                     uint32_t *hist, const AlignOut &out, int32_t *asv_start, int32_t *read_start,
                     uint32_t *cigar, uint32_t cigar_stride, uint32_t *cigar_cnt, const DeviceParams &p, cudaStream_t st);
int align_supported_band_max();
int align_words_for_band(int band);
int align_bits_for_band(int band);
size_t align_tb_hist_words(uint32_t n_pairs, uint32_t hist_cols, const DeviceParams &p);
void launch_assign(const FilterOut &fo, const PairPool &pool, const int32_t *nm,
                   const uint32_t *file_idx, uint32_t n_reads, uint32_t n_asv, int n_samples, int strand_mode,
                   const uint32_t *asv_len, uint32_t cells_per_row, unsigned long long *dp_cells,
                   uint8_t *status, int32_t *best_nm, uint32_t *class_cnt, uint32_t *class_slot,
                   unsigned long long *counters, cudaStream_t st);
void launch_best_pairs(const FilterOut &fo, const PairPool &pool, const int32_t *nm, const uint8_t *status, const int32_t *best_nm,
                       uint32_t n_reads, int strand_mode, uint32_t *flag, uint32_t *off, uint32_t *tmp, uint32_t *best_asv,
                       uint8_t *best_strand, uint32_t *p_read, uint32_t *p_asv, uint8_t *p_strand, cudaStream_t st);
void launch_cigar_compact(const uint32_t *slots, uint32_t stride, const uint32_t *cnt, const uint32_t *off, uint32_t n_pairs,
                          uint32_t *dst, cudaStream_t st);
void launch_gather_cand(const PairPool &pool, const uint32_t *pair_start, const uint32_t *pair_cnt, const uint32_t *low_mm,
                        const uint32_t *dst_off, uint32_t *asv, uint32_t *mm, uint32_t *mini, uint8_t *strand, uint32_t n,
                        cudaStream_t st);
// exclusive scan of counts[n] -> offsets[n+1] (u32), then gather variable-length runs
void launch_scan_u32(const uint32_t *counts, uint32_t *offsets, uint32_t n, uint32_t *tmp, cudaStream_t st);
void launch_gather_u32(const uint32_t *src, const uint32_t *src_start, const uint32_t *cnt, const uint32_t *dst_off,
                       uint32_t *dst, uint32_t n, cudaStream_t st);
void launch_gather_sketch(const SeqBatch &in, const SketchOut &sk, const uint32_t *mini_off, const uint32_t *snp_off,
                          uint32_t *mini_pos, uint64_t *mini_kmer, uint8_t *mini_fwd, uint32_t *snp_pos, uint64_t *snp_kmer,
                          const DeviceParams &p, cudaStream_t st);
void launch_pack_planes(const SeqBatch &in, uint2 *planes, cudaStream_t st);

// ---- Stage 4 pileups (pileup.cu) ----
struct PileupEntry {         // == savont_b200_pileup_entry (include/savont_b200.h), 12 bytes
    uint32_t read;
    uint8_t kind;            // 0 Base, 1 Deletion, 2 Insertion
    uint8_t n;
    uint8_t base[2], qual[2], hp[2];
};
struct PileupArgs {
    // aligned reads t = 0..n_aln (cluster order): read index in the batch, consensus (batch-local), strand, alignment
    uint32_t n_aln;
    const uint32_t *aln_read, *aln_cons;
    const uint8_t *strand;
    const int32_t *cons_start, *read_start, *read_end;
    const uint32_t *cigar_off;       // [n_aln + 1]
    const uint32_t *cigar;
    const uint64_t *row_off;         // [n_aln] first cell of the read's row in cnt / excl (row length = its consensus length)
    // compressed reads
    const uint64_t *hpc_offsets;     // [n_reads + 1]
    const uint8_t *hc, *hq, *hl;
    // consensus sequences of the batch
    const uint32_t *cons_len, *cons_pos_off, *cons_aln_off;   // [n_cons], [n_cons + 1] first position, [n_cons + 1] first aligned read
    const uint32_t *pos_cons;        // [n_pos] consensus of a global position
    uint32_t n_pos;
    uint32_t read_base;              // index of the batch's first read in the call (entries carry call-wide read indices)
    uint8_t *cnt; uint16_t *excl;
    uint32_t *col_total;             // [n_pos]
    const uint32_t *pos_off;         // [n_pos + 1] exclusive scan of col_total
    PileupEntry *entries;
};
void launch_hpc_count(const SeqBatch &in, int do_hpc, uint32_t *cnt, cudaStream_t st);
void launch_hpc_write(const SeqBatch &in, int do_hpc, const uint32_t *out_off, uint8_t *hc, uint8_t *hq, uint8_t *hl, cudaStream_t st);
void launch_pick_strand(const int32_t *nm2, const uint32_t *pr2, const uint32_t *pa2, uint32_t n, uint32_t *pr, uint32_t *pa,
                        uint8_t *ps, cudaStream_t st);
void launch_copy_to_host(const void *src_dev, void *dst_host, uint64_t bytes, cudaStream_t st);
void launch_pileup_count(const PileupArgs &a, cudaStream_t st);
void launch_pileup_rank(const PileupArgs &a, cudaStream_t st);
void launch_pileup_emit(const PileupArgs &a, cudaStream_t st);
void launch_pileup_median(const PileupArgs &a, uint8_t *ref_hp, cudaStream_t st);

// ---- Stage 2 greedy k-mer clustering (cluster.cu) ----
constexpr int kLshTables = 20;        // LSH_NUM_TABLES, reference src/constants.rs:67
constexpr int kLshBucket = 3;         // LSH_BUCKET_SIZE, src/constants.rs:68
constexpr int kClusterTopN = 10;      // top_n_candidates, src/asv_cluster.rs:85
struct ClusterResult {                // per read of a block: its candidates among the representatives known before the block
    uint32_t n_cand, max_hits, n_max;                     // representatives with >= 1 bucket hit; maximum hits; how many reach it
    uint32_t best_rep, best_hits, best_cnt, best_den;     // best containment of the window (max-hits group + first kClusterTopN); UINT32_MAX: none
    uint32_t cnt_h[kLshTables + 1];                       // representatives per hit count
    uint32_t top_rep[kClusterTopN], top_cnt[kClusterTopN], top_tot[kClusterTopN];   // first kClusterTopN by (hits, id) descending
    uint8_t top_hits[kClusterTopN]; uint8_t pad[2];
    uint32_t n_entries;                                   // bucket entries of the read's 20 buckets (each is walked twice)
};
struct ClusterQuery {
    const uint64_t *offsets;          // [n + 1] minimizer CSR
    const uint64_t *dk;               // sorted distinct minimizers per read (same offsets)
    const uint32_t *nd;               // [n] number of distinct minimizers
    const uint64_t *sig;              // [n * kLshTables]
    const uint8_t *valid;             // [n]
    const uint64_t *table_sig;        // the kLshTables bucket tables, each sorted by (signature, representative)
    const uint32_t *table_rep;
    const uint32_t *table_off;        // [kLshTables + 1]
    const uint64_t *delta_sig;        // second level: the representatives created since the main tables were last rebuilt (all
    const uint32_t *delta_rep;        // their ids are larger than every id of the main tables), same layout
    const uint32_t *delta_off;        // [kLshTables + 1]
    uint32_t first, n_block;          // reads [first, first + n_block)
    ClusterResult *out;               // [n_block]
};
void launch_lsh_sig(const uint64_t *kmers, const uint64_t *offsets, uint32_t n_reads, uint64_t *sig, uint8_t *valid, cudaStream_t st);
void launch_sort_dedup(const uint64_t *kmers, const uint64_t *offsets, uint32_t n_reads, uint64_t *dk, uint32_t *nd, int sm_count,
                       cudaStream_t st);
uint32_t cluster_sort_max();
void launch_gather_sig(const uint64_t *sig, const uint32_t *ids, uint32_t n, uint64_t *out, cudaStream_t st);
void launch_rep_query(const ClusterQuery &q, cudaStream_t st);
void launch_pair_stats(const ClusterQuery &q, const uint32_t *pair_i, const uint32_t *pair_j, uint32_t n_pairs, uint8_t *hits, uint32_t *cnt,
                       cudaStream_t st);
// pairs (read first + li, p0[pidx]) for pidx < pair_start[li + 1] - pair_start[li]; those sharing a bucket are appended to out as
// (li, pidx, hits, containment count), out_n = their number (zeroed by the caller)
void launch_block_pair_stats(const ClusterQuery &q, const uint32_t *pair_start, const uint32_t *p0, uint32_t n_pairs, uint4 *out,
                             uint32_t *out_n, cudaStream_t st);

// ---- Stage 3 reclustering: reads against the clusters' consensus SNPmers (reassign.cu) ----
struct SnpCompareArgs {
    int k;
    // inverted index over all clusters: splitmer -> bucket of (cluster << 2 | middle base), as AsvIndex::idx_*
    const uint64_t *idx_keys; const uint2 *idx_vals; uint32_t idx_mask, idx_shift; const uint32_t *idx_entries;
    uint32_t n_cons;
    const uint64_t *read_kmers; const uint64_t *read_offsets; uint32_t n_reads;
    uint32_t *best_cons, *best_mismatches, *best_matches;
    uint32_t *matrix;            // optional: rows of n_cons words, matches | mismatches << 16
    uint32_t first_row;
};
// reads [first, first + n) of the arrays in `a` (matrix = the rows of exactly these reads); returns 1 if n_cons is too large
int launch_snpmer_compare(const SnpCompareArgs &a, uint32_t first, uint32_t n, int sm_count, cudaStream_t st);


// ---- Stage 3 greedy SNPmer clustering inside the k-mer clusters (snpcluster.cu) ----
constexpr uint32_t kSnpTile = 512;                 // positions per tile of the round kernels (a tile never spans two clusters)
constexpr uint32_t kSnpTies = 6;                   // tied representatives kept per read
struct SnpClusterArgs {
    int k;
    const uint64_t *snp_kmers; const uint64_t *snp_offsets;                 // all reads' SNPmer lists (CSR)
    const uint32_t *cluster_reads; const uint64_t *cluster_offsets;         // the k-mer clusters: read ids in processing order (CSR)
    const uint32_t *order;                                                  // cluster indices, largest cluster first
    uint32_t n_clusters, iterative_above, n_memb, n_tiles;
    const uint32_t *poff; const unsigned long long *pkmers;                 // cluster-major copy of the lists: position p of the concatenated clusters
    const uint32_t *tile_cl, *tile_p0;                                      // tiles: cluster, first position
    uint32_t *next3;                                                        // [3][n_clusters]: the representative of round j at [j % 3]
    uint32_t *st_mt, *st_first, *st_tie;                                    // per position: best matches | ties << 16 | representative kind << 30; first compatible rep; [kSnpTies][n_memb] tied reps
    uint32_t *reps, *rep_size;                                              // per cluster segment: representative positions in creation order; per position: size
    uint32_t *rep_of;                                                       // [n_reads] out: representative read id
    uint32_t *work_ctr, *err; unsigned long long *total_reps; int *last_round;
    int prof;                                                               // SAVONT_B200_SNPCLUSTER_PROF: statistics of the resolve kernel
};
void launch_snp_init(const SnpClusterArgs &a, cudaStream_t st);
void launch_snp_round(const SnpClusterArgs &a, uint32_t j, cudaStream_t st);
void launch_snp_resolve(const SnpClusterArgs &a, int sm_count, cudaStream_t st);
void launch_snp_cluster_count(const SnpClusterArgs &a, uint32_t n_memb, uint32_t *cnt, cudaStream_t st);
void launch_snp_cluster_gather(const SnpClusterArgs &a, uint32_t n_memb, unsigned long long *dst, cudaStream_t st);

}  // namespace sb
