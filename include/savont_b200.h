/*
 * savont_b200.h -- C ABI of libsavont_b200.so: the B200 (sm_100a) implementation of savont's
 * read-vs-ASV assignment path (Stage 7 / 7b of `savont asv`, reference v0.6.4).
 *
 * The reference (a pure-Rust crate) has no FFI for this path; this header IS the seam a Rust
 * `extern "C"` shim binds.  Each entry point names the reference
 * interface it stands in for (paths relative to the reference repository root).
 *
 * Conventions
 *   - plain pointers and sizes only; all buffers are caller-owned HOST memory unless the name
 *     says `_dev`; the library owns all GPU memory and copies in/out itself.
 *   - every function returns 0 (SAVONT_B200_OK) or a negative error code; the message is in
 *     savont_b200_last_error() (thread-local).  No exceptions or panics cross the ABI.
 *   - there is NO CPU fallback: without a usable CUDA device every call fails with
 *     SAVONT_B200_ERR_CUDA.
 *   - one context = one GPU; use one context per device (one process per GPU under
 *     torch.distributed / NCCL), or the multi-GPU handle savont_b200_create_multi, which owns a
 *     context per device inside one process.
 *   - call order: savont_b200_set_params first, then load_snpmers / load_asvs (changing k later
 *     rebuilds the SNPmer set's Bloom filter and the ASV index from the host copies).
 *   - NM, best_nm, leq10 and every CIGAR follow the banded edit distance of DESIGN.md section 3,
 *     NOT minimap2's local affine NM that the reference reads (src/alignment.rs:1848-1863);
 *     DESIGN.md section 2 quantifies the difference in classes and depths.
 *   - sequences are concatenated ASCII (`bases`, and Phred+33 `quals` or NULL) with
 *     `offsets[n+1]`; k-mers are u64 in the reference's encoding (A=0 C=1 G=2 T=3, first base
 *     in the highest used bit pair, canonical-by-splitmer; low 48 bits == Kmer48).
 */
#ifndef SAVONT_B200_H
#define SAVONT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAVONT_B200_ABI_VERSION 9

typedef struct savont_b200_ctx savont_b200_ctx;

/* Parameters the path reads from the reference's Cli (src/cli.rs:83-84,99-100,153-158) plus
 * the two constants hard-coded in src/alignment.rs:1806,1829 and the DP band. */
typedef struct savont_b200_params {
    int32_t k;                  /* --kmer-size, odd, 13..23 (src/types.rs:580-609); default 17 */
    int32_t c;                  /* -c, open-syncmer window; default 11; s = k-c+1 */
    int32_t l;                  /* --blockmer-length; accepted and ignored (blockmers are off this path) */
    int32_t min_bq;             /* --minimum-base-quality; default 25 */
    int32_t band;               /* half band B of the DP: cells |i-j| <= B; default 64 */
    int32_t strand_mode;        /* 0: strand by minimizer seed vote; 1: align both strands, keep min NM */
    int32_t candidate_mode;     /* 0: SNPmer-index candidates (refine_asv_depths_with_em, src/alignment.rs:1786-1843);
                                   1: low-polymorphism mode (refine_asv_depths_with_minimap2, src/alignment.rs:1527-1719):
                                      every ASV sharing minimizers with the read is a candidate, no SNPmer index */
    int32_t reserved;
    double  ratio_thresh;       /* 0.0050  (src/alignment.rs:1829) */
    double  min_mini_frac_base; /* 0.950, raised to the k-th power (src/alignment.rs:1806) */
} savont_b200_params;

/* per-read status, shared by find_candidates and map_reads */
enum {
    SAVONT_B200_READ_ASSIGNED = 0,       /* eq-class emitted            src/alignment.rs:1890-1920 */
    SAVONT_B200_READ_NO_CANDIDATE = 1,   /* asv_scores empty            src/alignment.rs:1817-1820 */
    SAVONT_B200_READ_RATIO_FILTERED = 2, /* nothing with ratio<=thresh  src/alignment.rs:1835-1838 */
    SAVONT_B200_READ_NO_ALIGNMENT = 3,   /* best_alns empty             src/alignment.rs:1921-1924 */
    SAVONT_B200_READ_TOO_SHORT = 4,      /* len < k: no TwinRead        src/seeding.rs:339-341 */
    /* only with a read filter (savont_b200_set_read_filter): the read never becomes a TwinRead and is counted
     * neither as assigned nor as filtered */
    SAVONT_B200_READ_LENGTH_FILTERED = 5, /* outside [min,max] length   src/kmer_comp.rs:117-120 */
    SAVONT_B200_READ_REPETITIVE = 6,     /* #solid minimizers < len/c/20 src/kmer_comp.rs:185-188 */
    SAVONT_B200_READ_LOW_QUALITY = 7     /* est_id < quality cutoff     src/kmer_comp.rs:248 */
};

/* The Cli fields / constants the read-side wrapper kmer_comp::twin_reads_from_snpmers reads (src/kmer_comp.rs:68-258). */
typedef struct savont_b200_read_filter {
    uint32_t min_read_length;        /* --min-read-length  (src/cli.rs:87-88, default 1100); 0 = off */
    uint32_t max_read_length;        /* --max-read-length  (src/cli.rs:91-92, default 2000); 0 = no upper bound */
    uint32_t max_kmer_count_in_read; /* MAX_KMER_COUNT_IN_READ = 500 (src/constants.rs) */
    uint32_t reserved;
    double   quality_value_cutoff;   /* --quality-value-cutoff (src/cli.rs:95-96, default 98.0); <= 0 = off */
} savont_b200_read_filter;

enum {
    SAVONT_B200_OK = 0,
    SAVONT_B200_ERR_INVALID = -1,    /* bad argument / params / call order */
    SAVONT_B200_ERR_CUDA = -2,       /* CUDA runtime error or no device */
    SAVONT_B200_ERR_CAPACITY = -3,   /* a caller-provided output buffer is too small */
    SAVONT_B200_ERR_OVERFLOW = -4,   /* an internal per-read limit was exceeded (message says which) */
    SAVONT_B200_ERR_NOMEM = -5
};

/* timing / work counters of the last map_reads / map_resident call (CUDA events on the
 * context's stream; see DESIGN.md "Measurement") */
typedef struct savont_b200_stats {
    double   sketch_ms, filter_ms, align_ms, assign_ms; /* per-kernel-group device time, summed over batches */
    double   h2d_ms, d2h_ms, total_ms;                  /* total_ms: first enqueue to last completion (device) */
    uint64_t n_reads, n_bases, n_pairs, dp_cells;       /* dp_cells = sum over aligned pairs of L_asv*(2B+1) */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t kernel_launches;
    uint64_t n_batches;
    double   gather_ms;                                 /* host time spent gathering per-read buffers (the *_ptrs entry points) */
} savont_b200_stats;

int         savont_b200_abi_version(void);
const char *savont_b200_last_error(void);
void        savont_b200_default_params(savont_b200_params *p);

/* Context lifetime.  device_id is a CUDA ordinal. */
int savont_b200_create(int device_id, savont_b200_ctx **out);
int savont_b200_destroy(savont_b200_ctx *ctx);
/* Run all work of this context on an existing cudaStream_t (e.g. torch's current stream);
 * NULL restores the context's own stream. */
int savont_b200_set_stream(savont_b200_ctx *ctx, void *cuda_stream);
int savont_b200_set_params(savont_b200_ctx *ctx, const savont_b200_params *p);
/* Upper bound on reads per internal batch (0 = default). */
int savont_b200_set_batch_reads(savont_b200_ctx *ctx, size_t max_reads_per_batch);

/* The SNPmer set: both alleles of every SNPmer, exactly the u64 values the reference inserts
 * at src/kmer_comp.rs:41-48 (`split_kmer | mid_base << (k-1)`). */
int savont_b200_load_snpmers(savont_b200_ctx *ctx, const uint64_t *kmers, size_t n);

/* The high-frequency k-mer set of Stage 1 (KmerGlobalInfo.high_freq_kmers, Kmer48 values widened to u64), read by the
 * repetitive-read test of twin_reads_from_snpmers (src/kmer_comp.rs:163-188).  n = 0 clears it. */
int savont_b200_load_high_freq_kmers(savont_b200_ctx *ctx, const uint64_t *kmers, size_t n);
/* Switches the read tests of twin_reads_from_snpmers on (f != NULL) or off (NULL, the default: Stage 7 / 7b receive
 * reads that already passed them).  With a filter set, savont_b200_map_reads / upload_reads + map_resident /
 * find_candidates take raw FASTQ reads: a read failing a test gets status 5 / 6 / 7 and takes no part in the
 * assignment.  The TwinRead entry points ignore the filter (TwinReads exist only for reads that passed). */
int savont_b200_set_read_filter(savont_b200_ctx *ctx, const savont_b200_read_filter *f);
/* Replaces the per-read part of kmer_comp::twin_reads_from_snpmers (src/kmer_comp.rs:68-258) for a batch of FASTQ
 * reads, using the read filter and high-frequency set of the context (a filter must be set): status[i] in
 * {0 kept, 4, 5, 6, 7}, est_id[i] (NaN = None), and the number of minimizer / SNPmer positions that
 * retain_mini_indices / retain_snpmer_indices leave (src/types.rs:702-715; 0 for dropped reads).  The sort by read id
 * (:233) and by est_id (src/main.rs:538) that defines the Stage-7 read order is host work on these outputs
 * (savont_b200/host: twin_read_order). */
int savont_b200_twin_reads_from_snpmers(savont_b200_ctx *ctx, const uint8_t *bases, const uint8_t *quals,
                                        const uint64_t *offsets, size_t n_reads, uint8_t *status, double *est_id,
                                        uint32_t *n_solid_mini, uint32_t *n_solid_snp);
/* TwinRead.qual_seq of a batch of FASTQ reads (seeding::get_twin_read_syncmer, src/seeding.rs:578-602): the minimum quality byte of
 * every QUALITY_SEQ_BIN = 4 bases (src/constants.rs:48; the last bin may be shorter), rounded by QualCompact3::try_from_bits
 * (src/types.rs:437-459: bytes 0..34 -> code 0, 35..37 -> 1, ..., >= 77 -> 15), packed as bio-seq stores a Seq<QualCompact3>: 4 bits per
 * bin, bin i of a read in bits 4(i % 16).. of the read's word i / 16, unused bits zero, every read word aligned.  Outputs:
 * word_offsets[n_reads + 1] (in words, starting at 0) and the words themselves; SAVONT_B200_ERR_CAPACITY (message = words needed)
 * when words_cap is too small.  Stage 4 expands these codes back to qualities (src/alignment.rs:456-478: code * 3 + 33 per bin). */
int savont_b200_qual_seq(savont_b200_ctx *ctx, const uint8_t *quals, const uint64_t *offsets, size_t n_reads,
                         uint64_t *word_offsets, uint64_t *qual_words, size_t words_cap);

/* ---- Stage 1 counting: replaces the counting of seq_parse::read_to_split_kmers
 * (src/seq_parse.rs:12-83, bloom_filter_size == 0) over seeding::split_kmer_mid (src/seeding.rs:975-1068) for the reads
 * of one call = all input files.  k and min_bq come from savont_b200_set_params.  rc[i] != 0 (rc may be NULL) marks a read
 * whose id ends in the token "rc": it is reverse-complemented first (src/seq_parse.rs:361-369, src/utils.rs:51-65).
 * Output: the (k-mer, [c0, c1]) entries that pass the reference's filter (:30-44; c[flag] with flag = 1 iff the forward
 * strand is the canonical one), in UNSPECIFIED order (the reference's is hash-map order; get_snpmers_inplace_sort sorts
 * them itself, src/kmer_comp.rs:483).  out_counts holds 2 u32 per entry.  *n_out = number of entries; when it exceeds
 * out_cap the call fails with SAVONT_B200_ERR_CAPACITY and the message carries the required capacity.
 * stats->n_distinct is the reference's "Total SNPmers after bloom filter" (the number of distinct k-mers before the
 * filter, src/seq_parse.rs:26-28). */
typedef struct savont_b200_kmer_stats {
    double   count_ms, compact_ms, h2d_ms, total_ms;             /* device times (CUDA events); count_ms = all counting kernels */
    double   hist_ms, scatter_ms;                                /* partitioned algorithm: the first two of its three steps */
    uint64_t n_reads, n_bases;
    uint64_t table_slots, n_distinct;                            /* global table (16 B per slot); distinct k-mers before the filter */
    uint64_t h2d_bytes, d2h_bytes, kernel_launches, table_regrows;
    uint64_t partitions, spilled_distinct;                       /* partitioned algorithm (0 partitions = single-table algorithm) */
} savont_b200_kmer_stats;
int savont_b200_count_split_kmers(savont_b200_ctx *ctx, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets,
                                  const uint8_t *rc, size_t n_reads, int single_strand, uint64_t *out_kmers,
                                  uint32_t *out_counts, size_t out_cap, size_t *n_out, savont_b200_kmer_stats *stats);
/* The same over the reads left in HBM by savont_b200_upload_reads (no rc flags): kernel-only measurement. */
int savont_b200_count_split_kmers_resident(savont_b200_ctx *ctx, int single_strand, uint64_t *out_kmers, uint32_t *out_counts,
                                           size_t out_cap, size_t *n_out, savont_b200_kmer_stats *stats);

/* Replaces kmer_comp::twin_reads_from_fasta (src/kmer_comp.rs:39-66) + the SNPmer inverted
 * index build (src/alignment.rs:1754-1765): sketches the ASVs on the GPU (no qualities) and
 * builds the replicated device index.  ASV index = order given. */
int savont_b200_load_asvs(savont_b200_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, size_t n_asv);
int savont_b200_num_asvs(savont_b200_ctx *ctx, size_t *n_asv);

/* Replaces seeding::get_twin_read_syncmer (src/seeding.rs:317-327) for a batch: per sequence
 * the open-syncmer minimizers (position, canonical k-mer, forward flag), the quality-gated,
 * deduplicated SNPmers, est_id (NaN = None) and optionally one 2-bit code per base.
 * CSR outputs; mini_cap / snp_cap are the capacities of the flat arrays (total #bases is
 * always enough).  status[i] is 0 or SAVONT_B200_READ_TOO_SHORT. */
int savont_b200_sketch(savont_b200_ctx *ctx, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets,
                       size_t n_seqs, uint8_t *status, double *est_id,
                       uint64_t *mini_offsets, uint32_t *mini_pos, uint64_t *mini_kmer, uint8_t *mini_fwd, size_t mini_cap,
                       uint64_t *snp_offsets, uint32_t *snp_pos, uint64_t *snp_kmer, size_t snp_cap,
                       uint8_t *codes /* may be NULL */);

/* Replaces asv_cluster::find_compatible_candidates (src/asv_cluster.rs:356-383) plus the
 * minimizer-overlap / ratio / min-mismatch selection of src/alignment.rs:1796-1843.  Kept
 * candidates per read, sorted by ASV index. */
int savont_b200_find_candidates(savont_b200_ctx *ctx, const uint8_t *bases, const uint8_t *quals,
                                const uint64_t *offsets, size_t n_reads,
                                uint8_t *cand_status, uint64_t *cand_offsets,
                                uint32_t *cand_asv, uint32_t *cand_mm, uint32_t *cand_mini, uint8_t *cand_strand,
                                size_t cand_cap);

/* Replaces the minimap2 call of src/alignment.rs:1848-1863 (Aligner::with_seq(read) +
 * map(asv) -> alignment.nm / cigar) for explicit (read, ASV, strand) pairs, with the
 * north-star NM definition (banded, both-ends-free unit-cost edit distance).
 * Coordinates are in DP orientation (read reverse-complemented when strand==1);
 * CIGAR is BAM-encoded (len<<4|op; 0=M, 1=I read-only, 2=D ASV-only), CSR. */
int savont_b200_align_pairs(savont_b200_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, size_t n_reads,
                            const uint32_t *pair_read, const uint32_t *pair_asv, const uint8_t *pair_strand,
                            size_t n_pairs, int32_t *nm, int32_t *asv_end, int32_t *read_end,
                            int want_cigar, int32_t *asv_start, int32_t *read_start,
                            uint64_t *cigar_offsets, uint32_t *cigar, size_t cigar_cap);

/* The whole per-read closure of refine_asv_depths_with_em / compute_per_sample_depths
 * (src/alignment.rs:1786-1925, :2095-2171): sketch -> candidates -> NM -> best-NM ties ->
 * equivalence class and counters.  file_idx may be NULL (n_samples = 1).
 * Counter arrays are ADDED to: unambig/ambig/leq10 [n_asv], totals[2] = {assigned, filtered},
 * sample_singleton [n_samples*n_asv] (may be NULL).  class_* is the per-read sorted
 * equivalence class in CSR (empty for unassigned reads).  stats may be NULL. */
int savont_b200_map_reads(savont_b200_ctx *ctx, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets,
                          const uint32_t *file_idx, size_t n_reads, int n_samples,
                          uint8_t *status, int32_t *best_nm, uint64_t *class_offsets, uint32_t *class_asv,
                          size_t class_cap,
                          uint64_t *unambig, uint64_t *ambig, uint64_t *leq10, uint64_t *totals,
                          uint64_t *sample_singleton, savont_b200_stats *stats);

/* The same closure with the reads in the form the reference hands them to Stage 7 / 7b: `&[TwinRead]`
 * (refine_asv_depths_with_em, src/alignment.rs:1723-1729; compute_per_sample_depths, :2044-2051).  Per read the
 * shim passes three TwinRead fields (src/types.rs:386-412) without re-encoding them:
 *   dna_seq       bio-seq Seq<Dna> raw words, 2 bits per base, base i in bits 2(i%32).. of word i/32 (first base in
 *                 the LOW bits; A=0 C=1 G=2 T=3, src/types.rs:575-579), each read starting at a word boundary:
 *                 words [dna_word_offsets[r], dna_word_offsets[r+1]) hold the lengths[r] bases of read r
 *   snpmer_kmers  the quality-gated, de-duplicated SNPmers of the read (src/alignment.rs:1787, :1791), either as the
 *                 Vec<Kmer48> bytes themselves (snpmer_kmer_bytes = 6, little endian, src/types.rs:104-134) or
 *                 widened to u64 (snpmer_kmer_bytes = 8); read r owns [snpmer_offsets[r], snpmer_offsets[r+1])
 *   file_idx      sample index (NULL: one sample)
 * The read's minimizer k-mers are NOT passed: they are a pure function of dna_seq (get_twin_read_syncmer,
 * src/seeding.rs:526-543) and are recomputed on the GPU together with their strand flags; qualities are not needed
 * because they only gate SNPmers (src/seeding.rs:510-521), which arrive already gated.  Outputs as map_reads. */
int savont_b200_map_twin_reads(savont_b200_ctx *ctx, const uint64_t *dna_words, const uint64_t *dna_word_offsets,
                               const uint32_t *lengths, const void *snpmer_kmers, int snpmer_kmer_bytes,
                               const uint64_t *snpmer_offsets, const uint32_t *file_idx, size_t n_reads, int n_samples,
                               uint8_t *status, int32_t *best_nm, uint64_t *class_offsets, uint32_t *class_asv,
                               size_t class_cap,
                               uint64_t *unambig, uint64_t *ambig, uint64_t *leq10, uint64_t *totals,
                               uint64_t *sample_singleton, savont_b200_stats *stats);

/* Device-resident variant used for kernel-only measurement and by the multi-GPU host layer:
 * upload once, then map repeatedly without host<->device traffic of the inputs.
 * map_resident leaves the counters in the context's device counter vector
 *   [unambig(A) | ambig(A) | leq10(A) | assigned | filtered | sample_singleton(S*A)]  (u64)
 * which savont_b200_counts_dev exposes (for an in-place ncclAllReduce by the host layer) and
 * savont_b200_read_counts copies out.  Per-read outputs stay on the device until
 * savont_b200_fetch_resident. */
int savont_b200_upload_reads(savont_b200_ctx *ctx, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets,
                             const uint32_t *file_idx, size_t n_reads, int n_samples);
int savont_b200_upload_twin_reads(savont_b200_ctx *ctx, const uint64_t *dna_words, const uint64_t *dna_word_offsets,
                                  const uint32_t *lengths, const void *snpmer_kmers, int snpmer_kmer_bytes,
                                  const uint64_t *snpmer_offsets, const uint32_t *file_idx, size_t n_reads, int n_samples);
/* The resident reads are processed as n_parts parts pipelined over two internal streams, so that the sketch / filter kernels
 * of one part share the SMs with the DP kernel of the previous one (same results for every n_parts).  0 = default (one part:
 * measured on B200, 4-16 parts were never faster than one pass); takes effect at the next upload. */
int savont_b200_set_resident_parts(savont_b200_ctx *ctx, int n_parts);
int savont_b200_map_resident(savont_b200_ctx *ctx, int zero_counters, savont_b200_stats *stats);
int savont_b200_fetch_resident(savont_b200_ctx *ctx, uint8_t *status, int32_t *best_nm, uint64_t *class_offsets,
                               uint32_t *class_asv, size_t class_cap);
/* Best-hit alignment record of every read of the resident batch (after savont_b200_map_resident): what the reference's
 * Stage 7 holds in the minimap2 Mapping of the best ASV (src/alignment.rs:1858-1867: `map(asv_seq)` against the read's
 * index, first mapping, stable sort by NM).  best_asv = first member of the read's class (lowest ASV index), strand = the
 * orientation it was aligned in, coordinates in DP orientation as in savont_b200_align_pairs, identity =
 * 1 - nm / (asv_end - asv_start) with the ASV as the query (the reference's formula, src/chimera.rs:443-451; 0 for an
 * empty span), CIGAR BAM-encoded in CSR (cigar_offsets[n+1]).  Unassigned reads: best_asv = UINT32_MAX, strand = 0xff,
 * coordinates -1, identity NaN, empty CIGAR.  All per-read arrays have n_reads entries. */
int savont_b200_best_hits_resident(savont_b200_ctx *ctx, uint32_t *best_asv, uint8_t *strand, int32_t *asv_start,
                                   int32_t *asv_end, int32_t *read_start, int32_t *read_end, double *identity,
                                   uint64_t *cigar_offsets, uint32_t *cigar, size_t cigar_cap);
/* The rows the reference writes to temp/read_to_asv_mappings.tsv for every read of the resident batch (after savont_b200_map_resident;
 * src/alignment.rs:1866, :1875-1887): the read's aligned candidates stable-sorted by NM, the first max_rows_per_read of them (the
 * reference: 5), each with the ASV index, its NM and the read's SNPmer mismatch count against the kept candidates (all kept candidates
 * share it, :1839-1843).  Candidates with equal NM come in ascending ASV index (the reference: hash-map order, :1794).  Reads that were
 * not aligned have no rows.  CSR: row_offsets[n_reads + 1]; SAVONT_B200_ERR_CAPACITY when row_cap is too small (n_reads *
 * max_rows_per_read always suffices).  The text itself is host work: savont_host::write_read_to_asv_mappings. */
int savont_b200_mapping_rows_resident(savont_b200_ctx *ctx, uint32_t max_rows_per_read, uint64_t *row_offsets, uint32_t *row_asv,
                                      int32_t *row_nm, uint32_t *row_snp_mismatches, size_t row_cap);
int savont_b200_counts_dev(savont_b200_ctx *ctx, void **dev_ptr, size_t *n_u64);
int savont_b200_read_counts(savont_b200_ctx *ctx, uint64_t *unambig, uint64_t *ambig, uint64_t *leq10, uint64_t *totals,
                            uint64_t *sample_singleton);

/* ---- Stage 4 pileups: replaces alignment::generate_consensus_pileups, src/alignment.rs:415-641,
 * called from src/main.rs Stage 4 with the clusters' consensus sequences.  Reads arrive in cluster order
 * (`cluster[0..min(len, MAX_SEQS_CONSENSUS = 250)]`, src/alignment.rs:446-450, src/constants.rs:60), read_cons[r] = index
 * of the read's consensus, non-decreasing; read_bases = the upper-case ACGT the reference decodes from TwinRead.dna_seq
 * (:452-454), read_quals = the per-base qualities it expands from qual_seq (:456-478; NULL = 33 everywhere).  Each read
 * is homopolymer-compressed when use_hpc != 0 (utils::homopolymer_compress_with_quality, src/utils.rs:135-184; run
 * lengths cap at 255, quality = minimum of the run) and aligned to its (already compressed) consensus with the banded
 * both-ends-free edit-distance DP of this library on both strands (band = params.band; smaller NM wins, tie -> forward)
 * in place of minimap2 `map_ont` (:439-444,483): parity of the alignments with minimap2 is unpinned.  The CIGAR walk of
 * :524-569 then fills one column per consensus position.
 * Output: columns in CSR over the concatenated consensus positions -- pos_offsets[total_pos + 1], entries in the
 * reference's push order (read order, then CIGAR order); ref_hp_length[total_pos] = median run length of the column's
 * Base entries (:589-622; 1 when the column has none); read_strand[n_reads] 0/1 (0xff: read not aligned: empty read or
 * consensus), read_nm[n_reads] (-1 likewise).  SAVONT_B200_ERR_CAPACITY if entry_cap is too small (the sum over
 * the reads of read length + consensus length is always enough).
 * stats (device time per stage, summed over the internal chunks): sketch_ms = upload + compression + plane packing,
 * align_ms = the both-strand NM kernel, filter_ms = traceback + CIGAR compaction + their copies, assign_ms = the column
 * kernels; n_pairs / dp_cells count both strands. */
typedef struct savont_b200_pileup_entry {
    uint32_t read;        /* index of the read in this call */
    uint8_t  kind;        /* 0 = Base, 1 = Deletion, 2 = Insertion (PileupBase, src/alignment.rs:17-26) */
    uint8_t  n;           /* bases carried: 1, 0, 1..MAX_INSERTION_LENGTH (= 2, src/constants.rs:3) */
    uint8_t  base[2], qual[2], hp[2];   /* ASCII base, quality, homopolymer run length */
} savont_b200_pileup_entry;
int savont_b200_consensus_pileups(savont_b200_ctx *ctx, int use_hpc, const uint8_t *cons_bases, const uint64_t *cons_offsets,
                                  size_t n_cons, const uint8_t *read_bases, const uint8_t *read_quals,
                                  const uint64_t *read_offsets, const uint32_t *read_cons, size_t n_reads,
                                  uint64_t *pos_offsets, savont_b200_pileup_entry *entries, size_t entry_cap,
                                  uint64_t *n_entries, uint8_t *ref_hp_length, uint8_t *read_strand, int32_t *read_nm,
                                  savont_b200_stats *stats /* may be NULL */);

/* ---- Stage 2 greedy k-mer clustering.  Input of both calls: the reads' minimizer k-mers in CSR,
 * exactly TwinRead::minimizer_kmers() (src/types.rs: duplicates kept, position order; Kmer48 widened to u64), reads in the
 * order of `twin_reads` (Stage 1.5 output order, src/main.rs:538).
 * savont_b200_lsh_signatures replaces TwinRead::compute_lsh_signatures (src/types.rs:719-747; FxHasher64 of the fxhash crate
 *   0.2.1): sig[r * 20 + t] for the LSH_NUM_TABLES = 20 tables, valid[r] = 0 when the read has fewer than LSH_BUCKET_SIZE = 3
 *   minimizers (all its signatures are None).
 * savont_b200_cluster_reads_by_kmers replaces the greedy loop of asv_cluster::cluster_reads_by_kmers
 *   (src/asv_cluster.rs:72-190, use_bucketed = true; k = params.k, threshold = args.primary_clustering_threshold, default 0.95):
 *   rep_of[r] = the representative read r was assigned to, r itself when it became a representative.  The cluster list of
 *   :192-210 (group by representative, sort by size then first member, drop clusters below min_cluster_size) is host work on
 *   rep_of.  The result is identical to the sequential loop.  stats: sketch_ms = upload + signatures + sorted minimizer sets,
 *   filter_ms = bucket query kernels, align_ms = in-block pair kernels, n_batches = blocks, n_pairs = in-block pairs,
 *   dp_cells = bucket entries the query kernels walked (each twice). */
int savont_b200_lsh_signatures(savont_b200_ctx *ctx, const uint64_t *kmers, const uint64_t *offsets, size_t n_reads, uint64_t *sig,
                               uint8_t *valid);
int savont_b200_cluster_reads_by_kmers(savont_b200_ctx *ctx, const uint64_t *kmers, const uint64_t *offsets, size_t n_reads,
                                       double threshold, uint32_t *rep_of, uint64_t *n_reps, savont_b200_stats *stats /* may be NULL */);

/* ---- Sharded Stage 1 counting over several GPUs (one context and one process per GPU; reads left in HBM by
 * savont_b200_upload_reads, each rank its own shard).  Partition p of the P = savont_b200_kmer_shard_partitions() hash
 * partitions belongs to rank p * n_ranks / P.  Per call of the multi-GPU count (savont_b200/dist.py: count_split_kmers_sharded):
 *   1. savont_b200_kmer_shard_hist      hist[P] = this rank's occurrences per partition            -> all-gather of the histograms
 *   2. savont_b200_kmer_shard_alloc     entry buffer (4 B per occurrence) of the partitions this rank owns; the other ranks map
 *                                       it with savont_b200_ipc_export / _ipc_open (cudaIpc*, NVLink peer memory)
 *   3. savont_b200_kmer_shard_scatter   scatters this rank's reads; the level-2 scatter kernel writes the run of partition p
 *                                       straight into peer_entries[owner(p)] + dst_off[p] (dst_off[p] = start of the partition
 *                                       in the owner's buffer + the occurrences of the lower ranks): the exchange step is the
 *                                       kernel's own store, there is no separate all-to-all.  Returns after the stream has
 *                                       drained; the caller barriers.
 *   4. savont_b200_kmer_shard_count     counts the owned partitions (part_off[P + 1]: offsets in the own buffer, empty ranges
 *                                       for partitions of other ranks); output as savont_b200_count_split_kmers, restricted
 *                                       to the k-mers of the owned partitions (out_kmers = out_counts = NULL: count only).
 * Needs the two-level partitioned algorithm (k <= 19, reads <= 24000 bases). */
int savont_b200_kmer_shard_partitions(void);
int savont_b200_kmer_shard_hist(savont_b200_ctx *ctx, uint32_t *hist);
int savont_b200_kmer_shard_alloc(savont_b200_ctx *ctx, uint64_t n_entries, void **dev_ptr, uint64_t *capacity /* entries, may be NULL */);
int savont_b200_ipc_export(savont_b200_ctx *ctx, const void *dev_ptr, void *handle64 /* 64 bytes */);
int savont_b200_ipc_open(savont_b200_ctx *ctx, const void *handle64, void **dev_ptr);
int savont_b200_ipc_close(savont_b200_ctx *ctx, void *dev_ptr);
int savont_b200_kmer_shard_scatter(savont_b200_ctx *ctx, int n_ranks, void *const *peer_entries, const uint64_t *dst_off);
int savont_b200_kmer_shard_count(savont_b200_ctx *ctx, const uint64_t *part_off, int single_strand, uint64_t *out_kmers,
                                 uint32_t *out_counts, size_t out_cap, size_t *n_out, savont_b200_kmer_stats *stats);

/* ---- Stage 3 reclustering: reads -- or consensus lists -- against the clusters' consensus SNPmers.
 * cons_kmers / cons_offsets: per cluster the ConsensusSnpmer k-mers (`cs.kmer`; a later entry with the same splitmer replaces an
 * earlier one, as HashMap::insert at src/asv_cluster.rs:1031-1034); read_kmers / read_offsets: per read the k-mers of
 * TwinRead::snpmers_vec() (:1049).  Replaces the candidate loop of asv_cluster::reassign_reads_to_best_cluster (:1052-1098, SNPmer
 * markers): best_cons[r] = the first cluster with the fewest mismatches, then the most matches (UINT32_MAX without clusters),
 * best_mismatches / best_matches = its score.  matrix (optional, [n_reads][n_cons][2] u16) = (matches, mismatches) of every pair:
 * with consensus lists as "reads" this is compare_consensus_snpmers (:968-994) for all cluster pairs of a reclustering round.
 * k = params.k.  Limits: 51200 clusters, 65535 markers per read.  stats: filter_ms = the comparison kernel(s), h2d_ms = index +
 * markers upload, n_pairs = n_reads * n_cons. */
int savont_b200_snpmer_compare(savont_b200_ctx *ctx, const uint64_t *cons_kmers, const uint64_t *cons_offsets, size_t n_cons,
                               const uint64_t *read_kmers, const uint64_t *read_offsets, size_t n_reads, uint32_t *best_cons,
                               uint32_t *best_mismatches, uint32_t *best_matches, uint16_t *matrix /* may be NULL */,
                               savont_b200_stats *stats /* may be NULL */);

/* Stage 3 greedy SNPmer clustering inside the k-mer clusters: the per-cluster loop of asv_cluster::cluster_reads_by_snpmers
 * (src/asv_cluster.rs:561-700, blockmers off).  snp_kmers / snp_offsets: per read the k-mers of TwinRead::snpmer_kmers() (distinct
 * splitmers within a read, at most 512); cluster_reads / cluster_offsets: the k-mer clusters of Stage 2 as read ids in processing
 * order.  The reads of a cluster are visited in order: a read joins the representative it is compatible with (no mismatch, >= 1
 * match) that has the most matches, then the smallest cluster, then the smallest read id (:483-509), else it becomes a
 * representative.  Once a cluster has more than iterative_above representatives (the reference: 1000, :610) the reference asks
 * rayon for ANY compatible representative (:413-464, not deterministic); here, as in the oracle, the first one in creation order
 * wins.  rep_of[r] = the representative's read id (r itself for representatives, UINT32_MAX for reads in no cluster);
 * *n_representatives (may be NULL) = their number.  The grouping / size filter / ordering of the resulting clusters (:660-690) is
 * host work (savont_host::snpmer_clusters).  stats: filter_ms = the kernel, h2d_ms = the upload. */
int savont_b200_cluster_reads_by_snpmers(savont_b200_ctx *ctx, const uint64_t *snp_kmers, const uint64_t *snp_offsets, size_t n_reads,
                                         const uint32_t *cluster_reads, const uint64_t *cluster_offsets, size_t n_clusters,
                                         uint32_t iterative_above, uint32_t *rep_of, uint64_t *n_representatives /* may be NULL */,
                                         savont_b200_stats *stats /* may be NULL */);

/* The same closure from a Vec<TwinRead> as it lies in the reference's memory: an array of structs whose dna_seq and snpmer_kmers
 * are separate heap allocations (src/types.rs:386-412).  One pointer per read: dna_ptrs[r] -> the ceil(lengths[r] / 32) Seq<Dna>
 * words of read r, snpmer_ptrs[r] -> its snpmer_counts[r] k-mers (snpmer_kmer_bytes = 6: the Vec<Kmer48> bytes, 8: u64).  Per
 * internal batch the library gathers the buffers with host threads (savont_b200_set_host_threads, 0 = all cores up to 32) into
 * page-locked staging and uploads from there; the gather of batch i+1 runs under the kernels of batch i.  stats->gather_ms = the
 * host time of the gathers.  Outputs as map_reads. */
int savont_b200_set_host_threads(savont_b200_ctx *ctx, int n_threads);
int savont_b200_map_twin_read_ptrs(savont_b200_ctx *ctx, const uint64_t *const *dna_ptrs, const uint32_t *lengths,
                                   const void *const *snpmer_ptrs, const uint32_t *snpmer_counts, int snpmer_kmer_bytes,
                                   const uint32_t *file_idx, size_t n_reads, int n_samples, uint8_t *status, int32_t *best_nm,
                                   uint64_t *class_offsets, uint32_t *class_asv, size_t class_cap, uint64_t *unambig,
                                   uint64_t *ambig, uint64_t *leq10, uint64_t *totals, uint64_t *sample_singleton,
                                   savont_b200_stats *stats);

/* ---- One process, several GPUs.  `savont asv` is a single process whose rayon pool walks the reads
 * (src/alignment.rs:1786, src/main.rs:476-480): the drop-in for it on a multi-GPU box is ONE handle.  It owns a context per device
 * (device_ids may repeat a device: several contexts on one GPU), replicates parameters / SNPmer set / ASV index, shards the reads
 * of a call into contiguous ranges in read order (src/main.rs:538), runs every shard from its own host thread, sums the counter
 * vectors [unambig | ambig | leq10 | assigned | filtered | singleton S x A] across the devices and returns the per-read results
 * concatenated in read order -- the same outputs as the single-context calls.  The sum is the path's only collective:
 * savont_b200_multi_set_allreduce(0) (default) = one kernel on the first device that reads and rewrites all devices' vectors through
 * peer memory (NVLink), (1) = ncclAllReduce (libnccl.so.2 is loaded on demand).  A failing device reports its code and message. */
typedef struct savont_b200_multi savont_b200_multi;
int savont_b200_create_multi(const int *device_ids, int n_dev, savont_b200_multi **out);
int savont_b200_multi_destroy(savont_b200_multi *m);
int savont_b200_multi_num_devices(savont_b200_multi *m, int *n_dev);
int savont_b200_multi_ctx(savont_b200_multi *m, int i, savont_b200_ctx **ctx);       /* the i-th device's context (borrowed) */
int savont_b200_multi_set_allreduce(savont_b200_multi *m, int mode);
int savont_b200_multi_set_host_threads(savont_b200_multi *m, int n_threads);         /* gather threads over all devices; 0 = all cores */
int savont_b200_multi_set_params(savont_b200_multi *m, const savont_b200_params *p);
int savont_b200_multi_set_batch_reads(savont_b200_multi *m, size_t max_reads_per_batch);
int savont_b200_multi_load_snpmers(savont_b200_multi *m, const uint64_t *kmers, size_t n);
int savont_b200_multi_load_asvs(savont_b200_multi *m, const uint8_t *bases, const uint64_t *offsets, size_t n_asv);
int savont_b200_multi_map_reads(savont_b200_multi *m, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets,
                                const uint32_t *file_idx, size_t n_reads, int n_samples, uint8_t *status, int32_t *best_nm,
                                uint64_t *class_offsets, uint32_t *class_asv, size_t class_cap, uint64_t *unambig, uint64_t *ambig,
                                uint64_t *leq10, uint64_t *totals, uint64_t *sample_singleton, savont_b200_stats *stats);
int savont_b200_multi_map_twin_reads(savont_b200_multi *m, const uint64_t *dna_words, const uint64_t *dna_word_offsets,
                                     const uint32_t *lengths, const void *snpmer_kmers, int snpmer_kmer_bytes,
                                     const uint64_t *snpmer_offsets, const uint32_t *file_idx, size_t n_reads, int n_samples,
                                     uint8_t *status, int32_t *best_nm, uint64_t *class_offsets, uint32_t *class_asv, size_t class_cap,
                                     uint64_t *unambig, uint64_t *ambig, uint64_t *leq10, uint64_t *totals, uint64_t *sample_singleton,
                                     savont_b200_stats *stats);
int savont_b200_multi_map_twin_read_ptrs(savont_b200_multi *m, const uint64_t *const *dna_ptrs, const uint32_t *lengths,
                                         const void *const *snpmer_ptrs, const uint32_t *snpmer_counts, int snpmer_kmer_bytes,
                                         const uint32_t *file_idx, size_t n_reads, int n_samples, uint8_t *status, int32_t *best_nm,
                                         uint64_t *class_offsets, uint32_t *class_asv, size_t class_cap, uint64_t *unambig,
                                         uint64_t *ambig, uint64_t *leq10, uint64_t *totals, uint64_t *sample_singleton,
                                         savont_b200_stats *stats);

/* Pinned host memory helpers (inputs placed here are copied without a staging pass). */
int savont_b200_host_alloc(size_t bytes, void **out);
int savont_b200_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
