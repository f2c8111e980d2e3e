/*
 * savont_oracle.h -- CPU oracle for the read-vs-ASV assignment path of savont v0.6.4.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a scalar, sequential C restatement of the
 * reference's Rust logic (Stage 7 / 7b of `savont asv`).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product (libsavont_b200.so) never links or calls it.
 *
 * PARITY STATUS
 *   - sketch (a1-a3), ASV index (a5-a7), candidate filter (a8-a10), tie/eq-class
 *     counters (a12), EM + depths (a13-a14): restated line by line from the Rust
 *     source; pinned by the reference's own known-answer tests (src/types.rs:1108-1140)
 *     and by mm_hash64 / powi constants derived from the formulas.
 *   - alignment NM (a11): the reference calls minimap2 2.30 (`lr:hq`) through the
 *     minimap2 crate 0.1.31; that C library is NOT in /root/reference and no Rust or
 *     minimap2 toolchain exists in this image, so its per-read NM cannot be
 *     reproduced or checked.  The oracle implements the north-star definition instead
 *     (banded, both-ends-free unit-cost edit distance, SURVEY.md 8c "S2").
 *     ==> NM / CIGAR parity against the reference is UNPINNED ("parity unpinned").
 *         GPU-vs-oracle parity is bit-exact.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/).
 */
#ifndef SAVONT_ORACLE_H
#define SAVONT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct so_params {
    int32_t k;            /* src/cli.rs:153-154  default 17 (odd) */
    int32_t c;            /* src/cli.rs:83-84    default 11 */
    int32_t l;            /* src/cli.rs:157-158  default 3 (blockmers: unused on this path) */
    int32_t min_bq;       /* src/cli.rs:99-100   default 25 */
    int32_t band;         /* DP half band B (|i-j| <= B) */
    int32_t strand_mode;  /* 0 = orientation by minimizer seed vote, 1 = both strands, min NM */
    int32_t candidate_mode; /* 0 = SNPmer index candidates (src/alignment.rs:1786-1843);
                               1 = low-polymorphism mode: every ASV sharing minimizers (src/alignment.rs:1527-1719) */
    int32_t nm_mode;      /* 0 = the product's NM definition (banded S2).  ORACLE ONLY, sensitivity study of a11:
                             1 = S2 with band 4B, 2 = local dual-affine alignment with minimap2's lr:hq scoring */
    double  ratio_thresh;        /* src/alignment.rs:1829 0.0050 */
    double  min_mini_frac_base;  /* src/alignment.rs:1806 0.950 (raised to k) */
} so_params;

/* status codes of per-read assignment (shared with the product ABI) */
enum {
    SO_ASSIGNED = 0,   /* eq-class emitted                      src/alignment.rs:1890-1920 */
    SO_NO_CANDIDATE = 1, /* asv_scores empty                    src/alignment.rs:1817-1820 */
    SO_RATIO_FILTERED = 2, /* no candidate with ratio <= thresh src/alignment.rs:1835-1838 */
    SO_NO_ALIGNMENT = 3, /* best_alns empty                     src/alignment.rs:1921-1924 */
    SO_TOO_SHORT = 4,  /* get_twin_read_syncmer returned None   src/seeding.rs:339-341 */
    /* read-side wrapper twin_reads_from_snpmers (a4): the read never becomes a TwinRead */
    SO_LENGTH_FILTERED = 5, /* len < min_read_length || len > max_read_length   src/kmer_comp.rs:117-120 */
    SO_REPETITIVE = 6,      /* #solid minimizers < len / c / 20                 src/kmer_comp.rs:185-188 */
    SO_LOW_QUALITY = 7      /* est_id < quality_value_cutoff                    src/kmer_comp.rs:248 */
};

/* a4 options: the Cli fields / constants twin_reads_from_snpmers reads */
typedef struct so_read_filter {
    uint32_t min_read_length;        /* src/cli.rs:87-88   default 1100; 0 = off */
    uint32_t max_read_length;        /* src/cli.rs:91-92   default 2000; 0 = no upper bound */
    uint32_t max_kmer_count_in_read; /* MAX_KMER_COUNT_IN_READ = 500, src/constants.rs */
    uint32_t reserved;
    double   quality_value_cutoff;   /* src/cli.rs:95-96   default 98.0; <= 0 = off */
} so_read_filter;

/* ---- primitives with known answers ------------------------------------------------ */
uint8_t  so_encode_base(uint8_t b);                 /* src/types.rs:92-101 */
uint64_t so_mm_hash64(uint64_t key);                /* src/seeding.rs:18-28 */
uint64_t so_kmer_from_ascii(const char *s, int k);  /* src/types.rs:1118-1128 convention */
uint64_t so_kmer_revcomp(uint64_t kmer, int k);     /* src/types.rs:559-568, 654 */
void     so_kmer48_from_u64(uint64_t v, uint8_t out[6]); /* src/types.rs:112-127 */
uint64_t so_kmer48_to_u64(const uint8_t in[6]);
double   so_powi(double a, int b);                  /* compiler-rt __powidf2, used by f64::powi */
uint64_t so_split_mask(int k);                      /* src/seeding.rs:353 */

/* ---- a3: sketch a batch of sequences (src/seeding.rs:317-658) ------------------------
 * bases/quals: concatenated ASCII; offsets[n+1].  quals may be NULL (FASTA / ASV side).
 * Outputs are CSR; *_cap are capacities of the flat arrays.  mini_fwd[i]=1 iff the
 * canonical k-mer equals the forward-strand k-mer at that position (b200 extension used
 * for the strand vote).  codes: one 2-bit code per base (N->0).  est_id = NaN when None.
 * Returns 0, or -1 if a capacity was too small.
 */
int so_sketch_batch(const so_params *p, const uint64_t *snpmers, size_t n_snpmers,
                    const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets, size_t n_seqs,
                    uint8_t *status, double *est_id,
                    uint64_t *mini_offsets, uint32_t *mini_pos, uint64_t *mini_kmer, uint8_t *mini_fwd, size_t mini_cap,
                    uint64_t *snp_offsets, uint32_t *snp_pos, uint64_t *snp_kmer, size_t snp_cap,
                    uint8_t *codes);

/* ---- context: params + SNPmer set + ASV sketches/index (a5-a7) ---------------------- */
typedef struct so_ctx so_ctx;
so_ctx *so_create(const so_params *p, const uint64_t *snpmers, size_t n_snpmers);
void    so_destroy(so_ctx *ctx);
/* src/kmer_comp.rs:39-66 + src/alignment.rs:1754-1765 */
int     so_load_asvs(so_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, size_t n_asv);
size_t  so_num_asvs(const so_ctx *ctx);

/* ---- a8-a10: candidate filter (src/asv_cluster.rs:356-383, src/alignment.rs:1786-1843)
 * Per read: cand_status in {0 kept>0, SO_NO_CANDIDATE, SO_RATIO_FILTERED, SO_TOO_SHORT};
 * kept candidates (minimum SNPmer mismatches among ratio<=thresh) sorted by asv id:
 * cand_asv, cand_mm (snpmer mismatches), cand_mini (shared distinct minimizers),
 * cand_strand (0 fwd / 1 rev, by seed vote).
 */
int so_find_candidates(so_ctx *ctx, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets, size_t n_reads,
                       uint8_t *cand_status, uint64_t *cand_offsets,
                       uint32_t *cand_asv, uint32_t *cand_mm, uint32_t *cand_mini, uint8_t *cand_strand, size_t cand_cap);

/* ---- a11: banded both-ends-free edit distance (north-star definition) --
 * pair i aligns read pair_read[i] (strand pair_strand[i]) against ASV pair_asv[i].
 * nm, end cell (asv_end, read_end) always; with want_cigar also start cell and CIGAR
 * (BAM encoding len<<4|op, op 0=M 1=I(read only) 2=D(ASV only)), CSR via cigar_offsets.
 * Coordinates are in DP orientation (read reverse-complemented when strand==1).
 */
int so_align_pairs(so_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, size_t n_reads,
                   const uint32_t *pair_read, const uint32_t *pair_asv, const uint8_t *pair_strand, size_t n_pairs,
                   int32_t *nm, int32_t *asv_end, int32_t *read_end,
                   int want_cigar, int32_t *asv_start, int32_t *read_start,
                   uint64_t *cigar_offsets, uint32_t *cigar, size_t cigar_cap);

/* ---- whole path: a3 + a8-a12 (src/alignment.rs:1786-1925; Stage 7b :2095-2171) ------
 * file_idx may be NULL (single sample).  n_samples>=1.  Counters are ADDED to (caller zeroes).
 * unambig/ambig/leq10: [A]; totals[0]=assigned, totals[1]=filtered;
 * sample_singleton: [n_samples*A] counts of singleton classes per (sample, asv) (may be NULL).
 * n_threads: OpenMP threads (<=0: all).
 */
/* NM of one oriented pair of 2-bit code strings under p->nm_mode; returns 0 when the pair has no alignment (mode 2) */
int so_nm_pair(const so_params *p, const uint8_t *asv_codes, int n, const uint8_t *read_codes, int m, int32_t *nm);

int so_map_reads(so_ctx *ctx, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets,
                 const uint32_t *file_idx, size_t n_reads, int n_samples,
                 uint8_t *status, int32_t *best_nm, uint64_t *class_offsets, uint32_t *class_asv, size_t class_cap,
                 uint64_t *unambig, uint64_t *ambig, uint64_t *leq10, uint64_t *totals, uint64_t *sample_singleton,
                 uint64_t *dp_cells, int n_threads);

/* ---- a4: read-side wrapper kmer_comp::twin_reads_from_snpmers (src/kmer_comp.rs:68-258) and the read order of
 * src/main.rs:538.  Per read, in the reference's order of tests: length filter (:117-120), sketch (None -> SO_TOO_SHORT),
 * per-read minimizer multiplicities, "solid" = multiplicity <= max_kmer_count_in_read and not in high_freq (:163-183;
 * USE_SOLID_KMERS = false, src/constants.rs), repetitive-read drop (:185-188), then after the id sort the quality cutoff
 * (:248, only reads that have an est_id).  n_solid_mini / n_solid_snp = sizes of the position vectors that
 * retain_mini_indices / retain_snpmer_indices leave (src/types.rs:702-715; the k-mer vectors are NOT pruned).
 * ids: concatenated read ids with id_offsets[n+1] (may be NULL: order by input index).  order[0..*n_kept) = input indices
 * of the kept reads in the reference's Stage-7 read order: sort by id (:233, byte-wise String order, stable), then stable
 * sort by est_id descending with None as 100.0 (src/main.rs:538). */
int so_twin_reads_filter(const so_params *p, const so_read_filter *f, const uint64_t *snpmers, size_t n_snpmers,
                         const uint64_t *high_freq, size_t n_high_freq,
                         const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets, size_t n_reads,
                         const uint8_t *ids, const uint64_t *id_offsets,
                         uint8_t *status, double *est_id, uint32_t *n_solid_mini, uint32_t *n_solid_snp,
                         uint32_t *order, size_t *n_kept);
/* the same tests inside the whole path: reads that fail get their a4 status and are counted neither as assigned nor as
 * filtered (they do not exist in Stage 7).  f == NULL switches the tests off. */
void so_set_read_filter(so_ctx *ctx, const so_read_filter *f, const uint64_t *high_freq, size_t n_high_freq);

/* ---- Stage 1 counting (the integer part): seq_parse::read_to_split_kmers with
 * bloom_filter_size == 0 (src/seq_parse.rs:12-83, :315-497) over seeding::split_kmer_mid (src/seeding.rs:975-1068).
 * Every k-mer of every read whose splitmer is not palindromic (:1044-1046) and whose middle base has Phred >= min_bq
 * (:1009-1018; only when the read's qualities are not all equal) is counted per canonical k-mer and per strand flag
 * (counts[flag], flag = 1 iff the forward strand is the canonical one, :1053-1063; src/seq_parse.rs:455-460).
 * rc[r] != 0 (may be NULL): the last whitespace-separated token of the read id is "rc" (src/seq_parse.rs:361-369): the
 * read is reverse-complemented (and its qualities reversed) first.
 * Output: the entries that pass the filter of src/seq_parse.rs:30-44 (both strands seen and more than 2 occurrences; with
 * single_strand: counts[0] > 2), sorted by k-mer (the reference's order is hash-map order); out_counts = [c0, c1] per
 * entry.  *n_distinct = number of distinct k-mers before the filter ("Total SNPmers after bloom filter", :27).
 * Returns the number of entries, or -1 if cap is too small. */
long long so_count_split_kmers(int k, int min_bq, int single_strand, const uint8_t *bases, const uint8_t *quals,
                               const uint64_t *offsets, const uint8_t *rc, size_t n_reads, uint64_t *out_kmers,
                               uint32_t *out_counts, size_t cap, uint64_t *n_distinct, int n_threads);
/* src/kmer_comp.rs:459-500: high_freq_thresh = max(100, sorted(c0+c1)[len - len/100000 - 1]); the high-frequency set =
 * k-mers with c0 + c1 > high_freq_thresh (the set a4 reads), in input order.  Returns the threshold; n must be > 0. */
uint32_t so_high_freq_kmers(const uint64_t *kmers, const uint32_t *counts, size_t n, uint64_t *hf_out, size_t *n_hf);

/* ---- Stage 4 pileups, alignment::generate_consensus_pileups (src/alignment.rs:415-641).
 * Reads arrive in cluster order, at most MAX_SEQS_CONSENSUS = 250 per consensus (src/constants.rs:60), read_cons[r] = the
 * consensus of read r (non-decreasing); read_quals = the per-base qualities the reference expands from qual_seq
 * (:456-478; NULL = 33 everywhere).  Each read is homopolymer-compressed (src/utils.rs:135-184) and aligned to its
 * consensus (already compressed) with the north-star DP on both strands (replaces minimap2 `map_ont`: parity unpinned);
 * the CIGAR walk of :524-569 fills one column per consensus position.  Output: columns in CSR (pos_offsets over the
 * concatenated consensus positions), entries in push order (read order, then CIGAR order); ref_hp_length = the median
 * run length of :589-622 (1 when no base is aligned); read_strand 0/1 (0xff: not aligned), read_nm.
 * Returns the number of entries, or -1 if entry_cap is too small. */
typedef struct so_pileup_entry {
    uint32_t read;       /* index of the read in the call */
    uint8_t kind;        /* 0 = Base, 1 = Deletion, 2 = Insertion (src/alignment.rs:17-26) */
    uint8_t n;           /* bases carried: 1, 0, 1..MAX_INSERTION_LENGTH (= 2, src/constants.rs:3) */
    uint8_t base[2], qual[2], hp[2];
} so_pileup_entry;
long long so_consensus_pileups(int band, int use_hpc, const uint8_t *cons_bases, const uint64_t *cons_offsets, size_t n_cons,
                               const uint8_t *read_bases, const uint8_t *read_quals, const uint64_t *read_offsets,
                               const uint32_t *read_cons, size_t n_reads, uint64_t *pos_offsets, so_pileup_entry *entries,
                               size_t entry_cap, uint8_t *ref_hp_length, uint8_t *read_strand, int32_t *read_nm);

/* ---- SURVEY.md 8(f) row 3 (Stage 2): LSH signatures and greedy k-mer clustering.
 * Input: the reads' minimizer k-mers (TwinRead::minimizer_kmers(), duplicates kept, in position order) in CSR.
 * so_lsh_signatures: TwinRead::compute_lsh_signatures, src/types.rs:719-747: sig[r*20 + t], valid[r] = 0 for reads with fewer
 *   than LSH_BUCKET_SIZE = 3 minimizers (all their signatures are None).
 * so_cluster_reads_by_kmers: asv_cluster::cluster_reads_by_kmers, src/asv_cluster.rs:72-190 (use_bucketed = true):
 *   rep_of[r] = the representative read r was assigned to (r itself if it became a representative). */
int so_lsh_signatures(const uint64_t *kmers, const uint64_t *offsets, size_t n_reads, uint64_t *sig, uint8_t *valid);
int so_cluster_reads_by_kmers(int k, double threshold, const uint64_t *kmers, const uint64_t *offsets, size_t n_reads, uint32_t *rep_of,
                              uint64_t *n_reps);

/* ---- SURVEY.md 8(f) row 3 (Stage 3 reclustering): reads (or consensus lists) against cluster consensus SNPmers.
 * cons_kmers / cons_offsets: per cluster the ConsensusSnpmer k-mers (splitmer = kmer & split_mask); read_kmers / read_offsets:
 * per read the k-mers of TwinRead::snpmers_vec().  Per read: best_cons = the first cluster with the fewest mismatches, then the
 * most matches (reassign_reads_to_best_cluster, src/asv_cluster.rs:1052-1098; UINT32_MAX without clusters), its mismatches and
 * matches; matrix (optional) [n_reads][n_cons][2] = (matches, mismatches) of every pair, saturated at 65535
 * (compare_consensus_snpmers, :968-994, when the "reads" are consensus lists). */
int so_snpmer_compare(int k, const uint64_t *cons_kmers, const uint64_t *cons_offsets, size_t n_cons, const uint64_t *read_kmers,
                      const uint64_t *read_offsets, size_t n_reads, uint32_t *best_cons, uint32_t *best_mismatches, uint32_t *best_matches,
                      uint16_t *matrix);

/* ---- Stage 3 greedy loop for ONE k-mer cluster (asv_cluster::cluster_reads_by_snpmers, src/asv_cluster.rs:561-700; restated ahead of
 * a device implementation).  snp_kmers / snp_offsets: snpmer_kmers() of all reads in CSR by read id; cluster[n]: the read ids of the
 * k-mer cluster in order; rep_of[read] (only the cluster's reads are written) = the SNPmer-cluster representative.  iterative_above:
 * the representative count above which the reference switches to its parallel search (1000, :613); there the first compatible
 * representative in creation order is taken (the reference's find_any is not deterministic). */
int so_cluster_reads_by_snpmers(int k, const uint64_t *snp_kmers, const uint64_t *snp_offsets, const uint32_t *cluster, size_t n,
                                uint32_t iterative_above, uint32_t *rep_of);

/* ---- a13-a14: EM over equivalence classes (src/alignment.rs:1957-2038) ---------------
 * classes in CSR (cls_offsets[n_cls+1], cls_asv), counts[n_cls]; total_assigned;
 * out: abundances[A] (f64), depths[A] = round(ab*total).  Classes are iterated in the
 * given order (the reference iterates a randomly seeded std::HashMap; parity target is
 * the integer depths).  Returns the number of EM iterations.
 */
int so_em(size_t n_asv, const uint64_t *cls_offsets, const uint32_t *cls_asv, const uint64_t *counts, size_t n_cls,
          uint64_t total_assigned, double *abundances, uint64_t *depths);

#ifdef __cplusplus
}
#endif
#endif
