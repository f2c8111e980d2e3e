// savont_host.hpp -- host layer above the C ABI that mirrors the reference's Rust interface for the
// read-vs-ASV path (same names, argument meaning and error behaviour: errors are fatal, like the
// reference's expect()/unwrap()).  The reference is compiled Rust; Rust is absent from this image, so
// this layer is C++.  It owns NO compute on reads: every per-read step goes through
// libsavont_b200.so (include/savont_b200.h); what runs here is what savont also runs once per call on
// the host: the equivalence-class histogram, the EM and the writers.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/savont_b200.h"

namespace savont_host {

// the Cli fields the path reads (reference src/cli.rs:45-187)
struct Cli {
    int kmer_size = 17;            // src/cli.rs:153-154
    int c = 11;                    // src/cli.rs:83-84
    int blockmer_length = 3;       // src/cli.rs:157-158
    int minimum_base_quality = 25; // src/cli.rs:99-100
    int band = 64;                 // b200: DP half band
    int strand_mode = 0;
    bool low_polymorphism = false; // src/alignment.rs:1730 -> candidate_mode 1 (every ASV sharing minimizers is a candidate)
    size_t min_read_length = 1100; // src/cli.rs:87-88
    size_t max_read_length = 2000; // src/cli.rs:91-92
    double quality_value_cutoff = 98.0; // src/cli.rs:95-96
};

// KmerGlobalInfo reduced to what the path uses: both alleles of every SNPmer, i.e. the u64 values of
// reference src/kmer_comp.rs:41-48
struct KmerGlobalInfo {
    std::vector<uint64_t> snpmers;
    std::vector<uint64_t> high_freq_kmers;   // src/types.rs:801-824 (Kmer48 values), read by twin_reads_from_snpmers
};

// reference src/types.rs:162-190, fields used by Stage 7 / 7b and the writers
struct ConsensusSequence {
    std::vector<uint8_t> sequence;       // decompressed ASV sequence (ACGT)
    size_t depth = 0;
    size_t appended_depth = 0;
    size_t id = 0;
    long long chimera_score = 0;
    size_t unambig_best_read_map_count = 0;
    size_t ambig_read_map_count = 0;
    size_t num_map_leq_10nm = 0;
    std::vector<size_t> per_sample_depths;
};

// the reads of Stage 1.5 in twin_reads order (src/main.rs:538), kept as raw bases + Phred
struct ReadSet {
    const uint8_t *bases = nullptr;
    const uint8_t *quals = nullptr;      // may be null
    const uint64_t *offsets = nullptr;   // [n + 1]
    const uint32_t *file_idx = nullptr;  // TwinRead.file_idx, may be null
    size_t n = 0;
};

// the same reads as the reference hands them to Stage 7 / 7b: `&[TwinRead]` (src/types.rs:386-412), the three fields
// the path reads, laid out back to back without re-encoding (include/savont_b200.h, savont_b200_map_twin_reads)
struct TwinReadSet {
    const uint64_t *dna_words = nullptr;        // TwinRead.dna_seq raw words (bio-seq Seq<Dna>), each read word aligned
    const uint64_t *dna_word_offsets = nullptr; // [n + 1]
    const uint32_t *lengths = nullptr;          // [n] bases per read
    const void *snpmer_kmers = nullptr;         // TwinRead.snpmer_kmers: Kmer48 bytes (6) or u64 (8)
    int snpmer_kmer_bytes = 6;
    const uint64_t *snpmer_offsets = nullptr;   // [n + 1]
    const uint32_t *file_idx = nullptr;         // TwinRead.file_idx, may be null
    size_t n = 0;
};

// sorted equivalence classes with counts (the reference's HashMap<EquivalenceClass, usize>,
// src/alignment.rs:1517-1523,1915-1916, iterated here in lexicographic order)
struct EqClasses {
    std::vector<uint64_t> offsets{0};
    std::vector<uint32_t> asv;
    std::vector<uint64_t> counts;
    uint64_t total_assigned = 0;
    size_t size() const { return counts.size(); }
};

// per-read result of savont_b200_map_reads
struct Assignments {
    std::vector<uint8_t> status;
    std::vector<int32_t> best_nm;
    std::vector<uint64_t> class_offsets;
    std::vector<uint32_t> class_asv;
    std::vector<uint64_t> unambig, ambig, leq10, sample_singleton;
    uint64_t totals[2] = {0, 0};
    savont_b200_stats stats{};
};

// a Vec<TwinRead> as it lies in the reference's memory: one pointer per read into its dna_seq words and its snpmer_kmers
// (include/savont_b200.h, savont_b200_map_twin_read_ptrs); the library gathers the blocks itself
struct TwinReadPtrs {
    const uint64_t *const *dna_ptrs = nullptr;   // [n] -> ceil(lengths[r] / 32) Seq<Dna> words
    const uint32_t *lengths = nullptr;           // [n]
    const void *const *snpmer_ptrs = nullptr;    // [n] -> snpmer_counts[r] k-mers
    const uint32_t *snpmer_counts = nullptr;     // [n]
    int snpmer_kmer_bytes = 6;
    const uint32_t *file_idx = nullptr;
    size_t n = 0;
};

class Mapper {   // RAII over savont_b200_ctx (one GPU) or savont_b200_multi (one process, several GPUs)
public:
    Mapper(int device, const Cli &args, const KmerGlobalInfo &kmer_info);
    // `savont asv` is one process: all GPUs of the box behind one Mapper (reads sharded in read order, counters summed over NVLink)
    Mapper(const std::vector<int> &devices, const Cli &args, const KmerGlobalInfo &kmer_info);
    ~Mapper();
    Mapper(const Mapper &) = delete;
    Mapper &operator=(const Mapper &) = delete;
    void load_asvs(const std::vector<ConsensusSequence> &consensuses);
    Assignments map_reads(const ReadSet &reads, int n_samples);
    Assignments map_reads(const TwinReadSet &reads, int n_samples);
    Assignments map_reads(const TwinReadPtrs &reads, int n_samples);
    savont_b200_ctx *raw() { return ctx_; }          // the (first) device's context: the single-GPU entry points of the other stages
    size_t n_asv() const { return n_asv_; }
    int n_devices() const { return n_dev_; }

private:
    savont_b200_ctx *ctx_ = nullptr;
    savont_b200_multi *multi_ = nullptr;
    int n_dev_ = 1;
    size_t n_asv_ = 0;
};

// reference src/kmer_comp.rs:68-258 + src/main.rs:538 (Stage 1.5, the part Stage 7 depends on): which FASTQ reads become
// TwinReads and in which order Stage 7 sees them.  The per-read tests run on the GPU
// (savont_b200_twin_reads_from_snpmers); the sort by id (:233), the quality cutoff and the stable sort by est_id
// descending (None = 100.0, src/main.rs:538) are done here.  ids[i] is the read id (FASTQ header up to the first space).
struct TwinReadSelection {
    std::vector<uint8_t> status;          // per input read: 0 kept, 4 / 5 / 6 / 7 (include/savont_b200.h)
    std::vector<double> est_id;           // NaN = None
    std::vector<uint32_t> n_solid_mini, n_solid_snp;
    std::vector<uint32_t> order;          // input indices of the kept reads in Stage-7 order
    bool auto_low_polymorphism = false;   // > 75 % of the kept reads without SNPmer positions (src/main.rs:539-543)
};
TwinReadSelection twin_reads_from_snpmers(Mapper &mapper, const ReadSet &reads, const std::vector<std::string> &ids,
                                          const KmerGlobalInfo &kmer_info, const Cli &args);
// order only (also used by the flat C wrapper): stable by id, then stable by est_id descending
std::vector<uint32_t> twin_read_order(const std::vector<uint8_t> &status, const std::vector<double> &est_id,
                                      const uint8_t *ids, const uint64_t *id_offsets);

// reference src/seq_parse.rs:12-83 (Stage 1 counting, bloom_filter_size == 0): the filtered (k-mer, [c0, c1]) vector that
// get_snpmers_inplace_sort consumes.  rc[i] != 0 marks a read whose id ends in the token "rc" (:361-369); rc may be null.
struct SplitKmerCounts {
    std::vector<uint64_t> kmers;
    std::vector<uint32_t> counts;        // 2 per k-mer
    uint64_t n_distinct = 0;             // "Total SNPmers after bloom filter", :26-28
};
SplitKmerCounts read_to_split_kmers(Mapper &mapper, const ReadSet &reads, const uint8_t *rc, bool single_strand);
// reference src/kmer_comp.rs:459-500: high_freq_thresh = max(100, sorted(c0 + c1)[len - len/100000 - 1]); returns it and
// fills the high-frequency k-mers (c0 + c1 > thresh), the set twin_reads_from_snpmers reads
uint32_t high_freq_kmers(const SplitKmerCounts &counts, std::vector<uint64_t> &high_freq);

// sample < 0: all reads; otherwise only reads with file_idx == sample (src/alignment.rs:2095-2096)
EqClasses build_eq_classes(const Assignments &a, const uint32_t *file_idx, int sample);

// reference src/alignment.rs:1957-2009: returns the iteration count; depths = round(ab * total)
int run_em(size_t n_asv, const EqClasses &classes, std::vector<double> &abundances, std::vector<uint64_t> &depths);

// reference src/alignment.rs:1723-1729 (Stage 7).  Updates depth and the three counters, drops ASVs
// whose EM depth is 0 (:2031).  temp_dir receives final_asvs_for_em.fasta like the reference.
void refine_asv_depths_with_em(const ReadSet &twin_reads, std::vector<ConsensusSequence> &consensuses,
                               const KmerGlobalInfo &kmer_info, const Cli &args, const std::string &temp_dir, int device = 0);
void refine_asv_depths_with_em(const TwinReadSet &twin_reads, std::vector<ConsensusSequence> &consensuses,
                               const KmerGlobalInfo &kmer_info, const Cli &args, const std::string &temp_dir, int device = 0);

// reference src/alignment.rs:2044-2051 (Stage 7b): result[asv][sample]
std::vector<std::vector<size_t>> compute_per_sample_depths(const ReadSet &twin_reads, size_t n_samples,
                                                           const std::vector<ConsensusSequence> &consensuses,
                                                           const KmerGlobalInfo &kmer_info, const Cli &args, int device = 0);
std::vector<std::vector<size_t>> compute_per_sample_depths(const TwinReadSet &twin_reads, size_t n_samples,
                                                           const std::vector<ConsensusSequence> &consensuses,
                                                           const KmerGlobalInfo &kmer_info, const Cli &args, int device = 0);

// reference src/alignment.rs:830-860 and src/main.rs:381-400
void write_consensus_fasta(const std::vector<ConsensusSequence> &consensuses, const std::string &path, const std::string &prefix);
void write_feature_table(const std::vector<ConsensusSequence> &consensuses, const std::string &path,
                         const std::vector<std::string> &sample_names);
// reference src/alignment.rs:1766-1768, :1875-1887 (temp/read_to_asv_mappings.tsv): one line per row of
// savont_b200_mapping_rows_resident, "{read id}\tdebug_id:{consensus id}\t{SNPmer mismatches}\t{NM}" (the reference's variable
// names for the last two columns are swapped, the columns are these), reads in index order (the reference: rayon completion order)
void write_read_to_asv_mappings(const std::string &path, const std::vector<std::string> &read_ids, const std::vector<size_t> &consensus_ids,
                                const uint64_t *row_offsets, const uint32_t *row_asv, const int32_t *row_nm, const uint32_t *row_snp_mismatches);

// ---- the neighbouring stages: host halves around the device calls ----
// reference src/asv_cluster.rs:72-249 (cluster_reads_by_kmers): the greedy loop runs on the GPU
// (savont_b200_cluster_reads_by_kmers on TwinRead::minimizer_kmers() in CSR), the cluster lists of :192-210 are built here:
// grouped by representative, members ascending, larger clusters first (equal sizes: by the smallest member -- the reference
// orders them by `a.first()` of a list filled in std::HashMap iteration order), clusters below min_cluster_size dropped.
std::vector<std::vector<size_t>> clusters_from_assignment(const uint32_t *rep_of, size_t n_reads, size_t min_cluster_size);
std::vector<std::vector<size_t>> cluster_reads_by_kmers(Mapper &mapper, const uint64_t *minimizer_kmers, const uint64_t *minimizer_offsets,
                                                        size_t n_reads, double primary_clustering_threshold, size_t min_cluster_size);
// reference src/asv_cluster.rs:561-700 (cluster_reads_by_snpmers, per-cluster loop; blockmers off): the greedy loop of every
// k-mer cluster runs on the GPU (savont_b200_cluster_reads_by_snpmers), the lists of :660-690 are built here -- per k-mer cluster
// its SNPmer clusters: members ascending, larger first (equal sizes: smaller first member), clusters below min_cluster_size dropped.
std::vector<std::vector<std::vector<size_t>>> snpmer_clusters_from_assignment(const std::vector<std::vector<size_t>> &kmer_clusters,
                                                                              const uint32_t *rep_of, size_t min_cluster_size);
std::vector<std::vector<std::vector<size_t>>> cluster_reads_by_snpmers(Mapper &mapper, const uint64_t *snpmer_kmers,
                                                                       const uint64_t *snpmer_offsets, size_t n_reads,
                                                                       const std::vector<std::vector<size_t>> &kmer_clusters,
                                                                       size_t min_cluster_size, uint32_t iterative_above = 1000);
// reference src/asv_cluster.rs:1007-1131 (reassign_reads_to_best_cluster, SNPmer markers): `cons_*` = the ConsensusSnpmer k-mers
// of the current clusters, `snpmer_*` = snpmers_vec() k-mers of all reads in CSR by read id.  Returns the new clusters (members
// ascending, empty and small ones dropped) and the number of reads that changed cluster.
std::pair<std::vector<std::vector<size_t>>, size_t> clusters_from_reassignment(const std::vector<std::vector<size_t>> &clusters,
                                                                                const uint32_t *best_cons, size_t min_cluster_size);
std::pair<std::vector<std::vector<size_t>>, size_t> reassign_reads_to_best_cluster(
    Mapper &mapper, const std::vector<std::vector<size_t>> &clusters, const uint64_t *cons_kmers, const uint64_t *cons_offsets,
    const uint64_t *snpmer_kmers, const uint64_t *snpmer_offsets, size_t min_cluster_size);
// reference src/alignment.rs:17-60: one column of a consensus pileup
struct PileupBase { uint8_t kind; uint8_t n; uint8_t base[2], qual[2], hp[2]; };   // kind 0 Base, 1 Deletion, 2 Insertion (n bases)
struct Pileup { size_t position = 0; uint8_t ref_base = 0; uint8_t ref_hp_length = 1; std::vector<PileupBase> bases; };
// reference src/alignment.rs:415-641 (generate_consensus_pileups): reads in cluster order (read_cons non-decreasing)
std::vector<std::vector<Pileup>> generate_consensus_pileups(Mapper &mapper, const std::vector<ConsensusSequence> &consensuses,
                                                            const ReadSet &reads, const uint32_t *read_cons, bool use_hpc);

}  // namespace savont_host
