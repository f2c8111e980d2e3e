// savont_host.cpp -- see savont_host.hpp
#include "savont_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

namespace savont_host {

[[noreturn]] static void die(const std::string &what) {
    // the reference panics (expect/unwrap) on this path: src/alignment.rs:1853,1858
    throw std::runtime_error(what + ": " + savont_b200_last_error());
}
static void ck(int rc, const char *what) { if (rc != 0) die(what); }

static savont_b200_params params_of(const Cli &args) {
    savont_b200_params p;
    savont_b200_default_params(&p);
    p.k = args.kmer_size; p.c = args.c; p.l = args.blockmer_length; p.min_bq = args.minimum_base_quality;
    p.band = args.band; p.strand_mode = args.strand_mode; p.candidate_mode = args.low_polymorphism ? 1 : 0;
    return p;
}
Mapper::Mapper(int device, const Cli &args, const KmerGlobalInfo &kmer_info) {
    ck(savont_b200_create(device, &ctx_), "savont_b200_create");
    const savont_b200_params p = params_of(args);
    ck(savont_b200_set_params(ctx_, &p), "savont_b200_set_params");
    ck(savont_b200_load_snpmers(ctx_, kmer_info.snpmers.data(), kmer_info.snpmers.size()), "savont_b200_load_snpmers");
}
Mapper::Mapper(const std::vector<int> &devices, const Cli &args, const KmerGlobalInfo &kmer_info) {
    if (devices.empty()) throw std::runtime_error("Mapper: empty device list");
    ck(savont_b200_create_multi(devices.data(), (int)devices.size(), &multi_), "savont_b200_create_multi");
    n_dev_ = (int)devices.size();
    const savont_b200_params p = params_of(args);
    ck(savont_b200_multi_set_params(multi_, &p), "savont_b200_multi_set_params");
    ck(savont_b200_multi_load_snpmers(multi_, kmer_info.snpmers.data(), kmer_info.snpmers.size()), "savont_b200_multi_load_snpmers");
    ck(savont_b200_multi_ctx(multi_, 0, &ctx_), "savont_b200_multi_ctx");
}
Mapper::~Mapper() {
    if (multi_) savont_b200_multi_destroy(multi_);         // owns the per-device contexts, ctx_ included
    else if (ctx_) savont_b200_destroy(ctx_);
}

void Mapper::load_asvs(const std::vector<ConsensusSequence> &consensuses) {
    std::vector<uint8_t> bases; std::vector<uint64_t> off{0};
    for (const auto &c : consensuses) { bases.insert(bases.end(), c.sequence.begin(), c.sequence.end()); off.push_back(bases.size()); }
    if (multi_) ck(savont_b200_multi_load_asvs(multi_, bases.data(), off.data(), consensuses.size()), "savont_b200_multi_load_asvs");
    else ck(savont_b200_load_asvs(ctx_, bases.data(), off.data(), consensuses.size()), "savont_b200_load_asvs");
    n_asv_ = consensuses.size();
}

// one call into the library with growing class capacity; `call(cap)` returns the library's status
template <typename Call>
static Assignments map_with(size_t n, size_t A, int n_samples, Call call) {
    Assignments a;
    a.status.assign(n ? n : 1, 0); a.best_nm.assign(n ? n : 1, 0); a.class_offsets.assign(n + 1, 0);
    a.unambig.assign(A ? A : 1, 0); a.ambig.assign(A ? A : 1, 0); a.leq10.assign(A ? A : 1, 0);
    a.sample_singleton.assign(A * (size_t)n_samples + 1, 0);
    size_t cap = 4 * n + 1024;
    for (int attempt = 0; attempt < 4; ++attempt) {
        a.class_asv.assign(cap, 0);
        std::fill(a.unambig.begin(), a.unambig.end(), 0); std::fill(a.ambig.begin(), a.ambig.end(), 0);
        std::fill(a.leq10.begin(), a.leq10.end(), 0); std::fill(a.sample_singleton.begin(), a.sample_singleton.end(), 0);
        a.totals[0] = a.totals[1] = 0;
        const int rc = call(a, cap);
        if (rc == SAVONT_B200_ERR_CAPACITY) { cap *= 4; continue; }
        ck(rc, "savont_b200_map_reads");
        break;
    }
    a.status.resize(n); a.best_nm.resize(n);
    a.class_asv.resize(a.class_offsets[n]);
    a.unambig.resize(A); a.ambig.resize(A); a.leq10.resize(A); a.sample_singleton.resize(A * (size_t)n_samples);
    return a;
}

Assignments Mapper::map_reads(const ReadSet &r, int n_samples) {
    return map_with(r.n, n_asv_, n_samples, [&](Assignments &a, size_t cap) {
        if (multi_)
            return savont_b200_multi_map_reads(multi_, r.bases, r.quals, r.offsets, r.file_idx, r.n, n_samples, a.status.data(), a.best_nm.data(),
                                               a.class_offsets.data(), a.class_asv.data(), cap, a.unambig.data(), a.ambig.data(),
                                               a.leq10.data(), a.totals, a.sample_singleton.data(), &a.stats);
        return savont_b200_map_reads(ctx_, r.bases, r.quals, r.offsets, r.file_idx, r.n, n_samples, a.status.data(), a.best_nm.data(),
                                     a.class_offsets.data(), a.class_asv.data(), cap, a.unambig.data(), a.ambig.data(), a.leq10.data(),
                                     a.totals, a.sample_singleton.data(), &a.stats);
    });
}

Assignments Mapper::map_reads(const TwinReadSet &r, int n_samples) {
    return map_with(r.n, n_asv_, n_samples, [&](Assignments &a, size_t cap) {
        if (multi_)
            return savont_b200_multi_map_twin_reads(multi_, r.dna_words, r.dna_word_offsets, r.lengths, r.snpmer_kmers, r.snpmer_kmer_bytes,
                                                    r.snpmer_offsets, r.file_idx, r.n, n_samples, a.status.data(), a.best_nm.data(),
                                                    a.class_offsets.data(), a.class_asv.data(), cap, a.unambig.data(), a.ambig.data(),
                                                    a.leq10.data(), a.totals, a.sample_singleton.data(), &a.stats);
        return savont_b200_map_twin_reads(ctx_, r.dna_words, r.dna_word_offsets, r.lengths, r.snpmer_kmers, r.snpmer_kmer_bytes,
                                          r.snpmer_offsets, r.file_idx, r.n, n_samples, a.status.data(), a.best_nm.data(),
                                          a.class_offsets.data(), a.class_asv.data(), cap, a.unambig.data(), a.ambig.data(),
                                          a.leq10.data(), a.totals, a.sample_singleton.data(), &a.stats);
    });
}

Assignments Mapper::map_reads(const TwinReadPtrs &r, int n_samples) {
    return map_with(r.n, n_asv_, n_samples, [&](Assignments &a, size_t cap) {
        if (multi_)
            return savont_b200_multi_map_twin_read_ptrs(multi_, r.dna_ptrs, r.lengths, r.snpmer_ptrs, r.snpmer_counts, r.snpmer_kmer_bytes,
                                                        r.file_idx, r.n, n_samples, a.status.data(), a.best_nm.data(), a.class_offsets.data(),
                                                        a.class_asv.data(), cap, a.unambig.data(), a.ambig.data(), a.leq10.data(), a.totals,
                                                        a.sample_singleton.data(), &a.stats);
        return savont_b200_map_twin_read_ptrs(ctx_, r.dna_ptrs, r.lengths, r.snpmer_ptrs, r.snpmer_counts, r.snpmer_kmer_bytes, r.file_idx, r.n,
                                              n_samples, a.status.data(), a.best_nm.data(), a.class_offsets.data(), a.class_asv.data(), cap,
                                              a.unambig.data(), a.ambig.data(), a.leq10.data(), a.totals, a.sample_singleton.data(), &a.stats);
    });
}

std::vector<uint32_t> twin_read_order(const std::vector<uint8_t> &status, const std::vector<double> &est_id,
                                      const uint8_t *ids, const uint64_t *id_offsets) {
    std::vector<uint32_t> order;
    for (size_t i = 0; i < status.size(); ++i) if (status[i] == 0) order.push_back((uint32_t)i);
    if (ids) {   // twin_reads.sort_by(|a, b| a.id.cmp(&b.id)): byte-wise String order, stable (src/kmer_comp.rs:233)
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
            const size_t lx = id_offsets[x + 1] - id_offsets[x], ly = id_offsets[y + 1] - id_offsets[y];
            const int c = std::memcmp(ids + id_offsets[x], ids + id_offsets[y], std::min(lx, ly));
            return c != 0 ? c < 0 : lx < ly;
        });
    }
    // sort_by(|a, b| b.est_id.unwrap_or(100.0).partial_cmp(&a.est_id.unwrap_or(100.0))): descending, stable (src/main.rs:538)
    auto key = [&](uint32_t i) { return std::isnan(est_id[i]) ? 100.0 : est_id[i]; };
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return key(x) > key(y); });
    return order;
}

TwinReadSelection twin_reads_from_snpmers(Mapper &mapper, const ReadSet &r, const std::vector<std::string> &ids,
                                          const KmerGlobalInfo &kmer_info, const Cli &args) {
    savont_b200_read_filter f{};
    f.min_read_length = (uint32_t)args.min_read_length; f.max_read_length = (uint32_t)args.max_read_length;
    f.max_kmer_count_in_read = 500;                       // MAX_KMER_COUNT_IN_READ, src/constants.rs
    f.quality_value_cutoff = args.quality_value_cutoff;
    ck(savont_b200_load_high_freq_kmers(mapper.raw(), kmer_info.high_freq_kmers.data(), kmer_info.high_freq_kmers.size()),
       "savont_b200_load_high_freq_kmers");
    ck(savont_b200_set_read_filter(mapper.raw(), &f), "savont_b200_set_read_filter");
    TwinReadSelection s;
    s.status.assign(r.n ? r.n : 1, 0); s.est_id.assign(r.n ? r.n : 1, 0.0);
    s.n_solid_mini.assign(r.n ? r.n : 1, 0); s.n_solid_snp.assign(r.n ? r.n : 1, 0);
    ck(savont_b200_twin_reads_from_snpmers(mapper.raw(), r.bases, r.quals, r.offsets, r.n, s.status.data(), s.est_id.data(),
                                           s.n_solid_mini.data(), s.n_solid_snp.data()), "savont_b200_twin_reads_from_snpmers");
    ck(savont_b200_set_read_filter(mapper.raw(), nullptr), "savont_b200_set_read_filter");   // Stage 7 takes the kept reads as they are
    s.status.resize(r.n); s.est_id.resize(r.n); s.n_solid_mini.resize(r.n); s.n_solid_snp.resize(r.n);
    std::vector<uint8_t> idbuf; std::vector<uint64_t> idoff{0};
    for (const auto &id : ids) { idbuf.insert(idbuf.end(), id.begin(), id.end()); idoff.push_back(idbuf.size()); }
    const bool have_ids = ids.size() == r.n && r.n > 0;
    s.order = twin_read_order(s.status, s.est_id, have_ids ? idbuf.data() : nullptr, have_ids ? idoff.data() : nullptr);
    size_t without = 0;
    for (uint32_t i : s.order) if (s.n_solid_snp[i] == 0) ++without;      // tr.snpmers_vec().is_empty(), src/main.rs:539
    s.auto_low_polymorphism = !s.order.empty() && (double)without / (double)s.order.size() > 0.75;
    return s;
}

SplitKmerCounts read_to_split_kmers(Mapper &mapper, const ReadSet &r, const uint8_t *rc, bool single_strand) {
    SplitKmerCounts out;
    size_t cap = (size_t)((r.n ? r.offsets[r.n] - r.offsets[0] : 0) / 64) + 4096, n = 0;
    savont_b200_kmer_stats st{};
    for (int attempt = 0; attempt < 3; ++attempt) {
        out.kmers.assign(cap, 0); out.counts.assign(2 * cap, 0);
        const int rcode = savont_b200_count_split_kmers(mapper.raw(), r.bases, r.quals, r.offsets, rc, r.n, single_strand ? 1 : 0,
                                                        out.kmers.data(), out.counts.data(), cap, &n, &st);
        if (rcode == SAVONT_B200_ERR_CAPACITY && n > cap) { cap = n + 16; continue; }
        ck(rcode, "savont_b200_count_split_kmers");
        break;
    }
    out.kmers.resize(n); out.counts.resize(2 * n); out.n_distinct = st.n_distinct;
    return out;
}

uint32_t high_freq_kmers(const SplitKmerCounts &c, std::vector<uint64_t> &high_freq) {
    high_freq.clear();
    const size_t n = c.kmers.size();
    if (n == 0) throw std::runtime_error("No k-mers found");              // the reference exits here, src/kmer_comp.rs:469-472
    std::vector<uint32_t> kc(n);
    for (size_t i = 0; i < n; ++i) kc[i] = c.counts[2 * i] + c.counts[2 * i + 1];
    const size_t idx = n - (n / 100000) - 1;                               // :474
    std::nth_element(kc.begin(), kc.begin() + idx, kc.end());
    const uint32_t thresh = std::max<uint32_t>(kc[idx], 100u);
    for (size_t i = 0; i < n; ++i)
        if (c.counts[2 * i] + c.counts[2 * i + 1] > thresh) high_freq.push_back(c.kmers[i]);     // :497-499
    return thresh;
}

EqClasses build_eq_classes(const Assignments &a, const uint32_t *file_idx, int sample) {
    std::map<std::vector<uint32_t>, uint64_t> hist;
    uint64_t total = 0;
    const size_t n = a.status.size();
    for (size_t i = 0; i < n; ++i) {
        if (a.status[i] != SAVONT_B200_READ_ASSIGNED) continue;
        if (sample >= 0 && (file_idx ? file_idx[i] : 0u) != (uint32_t)sample) continue;
        std::vector<uint32_t> cls(a.class_asv.begin() + a.class_offsets[i], a.class_asv.begin() + a.class_offsets[i + 1]);
        ++hist[cls];
        ++total;
    }
    EqClasses e;
    e.total_assigned = total;
    for (const auto &kv : hist) {
        e.asv.insert(e.asv.end(), kv.first.begin(), kv.first.end());
        e.offsets.push_back(e.asv.size());
        e.counts.push_back(kv.second);
    }
    return e;
}

int run_em(size_t n_asv, const EqClasses &cls, std::vector<double> &ab, std::vector<uint64_t> &depths) {
    ab.assign(n_asv, n_asv ? 1.0 / (double)n_asv : 0.0);                    // :1959
    depths.assign(n_asv, 0);
    if (cls.size() == 0 || cls.total_assigned == 0) return 0;
    const double total_assigned = (double)cls.total_assigned;
    const double convergence_threshold = 0.01 / total_assigned;           // :1960
    std::vector<double> nw(n_asv);
    int iteration = 0;
    const int MAX_ITERATIONS = 10000;
    while (true) {
        ++iteration;
        std::fill(nw.begin(), nw.end(), 0.0);
        for (size_t q = 0; q < cls.size(); ++q) {                           // :1972-1983
            double denominator = 0.0;
            for (uint64_t e = cls.offsets[q]; e < cls.offsets[q + 1]; ++e) denominator += ab[cls.asv[e]];
            if (denominator > 0.0)
                for (uint64_t e = cls.offsets[q]; e < cls.offsets[q + 1]; ++e)
                    nw[cls.asv[e]] += (double)cls.counts[q] * ab[cls.asv[e]] / denominator;
        }
        double total_counts = 0.0;
        for (double v : nw) total_counts += v;
        if (total_counts > 0.0) for (double &v : nw) v /= total_assigned;    // :1986-1991
        double max_change = 0.0;
        for (size_t a = 0; a < n_asv; ++a) max_change = std::max(max_change, std::fabs(ab[a] - nw[a]));
        ab = nw;
        if (max_change < convergence_threshold || iteration >= MAX_ITERATIONS) break;   // :2001
    }
    for (size_t a = 0; a < n_asv; ++a) depths[a] = (uint64_t)std::llround(ab[a] * total_assigned);   // f64::round, :2015
    return iteration;
}

void write_consensus_fasta(const std::vector<ConsensusSequence> &cons, const std::string &path, const std::string &prefix) {
    std::ofstream f(path);
    if (!f) throw std::runtime_error("cannot create " + path);
    for (size_t i = 0; i < cons.size(); ++i) {
        const ConsensusSequence &c = cons[i];
        std::string depth_field;
        if (c.per_sample_depths.empty()) depth_field = std::to_string(c.depth + c.appended_depth);
        else for (size_t s = 0; s < c.per_sample_depths.size(); ++s) depth_field += (s ? "-" : "") + std::to_string(c.per_sample_depths[s]);
        // src/alignment.rs:848-850
        f << ">" << prefix << "_consensus_" << i << "_depth_" << depth_field << " debug_id:" << c.id << " chimera_score:" << c.chimera_score
          << " unambiguous_read_assignments:" << c.unambig_best_read_map_count << " ambig_read_assignments:" << c.ambig_read_map_count
          << " num_align_leq_10_mismatches:" << c.num_map_leq_10nm << "\n";
        size_t b = 0, e = c.sequence.size();                                // trim leading / trailing N (:842-843)
        while (b < e && c.sequence[b] == 'N') ++b;
        while (e > b && c.sequence[e - 1] == 'N') --e;
        f.write(reinterpret_cast<const char *>(c.sequence.data()) + b, (std::streamsize)(e - b));
        f << "\n";
    }
}

void write_feature_table(const std::vector<ConsensusSequence> &cons, const std::string &path, const std::vector<std::string> &sample_names) {
    std::ofstream f(path);
    if (!f) throw std::runtime_error("cannot create " + path);
    f << "#OTU ID";
    for (size_t s = 0; s < sample_names.size(); ++s) f << "\t" << sample_names[s];   // "#OTU ID\t" + join("\t"), src/main.rs:388
    if (sample_names.empty()) f << "\t";
    f << "\n";
    for (size_t i = 0; i < cons.size(); ++i) {
        const ConsensusSequence &c = cons[i];
        if (c.per_sample_depths.empty()) {
            const size_t d = c.depth + c.appended_depth;
            f << "final_consensus_" << i << "_depth_" << d << "\t" << d << "\n";
        } else {
            std::string dash, tab;
            for (size_t s = 0; s < c.per_sample_depths.size(); ++s) {
                dash += (s ? "-" : "") + std::to_string(c.per_sample_depths[s]);
                tab += (s ? "\t" : "") + std::to_string(c.per_sample_depths[s]);
            }
            f << "final_consensus_" << i << "_depth_" << dash << "\t" << tab << "\n";
        }
    }
}

void write_read_to_asv_mappings(const std::string &path, const std::vector<std::string> &read_ids, const std::vector<size_t> &consensus_ids,
                                const uint64_t *row_offsets, const uint32_t *row_asv, const int32_t *row_nm, const uint32_t *row_snp_mismatches) {
    std::ofstream f(path);
    if (!f) throw std::runtime_error("cannot create " + path);
    for (size_t r = 0; r < read_ids.size(); ++r)
        for (uint64_t i = row_offsets[r]; i < row_offsets[r + 1]; ++i)
            f << read_ids[r] << "\tdebug_id:" << consensus_ids.at(row_asv[i]) << "\t" << row_snp_mismatches[i] << "\t" << row_nm[i] << "\n";
}

template <typename Reads>
static void refine_impl(const Reads &twin_reads, std::vector<ConsensusSequence> &consensuses,
                        const KmerGlobalInfo &kmer_info, const Cli &args, const std::string &temp_dir, int device) {
    if (consensuses.empty()) return;                                        // :1734-1737
    if (!temp_dir.empty()) write_consensus_fasta(consensuses, temp_dir + "/final_asvs_for_em.fasta", "em_refinement");   // :1747-1749
    Mapper mapper(device, args, kmer_info);
    mapper.load_asvs(consensuses);
    const Assignments a = mapper.map_reads(twin_reads, 1);
    for (size_t i = 0; i < consensuses.size(); ++i) {                       // :1941-1945
        consensuses[i].unambig_best_read_map_count = (size_t)a.unambig[i];
        consensuses[i].ambig_read_map_count = (size_t)a.ambig[i];
        consensuses[i].num_map_leq_10nm = (size_t)a.leq10[i];
    }
    const EqClasses cls = build_eq_classes(a, nullptr, -1);
    if (cls.size() == 0) return;                                            // "Keeping original depths", :1953-1956
    std::vector<double> ab; std::vector<uint64_t> depths;
    run_em(consensuses.size(), cls, ab, depths);
    for (size_t i = 0; i < consensuses.size(); ++i) consensuses[i].depth = (size_t)depths[i];   // :2013-2027
    consensuses.erase(std::remove_if(consensuses.begin(), consensuses.end(), [](const ConsensusSequence &c) { return c.depth == 0; }),
                      consensuses.end());                                  // :2031
}

void refine_asv_depths_with_em(const ReadSet &twin_reads, std::vector<ConsensusSequence> &consensuses,
                               const KmerGlobalInfo &kmer_info, const Cli &args, const std::string &temp_dir, int device) {
    refine_impl(twin_reads, consensuses, kmer_info, args, temp_dir, device);
}
void refine_asv_depths_with_em(const TwinReadSet &twin_reads, std::vector<ConsensusSequence> &consensuses,
                               const KmerGlobalInfo &kmer_info, const Cli &args, const std::string &temp_dir, int device) {
    refine_impl(twin_reads, consensuses, kmer_info, args, temp_dir, device);
}

template <typename Reads>
static std::vector<std::vector<size_t>> per_sample_impl(const Reads &twin_reads, size_t n_samples,
                                                        const std::vector<ConsensusSequence> &consensuses,
                                                        const KmerGlobalInfo &kmer_info, const Cli &args, int device) {
    const size_t n_asvs = consensuses.size();
    std::vector<std::vector<size_t>> result(n_asvs, std::vector<size_t>(n_samples, 0));
    if (n_asvs == 0 || n_samples == 0) return result;                       // :2055-2057
    Mapper mapper(device, args, kmer_info);
    mapper.load_asvs(consensuses);
    const Assignments a = mapper.map_reads(twin_reads, (int)n_samples);    // one pass for all samples
    for (size_t s = 0; s < n_samples; ++s) {                                // :2080-2211
        const EqClasses cls = build_eq_classes(a, twin_reads.file_idx, (int)s);
        if (cls.size() == 0 || cls.total_assigned == 0) continue;
        std::vector<double> ab; std::vector<uint64_t> depths;
        run_em(n_asvs, cls, ab, depths);
        for (size_t i = 0; i < n_asvs; ++i) result[i][s] = (size_t)depths[i];
    }
    return result;
}
std::vector<std::vector<size_t>> compute_per_sample_depths(const ReadSet &twin_reads, size_t n_samples,
                                                           const std::vector<ConsensusSequence> &consensuses,
                                                           const KmerGlobalInfo &kmer_info, const Cli &args, int device) {
    return per_sample_impl(twin_reads, n_samples, consensuses, kmer_info, args, device);
}
std::vector<std::vector<size_t>> compute_per_sample_depths(const TwinReadSet &twin_reads, size_t n_samples,
                                                           const std::vector<ConsensusSequence> &consensuses,
                                                           const KmerGlobalInfo &kmer_info, const Cli &args, int device) {
    return per_sample_impl(twin_reads, n_samples, consensuses, kmer_info, args, device);
}


// ------------------------------------------------------------------------------------------------
// the neighbouring stages: cluster lists, reassignment, pileups
// ------------------------------------------------------------------------------------------------
std::vector<std::vector<size_t>> clusters_from_assignment(const uint32_t *rep_of, size_t n, size_t min_cluster_size) {
    std::vector<std::vector<size_t>> by_rep(n);
    for (size_t i = 0; i < n; ++i) {
        if (rep_of[i] >= n) die("clusters_from_assignment: representative out of range");
        by_rep[rep_of[i]].push_back(i);                       // ascending read ids
    }
    std::vector<std::vector<size_t>> out;
    for (auto &c : by_rep) if (!c.empty()) out.push_back(std::move(c));
    std::stable_sort(out.begin(), out.end(), [](const std::vector<size_t> &a, const std::vector<size_t> &b) {
        return a.size() != b.size() ? a.size() > b.size() : a.front() < b.front();
    });
    out.erase(std::remove_if(out.begin(), out.end(), [&](const std::vector<size_t> &c) { return c.size() < min_cluster_size; }), out.end());
    return out;
}

std::vector<std::vector<size_t>> cluster_reads_by_kmers(Mapper &mapper, const uint64_t *kmers, const uint64_t *offsets, size_t n,
                                                        double threshold, size_t min_cluster_size) {
    std::vector<uint32_t> rep(n ? n : 1);
    uint64_t n_reps = 0;
    ck(savont_b200_cluster_reads_by_kmers(mapper.raw(), kmers, offsets, n, threshold, rep.data(), &n_reps, nullptr), "cluster_reads_by_kmers");
    return clusters_from_assignment(rep.data(), n, min_cluster_size);
}

std::vector<std::vector<std::vector<size_t>>> snpmer_clusters_from_assignment(const std::vector<std::vector<size_t>> &kmer_clusters,
                                                                              const uint32_t *rep_of, size_t min_cluster_size) {
    std::vector<std::vector<std::vector<size_t>>> out(kmer_clusters.size());
    for (size_t kc = 0; kc < kmer_clusters.size(); ++kc) {
        std::vector<std::pair<uint32_t, size_t>> pairs;                 // (representative, read)
        pairs.reserve(kmer_clusters[kc].size());
        for (size_t read : kmer_clusters[kc]) pairs.push_back({rep_of[read], read});
        std::sort(pairs.begin(), pairs.end());
        std::vector<std::vector<size_t>> local;
        for (size_t i = 0; i < pairs.size();) {
            size_t j = i;
            std::vector<size_t> c;
            while (j < pairs.size() && pairs[j].first == pairs[i].first) c.push_back(pairs[j++].second);   // ascending (:681)
            local.push_back(std::move(c));
            i = j;
        }
        std::stable_sort(local.begin(), local.end(), [](const std::vector<size_t> &a, const std::vector<size_t> &b) {
            return a.size() != b.size() ? a.size() > b.size() : a.front() < b.front();                       // :683-684
        });
        local.erase(std::remove_if(local.begin(), local.end(), [&](const std::vector<size_t> &c) { return c.size() < min_cluster_size; }),
                    local.end());                                                                            // :688
        out[kc] = std::move(local);
    }
    return out;
}

std::vector<std::vector<std::vector<size_t>>> cluster_reads_by_snpmers(Mapper &mapper, const uint64_t *snp_kmers, const uint64_t *snp_offsets,
                                                                       size_t n_reads, const std::vector<std::vector<size_t>> &kmer_clusters,
                                                                       size_t min_cluster_size, uint32_t iterative_above) {
    std::vector<uint32_t> members; std::vector<uint64_t> offs{0};
    for (const auto &c : kmer_clusters) { for (size_t r : c) members.push_back((uint32_t)r); offs.push_back(members.size()); }
    std::vector<uint32_t> rep(n_reads ? n_reads : 1);
    ck(savont_b200_cluster_reads_by_snpmers(mapper.raw(), snp_kmers, snp_offsets, n_reads, members.data(), offs.data(), kmer_clusters.size(),
                                            iterative_above, rep.data(), nullptr, nullptr), "cluster_reads_by_snpmers");
    return snpmer_clusters_from_assignment(kmer_clusters, rep.data(), min_cluster_size);
}

std::pair<std::vector<std::vector<size_t>>, size_t> clusters_from_reassignment(const std::vector<std::vector<size_t>> &clusters,
                                                                                const uint32_t *best_cons, size_t min_cluster_size) {
    std::vector<std::vector<size_t>> fresh(clusters.size());
    size_t moved = 0, at = 0;
    for (size_t ci = 0; ci < clusters.size(); ++ci)
        for (size_t read : clusters[ci]) {
            const uint32_t b = best_cons[at++];
            if (b >= clusters.size()) die("clusters_from_reassignment: cluster out of range");
            fresh[b].push_back(read);
            if (b != ci) ++moved;
        }
    std::vector<std::vector<size_t>> out;
    for (auto &c : fresh) {
        if (c.empty() || c.size() < min_cluster_size) continue;
        std::sort(c.begin(), c.end());
        out.push_back(std::move(c));
    }
    return {out, moved};
}

std::pair<std::vector<std::vector<size_t>>, size_t> reassign_reads_to_best_cluster(
    Mapper &mapper, const std::vector<std::vector<size_t>> &clusters, const uint64_t *cons_kmers, const uint64_t *cons_offsets,
    const uint64_t *snp_kmers, const uint64_t *snp_offsets, size_t min_cluster_size) {
    std::vector<uint64_t> rk, ro{0};
    for (const auto &c : clusters)
        for (size_t read : c) {
            rk.insert(rk.end(), snp_kmers + snp_offsets[read], snp_kmers + snp_offsets[read + 1]);
            ro.push_back(rk.size());
        }
    const size_t n = ro.size() - 1;
    std::vector<uint32_t> best(n ? n : 1), mm(n ? n : 1), ma(n ? n : 1);
    ck(savont_b200_snpmer_compare(mapper.raw(), cons_kmers, cons_offsets, clusters.size(), rk.data(), ro.data(), n, best.data(), mm.data(),
                                  ma.data(), nullptr, nullptr), "snpmer_compare");
    return clusters_from_reassignment(clusters, best.data(), min_cluster_size);
}

std::vector<std::vector<Pileup>> generate_consensus_pileups(Mapper &mapper, const std::vector<ConsensusSequence> &consensuses,
                                                            const ReadSet &reads, const uint32_t *read_cons, bool use_hpc) {
    std::vector<uint8_t> cb; std::vector<uint64_t> co{0};
    for (const auto &c : consensuses) { cb.insert(cb.end(), c.sequence.begin(), c.sequence.end()); co.push_back(cb.size()); }
    size_t cap = 64;
    for (size_t r = 0; r < reads.n; ++r) {
        if (read_cons[r] >= consensuses.size()) die("generate_consensus_pileups: consensus out of range");
        cap += (size_t)(reads.offsets[r + 1] - reads.offsets[r]) + consensuses[read_cons[r]].sequence.size();
    }
    std::vector<savont_b200_pileup_entry> ent(cap);
    std::vector<uint64_t> po(cb.size() + 1);
    std::vector<uint8_t> hp(cb.size() ? cb.size() : 1), strand(reads.n ? reads.n : 1);
    std::vector<int32_t> nm(reads.n ? reads.n : 1);
    uint64_t n_ent = 0;
    ck(savont_b200_consensus_pileups(mapper.raw(), use_hpc ? 1 : 0, cb.data(), co.data(), consensuses.size(), reads.bases, reads.quals,
                                     reads.offsets, read_cons, reads.n, po.data(), ent.data(), cap, &n_ent, hp.data(), strand.data(), nm.data(),
                                     nullptr), "consensus_pileups");
    std::vector<std::vector<Pileup>> out(consensuses.size());
    for (size_t ci = 0; ci < consensuses.size(); ++ci) {
        const size_t len = consensuses[ci].sequence.size();
        out[ci].resize(len);
        for (size_t p = 0; p < len; ++p) {
            Pileup &col = out[ci][p];
            const size_t g = co[ci] + p;
            col.position = p; col.ref_base = consensuses[ci].sequence[p]; col.ref_hp_length = hp[g];
            col.bases.reserve(po[g + 1] - po[g]);
            for (uint64_t e = po[g]; e < po[g + 1]; ++e) {
                PileupBase b{};
                b.kind = ent[e].kind; b.n = ent[e].n;
                for (int j = 0; j < 2; ++j) { b.base[j] = ent[e].base[j]; b.qual[j] = ent[e].qual[j]; b.hp[j] = ent[e].hp[j]; }
                col.bases.push_back(b);
            }
        }
    }
    return out;
}

}  // namespace savont_host

// ---- flat C wrappers so that the Python tests / host layer can drive this library with ctypes ----
extern "C" {

static thread_local std::string g_host_error;
const char *savont_host_last_error(void) { return g_host_error.c_str(); }

// cluster lists from an assignment vector (host half of cluster_reads_by_kmers): CSR out, returns the number of clusters or -1
long long savont_host_clusters_from_assignment(const uint32_t *rep_of, size_t n, size_t min_cluster_size, uint64_t *offsets, uint64_t *members) {
    try {
        const auto cl = savont_host::clusters_from_assignment(rep_of, n, min_cluster_size);
        size_t at = 0; offsets[0] = 0;
        for (size_t c = 0; c < cl.size(); ++c) { for (size_t r : cl[c]) members[at++] = r; offsets[c + 1] = at; }
        return (long long)cl.size();
    } catch (const std::exception &e) { g_host_error = e.what(); return -1; }
}
// host half of cluster_reads_by_snpmers: k-mer clusters in CSR + rep_of in; the SNPmer clusters of all k-mer clusters out, flattened in
// k-mer cluster order (CSR) with kc_of[c] = k-mer cluster of SNPmer cluster c; returns their number or -1
long long savont_host_snpmer_clusters_from_assignment(const uint64_t *kc_offsets, const uint32_t *kc_members, size_t n_kc, const uint32_t *rep_of,
                                                      size_t min_cluster_size, uint64_t *offsets, uint64_t *members, uint32_t *kc_of) {
    try {
        std::vector<std::vector<size_t>> kcs(n_kc);
        for (size_t c = 0; c < n_kc; ++c) kcs[c].assign(kc_members + kc_offsets[c], kc_members + kc_offsets[c + 1]);
        const auto res = savont_host::snpmer_clusters_from_assignment(kcs, rep_of, min_cluster_size);
        size_t at = 0, nc = 0; offsets[0] = 0;
        for (size_t kc = 0; kc < res.size(); ++kc)
            for (const auto &c : res[kc]) { for (size_t r : c) members[at++] = r; kc_of[nc] = (uint32_t)kc; offsets[++nc] = at; }
        return (long long)nc;
    } catch (const std::exception &e) { g_host_error = e.what(); return -1; }
}
// temp/read_to_asv_mappings.tsv: read ids joined by '\n'; returns 0 or -1
int savont_host_write_read_to_asv_mappings(const char *path, const char *read_ids_nl, size_t n_reads, const uint64_t *consensus_ids, size_t n_cons,
                                           const uint64_t *row_offsets, const uint32_t *row_asv, const int32_t *row_nm, const uint32_t *row_mm) {
    try {
        std::vector<std::string> ids; ids.reserve(n_reads);
        const char *p = read_ids_nl;
        for (size_t r = 0; r < n_reads; ++r) { const char *e = strchr(p, '\n'); if (!e) e = p + strlen(p); ids.emplace_back(p, e); p = *e ? e + 1 : e; }
        std::vector<size_t> cons(consensus_ids, consensus_ids + n_cons);
        savont_host::write_read_to_asv_mappings(path, ids, cons, row_offsets, row_asv, row_nm, row_mm);
        return 0;
    } catch (const std::exception &e) { g_host_error = e.what(); return -1; }
}
// host half of reassign_reads_to_best_cluster: clusters in CSR in, CSR out; returns the number of new clusters or -1; *moved = reassignments
long long savont_host_clusters_from_reassignment(const uint64_t *cl_offsets, const uint64_t *cl_members, size_t n_clusters, const uint32_t *best_cons,
                                                 size_t min_cluster_size, uint64_t *offsets, uint64_t *members, uint64_t *moved) {
    try {
        std::vector<std::vector<size_t>> cl(n_clusters);
        for (size_t c = 0; c < n_clusters; ++c) cl[c].assign(cl_members + cl_offsets[c], cl_members + cl_offsets[c + 1]);
        const auto res = savont_host::clusters_from_reassignment(cl, best_cons, min_cluster_size);
        size_t at = 0; offsets[0] = 0;
        for (size_t c = 0; c < res.first.size(); ++c) { for (size_t r : res.first[c]) members[at++] = r; offsets[c + 1] = at; }
        *moved = res.second;
        return (long long)res.first.size();
    } catch (const std::exception &e) { g_host_error = e.what(); return -1; }
}

// high-frequency rule on (k-mer, [c0, c1]) entries: returns the threshold (or -1), *n_hf entries written to hf_out
long long savont_host_high_freq_kmers(const uint64_t *kmers, const uint32_t *counts, size_t n, uint64_t *hf_out, size_t *n_hf) {
    try {
        savont_host::SplitKmerCounts c;
        c.kmers.assign(kmers, kmers + n); c.counts.assign(counts, counts + 2 * n);
        std::vector<uint64_t> hf;
        const uint32_t th = savont_host::high_freq_kmers(c, hf);
        for (size_t i = 0; i < hf.size(); ++i) hf_out[i] = hf[i];
        *n_hf = hf.size();
        return (long long)th;
    } catch (const std::exception &e) { g_host_error = e.what(); return -1; }
}

// Stage-7 read order of the kept reads (status == 0); returns the number of kept reads, or -1 on error
long long savont_host_twin_read_order(const uint8_t *status, const double *est_id, size_t n, const uint8_t *ids,
                                      const uint64_t *id_offsets, uint32_t *order) {
    try {
        std::vector<uint8_t> st(status, status + n); std::vector<double> est(est_id, est_id + n);
        const std::vector<uint32_t> o = savont_host::twin_read_order(st, est, ids, id_offsets);
        for (size_t i = 0; i < o.size(); ++i) order[i] = o[i];
        return (long long)o.size();
    } catch (const std::exception &e) { g_host_error = e.what(); return -1; }
}

static savont_host::Cli cli_from(const savont_b200_params *p) {
    savont_host::Cli args;
    args.kmer_size = p->k; args.c = p->c; args.blockmer_length = p->l; args.minimum_base_quality = p->min_bq;
    args.band = p->band; args.strand_mode = p->strand_mode; args.low_polymorphism = p->candidate_mode == 1;
    return args;
}
static std::vector<savont_host::ConsensusSequence> cons_from(const uint8_t *asv_bases, const uint64_t *asv_offsets, size_t n_asv) {
    std::vector<savont_host::ConsensusSequence> cons(n_asv);
    for (size_t a = 0; a < n_asv; ++a) { cons[a].sequence.assign(asv_bases + asv_offsets[a], asv_bases + asv_offsets[a + 1]); cons[a].id = a; cons[a].depth = 1; }
    return cons;
}

// Stage 7 / 7b with the reads as TwinRead fields (savont_host::TwinReadSet); same outputs as savont_host_refine /
// savont_host_per_sample
long long savont_host_refine_twin(int device, const savont_b200_params *p, const uint64_t *snpmers, size_t n_snpmers,
                                  const uint8_t *asv_bases, const uint64_t *asv_offsets, size_t n_asv,
                                  const uint64_t *dna_words, const uint64_t *dna_word_offsets, const uint32_t *lengths,
                                  const void *snpmer_kmers, int snpmer_kmer_bytes, const uint64_t *snpmer_offsets, size_t n_reads,
                                  const char *temp_dir, uint64_t *depth, uint64_t *unambig, uint64_t *ambig, uint64_t *leq10, uint64_t *orig_idx) {
    try {
        savont_host::KmerGlobalInfo ki; ki.snpmers.assign(snpmers, snpmers + n_snpmers);
        auto cons = cons_from(asv_bases, asv_offsets, n_asv);
        savont_host::TwinReadSet rs;
        rs.dna_words = dna_words; rs.dna_word_offsets = dna_word_offsets; rs.lengths = lengths; rs.snpmer_kmers = snpmer_kmers;
        rs.snpmer_kmer_bytes = snpmer_kmer_bytes; rs.snpmer_offsets = snpmer_offsets; rs.n = n_reads;
        savont_host::refine_asv_depths_with_em(rs, cons, ki, cli_from(p), temp_dir ? temp_dir : "", device);
        for (size_t i = 0; i < cons.size(); ++i) {
            depth[i] = cons[i].depth; unambig[i] = cons[i].unambig_best_read_map_count; ambig[i] = cons[i].ambig_read_map_count;
            leq10[i] = cons[i].num_map_leq_10nm; orig_idx[i] = cons[i].id;
        }
        return (long long)cons.size();
    } catch (const std::exception &e) { g_host_error = e.what(); return -1; }
}
int savont_host_per_sample_twin(int device, const savont_b200_params *p, const uint64_t *snpmers, size_t n_snpmers,
                                const uint8_t *asv_bases, const uint64_t *asv_offsets, size_t n_asv,
                                const uint64_t *dna_words, const uint64_t *dna_word_offsets, const uint32_t *lengths,
                                const void *snpmer_kmers, int snpmer_kmer_bytes, const uint64_t *snpmer_offsets,
                                const uint32_t *file_idx, size_t n_reads, size_t n_samples, uint64_t *out) {
    try {
        savont_host::KmerGlobalInfo ki; ki.snpmers.assign(snpmers, snpmers + n_snpmers);
        const auto cons = cons_from(asv_bases, asv_offsets, n_asv);
        savont_host::TwinReadSet rs;
        rs.dna_words = dna_words; rs.dna_word_offsets = dna_word_offsets; rs.lengths = lengths; rs.snpmer_kmers = snpmer_kmers;
        rs.snpmer_kmer_bytes = snpmer_kmer_bytes; rs.snpmer_offsets = snpmer_offsets; rs.file_idx = file_idx; rs.n = n_reads;
        const auto res = savont_host::compute_per_sample_depths(rs, n_samples, cons, ki, cli_from(p), device);
        for (size_t a = 0; a < n_asv; ++a) for (size_t s_ = 0; s_ < n_samples; ++s_) out[a * n_samples + s_] = res[a][s_];
        return 0;
    } catch (const std::exception &e) { g_host_error = e.what(); return -1; }
}

// EM over CSR classes; returns iterations, or -1 on error
int savont_host_em(size_t n_asv, const uint64_t *cls_offsets, const uint32_t *cls_asv, const uint64_t *counts, size_t n_cls,
                   uint64_t total_assigned, double *abundances, uint64_t *depths) {
    try {
        savont_host::EqClasses c;
        c.offsets.assign(cls_offsets, cls_offsets + n_cls + 1);
        c.asv.assign(cls_asv, cls_asv + cls_offsets[n_cls]);
        c.counts.assign(counts, counts + n_cls);
        c.total_assigned = total_assigned;
        std::vector<double> ab; std::vector<uint64_t> dp;
        const int it = savont_host::run_em(n_asv, c, ab, dp);
        for (size_t a = 0; a < n_asv; ++a) { if (abundances) abundances[a] = ab[a]; if (depths) depths[a] = dp[a]; }
        return it;
    } catch (const std::exception &e) { g_host_error = e.what(); return -1; }
}

// Stage 7 on flat arrays: ASVs (bases/offsets) with initial depths; returns the number of ASVs kept and
// writes depth / counters / original index of every kept ASV.  Returns -1 on error.
long long savont_host_refine(int device, const savont_b200_params *p, const uint64_t *snpmers, size_t n_snpmers,
                             const uint8_t *asv_bases, const uint64_t *asv_offsets, size_t n_asv,
                             const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets, size_t n_reads,
                             const char *temp_dir, uint64_t *depth, uint64_t *unambig, uint64_t *ambig, uint64_t *leq10, uint64_t *orig_idx) {
    try {
        savont_host::Cli args;
        args.kmer_size = p->k; args.c = p->c; args.blockmer_length = p->l; args.minimum_base_quality = p->min_bq;
        args.band = p->band; args.strand_mode = p->strand_mode; args.low_polymorphism = p->candidate_mode == 1;
        savont_host::KmerGlobalInfo ki; ki.snpmers.assign(snpmers, snpmers + n_snpmers);
        std::vector<savont_host::ConsensusSequence> cons(n_asv);
        for (size_t a = 0; a < n_asv; ++a) { cons[a].sequence.assign(asv_bases + asv_offsets[a], asv_bases + asv_offsets[a + 1]); cons[a].id = a; cons[a].depth = 1; }
        savont_host::ReadSet rs; rs.bases = bases; rs.quals = quals; rs.offsets = offsets; rs.n = n_reads;
        savont_host::refine_asv_depths_with_em(rs, cons, ki, args, temp_dir ? temp_dir : "", device);
        for (size_t i = 0; i < cons.size(); ++i) {
            depth[i] = cons[i].depth; unambig[i] = cons[i].unambig_best_read_map_count; ambig[i] = cons[i].ambig_read_map_count;
            leq10[i] = cons[i].num_map_leq_10nm; orig_idx[i] = cons[i].id;
        }
        return (long long)cons.size();
    } catch (const std::exception &e) { g_host_error = e.what(); return -1; }
}

// Stage 7b on flat arrays: out[asv * n_samples + sample]
int savont_host_per_sample(int device, const savont_b200_params *p, const uint64_t *snpmers, size_t n_snpmers,
                           const uint8_t *asv_bases, const uint64_t *asv_offsets, size_t n_asv,
                           const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets, const uint32_t *file_idx, size_t n_reads,
                           size_t n_samples, uint64_t *out) {
    try {
        savont_host::Cli args;
        args.kmer_size = p->k; args.c = p->c; args.blockmer_length = p->l; args.minimum_base_quality = p->min_bq;
        args.band = p->band; args.strand_mode = p->strand_mode; args.low_polymorphism = p->candidate_mode == 1;
        savont_host::KmerGlobalInfo ki; ki.snpmers.assign(snpmers, snpmers + n_snpmers);
        std::vector<savont_host::ConsensusSequence> cons(n_asv);
        for (size_t a = 0; a < n_asv; ++a) { cons[a].sequence.assign(asv_bases + asv_offsets[a], asv_bases + asv_offsets[a + 1]); cons[a].id = a; }
        savont_host::ReadSet rs; rs.bases = bases; rs.quals = quals; rs.offsets = offsets; rs.file_idx = file_idx; rs.n = n_reads;
        const auto res = savont_host::compute_per_sample_depths(rs, n_samples, cons, ki, args, device);
        for (size_t a = 0; a < n_asv; ++a) for (size_t s = 0; s < n_samples; ++s) out[a * n_samples + s] = res[a][s];
        return 0;
    } catch (const std::exception &e) { g_host_error = e.what(); return -1; }
}

// writers on flat arrays (format checks in tests/test_host_api.py)
int savont_host_write_outputs(const uint8_t *asv_bases, const uint64_t *asv_offsets, size_t n_asv, const uint64_t *depth,
                              const uint64_t *ids, const uint64_t *unambig, const uint64_t *ambig, const uint64_t *leq10,
                              const uint64_t *per_sample, size_t n_samples, const char *fasta_path, const char *prefix,
                              const char *table_path, const char *sample_names_tab) {
    try {
        std::vector<savont_host::ConsensusSequence> cons(n_asv);
        for (size_t a = 0; a < n_asv; ++a) {
            cons[a].sequence.assign(asv_bases + asv_offsets[a], asv_bases + asv_offsets[a + 1]);
            cons[a].depth = depth[a]; cons[a].id = ids[a];
            cons[a].unambig_best_read_map_count = unambig[a]; cons[a].ambig_read_map_count = ambig[a]; cons[a].num_map_leq_10nm = leq10[a];
            if (per_sample) for (size_t s = 0; s < n_samples; ++s) cons[a].per_sample_depths.push_back(per_sample[a * n_samples + s]);
        }
        if (fasta_path) savont_host::write_consensus_fasta(cons, fasta_path, prefix ? prefix : "final");
        if (table_path) {
            std::vector<std::string> names; std::stringstream ss(sample_names_tab ? sample_names_tab : "sample"); std::string t;
            while (std::getline(ss, t, '\t')) names.push_back(t);
            savont_host::write_feature_table(cons, table_path, names);
        }
        return 0;
    } catch (const std::exception &e) { g_host_error = e.what(); return -1; }
}

}  // extern "C"
