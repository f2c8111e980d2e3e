// api_internal.cuh -- types and helpers shared by the host-side translation units of libsavont_b200.so
// (api.cu: the read-vs-ASV path; api_kmer.cu: Stage 1 split-k-mer counting).  Not part of the ABI.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/savont_b200.h"
#include "common.cuh"

namespace sb {


std::string &last_error_ref();          // thread-local message behind savont_b200_last_error()

inline int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(SAVONT_B200_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,            \
                        cudaGetErrorString(e__));                                                             \
    } while (0)
#define TRY(call)                    \
    do {                             \
        int rc__ = (call);           \
        if (rc__ != 0) return rc__;  \
    } while (0)

// A small persistent pool of host threads (the gather of the *_ptrs entry points runs on it once per batch: creating threads per
// batch costs more than the gather of a small batch).  run(n, f) executes f(0..n-1), f(0) on the caller, and returns when all are done.
class HostPool {
public:
    ~HostPool() { stop(); }
    void run(int n, const std::function<void(int)> &f) {
        if (n <= 1) { if (n == 1) f(0); return; }
        while ((int)th_.size() < n - 1) th_.emplace_back([this] { loop(); });
        std::unique_lock<std::mutex> lk(mu_);
        fn_ = &f; n_ = n; next_ = 0; pending_ = n; ++gen_;
        cv_.notify_all();
        while (next_ < n_) {                         // the caller takes indices too
            const int idx = next_++;
            lk.unlock(); f(idx); lk.lock();
            --pending_;
        }
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }
    void stop() {
        { std::unique_lock<std::mutex> lk(mu_); quit_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
        th_.clear(); quit_ = false;
    }
private:
    void loop() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu_);
        while (true) {
            cv_.wait(lk, [&] { return quit_ || gen_ != seen; });
            if (quit_) return;
            seen = gen_;
            while (next_ < n_) {
                const int idx = next_++;
                const std::function<void(int)> *f = fn_;
                lk.unlock(); (*f)(idx); lk.lock();
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_; std::condition_variable cv_, done_;
    const std::function<void(int)> *fn_ = nullptr;
    int n_ = 0, next_ = 0, pending_ = 0; uint64_t gen_ = 0; bool quit_ = false;
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool borrowed = false;        // a view into another buffer (parts of the resident reads): never freed, never grown
    void borrow(void *ptr, size_t bytes) { release(); p = ptr; cap = bytes; borrowed = true; }
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (borrowed) return fail(SAVONT_B200_ERR_INVALID, "internal: a borrowed device buffer of %zu bytes cannot grow to %zu", cap, bytes);
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(SAVONT_B200_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return 0;
    }
    void release() { if (p && !borrowed) cudaFree(p); p = nullptr; cap = 0; borrowed = false; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) { p = nullptr; return fail(SAVONT_B200_ERR_NOMEM, "cudaMallocHost(%zu) failed", want); }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// device-resident inputs, intermediates and outputs of one batch of sequences
struct Batch {
    DevBuf bases, quals, offsets, file_idx;
    DevBuf solid_mini, solid_snp;                        // a4 outputs (savont_b200_twin_reads_from_snpmers)
    DevBuf dna_words, word_offsets, snp_raw, snp_dummy;   // TwinRead input: Seq<Dna> words; Kmer48 staging; scratch counters
    bool has_quals = false, has_file_idx = false;
    bool twin = false;            // sequences are TwinRead.dna_seq words and the SNPmers come from the caller
    bool force_generic = false;   // the fast sketch kernel overflowed one of its shared-memory lists on this batch
    uint32_t n = 0, max_len = 0;
    uint64_t total_bases = 0;
    // sketch
    DevBuf planes, sk_status, flatq, mini_cnt, mini_kmer, mini_pos, snp_start, snp_cnt, snp_kmer, snp_pos, est_id;
    DevBuf spill_k, spill_p; uint32_t spill_stride = 0;
    uint32_t snp_cap = 0;
    DevBuf ctrs;   // u32[4]: snp pool counter, pair counter, error bits, spare
    // filter
    DevBuf cand_status, pair_start, pair_cnt, low_mm, p_read, p_asv, p_mini, p_strand;
    uint32_t pair_cap = 0;
    // align
    DevBuf nm, asv_end, read_end;
    // assign
    DevBuf rstatus, best_nm, class_cnt, class_slot, class_off, class_out, scan_tmp;
    bool mapped = false;
    PinBuf pin_offsets;
    void release() {
        pin_offsets.release();
        DevBuf *all[] = {&solid_mini, &solid_snp, &dna_words, &word_offsets, &snp_raw, &snp_dummy, &bases, &quals, &offsets, &file_idx, &planes, &sk_status, &flatq, &mini_cnt, &mini_kmer, &mini_pos,
                         &snp_start, &snp_cnt, &snp_kmer, &snp_pos, &est_id, &spill_k, &spill_p, &ctrs, &cand_status, &pair_start, &pair_cnt,
                         &low_mm, &p_read, &p_asv, &p_mini, &p_strand, &nm, &asv_end, &read_end, &rstatus, &best_nm,
                         &class_cnt, &class_slot, &class_off, &class_out, &scan_tmp};
        for (DevBuf *b : all) b->release();
    }
};

struct Context {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    savont_b200_params params;
    DeviceParams dp;
    size_t batch_reads = 131072;

    // tables
    DevBuf d_rank_table; bool use_rank_table = false;
    DevBuf d_snp_keys, d_snp_bloom; uint32_t snp_mask = 0, snp_shift = 0, bloom_shift = 0, bloom_words = 0;
    std::vector<uint64_t> h_snpmers;
    DevBuf d_phred, d_thresh; bool thresh_mm0_fails = false;

    // ASVs (replicated index)
    std::vector<uint8_t> asv_bases_h; std::vector<uint64_t> asv_offsets_h;
    size_t n_asv = 0; bool index_dirty = true; uint32_t max_asv_len = 0;
    DevBuf d_idx_keys, d_idx_vals, d_idx_entries, d_mset, d_mset_off, d_asv_planes, d_asv_plane_off, d_asv_len;
    DevBuf d_mx_keys, d_mx_vals, d_mx_entries;
    uint32_t idx_mask = 0, idx_shift = 0, mx_mask = 0, mx_shift = 0;
    std::vector<uint32_t> asv_len_h;

    // counters [unambig A | ambig A | leq10 A | assigned | filtered | singleton S*A]
    DevBuf d_counters; size_t counters_len = 0; int counters_samples = 0;

    // K5 (Stage 1 split-k-mer counting): global table, counters, output staging, partition entries, read chunks
    DevBuf d_kc_table, d_kc_ctr, d_kc_out_k, d_kc_out_c, d_kp_small, d_kp_entries, d_kp_entries1;
    struct KcChunk { DevBuf bases, quals, offsets, rc; PinBuf pin; uint32_t n = 0, max_len = 0; uint64_t total = 0; bool has_quals = false, has_rc = false; };
    std::vector<KcChunk> kc_chunks;
    bool kc_partitioned = true;                          // SAVONT_B200_KC_PARTITIONED=0 forces the single-table algorithm
    uint32_t kc_smem_limit = 0;                          // SAVONT_B200_KC_SMEM_LIMIT: distinct k-mers per shared-memory round (tests)
    cudaEvent_t kc_ev_up[2] = {nullptr, nullptr}, kc_ev_done[2] = {nullptr, nullptr};
    // a4 read filter (twin_reads_from_snpmers): options + high-frequency k-mer set
    bool rf_on = false; ReadFilter rf{}; std::vector<uint64_t> h_hf; DevBuf d_hf_keys; uint32_t hf_mask = 0, hf_shift = 0;
    Batch work, work2, work3, work4, resident, asvb;
    int resident_samples = 1;
    // map_resident pipeline: the resident reads cut into parts (views of the resident inputs with their own intermediates) that
    // run on the two compute lanes, so that the sketch / filter kernels of part i+1 share the SMs with the DP kernel of part i
    std::vector<Batch> res_parts; std::vector<size_t> res_cuts; int resident_parts_req = 0;
    PinBuf pin_small, pin_out;
    // savont_b200_map_twin_read_ptrs: page-locked staging of the gathered reads, one slot per batch in flight
    PinBuf gather_words[8], gather_snp[8], gather_meta[8]; int host_threads = 0; HostPool pool;
    cudaStream_t copy_stream = nullptr;          // H2D of batch i+1 overlaps the kernels of batch i
    cudaStream_t d2h_stream = nullptr;           // D2H of batch i-1's results overlaps the kernels of batch i
    cudaEvent_t ev_d2h[2] = {nullptr, nullptr};
    // map_reads / map_twin_reads run consecutive batches on two internal compute lanes, so the tail of batch i's DP
    // kernel (a partial last wave) overlaps the sketch / filter kernels of batch i+1
    cudaStream_t lane_stream[2] = {nullptr, nullptr};
    cudaEvent_t lane_ev[2][5] = {{nullptr}};
    cudaEvent_t *pev = nullptr;                  // the 5 stage events enqueue_path / finish_path use (ev or lane_ev[i])
    cudaEvent_t ev[8] = {nullptr};
    cudaEvent_t ev_up[4] = {nullptr, nullptr, nullptr, nullptr}, ev_h0[4] = {nullptr, nullptr, nullptr, nullptr};
    savont_b200_stats last_stats;
    // traceback workspace (history, CIGAR slots, compacted ops, ...) and the Stage 4 pileup scratch: kept between calls
    DevBuf tb_ws[7]; PinBuf tb_pin;
    DevBuf bh_ws[8];                  // best_hits_batch: flags / offsets / best pairs (kept between calls: cudaMalloc + cudaFree cost more than the kernels)
    std::vector<uint32_t> bh_off; std::vector<int32_t> bh_rec; std::vector<uint64_t> bh_coff;
    DevBuf pileup_ws[40];
    DevBuf cluster_ws[16]; PinBuf cluster_pin[4];
    PinBuf stage_pin;                 // upload_host: ring of pinned chunks for pageable sources
    // sharded k-mer counting (savont_b200_kmer_shard_*): state between the phases
    struct KmerShard {
        bool active = false; KmerPartPlan pl{}; std::vector<unsigned long long> off_local; uint32_t *cursor1 = nullptr, *d_tile_start = nullptr,
             *tail = nullptr; uint64_t slots = 0; int n_ranks = 0; int count_runs = 0;
    } ks;
    DevBuf d_ks_entries, d_ks_dstptr, d_ks_partoff;
};

struct Context;
inline uint32_t next_pow2_u32(uint64_t x) { uint32_t p = 2; while ((uint64_t)p < x) p <<= 1; return p; }
inline uint32_t shift_for_cap(uint32_t cap) { uint32_t b = 0; while ((1u << b) < cap) ++b; return 64u - b; }
inline float ev_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
int use_device(Context *c);
// api.cu helpers shared with api_pileup.cu
int upload_host(Context *c, void *dst, const void *src, size_t bytes, cudaStream_t st);
int batch_upload(Context *c, Batch &b, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets, const uint32_t *file_idx,
                 size_t n, uint64_t *h2d_bytes, cudaStream_t st = nullptr);
void build_dp_planes(const DeviceParams &dp, const uint8_t *bases, const uint64_t *offsets, size_t A, std::vector<uint2> &planes,
                     std::vector<uint32_t> &plane_off, std::vector<uint32_t> &lens, uint32_t *max_len);
// device-resident result of a traceback that fitted one sub-batch (valid): compacted ops + offsets, and the five int32 arrays
// nm | asv_end | read_end | asv_start | read_start of n pairs each; they live in Context::tb_ws until the next traceback
struct TbDevice { bool valid = false; const uint32_t *ops = nullptr; const uint32_t *off = nullptr; const int32_t *out = nullptr; uint32_t n = 0; };
int traceback_pairs(Context *c, Batch &b, const AsvIndex &ix, uint32_t max_ref_len, const uint32_t *d_pr, const uint32_t *d_pa,
                    const uint8_t *d_ps, size_t n_pairs, int32_t *nm, int32_t *asv_end, int32_t *read_end, int32_t *asv_start,
                    int32_t *read_start, uint64_t *cigar_offsets, uint32_t *cigar, size_t cigar_cap,
                    std::vector<uint32_t> *cigar_vec = nullptr, TbDevice *dev = nullptr);
SeqBatch seq_view(const Batch &b);

}  // namespace sb

struct savont_b200_ctx { sb::Context c; };
