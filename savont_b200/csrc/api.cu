// api.cu -- host side of libsavont_b200.so: context, replicated ASV index, batch pipeline and the
// extern "C" entry points declared in include/savont_b200.h.
//
// Everything that computes on reads runs in the CUDA kernels (sketch_fast.cu, sketch.cu, prefilter.cu, filter.cu,
// align.cu, assign.cu; Stage 1 k-mer counting: kmercount.cu with its host side in api_kmer.cu).  The host only orchestrates: stages buffers, builds the (small, once per call)
// replicated ASV index from the GPU-produced ASV sketches, and sums counters.  There is no CPU
// fallback: without a device every entry point fails with SAVONT_B200_ERR_CUDA.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>

#include <emmintrin.h>

#include "api_internal.cuh"

namespace sb {

std::string &last_error_ref() { static thread_local std::string e; return e; }

// ------------------------------------------------------------------------------------------------

static int validate_params(const savont_b200_params &p) {
    if (p.k < 13 || p.k > 23 || (p.k % 2) == 0)
        return fail(SAVONT_B200_ERR_INVALID, "k=%d unsupported: must be odd, 13..23 (reference src/types.rs:580-609)", p.k);
    if (p.c < 1 || p.c > p.k) return fail(SAVONT_B200_ERR_INVALID, "c=%d must be in 1..k", p.c);
    const int s = p.k - p.c + 1;
    if (s < 1 || s > 16) return fail(SAVONT_B200_ERR_INVALID, "s=k-c+1=%d must be in 1..16", s);
    if (p.band < 0 || p.band > align_supported_band_max())
        return fail(SAVONT_B200_ERR_INVALID, "band=%d unsupported (0..%d)", p.band, align_supported_band_max());
    if (p.min_bq < 0 || p.min_bq > 93) return fail(SAVONT_B200_ERR_INVALID, "min_bq=%d out of range", p.min_bq);
    if (p.strand_mode != 0 && p.strand_mode != 1) return fail(SAVONT_B200_ERR_INVALID, "strand_mode must be 0 or 1");
    if (p.candidate_mode != 0 && p.candidate_mode != 1) return fail(SAVONT_B200_ERR_INVALID, "candidate_mode must be 0 or 1");
    if (!(p.ratio_thresh >= 0.0) || !(p.min_mini_frac_base > 0.0)) return fail(SAVONT_B200_ERR_INVALID, "bad thresholds");
    return 0;
}

// compiler-rt __powidf2: what Rust's f64::powi lowers to
static double powi_rt(double a, int b) {
    const bool recip = b < 0;
    double r = 1;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1 / r : r;
}

static DeviceParams make_device_params(const savont_b200_params &p) {
    DeviceParams d{};
    d.k = p.k; d.c = p.c; d.s = p.k - p.c + 1;
    d.min_bq = p.min_bq; d.band = p.band; d.strand_mode = p.strand_mode; d.candidate_mode = p.candidate_mode;
    d.asv_padl = p.band + 32;      // window of DP column j starts at padded bit 32 + (j-1), see align.cu
    d.mini_spacing = (p.c - 1) / 2 + 1;
    d.split_mask = ~(3ull << (p.k - 1));
    d.kmer_mask = (p.k == 32) ? ~0ull : ((1ull << (2 * p.k)) - 1ull);
    d.ratio_thresh = p.ratio_thresh;
    d.min_frac = powi_rt(p.min_mini_frac_base, p.k);
    return d;
}

static uint64_t revcomp_kmer_host(uint64_t x, int k) {
    uint64_t out = 0;
    for (int i = 0; i < k; ++i) { out = (out << 2) | (3 - (x & 3)); x >>= 2; }
    return out;
}
static uint8_t enc_host(uint8_t b) {
    switch (b) {
        case 'C': case 'c': case 1: return 1;
        case 'G': case 'g': case 2: return 2;
        case 'T': case 't': case 'U': case 'u': case 3: return 3;
        default: return 0;
    }
}

static int build_tables(Context *c) {
    const int s = c->dp.s;
    c->use_rank_table = s <= 7;
    if (c->use_rank_table) {
        // rank of mm_hash64(canonical s-mer) among all s-mers; indexed by the FORWARD s-mer
        const uint32_t n = 1u << (2 * s);
        std::vector<std::pair<uint64_t, uint32_t>> hs;
        hs.reserve(n);
        std::vector<uint32_t> canon(n);
        for (uint32_t f = 0; f < n; ++f) {
            const uint32_t r = (uint32_t)revcomp_kmer_host(f, s);
            canon[f] = f < r ? f : r;
            if (canon[f] == f) hs.emplace_back(mm_hash64((uint64_t)f), f);
        }
        std::sort(hs.begin(), hs.end());
        std::vector<uint16_t> rank_of(n, 0);
        for (size_t i = 0; i < hs.size(); ++i) rank_of[hs[i].second] = (uint16_t)i;
        std::vector<uint16_t> table(n);
        for (uint32_t f = 0; f < n; ++f) table[f] = rank_of[canon[f]];
        TRY(c->d_rank_table.ensure(n * sizeof(uint16_t)));
        CK(cudaMemcpyAsync(c->d_rank_table.p, table.data(), n * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    // est_id table: 10^(-(b-33)/10) with the host libm (reference src/seeding.rs:812-813)
    std::vector<double> ph(256);
    for (int b = 0; b < 256; ++b) { const double q = (double)(b - 33); ph[b] = pow(10.0, -q / 10.0); }
    TRY(c->d_phred.ensure(256 * sizeof(double)));
    CK(cudaMemcpyAsync(c->d_phred.p, ph.data(), 256 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    // integer thresholds of the filter's two f64 tests, in the reference's operation order (both are monotonic in the count)
    std::vector<uint16_t> th(2 * kThreshTable, 0);
    for (uint32_t mn = 1; mn < kThreshTable; ++mn) {
        uint32_t lo = 0, hi = mn + 1;                       // smallest m in [0, mn + 1] with !(m / mn < min_frac); mn + 1 = none
        while (lo < hi) { const uint32_t mid = (lo + hi) / 2; if ((double)mid / (double)mn < c->dp.min_frac) lo = mid + 1; else hi = mid; }
        th[mn] = (uint16_t)std::min<uint32_t>(lo, 0xffffu);
    }
    th[0] = 0xffffu;
    for (uint32_t m = 1; m < kThreshTable; ++m) {
        uint32_t lo = 0, hi = 0xffffu;                      // largest mm in [0, 65535] with (mm / m) / c <= ratio_thresh; the test holds for mm = 0 iff 0 <= thresh
        if (!((double)0 / (double)m / (double)c->dp.c <= c->dp.ratio_thresh)) { th[kThreshTable + m] = 0; continue; }
        while (lo < hi) { const uint32_t mid = (lo + hi + 1) / 2; if ((double)mid / (double)m / (double)c->dp.c <= c->dp.ratio_thresh) lo = mid; else hi = mid - 1; }
        th[kThreshTable + m] = (uint16_t)lo;
    }
    c->thresh_mm0_fails = !(0.0 <= c->dp.ratio_thresh);
    TRY(c->d_thresh.ensure(th.size() * 2));
    CK(cudaMemcpyAsync(c->d_thresh.p, th.data(), th.size() * 2, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->dp.frac_min = c->d_thresh.as<uint16_t>(); c->dp.mm_max = c->d_thresh.as<uint16_t>() + kThreshTable;
    c->dp.thresh_n = c->thresh_mm0_fails ? 0u : kThreshTable;      // a negative ratio threshold (nothing passes): keep the f64 path
    return 0;
}

static int upload_snp_set(Context *c) {
    const size_t n = c->h_snpmers.size();
    if (n == 0) { c->snp_mask = 0; return 0; }
    const uint32_t cap = next_pow2_u32(4 * n + 2);           // load factor <= 0.25: ~1.2 probes per lookup
    const uint32_t shift = shift_for_cap(cap);
    std::vector<uint64_t> keys(cap, kEmpty64);
    // blocked two-bit Bloom filter over the same Fibonacci product, staged in shared memory by the sketch
    // kernel: >= 16 bits per key, between 4 KiB and 32 KiB
    uint32_t bloom_bits = next_pow2_u32(16 * n);
    if (bloom_bits < (1u << 15)) bloom_bits = 1u << 15;
    if (bloom_bits > (1u << 18)) bloom_bits = 1u << 18;
    const uint32_t bshift = shift_for_cap(bloom_bits / 32) - 32u;     // 32 - log2(words)
    std::vector<uint32_t> bloom(bloom_bits / 32, 0u);
    for (uint64_t key : c->h_snpmers) {
        uint32_t slot = hash_slot(key, shift);
        while (keys[slot] != kEmpty64 && keys[slot] != key) slot = (slot + 1) & (cap - 1);
        keys[slot] = key;
        // strand-symmetric Bloom key: k-mer + its reverse complement (the kernels probe F + Rv without choosing the
        // canonical strand first); blocked: both bits in one word
        const uint32_t h = bloom_hash32_sym((uint32_t)key, (uint32_t)revcomp_kmer_host(key, c->dp.k));
        bloom[h >> bshift] |= (1u << (h & 31)) | (1u << ((h >> 5) & 31));
    }
    TRY(c->d_snp_keys.ensure(cap * sizeof(uint64_t)));
    TRY(c->d_snp_bloom.ensure(bloom.size() * 4));
    CK(cudaMemcpyAsync(c->d_snp_keys.p, keys.data(), cap * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_snp_bloom.p, bloom.data(), bloom.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->snp_mask = cap - 1; c->snp_shift = shift; c->bloom_shift = bshift; c->bloom_words = (uint32_t)bloom.size();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// batch staging
// ------------------------------------------------------------------------------------------------
constexpr size_t kCtrBytes = 32;   // u32 snp pool counter, pair counter, error bits, sketch work counter; u64 dp_cells; u32 filter work counters [2]

// Host memory -> device, ordered on `st` like a copy.  Page-locked sources go straight to the DMA engine; pageable ones (what a caller's
// Vec is) through a ring of pinned chunks that the host pool fills: a cudaMemcpyAsync straight from pageable memory is staged by the
// driver on one thread (~11 GB/s measured here), the pool reaches 31 GB/s.  Must not run concurrently with another user of c->pool.
int upload_host(Context *c, void *dst, const void *src, size_t bytes, cudaStream_t st) {
    constexpr size_t kChunk = 8u << 20;
    constexpr int kRing = 3;
    bool direct = bytes < 2 * kChunk;
    if (!direct) {                                            // page-locked (cudaMallocHost / cudaHostRegister) memory is read by the DMA engine as it is
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, src) == cudaSuccess) direct = at.type != cudaMemoryTypeUnregistered;
        else cudaGetLastError();
    }
    if (direct) { if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st)); return 0; }
    TRY(c->stage_pin.ensure(kRing * kChunk));
    cudaEvent_t ev[kRing];
    for (auto &e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const int T = c->host_threads > 0 ? c->host_threads : (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    int rc = 0;
    size_t i = 0;
    for (size_t off = 0; off < bytes && rc == 0; off += kChunk, ++i) {
        const size_t len = std::min(kChunk, bytes - off);
        char *pin = c->stage_pin.as<char>() + (i % kRing) * kChunk;
        if (i >= (size_t)kRing && cudaEventSynchronize(ev[i % kRing]) != cudaSuccess) { rc = fail(SAVONT_B200_ERR_CUDA, "staged upload: event wait failed"); break; }
        const size_t per = (len + T - 1) / T;
        c->pool.run(T, [&](int t) {
            const size_t a0 = std::min(len, (size_t)t * per), a1 = std::min(len, a0 + per);
            if (a1 > a0) memcpy(pin + a0, (const char *)src + off + a0, a1 - a0);
        });
        if (cudaMemcpyAsync((char *)dst + off, pin, len, cudaMemcpyHostToDevice, st) != cudaSuccess || cudaEventRecord(ev[i % kRing], st) != cudaSuccess)
            rc = fail(SAVONT_B200_ERR_CUDA, "staged upload: copy failed");
    }
    // the ring is reused by the next call: its last copies must have left the pinned chunks
    for (int r = 0; r < kRing && (size_t)r < i; ++r) cudaEventSynchronize(ev[r]);
    for (auto &e : ev) cudaEventDestroy(e);
    return rc;
}


int batch_upload(Context *c, Batch &b, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets,
                 const uint32_t *file_idx, size_t n, uint64_t *h2d_bytes, cudaStream_t st) {
    if (st == nullptr) st = c->stream;
    const uint64_t base0 = offsets[0];
    const uint64_t total = offsets[n] - base0;
    if (total >= (1ull << 32) - 64) return fail(SAVONT_B200_ERR_INVALID, "batch of %llu bases exceeds 4 Gi", (unsigned long long)total);
    TRY(b.pin_offsets.ensure((n + 1) * sizeof(uint64_t)));
    uint64_t *rel = b.pin_offsets.as<uint64_t>();
    uint32_t max_len = 0;
    for (size_t i = 0; i <= n; ++i) {
        if (i > 0 && offsets[i] < offsets[i - 1]) return fail(SAVONT_B200_ERR_INVALID, "offsets not monotonic at %zu", i);
        rel[i] = offsets[i] - base0;
        if (i > 0) { const uint64_t len = offsets[i] - offsets[i - 1]; if (len > max_len) max_len = (uint32_t)std::min<uint64_t>(len, 0xffffffffu); }
    }
    b.n = (uint32_t)n; b.total_bases = total; b.max_len = max_len; b.mapped = false;
    b.has_quals = quals != nullptr; b.has_file_idx = file_idx != nullptr;
    b.twin = false; b.force_generic = false;
    TRY(b.bases.ensure(total + 64));
    TRY(b.offsets.ensure((n + 1) * sizeof(uint64_t)));
    CK(cudaMemcpyAsync(b.offsets.p, rel, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    TRY(upload_host(c, b.bases.p, bases + base0, total, st));
    uint64_t bytes = (n + 1) * sizeof(uint64_t) + total;
    if (quals) {
        TRY(b.quals.ensure(total + 64));
        TRY(upload_host(c, b.quals.p, quals + base0, total, st));
        bytes += total;
    }
    if (file_idx) {
        TRY(b.file_idx.ensure(n * sizeof(uint32_t) + 4));
        if (n) CK(cudaMemcpyAsync(b.file_idx.p, file_idx, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        bytes += n * sizeof(uint32_t);
    }
    if (h2d_bytes) *h2d_bytes += bytes;
    return 0;
}

// TwinRead batch (savont_b200_map_twin_reads): Seq<Dna> words + the caller's SNPmer k-mers.  The per-read SNPmer
// chunks (snp_start / snp_cnt / snp_kmer) are filled here instead of by the sketch kernel.
struct TwinInput {
    const uint64_t *dna_words, *word_offsets;
    const uint32_t *lengths;
    const void *snp_kmers; int kmer_bytes;
    const uint64_t *snp_offsets;
    const uint32_t *file_idx;
};
// Host half of a TwinRead batch: the per-read offset arrays relative to the batch (page-locked, `meta`), sizes and checks.
struct TwinPrep {
    uint64_t *rel = nullptr, *wrel = nullptr; uint32_t *sstart = nullptr, *scnt = nullptr;
    uint64_t total = 0, n_words = 0, n_snp = 0, w0 = 0, s0 = 0; uint32_t max_len = 0; size_t n = 0;
};
static int prepare_twin(const TwinInput &in, size_t r0, size_t n, PinBuf &meta, TwinPrep *p) {
    TRY(meta.ensure((n + 1) * 16 + n * 8 + 64));
    uint64_t *rel = meta.as<uint64_t>();
    uint64_t *wrel = rel + (n + 1);
    uint32_t *sstart = reinterpret_cast<uint32_t *>(wrel + (n + 1));
    uint32_t *scnt = sstart + n;
    const uint64_t w0 = in.word_offsets[r0], s0 = in.snp_offsets[r0];
    uint64_t total = 0; uint32_t max_len = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t len = in.lengths[r0 + i];
        const uint64_t wa = in.word_offsets[r0 + i], wb = in.word_offsets[r0 + i + 1];
        const uint64_t sa = in.snp_offsets[r0 + i], sb_ = in.snp_offsets[r0 + i + 1];
        if (wb < wa || wb - wa < ((uint64_t)len + 31u) / 32u)
            return fail(SAVONT_B200_ERR_INVALID, "read %zu: %llu words cannot hold %u bases", r0 + i, (unsigned long long)(wb - wa), len);
        if (sb_ < sa || sb_ - s0 >= (1ull << 32)) return fail(SAVONT_B200_ERR_INVALID, "snpmer_offsets not monotonic / too large at read %zu", r0 + i);
        rel[i] = total; wrel[i] = wa - w0;
        sstart[i] = (uint32_t)(sa - s0); scnt[i] = (uint32_t)(sb_ - sa);
        total += len;
        if (len > max_len) max_len = len;
    }
    rel[n] = total; wrel[n] = in.word_offsets[r0 + n] - w0;
    if (total >= (1ull << 32) - 64) return fail(SAVONT_B200_ERR_INVALID, "batch of %llu bases exceeds 4 Gi", (unsigned long long)total);
    p->rel = rel; p->wrel = wrel; p->sstart = sstart; p->scnt = scnt;
    p->total = total; p->n_words = wrel[n]; p->n_snp = in.snp_offsets[r0 + n] - s0; p->w0 = w0; p->s0 = s0; p->max_len = max_len; p->n = n;
    return 0;
}
// Device half: sizes the batch's buffers and enqueues the copies on `st`
static int issue_twin(Context *c, Batch &b, const TwinInput &in, size_t r0, const TwinPrep &p, uint64_t *h2d_bytes, cudaStream_t st) {
    const size_t n = p.n;
    b.n = (uint32_t)n; b.total_bases = p.total; b.max_len = p.max_len; b.mapped = false;
    b.has_quals = false; b.has_file_idx = in.file_idx != nullptr;
    b.twin = true; b.force_generic = false;
    TRY(b.offsets.ensure((n + 1) * 8)); TRY(b.word_offsets.ensure((n + 1) * 8)); TRY(b.dna_words.ensure((p.n_words + 2) * 8));
    TRY(b.snp_start.ensure((n + 1) * 4)); TRY(b.snp_cnt.ensure((n + 1) * 4)); TRY(b.snp_kmer.ensure((p.n_snp + 1) * 8));
    CK(cudaMemcpyAsync(b.offsets.p, p.rel, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b.word_offsets.p, p.wrel, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n) CK(cudaMemcpyAsync(b.snp_start.p, p.sstart, n * 4, cudaMemcpyHostToDevice, st));
    if (n) CK(cudaMemcpyAsync(b.snp_cnt.p, p.scnt, n * 4, cudaMemcpyHostToDevice, st));
    TRY(upload_host(c, b.dna_words.p, in.dna_words + p.w0, p.n_words * 8, st));      // pageable sources: through the pinned ring
    uint64_t bytes = (n + 1) * 16 + n * 8 + p.n_words * 8;
    if (p.n_snp) {
        if (in.kmer_bytes == 8) {
            TRY(upload_host(c, b.snp_kmer.p, static_cast<const uint64_t *>(in.snp_kmers) + p.s0, p.n_snp * 8, st));
        } else {
            TRY(b.snp_raw.ensure(p.n_snp * 6 + 16));
            TRY(upload_host(c, b.snp_raw.p, static_cast<const uint8_t *>(in.snp_kmers) + p.s0 * 6, p.n_snp * 6, st));
            launch_widen_kmer48(b.snp_raw.as<uint8_t>(), b.snp_kmer.as<uint64_t>(), p.n_snp, st);
            CK(cudaGetLastError());
            c->last_stats.kernel_launches += 1;
        }
        bytes += p.n_snp * (uint64_t)in.kmer_bytes;
    }
    if (in.file_idx) {
        TRY(b.file_idx.ensure(n * 4 + 4));
        if (n) CK(cudaMemcpyAsync(b.file_idx.p, in.file_idx + r0, n * 4, cudaMemcpyHostToDevice, st));
        bytes += n * 4;
    }
    if (h2d_bytes) *h2d_bytes += bytes;
    return 0;
}
static int batch_upload_twin(Context *c, Batch &b, const TwinInput &in, size_t r0, size_t n, uint64_t *h2d_bytes,
                             cudaStream_t st = nullptr) {
    if (st == nullptr) st = c->stream;
    TwinPrep p;
    TRY(prepare_twin(in, r0, n, b.pin_offsets, &p));
    return issue_twin(c, b, in, r0, p, h2d_bytes, st);
}
static int validate_twin(const TwinInput &in, size_t n) {
    if (!in.word_offsets || !in.lengths || !in.snp_offsets) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    if (in.kmer_bytes != 6 && in.kmer_bytes != 8) return fail(SAVONT_B200_ERR_INVALID, "snpmer_kmer_bytes must be 6 (Kmer48) or 8 (u64)");
    if (n && !in.dna_words && in.word_offsets[n] > in.word_offsets[0]) return fail(SAVONT_B200_ERR_INVALID, "dna_words is null");
    if (n && !in.snp_kmers && in.snp_offsets[n] > in.snp_offsets[0]) return fail(SAVONT_B200_ERR_INVALID, "snpmer_kmers is null");
    return 0;
}

// sample ids index the S x A singleton matrix on the device: reject anything outside [0, n_samples)
static int validate_file_idx(const uint32_t *file_idx, size_t n, int n_samples) {
    if (!file_idx) return 0;
    for (size_t i = 0; i < n; ++i)
        if (file_idx[i] >= (uint32_t)n_samples)
            return fail(SAVONT_B200_ERR_INVALID, "file_idx[%zu] = %u is not below n_samples = %d", i, file_idx[i], n_samples);
    return 0;
}

// Cuts the resident reads (c->resident, uploaded by batch_upload / batch_upload_twin whose host-side offset arrays are still in
// its pin_offsets) into parts: views of the resident input buffers plus part-relative offset arrays.
static int resident_part_count(const Context *c, size_t n) {
    if (c->resident_parts_req > 0) return c->resident_parts_req;
    static const int env = [] { const char *e = getenv("SAVONT_B200_RESIDENT_PARTS"); return e ? atoi(e) : 0; }();
    if (env > 0) return env;
    (void)n;
    return 1;     // measured on B200 (profiles/r02b_parts_sweep.txt): 4-16 parts never beat one pass -- every kernel of the path is issue bound
}
static int build_resident_parts(Context *c) {
    Batch &R = c->resident;
    for (Batch &pb : c->res_parts) pb.release();
    c->res_parts.clear(); c->res_cuts.clear();
    const size_t n = R.n;
    size_t P = (size_t)resident_part_count(c, n);
    const size_t min_reads = c->resident_parts_req > 0 ? 8 : 1024;      // per part; an explicit request may cut small inputs (tests)
    if (P > n / min_reads) P = n / min_reads;
    if (P <= 1) return 0;
    const uint64_t *rel = R.pin_offsets.as<uint64_t>();
    const uint64_t *wrel = rel + (n + 1);                              // twin only
    c->res_parts.resize(P);
    c->res_cuts.assign(P + 1, 0);
    for (size_t q = 0; q <= P; ++q) c->res_cuts[q] = n * q / P;
    std::vector<uint64_t> tmp;
    for (size_t q = 0; q < P; ++q) {
        Batch &pb = c->res_parts[q];
        const size_t r0 = c->res_cuts[q], cnt = c->res_cuts[q + 1] - r0;
        pb.n = (uint32_t)cnt; pb.twin = R.twin; pb.has_quals = R.has_quals; pb.has_file_idx = R.has_file_idx;
        pb.force_generic = false; pb.mapped = false;
        pb.total_bases = rel[r0 + cnt] - rel[r0];
        uint32_t mx = 0;
        tmp.resize((cnt + 1) * (R.twin ? 2 : 1));
        for (size_t i = 0; i <= cnt; ++i) {
            tmp[i] = rel[r0 + i] - rel[r0];
            if (i > 0) { const uint64_t len = tmp[i] - tmp[i - 1]; if (len > mx) mx = (uint32_t)len; }
        }
        pb.max_len = mx;
        TRY(pb.offsets.ensure((cnt + 1) * 8));
        CK(cudaMemcpyAsync(pb.offsets.p, tmp.data(), (cnt + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        if (R.twin) {
            for (size_t i = 0; i <= cnt; ++i) tmp[cnt + 1 + i] = wrel[r0 + i] - wrel[r0];
            TRY(pb.word_offsets.ensure((cnt + 1) * 8));
            CK(cudaMemcpyAsync(pb.word_offsets.p, tmp.data() + cnt + 1, (cnt + 1) * 8, cudaMemcpyHostToDevice, c->stream));
            pb.dna_words.borrow(R.dna_words.as<uint64_t>() + wrel[r0], (size_t)(wrel[r0 + cnt] - wrel[r0] + 2) * 8);
            pb.snp_start.borrow(R.snp_start.as<uint32_t>() + r0, (cnt + 1) * 4);
            pb.snp_cnt.borrow(R.snp_cnt.as<uint32_t>() + r0, (cnt + 1) * 4);
            pb.snp_kmer.borrow(R.snp_kmer.p, R.snp_kmer.cap);
        } else {
            pb.bases.borrow(R.bases.as<uint8_t>() + rel[r0], (size_t)pb.total_bases + 64);
            if (R.has_quals) pb.quals.borrow(R.quals.as<uint8_t>() + rel[r0], (size_t)pb.total_bases + 64);
        }
        if (R.has_file_idx) pb.file_idx.borrow(R.file_idx.as<uint32_t>() + r0, cnt * 4 + 4);
        CK(cudaStreamSynchronize(c->stream));                          // tmp is reused
    }
    return 0;
}

static SnpSet snp_view(const Context *c) {
    SnpSet set{};
    if (c->snp_mask) {
        set.keys = c->d_snp_keys.as<uint64_t>(); set.mask = c->snp_mask; set.shift = c->snp_shift;
        set.bloom = c->d_snp_bloom.as<uint32_t>(); set.bloom_shift = c->bloom_shift; set.bloom_words = c->bloom_words;
    }
    return set;
}

SeqBatch seq_view(const Batch &b) {
    SeqBatch v{};
    v.bases = b.twin ? nullptr : b.bases.as<uint8_t>();
    v.quals = b.has_quals ? b.quals.as<uint8_t>() : nullptr;
    v.offsets = b.offsets.as<uint64_t>();
    v.n = b.n; v.max_len = b.max_len;
    v.dna_words = b.twin ? b.dna_words.as<uint64_t>() : nullptr;
    v.word_offsets = b.twin ? b.word_offsets.as<uint64_t>() : nullptr;
    return v;
}

static SketchOut sketch_view(const Batch &b, bool with_pos) {
    SketchOut o{};
    o.status = b.sk_status.as<uint8_t>(); o.flatq = b.flatq.as<uint8_t>();
    o.planes = b.planes.as<uint2>();
    o.mini_cnt = b.mini_cnt.as<uint32_t>(); o.mini_kmer = b.mini_kmer.as<uint64_t>();
    o.mini_pos = with_pos ? b.mini_pos.as<uint32_t>() : nullptr;
    o.snp_start = b.snp_start.as<uint32_t>(); o.snp_cnt = b.snp_cnt.as<uint32_t>();
    o.snp_kmer = b.snp_kmer.as<uint64_t>(); o.snp_pos = b.snp_pos.as<uint32_t>();
    o.snp_pool_ctr = b.ctrs.as<uint32_t>(); o.snp_pool_cap = b.snp_cap;
    o.spill_kmer = b.spill_k.as<uint64_t>(); o.spill_pos = b.spill_p.as<uint32_t>(); o.spill_stride = b.spill_stride;
    o.err = b.ctrs.as<uint32_t>() + 2;
    o.work_ctr = b.ctrs.as<uint32_t>() + 3;
    return o;
}

static int stage_sketch(Context *c, Batch &b, bool with_pos) {
    const size_t n = b.n;
    const uint64_t total = b.total_bases;
    const size_t plane_words = (size_t)plane_word_off(total, n) + 4;
    const size_t mini_slots = (size_t)mini_slot_off(total, n, (uint32_t)c->dp.mini_spacing) + 4;
    TRY(b.planes.ensure(plane_words * sizeof(uint2)));
    TRY(b.sk_status.ensure(n + 4)); TRY(b.flatq.ensure(n + 4));
    TRY(b.mini_cnt.ensure((n + 1) * 4)); TRY(b.mini_kmer.ensure(mini_slots * 8));
    if (with_pos) TRY(b.mini_pos.ensure(mini_slots * 4));
    TRY(b.snp_start.ensure((n + 1) * 4)); TRY(b.snp_cnt.ensure((n + 1) * 4));
    TRY(b.ctrs.ensure(kCtrBytes));
    const uint16_t *rank = c->use_rank_table ? c->d_rank_table.as<uint16_t>() : nullptr;
    const bool fast = !b.force_generic && sketch_fast_supported(b.max_len, c->dp, c->use_rank_table, snp_view(c));
    if (b.twin) {
        // the caller's SNPmers are already in the pool (batch_upload_twin); only planes + minimizers are computed
        CK(cudaMemsetAsync(b.ctrs.p, 0, kCtrBytes, c->stream));
        SketchOut so = sketch_view(b, with_pos);
        if (fast) {
            launch_sketch_fast(seq_view(b), so, SnpSet{}, rank, c->dp, c->sm_count, c->stream);
            c->last_stats.kernel_launches += 1;
        } else {
            // generic parameters: unpack to one code byte per base and run the generic kernel without a SNPmer set;
            // its (empty) SNPmer bookkeeping goes to scratch arrays
            TRY(b.bases.ensure(total + 64)); TRY(b.snp_dummy.ensure((n + 1) * 8));
            launch_unpack_twin(seq_view(b), b.bases.as<uint8_t>(), c->stream);
            SeqBatch sv = seq_view(b); sv.bases = b.bases.as<uint8_t>(); sv.quals = nullptr; sv.dna_words = nullptr; sv.word_offsets = nullptr;
            so.snp_start = b.snp_dummy.as<uint32_t>(); so.snp_cnt = b.snp_dummy.as<uint32_t>() + (n + 1);
            const uint32_t grid = sketch_grid(b.n, b.max_len, c->use_rank_table, c->dp, SnpSet{}, c->sm_count);
            b.spill_stride = b.max_len + 1;
            const size_t entries = (size_t)grid * kSketchWarps * b.spill_stride;
            TRY(b.spill_k.ensure(entries * 8)); TRY(b.spill_p.ensure(entries * 4));
            so.spill_kmer = b.spill_k.as<uint64_t>(); so.spill_pos = b.spill_p.as<uint32_t>(); so.spill_stride = b.spill_stride;
            launch_sketch(sv, so, SnpSet{}, rank, c->dp, c->sm_count, c->stream);
            c->last_stats.kernel_launches += 2;
        }
        CK(cudaGetLastError());
        return 0;
    }
    if (b.snp_cap == 0 || b.snp_cap < total / 8 + 4096) b.snp_cap = (uint32_t)std::min<uint64_t>(total / 8 + 4096, 0xfffffff0u);
    if (!fast) {   // per-warp spill region for sequences with more than kSnpHitCap SNPmer hits
        const uint32_t grid = sketch_grid(b.n, b.max_len, c->use_rank_table, c->dp, snp_view(c), c->sm_count);
        b.spill_stride = b.max_len + 1;
        const size_t entries = (size_t)grid * kSketchWarps * b.spill_stride;
        TRY(b.spill_k.ensure(entries * 8)); TRY(b.spill_p.ensure(entries * 4));
    }
    TRY(b.snp_kmer.ensure((size_t)b.snp_cap * 8)); TRY(b.snp_pos.ensure((size_t)b.snp_cap * 4));
    CK(cudaMemsetAsync(b.ctrs.p, 0, kCtrBytes, c->stream));
    if (fast) launch_sketch_fast(seq_view(b), sketch_view(b, with_pos), snp_view(c), rank, c->dp, c->sm_count, c->stream);
    else launch_sketch(seq_view(b), sketch_view(b, with_pos), snp_view(c), rank, c->dp, c->sm_count, c->stream);
    CK(cudaGetLastError());
    c->last_stats.kernel_launches += 1;
    return 0;
}

// Reads the batch counters after the stream drained.  Pool overflows are not errors: the kernels downstream of
// an overflow do nothing (they test the error word), the pools are grown to the exact need and *retry is set.
struct BatchCounters { uint32_t snp_used, n_pairs, err, spare; unsigned long long dp_cells; };
static int check_batch(Context *c, Batch &b, BatchCounters *out, bool *retry) {
    BatchCounters h;
    CK(cudaMemcpyAsync(&h, b.ctrs.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *retry = false;
    if (h.err & kErrSnpPool) { b.snp_cap = h.snp_used + 1024; *retry = true; }
    if (h.err & kErrPairPool) { b.pair_cap = h.n_pairs + 1024; *retry = true; }
    if ((h.err & kErrSnpHits) && !b.force_generic) { b.force_generic = true; *retry = true; h.err &= ~kErrSnpHits; }
    if (h.err & ~(kErrSnpPool | kErrPairPool)) return fail(SAVONT_B200_ERR_OVERFLOW, "device error flags 0x%x", h.err);
    *out = h;
    return 0;
}

// sketch + error check, re-run with a larger SNPmer pool when it overflowed
static int sketch_checked(Context *c, Batch &b, bool with_pos) {
    for (int attempt = 0;; ++attempt) {
        TRY(stage_sketch(c, b, with_pos));
        BatchCounters h{}; bool retry = false;
        TRY(check_batch(c, b, &h, &retry));
        if (!retry) return 0;
        if (attempt == 3) return fail(SAVONT_B200_ERR_OVERFLOW, "SNPmer pool sizing did not converge");
    }
}

// a4: the read tests of twin_reads_from_snpmers on the sketches of an ASCII batch (no-op without a read filter; TwinRead
// batches are reads that already passed)
static int stage_prefilter(Context *c, Batch &b, bool want_counts) {
    if (!c->rf_on || b.twin || b.n == 0) return 0;
    const double *est = nullptr;
    if (c->rf.use_cutoff && b.has_quals) {
        TRY(b.est_id.ensure((b.n + 1) * 8));
        launch_est_id(seq_view(b), b.sk_status.as<uint8_t>(), b.flatq.as<uint8_t>(), c->d_phred.as<double>(), b.est_id.as<double>(), c->stream);
        est = b.est_id.as<double>();
        c->last_stats.kernel_launches += 1;
    }
    if (want_counts) { TRY(b.solid_mini.ensure((b.n + 1) * 4)); TRY(b.solid_snp.ensure((b.n + 1) * 4)); }
    HfSet hf{c->hf_mask ? c->d_hf_keys.as<uint64_t>() : nullptr, c->hf_mask, c->hf_shift};
    launch_read_filter(seq_view(b), sketch_view(b, false), hf, est, c->rf, c->dp, want_counts ? b.solid_mini.as<uint32_t>() : nullptr,
                       want_counts ? b.solid_snp.as<uint32_t>() : nullptr, c->sm_count, c->stream);
    CK(cudaGetLastError());
    c->last_stats.kernel_launches += 1;
    return 0;
}

static AsvIndex index_view(const Context *c) {
    AsvIndex ix{};
    ix.idx_keys = c->idx_mask ? c->d_idx_keys.as<uint64_t>() : nullptr;
    ix.idx_vals = c->d_idx_vals.as<uint2>(); ix.idx_mask = c->idx_mask; ix.idx_shift = c->idx_shift; ix.idx_entries = c->d_idx_entries.as<uint32_t>();
    ix.mset = c->d_mset.as<uint64_t>(); ix.mset_off = c->d_mset_off.as<uint32_t>();
    ix.mx_keys = c->mx_mask ? c->d_mx_keys.as<uint64_t>() : nullptr; ix.mx_vals = c->d_mx_vals.as<uint2>();
    ix.mx_mask = c->mx_mask; ix.mx_shift = c->mx_shift; ix.mx_entries = c->d_mx_entries.as<uint32_t>();
    ix.planes = c->d_asv_planes.as<uint2>(); ix.plane_off = c->d_asv_plane_off.as<uint32_t>(); ix.len = c->d_asv_len.as<uint32_t>();
    ix.n_asv = (uint32_t)c->n_asv;
    return ix;
}

static PairPool pool_view(const Batch &b) {
    PairPool p{};
    p.read = b.p_read.as<uint32_t>(); p.asv = b.p_asv.as<uint32_t>(); p.mini = b.p_mini.as<uint32_t>();
    p.strand = b.p_strand.as<uint8_t>(); p.counter = b.ctrs.as<uint32_t>() + 1; p.cap = b.pair_cap;
    return p;
}
static FilterOut filter_view(const Batch &b) {
    FilterOut f{};
    f.cand_status = b.cand_status.as<uint8_t>(); f.pair_start = b.pair_start.as<uint32_t>(); f.pair_cnt = b.pair_cnt.as<uint32_t>();
    f.low_mm = b.low_mm.as<uint32_t>(); f.err = b.ctrs.as<uint32_t>() + 2; f.work_ctr = b.ctrs.as<uint32_t>() + 6;
    return f;
}

static int ensure_pair_pool(Batch &b) {
    TRY(b.p_read.ensure((size_t)b.pair_cap * 4)); TRY(b.p_asv.ensure((size_t)b.pair_cap * 4));
    TRY(b.p_mini.ensure((size_t)b.pair_cap * 4)); TRY(b.p_strand.ensure((size_t)b.pair_cap + 4));
    TRY(b.nm.ensure((size_t)b.pair_cap * 4)); TRY(b.asv_end.ensure((size_t)b.pair_cap * 4)); TRY(b.read_end.ensure((size_t)b.pair_cap * 4));
    TRY(b.class_slot.ensure((size_t)b.pair_cap * 4));
    return 0;
}

// candidate filter (launch only; the pair count stays on the device)
static int stage_filter(Context *c, Batch &b, int strand_mode) {
    const size_t n = b.n;
    if (n > 0 && (uint64_t)b.max_len / (uint64_t)c->dp.mini_spacing + 2 >= 32768u)
        return fail(SAVONT_B200_ERR_OVERFLOW, "read of %u bases is too long for the candidate filter", b.max_len);
    TRY(b.cand_status.ensure(n + 4)); TRY(b.pair_start.ensure((n + 1) * 4)); TRY(b.pair_cnt.ensure((n + 1) * 4));
    TRY(b.low_mm.ensure((n + 1) * 4));
    const uint32_t per = strand_mode == 1 ? 2u : 1u;
    if (b.pair_cap < 2u * per * (uint32_t)n + 1024u) b.pair_cap = 2u * per * (uint32_t)n + 1024u;
    DeviceParams dp = c->dp; dp.strand_mode = strand_mode;
    TRY(ensure_pair_pool(b));
    c->last_stats.kernel_launches += (uint64_t)launch_filter(seq_view(b), sketch_view(b, false), index_view(c), pool_view(b), filter_view(b), dp,
                                                             c->sm_count, c->stream);
    CK(cudaGetLastError());
    return 0;
}

// n_pairs_dev: device-side pair count written by the filter (null: exactly n_pairs_max pairs)
static int stage_align(Context *c, Batch &b, const uint32_t *n_pairs_dev, uint32_t n_pairs_max) {
    AlignOut ao{b.nm.as<int32_t>(), b.asv_end.as<int32_t>(), b.read_end.as<int32_t>()};
    launch_align_nm(seq_view(b), b.planes.as<uint2>(), index_view(c), b.p_read.as<uint32_t>(), b.p_asv.as<uint32_t>(),
                    b.p_strand.as<uint8_t>(), n_pairs_dev, n_pairs_max, n_pairs_dev ? b.ctrs.as<uint32_t>() + 2 : nullptr, ao, c->dp,
                    c->sm_count, c->stream);
    CK(cudaGetLastError());
    if (n_pairs_max) c->last_stats.kernel_launches += 1;
    return 0;
}

static int ensure_counters(Context *c, int n_samples, bool zero) {
    const size_t len = 3 * c->n_asv + 2 + (size_t)n_samples * c->n_asv;
    if (len != c->counters_len || n_samples != c->counters_samples) {
        TRY(c->d_counters.ensure(len * sizeof(uint64_t)));
        c->counters_len = len; c->counters_samples = n_samples;
        zero = true;
    }
    if (zero) CK(cudaMemsetAsync(c->d_counters.p, 0, len * sizeof(uint64_t), c->stream));
    return 0;
}

static int stage_assign(Context *c, Batch &b, int n_samples, int strand_mode) {
    const size_t n = b.n;
    TRY(b.rstatus.ensure(n + 4)); TRY(b.best_nm.ensure((n + 1) * 4)); TRY(b.class_cnt.ensure((n + 1) * 4));
    TRY(b.class_off.ensure((n + 2) * 4)); TRY(b.class_out.ensure((size_t)b.pair_cap * 4 + 4));
    TRY(b.scan_tmp.ensure((n / 2048 + 4) * 4));
    launch_assign(filter_view(b), pool_view(b), b.nm.as<int32_t>(), b.has_file_idx ? b.file_idx.as<uint32_t>() : nullptr, (uint32_t)n,
                  (uint32_t)c->n_asv, n_samples, strand_mode, c->d_asv_len.as<uint32_t>(), (uint32_t)(2 * c->dp.band + 1),
                  reinterpret_cast<unsigned long long *>(b.ctrs.as<uint32_t>() + 4), b.rstatus.as<uint8_t>(), b.best_nm.as<int32_t>(),
                  b.class_cnt.as<uint32_t>(), b.class_slot.as<uint32_t>(), c->d_counters.as<unsigned long long>(), c->stream);
    launch_scan_u32(b.class_cnt.as<uint32_t>(), b.class_off.as<uint32_t>(), (uint32_t)n, b.scan_tmp.as<uint32_t>(), c->stream);
    launch_gather_u32(b.class_slot.as<uint32_t>(), b.pair_start.as<uint32_t>(), b.class_cnt.as<uint32_t>(), b.class_off.as<uint32_t>(),
                      b.class_out.as<uint32_t>(), (uint32_t)n, c->stream);
    CK(cudaGetLastError());
    if (n) c->last_stats.kernel_launches += 5;
    b.mapped = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// ASV index (reference src/kmer_comp.rs:39-66 + src/alignment.rs:1754-1765)
// ------------------------------------------------------------------------------------------------
// DP planes of the sequences the reads are aligned against (ASVs; consensus sequences in the pileup path): per sequence
// padl zero bits, then the bases (bit i + padl of plane word (i + padl) / 32), then >= 32 * (NW + 4) zero bits
void build_dp_planes(const DeviceParams &dp, const uint8_t *bases, const uint64_t *offsets, size_t A, std::vector<uint2> &planes,
                     std::vector<uint32_t> &plane_off, std::vector<uint32_t> &lens, uint32_t *max_len) {
    const int nw = align_words_for_band(dp.band);
    const uint32_t padl = (uint32_t)dp.asv_padl;
    plane_off.assign(A, 0); lens.assign(A, 0);
    size_t words = 0;
    *max_len = 0;
    for (size_t a = 0; a < A; ++a) {
        const uint32_t len = (uint32_t)(offsets[a + 1] - offsets[a]);
        lens[a] = len; plane_off[a] = (uint32_t)words;
        words += (padl + len) / 32 + (size_t)nw + 6;
        if (len > *max_len) *max_len = len;
    }
    planes.assign(words + 8, make_uint2(0, 0));
    for (size_t a = 0; a < A; ++a) {
        const uint8_t *sq = bases + offsets[a];
        uint2 *pw = planes.data() + plane_off[a];
        for (uint32_t i = 0; i < lens[a]; ++i) {
            const uint32_t code = enc_host(sq[i]), bit = i + padl;
            if (code & 1) pw[bit >> 5].x |= 1u << (bit & 31);
            if (code & 2) pw[bit >> 5].y |= 1u << (bit & 31);
        }
    }
}

static int build_asv_index(Context *c) {
    const size_t A = c->asv_offsets_h.empty() ? 0 : c->asv_offsets_h.size() - 1;
    c->n_asv = A;
    c->index_dirty = false;
    c->counters_len = 0;
    if (A == 0) { c->idx_mask = 0; return 0; }
    const int k = c->dp.k;
    Batch &b = c->asvb;
    TRY(batch_upload(c, b, c->asv_bases_h.data(), nullptr, c->asv_offsets_h.data(), nullptr, A, nullptr));
    TRY(sketch_checked(c, b, true));
    // download the sketches
    const size_t mini_slots = (size_t)mini_slot_off(b.total_bases, A, (uint32_t)c->dp.mini_spacing) + 4;
    std::vector<uint8_t> st(A);
    std::vector<uint32_t> mini_cnt(A), snp_start(A), snp_cnt(A);
    std::vector<uint64_t> mini_kmer(mini_slots);
    uint32_t ctr[4];
    CK(cudaMemcpyAsync(ctr, b.ctrs.p, sizeof(ctr), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(st.data(), b.sk_status.p, A, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(mini_cnt.data(), b.mini_cnt.p, A * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(snp_start.data(), b.snp_start.p, A * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(snp_cnt.data(), b.snp_cnt.p, A * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(mini_kmer.data(), b.mini_kmer.p, mini_slots * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::vector<uint32_t> snp_pos(ctr[0] + 1);
    if (ctr[0]) CK(cudaMemcpy(snp_pos.data(), b.snp_pos.p, (size_t)ctr[0] * 4, cudaMemcpyDeviceToHost));
    for (size_t a = 0; a < A; ++a)
        if (st[a] != 0) return fail(SAVONT_B200_ERR_INVALID, "ASV %zu is shorter than k=%d", a, k);

    // (1) distinct minimizer sets, sorted, forward-any flag in bit 63
    std::vector<uint64_t> mset; std::vector<uint32_t> mset_off(A + 1, 0);
    std::vector<uint64_t> tmp;
    for (size_t a = 0; a < A; ++a) {
        const uint64_t mo = mini_slot_off(c->asv_offsets_h[a] - c->asv_offsets_h[0], a, (uint32_t)c->dp.mini_spacing);
        tmp.assign(mini_kmer.begin() + mo, mini_kmer.begin() + mo + mini_cnt[a]);
        for (uint64_t &v : tmp) v = (v << 1) | (v >> 63);            // key in the high bits, flag in bit 0
        std::sort(tmp.begin(), tmp.end());
        size_t first = mset.size();
        for (uint64_t v : tmp) {
            const uint64_t key = v >> 1, f = v & 1;
            if (mset.size() > first && (mset.back() & ~kFwdBit) == key) { if (f) mset.back() |= kFwdBit; }
            else mset.push_back(key | (f ? kFwdBit : 0));
        }
        mset_off[a + 1] = (uint32_t)mset.size();
    }
    // (1b) minimizer inverted index for the low-polymorphism mode: k-mer -> (asv << 1 | forward-any flag)
    {
        struct ME { uint64_t key; uint32_t ent; };
        std::vector<ME> me; me.reserve(mset.size());
        for (size_t a = 0; a < A; ++a)
            for (uint32_t t = mset_off[a]; t < mset_off[a + 1]; ++t)
                me.push_back({mset[t] & ~kFwdBit, (uint32_t)(a << 1) | (uint32_t)(mset[t] >> 63)});
        std::sort(me.begin(), me.end(), [](const ME &x, const ME &y) { return x.key != y.key ? x.key < y.key : x.ent < y.ent; });
        size_t nk = 0;
        for (size_t i = 0; i < me.size(); ++i) if (i == 0 || me[i].key != me[i - 1].key) ++nk;
        const uint32_t mcap = nk ? next_pow2_u32(4 * nk + 2) : 0;
        const uint32_t mshift = mcap ? shift_for_cap(mcap) : 0;
        std::vector<uint64_t> mkeys(mcap ? mcap : 1, kEmpty64);
        std::vector<uint2> mvals(mcap ? mcap : 1, make_uint2(0, 0));
        std::vector<uint32_t> ments(me.size() + 1);
        for (size_t i = 0; i < me.size();) {
            size_t j = i;
            while (j < me.size() && me[j].key == me[i].key) { ments[j] = me[j].ent; ++j; }
            uint32_t slot = hash_slot(me[i].key, mshift);
            while (mkeys[slot] != kEmpty64) slot = (slot + 1) & (mcap - 1);
            mkeys[slot] = me[i].key; mvals[slot] = make_uint2((uint32_t)i, (uint32_t)(j - i));
            i = j;
        }
        c->mx_mask = mcap ? mcap - 1 : 0; c->mx_shift = mshift;
        TRY(c->d_mx_keys.ensure(mkeys.size() * 8)); TRY(c->d_mx_vals.ensure(mvals.size() * 8)); TRY(c->d_mx_entries.ensure(ments.size() * 4));
        CK(cudaMemcpy(c->d_mx_keys.p, mkeys.data(), mkeys.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->d_mx_vals.p, mvals.data(), mvals.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->d_mx_entries.p, ments.data(), ments.size() * 4, cudaMemcpyHostToDevice));
    }
    // (2) SNPmer inverted index from asv.snpmers_vec(): k-mers re-extracted from positions with
    //     kmer_from_position's canonical rule (tie -> forward), src/types.rs:622-663
    struct Ent { uint64_t sp; uint32_t asv; uint32_t mid; };
    std::vector<Ent> ents;
    const uint64_t split_mask = c->dp.split_mask;
    for (size_t a = 0; a < A; ++a) {
        const uint8_t *sq = c->asv_bases_h.data() + c->asv_offsets_h[a];
        for (uint32_t t = 0; t < snp_cnt[a]; ++t) {
            const uint32_t pos = snp_pos[snp_start[a] + t];
            uint64_t kmer = 0;
            for (int q = 0; q < k; ++q) kmer = (kmer << 2) | enc_host(sq[pos + q]);
            const uint64_t rc = revcomp_kmer_host(kmer, k);
            const uint64_t canon = ((rc & split_mask) < (kmer & split_mask)) ? rc : kmer;
            ents.push_back({canon & split_mask, (uint32_t)a, (uint32_t)((canon >> (k - 1)) & 3)});
        }
    }
    std::sort(ents.begin(), ents.end(), [](const Ent &x, const Ent &y) { return x.sp != y.sp ? x.sp < y.sp : x.asv < y.asv; });
    size_t n_keys = 0;
    for (size_t i = 0; i < ents.size(); ++i) if (i == 0 || ents[i].sp != ents[i - 1].sp) ++n_keys;
    std::vector<uint32_t> entries(ents.size() + 1);
    const uint32_t cap = n_keys ? next_pow2_u32(4 * n_keys + 2) : 0;
    const uint32_t ishift = cap ? shift_for_cap(cap) : 0;
    std::vector<uint64_t> keys(cap ? cap : 1, kEmpty64);
    std::vector<uint2> vals(cap ? cap : 1, make_uint2(0, 0));
    for (size_t i = 0; i < ents.size();) {
        size_t j = i;
        while (j < ents.size() && ents[j].sp == ents[i].sp) { entries[j] = (ents[j].asv << 2) | ents[j].mid; ++j; }
        uint32_t slot = hash_slot(ents[i].sp, ishift);
        while (keys[slot] != kEmpty64) slot = (slot + 1) & (cap - 1);
        keys[slot] = ents[i].sp; vals[slot] = make_uint2((uint32_t)i, (uint32_t)(j - i));
        i = j;
    }
    c->idx_mask = cap ? cap - 1 : 0; c->idx_shift = ishift;
    // (3) DP planes
    std::vector<uint32_t> plane_off, lens;
    std::vector<uint2> planes;
    build_dp_planes(c->dp, c->asv_bases_h.data(), c->asv_offsets_h.data(), A, planes, plane_off, lens, &c->max_asv_len);
    c->asv_len_h = lens;
    // upload
    TRY(c->d_mset.ensure((mset.size() + 1) * 8)); TRY(c->d_mset_off.ensure((A + 1) * 4));
    TRY(c->d_idx_keys.ensure(keys.size() * 8)); TRY(c->d_idx_vals.ensure(vals.size() * 8)); TRY(c->d_idx_entries.ensure(entries.size() * 4));
    TRY(c->d_asv_planes.ensure(planes.size() * 8)); TRY(c->d_asv_plane_off.ensure(A * 4)); TRY(c->d_asv_len.ensure(A * 4));
    if (!mset.empty()) CK(cudaMemcpy(c->d_mset.p, mset.data(), mset.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_mset_off.p, mset_off.data(), (A + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_idx_keys.p, keys.data(), keys.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_idx_vals.p, vals.data(), vals.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_idx_entries.p, entries.data(), entries.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_asv_planes.p, planes.data(), planes.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_asv_plane_off.p, plane_off.data(), A * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_asv_len.p, lens.data(), A * 4, cudaMemcpyHostToDevice));
    return 0;
}

static int ensure_index(Context *c) {
    if (c->index_dirty) TRY(build_asv_index(c));
    if (c->n_asv == 0) return fail(SAVONT_B200_ERR_INVALID, "no ASVs loaded (call savont_b200_load_asvs first)");
    return 0;
}

int use_device(Context *c) {
    if (!c) return fail(SAVONT_B200_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    return 0;
}


// sketch -> filter -> align -> assign on a device-resident batch: four kernel groups enqueued back to back; the ONE
// host synchronisation happens in finish_path; per-group device times from CUDA events on the stream
static int enqueue_path(Context *c, Batch &b, int n_samples) {
    CK(cudaEventRecord(c->pev[0], c->stream));
    TRY(stage_sketch(c, b, false));
    TRY(stage_prefilter(c, b, false));
    CK(cudaEventRecord(c->pev[1], c->stream));
    TRY(stage_filter(c, b, c->dp.strand_mode));
    CK(cudaEventRecord(c->pev[2], c->stream));
    TRY(stage_align(c, b, b.ctrs.as<uint32_t>() + 1, b.pair_cap));
    CK(cudaEventRecord(c->pev[3], c->stream));
    TRY(stage_assign(c, b, n_samples, c->dp.strand_mode));
    CK(cudaEventRecord(c->pev[4], c->stream));
    return 0;
}
// waits for the batch enqueued last, re-runs it with larger pools when one overflowed, accumulates the statistics
static int finish_path(Context *c, Batch &b, int n_samples, savont_b200_stats &st) {
    BatchCounters h{};
    for (int attempt = 0;; ++attempt) {
        bool retry = false;
        TRY(check_batch(c, b, &h, &retry));
        if (!retry) break;
        if (attempt == 3) return fail(SAVONT_B200_ERR_OVERFLOW, "pool sizing did not converge");
        TRY(enqueue_path(c, b, n_samples));
    }
    st.sketch_ms += ev_ms(c->pev[0], c->pev[1]);
    st.filter_ms += ev_ms(c->pev[1], c->pev[2]);
    st.align_ms += ev_ms(c->pev[2], c->pev[3]);
    st.assign_ms += ev_ms(c->pev[3], c->pev[4]);
    st.n_pairs += h.n_pairs;
    st.n_reads += b.n; st.n_bases += b.total_bases;
    st.dp_cells += h.dp_cells;
    return 0;
}
// runs the stage launchers / the batch check on one of the internal compute lanes instead of the caller's stream
struct LaneGuard {
    Context *c; cudaStream_t s0; cudaEvent_t *e0;
    LaneGuard(Context *ctx, int lane) : c(ctx), s0(ctx->stream), e0(ctx->pev) { c->stream = c->lane_stream[lane]; c->pev = c->lane_ev[lane]; }
    ~LaneGuard() { c->stream = s0; c->pev = e0; }
};
static int run_path(Context *c, Batch &b, int n_samples, savont_b200_stats &st) {
    TRY(enqueue_path(c, b, n_samples));
    return finish_path(c, b, n_samples, st);
}

// results of a finished batch of n reads -> caller memory, on stream `fs` (the batch's kernels must have completed)
static int fetch_batch(Context *c, Batch &b, size_t n, cudaStream_t fs, uint8_t *status, int32_t *best_nm, uint64_t *class_offsets,
                       uint64_t class_base, uint32_t *class_asv, size_t class_cap, uint64_t *n_class_out, uint64_t *d2h_bytes) {
    TRY(c->pin_small.ensure((n + 2) * 4));
    uint32_t *off = c->pin_small.as<uint32_t>();
    CK(cudaMemcpyAsync(off, b.class_off.p, (n + 1) * 4, cudaMemcpyDeviceToHost, fs));
    if (n) CK(cudaMemcpyAsync(status, b.rstatus.p, n, cudaMemcpyDeviceToHost, fs));
    if (n) CK(cudaMemcpyAsync(best_nm, b.best_nm.p, n * 4, cudaMemcpyDeviceToHost, fs));
    CK(cudaStreamSynchronize(fs));
    const uint32_t total = off[n];
    if (class_base + total > class_cap)
        return fail(SAVONT_B200_ERR_CAPACITY, "class_asv capacity %zu too small (need >= %llu)", class_cap,
                    (unsigned long long)(class_base + total));
    if (total) CK(cudaMemcpyAsync(class_asv + class_base, b.class_out.p, (size_t)total * 4, cudaMemcpyDeviceToHost, fs));
    for (size_t i = 0; i <= n; ++i) class_offsets[i] = class_base + off[i];
    CK(cudaStreamSynchronize(fs));
    *n_class_out = total;
    if (d2h_bytes) *d2h_bytes += (n + 1) * 4 + n * 5 + (uint64_t)total * 4;
    return 0;
}

static int read_counters(Context *c, uint64_t *unambig, uint64_t *ambig, uint64_t *leq10, uint64_t *totals, uint64_t *single,
                         bool add) {
    if (c->counters_len == 0) return fail(SAVONT_B200_ERR_INVALID, "no counters: nothing was mapped yet");
    std::vector<uint64_t> h(c->counters_len);
    CK(cudaMemcpyAsync(h.data(), c->d_counters.p, c->counters_len * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const size_t A = c->n_asv;
    auto put = [&](uint64_t *dst, const uint64_t *src, size_t n) {
        if (!dst) return;
        for (size_t i = 0; i < n; ++i) dst[i] = (add ? dst[i] : 0) + src[i];
    };
    put(unambig, h.data(), A); put(ambig, h.data() + A, A); put(leq10, h.data() + 2 * A, A);
    put(totals, h.data() + 3 * A, 2);
    put(single, h.data() + 3 * A + 2, (size_t)c->counters_samples * A);
    return 0;
}

}  // namespace sb


// ------------------------------------------------------------------------------------------------
// K5: Stage 1 split-k-mer counting (kmercount.cu)

// ================================================================================================
// extern "C"
// ================================================================================================
using namespace sb;


namespace sb {
// Traceback (align_tb_kernel) of explicit device-resident pairs against the batch's read planes, in sub-batches bounded by
// a 4 GiB history workspace; the CIGAR slots are compacted on the device (scan of the op counts + warp-per-pair copy) so that
// only the ops cross PCIe.  cigar_offsets[0] must be set by the caller (base of this call's ops in `cigar`).
int traceback_pairs(Context *c, Batch &b, const AsvIndex &ix, uint32_t max_ref_len, const uint32_t *d_pr, const uint32_t *d_pa,
                    const uint8_t *d_ps, size_t n_pairs, int32_t *nm, int32_t *asv_end, int32_t *read_end, int32_t *asv_start,
                    int32_t *read_start, uint64_t *cigar_offsets, uint32_t *cigar, size_t cigar_cap, std::vector<uint32_t> *cigar_vec,
                    TbDevice *dev) {
    const uint32_t hist_cols = std::max<uint32_t>(b.max_len, 1u);
    const uint32_t cigar_stride = b.max_len + max_ref_len + 2;
    const size_t per_pair_words = align_tb_hist_words(1, hist_cols, c->dp);
    size_t sub = std::max<size_t>(128, ((size_t)1 << 30) / std::max<size_t>(per_pair_words, 1));
    sub = std::min<size_t>(sub, 1u << 17);
    DevBuf &d_hist = c->tb_ws[0], &d_slots = c->tb_ws[1], &d_cnt = c->tb_ws[2], &d_off = c->tb_ws[3], &d_tmp = c->tb_ws[4],
           &d_ops = c->tb_ws[5], &d_out = c->tb_ws[6];
    auto cleanup = [&]() {};                     // the workspace stays with the context (released by savont_b200_destroy)
    uint64_t cig_base = cigar_offsets[0];
    std::vector<uint32_t> h_off;
    std::vector<int32_t> h_out;
    int rc = 0;
    static const bool tb_trace = getenv("SAVONT_B200_TRACE") != nullptr;
    auto tnow = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_kern = 0, t_copy = 0, t_host = 0;
    for (size_t p0 = 0; p0 < n_pairs && rc == 0; p0 += sub) {
        const double ta = tnow();
        const uint32_t np = (uint32_t)std::min(sub, n_pairs - p0);
        if ((rc = d_hist.ensure(align_tb_hist_words(np, hist_cols, c->dp) * 4)) != 0) break;
        if ((rc = d_slots.ensure((size_t)np * cigar_stride * 4)) != 0) break;
        if ((rc = d_ops.ensure((size_t)np * cigar_stride * 4)) != 0) break;
        if ((rc = d_cnt.ensure((size_t)np * 4)) != 0) break;
        if ((rc = d_off.ensure(((size_t)np + 1) * 4)) != 0) break;
        if ((rc = d_tmp.ensure(((size_t)np / 2048 + 4) * 4)) != 0) break;
        if ((rc = d_out.ensure((size_t)np * 5 * 4)) != 0) break;
        int32_t *o = d_out.as<int32_t>();
        AlignOut ao{o, o + np, o + 2 * (size_t)np};
        launch_align_tb(seq_view(b), b.planes.as<uint2>(), ix, d_pr + p0, d_pa + p0, d_ps + p0, np, d_hist.as<uint32_t>(), ao,
                        o + 3 * (size_t)np, o + 4 * (size_t)np, d_slots.as<uint32_t>(), cigar_stride, d_cnt.as<uint32_t>(), c->dp, c->stream);
        launch_scan_u32(d_cnt.as<uint32_t>(), d_off.as<uint32_t>(), np, d_tmp.as<uint32_t>(), c->stream);
        launch_cigar_compact(d_slots.as<uint32_t>(), cigar_stride, d_cnt.as<uint32_t>(), d_off.as<uint32_t>(), np, d_ops.as<uint32_t>(),
                             c->stream);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { rc = fail(SAVONT_B200_ERR_CUDA, "align_tb launch: %s", cudaGetErrorString(e)); break; }
        c->last_stats.kernel_launches += 5;
        h_off.resize((size_t)np + 1); h_out.resize((size_t)np * 5);
        cudaMemcpyAsync(h_off.data(), d_off.p, ((size_t)np + 1) * 4, cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(h_out.data(), d_out.p, (size_t)np * 5 * 4, cudaMemcpyDeviceToHost, c->stream);
        e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { rc = fail(SAVONT_B200_ERR_CUDA, "align_tb: %s", cudaGetErrorString(e)); break; }
        const double tb_ = tnow();
        const uint32_t total = h_off[np];
        const bool keep_on_device = dev && p0 == 0 && np == n_pairs;       // the caller consumes the ops on the device
        if (dev) *dev = keep_on_device ? TbDevice{true, d_ops.as<uint32_t>(), d_off.as<uint32_t>(), d_out.as<int32_t>(), np} : TbDevice{};
        if (keep_on_device) { /* no copy of the ops */ }
        else if (cigar_vec) { cigar_vec->resize(cig_base + total + 1); cigar = cigar_vec->data(); }   // growable destination
        else if (cig_base + total > cigar_cap) { rc = fail(SAVONT_B200_ERR_CAPACITY, "cigar capacity %zu too small", cigar_cap); break; }
        if (total && !keep_on_device) cudaMemcpyAsync(cigar + cig_base, d_ops.p, (size_t)total * 4, cudaMemcpyDeviceToHost, c->stream);
        memcpy(nm + p0, h_out.data(), (size_t)np * 4);
        memcpy(asv_end + p0, h_out.data() + np, (size_t)np * 4);
        memcpy(read_end + p0, h_out.data() + 2 * (size_t)np, (size_t)np * 4);
        memcpy(asv_start + p0, h_out.data() + 3 * (size_t)np, (size_t)np * 4);
        memcpy(read_start + p0, h_out.data() + 4 * (size_t)np, (size_t)np * 4);
        for (uint32_t t = 0; t < np; ++t) cigar_offsets[p0 + t + 1] = cig_base + h_off[t + 1];
        const double tc = tnow();
        e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { rc = fail(SAVONT_B200_ERR_CUDA, "align_tb copy: %s", cudaGetErrorString(e)); break; }
        cig_base += total;
        const double td = tnow();
        t_kern += tb_ - ta; t_host += tc - tb_; t_copy += td - tc;
    }
    if (tb_trace) fprintf(stderr, "[trace] traceback_pairs: %zu pairs in sub-batches of %zu: kernels + record copy %.2f ms, host %.2f ms, CIGAR copy wait %.2f ms\n", n_pairs, sub, t_kern, t_host, t_copy);
    cleanup();
    return rc;
}

}  // namespace sb

extern "C" {

static int ctx_init(savont_b200_ctx *h, int device_id);

int savont_b200_abi_version(void) { return SAVONT_B200_ABI_VERSION; }
const char *savont_b200_last_error(void) { return last_error_ref().c_str(); }

void savont_b200_default_params(savont_b200_params *p) {
    if (!p) return;
    p->k = 17; p->c = 11; p->l = 3; p->min_bq = 25; p->band = 64; p->strand_mode = 0; p->candidate_mode = 0; p->reserved = 0;
    p->ratio_thresh = 0.0050; p->min_mini_frac_base = 0.950;
}

int savont_b200_create(int device_id, savont_b200_ctx **out) {
    if (!out) return fail(SAVONT_B200_ERR_INVALID, "out is null");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(SAVONT_B200_ERR_CUDA, "no CUDA device available (%s); libsavont_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device_id < 0 || device_id >= n_dev) return fail(SAVONT_B200_ERR_INVALID, "device %d out of range (0..%d)", device_id, n_dev - 1);
    CK(cudaSetDevice(device_id));
    savont_b200_ctx *h = new savont_b200_ctx();
    const int rc_init = ctx_init(h, device_id);
    if (rc_init != 0) { savont_b200_destroy(h); return rc_init; }      // releases whatever was created before the failure
    *out = h;
    return 0;
}

static int ctx_init(savont_b200_ctx *h, int device_id) {
    Context *c = &h->c;
    c->device = device_id;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device_id));
    c->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    for (auto &ev : c->ev_d2h) CK(cudaEventCreate(&ev));
    for (int i = 0; i < 2; ++i) { CK(cudaEventCreate(&c->kc_ev_up[i])); CK(cudaEventCreate(&c->kc_ev_done[i])); }
    for (int i = 0; i < 2; ++i) {
        CK(cudaStreamCreateWithFlags(&c->lane_stream[i], cudaStreamNonBlocking));
        for (auto &ev : c->lane_ev[i]) CK(cudaEventCreate(&ev));
    }
    c->pev = c->ev;
    if (const char *e = getenv("SAVONT_B200_KC_PARTITIONED")) c->kc_partitioned = atoi(e) != 0;
    if (const char *e = getenv("SAVONT_B200_KC_SMEM_LIMIT")) c->kc_smem_limit = (uint32_t)atoi(e);
    for (auto &ev : c->ev) CK(cudaEventCreate(&ev));
    for (int i = 0; i < 4; ++i) { CK(cudaEventCreate(&c->ev_up[i])); CK(cudaEventCreate(&c->ev_h0[i])); }
    savont_b200_default_params(&c->params);
    c->dp = make_device_params(c->params);
    memset(&c->last_stats, 0, sizeof(c->last_stats));
    return build_tables(c);
}

int savont_b200_destroy(savont_b200_ctx *h) {
    if (!h) return 0;
    Context *c = &h->c;
    cudaSetDevice(c->device);
    // c->stream may be the caller's (savont_b200_set_stream); the lanes / copy streams may still hold work of a failed call
    cudaStreamSynchronize(c->stream);
    for (cudaStream_t s : {c->own_stream, c->lane_stream[0], c->lane_stream[1], c->copy_stream, c->d2h_stream}) if (s) cudaStreamSynchronize(s);
    for (Batch &pb : c->res_parts) pb.release();
    c->work.release(); c->work2.release(); c->work3.release(); c->work4.release(); c->resident.release(); c->asvb.release();
    DevBuf *bufs[] = {&c->d_rank_table, &c->d_snp_keys, &c->d_snp_bloom, &c->d_phred, &c->d_thresh, &c->d_mx_keys, &c->d_mx_vals, &c->d_mx_entries, &c->d_idx_keys, &c->d_idx_vals, &c->d_idx_entries, &c->d_mset,
                      &c->d_mset_off, &c->d_asv_planes, &c->d_asv_plane_off, &c->d_asv_len, &c->d_counters, &c->d_hf_keys};
    for (DevBuf *b : bufs) b->release();
    c->pin_small.release(); c->pin_out.release(); c->tb_pin.release();
    for (int i = 0; i < 8; ++i) { c->gather_words[i].release(); c->gather_snp[i].release(); c->gather_meta[i].release(); }
    c->pool.stop();
    for (DevBuf &b : c->tb_ws) b.release();
    for (DevBuf &b : c->pileup_ws) b.release();
    for (DevBuf &b : c->cluster_ws) b.release();
    c->d_ks_entries.release(); c->d_ks_dstptr.release(); c->d_ks_partoff.release();
    for (PinBuf &b : c->cluster_pin) b.release();
    c->stage_pin.release();
    for (DevBuf &b : c->bh_ws) b.release();
    for (auto &ev : c->ev) if (ev) cudaEventDestroy(ev);
    for (int i = 0; i < 4; ++i) { if (c->ev_up[i]) cudaEventDestroy(c->ev_up[i]); if (c->ev_h0[i]) cudaEventDestroy(c->ev_h0[i]); }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (auto &ev : c->ev_d2h) if (ev) cudaEventDestroy(ev);
    for (int i = 0; i < 2; ++i) { if (c->kc_ev_up[i]) cudaEventDestroy(c->kc_ev_up[i]); if (c->kc_ev_done[i]) cudaEventDestroy(c->kc_ev_done[i]); }
    for (auto &ch : c->kc_chunks) { ch.bases.release(); ch.quals.release(); ch.offsets.release(); ch.rc.release(); ch.pin.release(); }
    c->kc_chunks.clear();
    c->d_kc_table.release(); c->d_kc_ctr.release(); c->d_kc_out_k.release(); c->d_kc_out_c.release(); c->d_kp_small.release(); c->d_kp_entries.release(); c->d_kp_entries1.release();
    for (int i = 0; i < 2; ++i) {
        for (auto &ev : c->lane_ev[i]) if (ev) cudaEventDestroy(ev);
        if (c->lane_stream[i]) cudaStreamDestroy(c->lane_stream[i]);
    }
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete h;
    return 0;
}

int savont_b200_set_stream(savont_b200_ctx *h, void *cuda_stream) {
    TRY(use_device(h ? &h->c : nullptr));
    h->c.stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->c.own_stream;
    return 0;
}

int savont_b200_set_params(savont_b200_ctx *h, const savont_b200_params *p) {
    TRY(use_device(h ? &h->c : nullptr));
    if (!p) return fail(SAVONT_B200_ERR_INVALID, "params is null");
    TRY(validate_params(*p));
    Context *c = &h->c;
    const int old_k = c->dp.k;
    c->params = *p;
    c->dp = make_device_params(*p);
    c->index_dirty = true;
    TRY(build_tables(c));
    // the SNPmer Bloom filter is keyed by the k-mer and its reverse complement, i.e. by k: rebuild it from the host copy when k
    // changed after savont_b200_load_snpmers
    if (c->dp.k != old_k && !c->h_snpmers.empty()) TRY(upload_snp_set(c));
    return 0;
}

int savont_b200_set_batch_reads(savont_b200_ctx *h, size_t n) {
    if (!h) return fail(SAVONT_B200_ERR_INVALID, "null context");
    h->c.batch_reads = n ? n : 131072;
    return 0;
}

int savont_b200_load_snpmers(savont_b200_ctx *h, const uint64_t *kmers, size_t n) {
    TRY(use_device(h ? &h->c : nullptr));
    if (n && !kmers) return fail(SAVONT_B200_ERR_INVALID, "kmers is null");
    Context *c = &h->c;
    c->h_snpmers.assign(kmers, kmers + n);
    std::sort(c->h_snpmers.begin(), c->h_snpmers.end());
    c->h_snpmers.erase(std::unique(c->h_snpmers.begin(), c->h_snpmers.end()), c->h_snpmers.end());
    for (uint64_t v : c->h_snpmers)
        if (v == kEmpty64) return fail(SAVONT_B200_ERR_INVALID, "0xffffffffffffffff is not a valid k-mer");
    c->index_dirty = true;
    return upload_snp_set(c);
}

int savont_b200_load_high_freq_kmers(savont_b200_ctx *h, const uint64_t *kmers, size_t n) {
    TRY(use_device(h ? &h->c : nullptr));
    if (n && !kmers) return fail(SAVONT_B200_ERR_INVALID, "kmers is null");
    Context *c = &h->c;
    c->h_hf.assign(kmers, kmers + n);
    std::sort(c->h_hf.begin(), c->h_hf.end());
    c->h_hf.erase(std::unique(c->h_hf.begin(), c->h_hf.end()), c->h_hf.end());
    for (uint64_t v : c->h_hf)
        if (v == kEmpty64) return fail(SAVONT_B200_ERR_INVALID, "0xffffffffffffffff is not a valid k-mer");
    if (c->h_hf.empty()) { c->hf_mask = 0; return 0; }
    const uint32_t cap = next_pow2_u32(4 * c->h_hf.size() + 2);
    const uint32_t shift = shift_for_cap(cap);
    std::vector<uint64_t> keys(cap, kEmpty64);
    for (uint64_t key : c->h_hf) {
        uint32_t slot = hash_slot(key, shift);
        while (keys[slot] != kEmpty64) slot = (slot + 1) & (cap - 1);
        keys[slot] = key;
    }
    TRY(c->d_hf_keys.ensure((size_t)cap * 8));
    CK(cudaMemcpyAsync(c->d_hf_keys.p, keys.data(), (size_t)cap * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->hf_mask = cap - 1; c->hf_shift = shift;
    return 0;
}

int savont_b200_set_read_filter(savont_b200_ctx *h, const savont_b200_read_filter *f) {
    if (!h) return fail(SAVONT_B200_ERR_INVALID, "null context");
    Context *c = &h->c;
    if (!f) { c->rf_on = false; return 0; }
    if (f->max_read_length && f->max_read_length < f->min_read_length)
        return fail(SAVONT_B200_ERR_INVALID, "max_read_length %u < min_read_length %u", f->max_read_length, f->min_read_length);
    if (f->quality_value_cutoff != f->quality_value_cutoff) return fail(SAVONT_B200_ERR_INVALID, "quality_value_cutoff is NaN");
    c->rf.min_len = f->min_read_length; c->rf.max_len = f->max_read_length; c->rf.max_count = f->max_kmer_count_in_read;
    c->rf.use_cutoff = f->quality_value_cutoff > 0.0 ? 1u : 0u; c->rf.cutoff = f->quality_value_cutoff;
    c->rf_on = true;
    return 0;
}

int savont_b200_twin_reads_from_snpmers(savont_b200_ctx *h, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets,
                                        size_t n_reads, uint8_t *status, double *est_id, uint32_t *n_solid_mini,
                                        uint32_t *n_solid_snp) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!offsets || !status) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    if (!c->rf_on) return fail(SAVONT_B200_ERR_INVALID, "no read filter set (call savont_b200_set_read_filter first)");
    for (size_t r0 = 0; r0 < n_reads;) {
        size_t r1 = std::min(n_reads, r0 + c->batch_reads);
        while (r1 > r0 + 1 && offsets[r1] - offsets[r0] > (3ull << 30)) r1 = r0 + (r1 - r0) / 2;
        const size_t n = r1 - r0;
        Batch &b = c->work;
        TRY(batch_upload(c, b, bases, quals, offsets + r0, nullptr, n, nullptr));
        TRY(sketch_checked(c, b, false));
        TRY(b.est_id.ensure((n + 1) * 8));
        launch_est_id(seq_view(b), b.sk_status.as<uint8_t>(), b.flatq.as<uint8_t>(), c->d_phred.as<double>(), b.est_id.as<double>(), c->stream);
        if (est_id) CK(cudaMemcpyAsync(est_id + r0, b.est_id.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
        TRY(stage_prefilter(c, b, true));
        CK(cudaMemcpyAsync(status + r0, b.sk_status.p, n, cudaMemcpyDeviceToHost, c->stream));
        if (n_solid_mini) CK(cudaMemcpyAsync(n_solid_mini + r0, b.solid_mini.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
        if (n_solid_snp) CK(cudaMemcpyAsync(n_solid_snp + r0, b.solid_snp.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        r0 = r1;
    }
    return 0;
}

int savont_b200_qual_seq(savont_b200_ctx *h, const uint8_t *quals, const uint64_t *offsets, size_t n_reads, uint64_t *word_offsets,
                         uint64_t *qual_words, size_t words_cap) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!offsets || !word_offsets || (n_reads && !quals)) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    word_offsets[0] = 0;
    for (size_t r = 0; r < n_reads; ++r) {
        if (offsets[r + 1] < offsets[r]) return fail(SAVONT_B200_ERR_INVALID, "offsets not monotonic at %zu", r + 1);
        const uint64_t bins = (offsets[r + 1] - offsets[r] + 3u) / 4u;               // QUALITY_SEQ_BIN = 4 (src/constants.rs:48)
        word_offsets[r + 1] = word_offsets[r] + (bins + 15u) / 16u;
    }
    if (word_offsets[n_reads] > words_cap) return fail(SAVONT_B200_ERR_CAPACITY, "qual_words holds %zu words, %llu needed", words_cap, (unsigned long long)word_offsets[n_reads]);
    if (word_offsets[n_reads] && !qual_words) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    DevBuf *w = c->cluster_ws;                                                         // the clustering scratch is free between calls
    cudaStream_t st = c->stream;
    for (size_t r0 = 0; r0 < n_reads;) {
        size_t r1 = std::min(n_reads, r0 + (size_t)(1u << 20));
        while (r1 > r0 + 1 && offsets[r1] - offsets[r0] > (1ull << 30)) r1 = r0 + (r1 - r0) / 2;
        const size_t n = r1 - r0;
        const uint64_t nbytes = offsets[r1] - offsets[r0], nwords = word_offsets[r1] - word_offsets[r0];
        std::vector<uint64_t> rel(2 * (n + 1));
        for (size_t i = 0; i <= n; ++i) { rel[i] = offsets[r0 + i] - offsets[r0]; rel[n + 1 + i] = word_offsets[r0 + i] - word_offsets[r0]; }
        TRY(w[0].ensure(nbytes + 16)); TRY(w[1].ensure(rel.size() * 8)); TRY(w[2].ensure(nwords * 8 + 16));
        if (nbytes) CK(cudaMemcpyAsync(w[0].p, quals + offsets[r0], nbytes, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(w[1].p, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, st));
        launch_qual_seq(w[0].as<uint8_t>(), w[1].as<uint64_t>(), w[1].as<uint64_t>() + n + 1, (uint32_t)n, w[2].as<uint64_t>(), c->sm_count, st);
        CK(cudaGetLastError());
        if (nwords) CK(cudaMemcpyAsync(qual_words + word_offsets[r0], w[2].p, nwords * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        r0 = r1;
    }
    return 0;
}

int savont_b200_load_asvs(savont_b200_ctx *h, const uint8_t *bases, const uint64_t *offsets, size_t n_asv) {
    TRY(use_device(h ? &h->c : nullptr));
    if (n_asv && (!bases || !offsets)) return fail(SAVONT_B200_ERR_INVALID, "null ASV input");
    Context *c = &h->c;
    if (n_asv >= (1u << 30)) return fail(SAVONT_B200_ERR_INVALID, "too many ASVs");
    c->asv_offsets_h.assign(n_asv + 1, 0);
    const uint64_t base0 = n_asv ? offsets[0] : 0;
    for (size_t i = 0; i <= n_asv && n_asv; ++i) c->asv_offsets_h[i] = offsets[i] - base0;
    c->asv_bases_h.assign(bases + base0, bases + base0 + (n_asv ? offsets[n_asv] - base0 : 0));
    c->index_dirty = true;
    return build_asv_index(c);
}

int savont_b200_num_asvs(savont_b200_ctx *h, size_t *n_asv) {
    if (!h || !n_asv) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    *n_asv = h->c.n_asv;
    return 0;
}

int savont_b200_sketch(savont_b200_ctx *h, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets, size_t n_seqs,
                       uint8_t *status, double *est_id, uint64_t *mini_offsets, uint32_t *mini_pos, uint64_t *mini_kmer,
                       uint8_t *mini_fwd, size_t mini_cap, uint64_t *snp_offsets, uint32_t *snp_pos, uint64_t *snp_kmer,
                       size_t snp_cap, uint8_t *codes) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!offsets || !status || !mini_offsets || !snp_offsets) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    mini_offsets[0] = 0; snp_offsets[0] = 0;
    size_t mini_base = 0, snp_base = 0;
    for (size_t r0 = 0; r0 < n_seqs;) {
        size_t r1 = std::min(n_seqs, r0 + c->batch_reads);
        while (r1 > r0 + 1 && offsets[r1] - offsets[r0] > (3ull << 30)) r1 = r0 + (r1 - r0) / 2;
        const size_t n = r1 - r0;
        Batch &b = c->work;
        TRY(batch_upload(c, b, bases, quals, offsets + r0, nullptr, n, nullptr));
        TRY(sketch_checked(c, b, true));
        // compact CSR on the device
        TRY(b.class_off.ensure((n + 2) * 4)); TRY(b.pair_start.ensure((n + 2) * 4)); TRY(b.scan_tmp.ensure((n / 2048 + 4) * 4));
        uint32_t *d_moff = b.class_off.as<uint32_t>(), *d_soff = b.pair_start.as<uint32_t>();
        launch_scan_u32(b.mini_cnt.as<uint32_t>(), d_moff, (uint32_t)n, b.scan_tmp.as<uint32_t>(), c->stream);
        launch_scan_u32(b.snp_cnt.as<uint32_t>(), d_soff, (uint32_t)n, b.scan_tmp.as<uint32_t>(), c->stream);
        std::vector<uint32_t> moff(n + 1), soff(n + 1);
        CK(cudaMemcpyAsync(moff.data(), d_moff, (n + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(soff.data(), d_soff, (n + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        const size_t nm = moff[n], ns = soff[n];
        if (mini_base + nm > mini_cap || snp_base + ns > snp_cap)
            return fail(SAVONT_B200_ERR_CAPACITY, "sketch output capacity too small (need mini>=%zu snp>=%zu)", mini_base + nm, snp_base + ns);
        DevBuf &o_mk = b.p_read, &o_mp = b.p_asv, &o_mf = b.p_strand, &o_sk = b.p_mini, &o_sp = b.nm;   // reuse as scratch
        TRY(o_mk.ensure((nm + 1) * 8)); TRY(o_mp.ensure((nm + 1) * 4)); TRY(o_mf.ensure(nm + 4));
        TRY(o_sk.ensure((ns + 1) * 8)); TRY(o_sp.ensure((ns + 1) * 4));
        launch_gather_sketch(seq_view(b), sketch_view(b, true), d_moff, d_soff, o_mp.as<uint32_t>(), o_mk.as<uint64_t>(),
                             o_mf.as<uint8_t>(), o_sp.as<uint32_t>(), o_sk.as<uint64_t>(), c->dp, c->stream);
        if (est_id) {
            TRY(b.est_id.ensure((n + 1) * 8));
            launch_est_id(seq_view(b), b.sk_status.as<uint8_t>(), b.flatq.as<uint8_t>(), c->d_phred.as<double>(), b.est_id.as<double>(), c->stream);
            CK(cudaMemcpyAsync(est_id + r0, b.est_id.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(status + r0, b.sk_status.p, n, cudaMemcpyDeviceToHost, c->stream));
        if (nm) {
            CK(cudaMemcpyAsync(mini_pos + mini_base, o_mp.p, nm * 4, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(mini_kmer + mini_base, o_mk.p, nm * 8, cudaMemcpyDeviceToHost, c->stream));
            if (mini_fwd) CK(cudaMemcpyAsync(mini_fwd + mini_base, o_mf.p, nm, cudaMemcpyDeviceToHost, c->stream));
        }
        if (ns) {
            CK(cudaMemcpyAsync(snp_pos + snp_base, o_sp.p, ns * 4, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(snp_kmer + snp_base, o_sk.p, ns * 8, cudaMemcpyDeviceToHost, c->stream));
        }
        if (codes && b.total_bases) {
            TRY(b.class_out.ensure(b.total_bases + 16));
            launch_codes(b.bases.as<uint8_t>(), b.class_out.as<uint8_t>(), b.total_bases, c->stream);
            CK(cudaMemcpyAsync(codes + offsets[r0], b.class_out.p, b.total_bases, cudaMemcpyDeviceToHost, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < n; ++i) { mini_offsets[r0 + i + 1] = mini_base + moff[i + 1]; snp_offsets[r0 + i + 1] = snp_base + soff[i + 1]; }
        mini_base += nm; snp_base += ns;
        r0 = r1;
    }
    return 0;
}

int savont_b200_find_candidates(savont_b200_ctx *h, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets,
                                size_t n_reads, uint8_t *cand_status, uint64_t *cand_offsets, uint32_t *cand_asv,
                                uint32_t *cand_mm, uint32_t *cand_mini, uint8_t *cand_strand, size_t cand_cap) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    TRY(ensure_index(c));
    if (!offsets || !cand_status || !cand_offsets) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    cand_offsets[0] = 0;
    size_t base = 0;
    for (size_t r0 = 0; r0 < n_reads;) {
        size_t r1 = std::min(n_reads, r0 + c->batch_reads);
        while (r1 > r0 + 1 && offsets[r1] - offsets[r0] > (3ull << 30)) r1 = r0 + (r1 - r0) / 2;
        const size_t n = r1 - r0;
        Batch &b = c->work;
        uint32_t n_pairs = 0;
        TRY(batch_upload(c, b, bases, quals, offsets + r0, nullptr, n, nullptr));
        for (int attempt = 0;; ++attempt) {
            TRY(stage_sketch(c, b, false));
            TRY(stage_prefilter(c, b, false));
            TRY(stage_filter(c, b, 0));
            BatchCounters h{}; bool retry = false;
            TRY(check_batch(c, b, &h, &retry));
            n_pairs = h.n_pairs;
            if (!retry) break;
            if (attempt == 3) return fail(SAVONT_B200_ERR_OVERFLOW, "pool sizing did not converge");
        }
        if (base + n_pairs > cand_cap) return fail(SAVONT_B200_ERR_CAPACITY, "candidate capacity too small (need >= %zu)", base + n_pairs);
        TRY(b.class_off.ensure((n + 2) * 4)); TRY(b.scan_tmp.ensure((n / 2048 + 4) * 4));
        TRY(b.class_out.ensure(((size_t)n_pairs + 1) * 4)); TRY(b.class_slot.ensure(((size_t)n_pairs + 1) * 4));
        TRY(b.nm.ensure(((size_t)n_pairs + 1) * 4)); TRY(b.asv_end.ensure((size_t)n_pairs + 4));
        launch_scan_u32(b.pair_cnt.as<uint32_t>(), b.class_off.as<uint32_t>(), (uint32_t)n, b.scan_tmp.as<uint32_t>(), c->stream);
        launch_gather_cand(pool_view(b), b.pair_start.as<uint32_t>(), b.pair_cnt.as<uint32_t>(), b.low_mm.as<uint32_t>(),
                           b.class_off.as<uint32_t>(), b.class_out.as<uint32_t>(), b.class_slot.as<uint32_t>(), b.nm.as<uint32_t>(),
                           b.asv_end.as<uint8_t>(), (uint32_t)n, c->stream);
        CK(cudaGetLastError());
        std::vector<uint32_t> off(n + 1);
        CK(cudaMemcpyAsync(off.data(), b.class_off.p, (n + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(cand_status + r0, b.cand_status.p, n, cudaMemcpyDeviceToHost, c->stream));
        if (n_pairs) {
            CK(cudaMemcpyAsync(cand_asv + base, b.class_out.p, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(cand_mm + base, b.class_slot.p, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(cand_mini + base, b.nm.p, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(cand_strand + base, b.asv_end.p, (size_t)n_pairs, cudaMemcpyDeviceToHost, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < n; ++i) cand_offsets[r0 + i + 1] = base + off[i + 1];
        base += n_pairs;
        r0 = r1;
    }
    return 0;
}

int savont_b200_align_pairs(savont_b200_ctx *h, const uint8_t *bases, const uint64_t *offsets, size_t n_reads,
                            const uint32_t *pair_read, const uint32_t *pair_asv, const uint8_t *pair_strand, size_t n_pairs,
                            int32_t *nm, int32_t *asv_end, int32_t *read_end, int want_cigar, int32_t *asv_start,
                            int32_t *read_start, uint64_t *cigar_offsets, uint32_t *cigar, size_t cigar_cap) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    TRY(ensure_index(c));
    if (!offsets || (n_pairs && (!pair_read || !pair_asv || !pair_strand || !nm || !asv_end || !read_end)))
        return fail(SAVONT_B200_ERR_INVALID, "null argument");
    if (want_cigar && (!asv_start || !read_start || !cigar_offsets)) return fail(SAVONT_B200_ERR_INVALID, "null cigar argument");
    for (size_t i = 0; i < n_pairs; ++i)
        if (pair_read[i] >= n_reads || pair_asv[i] >= c->n_asv) return fail(SAVONT_B200_ERR_INVALID, "pair %zu out of range", i);
    if (n_pairs >= (1ull << 31)) return fail(SAVONT_B200_ERR_INVALID, "too many pairs");
    Batch &b = c->work;
    TRY(batch_upload(c, b, bases, nullptr, offsets, nullptr, n_reads, nullptr));
    TRY(b.planes.ensure(((size_t)plane_word_off(b.total_bases, n_reads) + 4) * sizeof(uint2)));
    launch_pack_planes(seq_view(b), b.planes.as<uint2>(), c->stream);
    b.pair_cap = std::max<uint32_t>(b.pair_cap, (uint32_t)n_pairs + 16);
    TRY(ensure_pair_pool(b));
    if (n_pairs) {
        CK(cudaMemcpyAsync(b.p_read.p, pair_read, n_pairs * 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(b.p_asv.p, pair_asv, n_pairs * 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(b.p_strand.p, pair_strand, n_pairs, cudaMemcpyHostToDevice, c->stream));
    }
    if (cigar_offsets) cigar_offsets[0] = 0;
    if (!want_cigar) {
        TRY(stage_align(c, b, nullptr, (uint32_t)n_pairs));
        if (n_pairs) {
            CK(cudaMemcpyAsync(nm, b.nm.p, n_pairs * 4, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(asv_end, b.asv_end.p, n_pairs * 4, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(read_end, b.read_end.p, n_pairs * 4, cudaMemcpyDeviceToHost, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
        return 0;
    }
    return traceback_pairs(c, b, index_view(c), c->max_asv_len, b.p_read.as<uint32_t>(), b.p_asv.as<uint32_t>(), b.p_strand.as<uint8_t>(),
                           n_pairs, nm, asv_end, read_end, asv_start, read_start, cigar_offsets, cigar, cigar_cap);
}

// shared body of savont_b200_map_reads / savont_b200_map_twin_reads: double-buffered batches, the H2D copies of
// batch i+1 run on the copy stream while batch i computes
// the first batches are smaller (1/8, 1/4, 1/2 of batch_reads) so that the pipeline fills quickly: nothing overlaps
// the upload of batch 0
// (shrinking the last batches as well was measured: +2.7 % from FASTQ bytes, -1.5 % from TwinReads -- more batches, more host
// synchronisations -- and left out)
static size_t ramp_batch(const Context *c, size_t bi, size_t /*remaining*/) {
    const size_t full = c->batch_reads;
    static const int steps = [] { const char *e = getenv("SAVONT_B200_RAMP"); return e ? std::max(0, std::min(5, atoi(e))) : 3; }();   // tuning
    if (bi >= (size_t)steps || full < 8 * 4096) return full;
    return full >> (steps - bi);
}
using UploadFn = std::function<int(Batch &, size_t, size_t, uint64_t *, cudaStream_t)>;
// SAVONT_B200_TRACE=1: per-batch host timeline of the e2e pipeline on stderr (ms since the call started)
static bool trace_on() { static const bool on = [] { const char *e = getenv("SAVONT_B200_TRACE"); return e && atoi(e) != 0; }(); return on; }
static std::chrono::steady_clock::time_point g_trace_t0;
static double trace_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g_trace_t0).count(); }
#define TRACE(...) do { if (trace_on()) { fprintf(stderr, "[trace %8.3f] ", trace_ms()); fprintf(stderr, __VA_ARGS__); fputc('\n', stderr); } } while (0)
static int map_batches(Context *c, const std::vector<size_t> &cuts, const UploadFn &upload_batch, int n_samples, uint8_t *status,
                       int32_t *best_nm, uint64_t *class_offsets, uint32_t *class_asv, size_t class_cap, uint64_t *unambig,
                       uint64_t *ambig, uint64_t *leq10, uint64_t *totals, uint64_t *sample_singleton, savont_b200_stats *stats,
                       const std::function<void(size_t)> &on_finished = nullptr) {
    savont_b200_stats st; memset(&st, 0, sizeof(st));
    memset(&c->last_stats, 0, sizeof(c->last_stats));
    if (trace_on()) g_trace_t0 = std::chrono::steady_clock::now();
    TRY(ensure_counters(c, n_samples, true));
    CK(cudaEventRecord(c->ev[6], c->stream));
    class_offsets[0] = 0;
    uint64_t class_base = 0;
    const size_t nb = cuts.size() - 1;
    // four batch slots, two compute lanes: batch bi lives in slot bi % 4 and runs on lane bi & 1.  While batch bi
    // computes, batch bi-1 finishes on the other lane (the tail of its DP kernel overlaps bi's sketch / filter kernels) and
    // has its results fetched on the D2H stream, and batches bi+1, bi+2 are uploaded on the copy stream: the upload runs two
    // batches ahead of the kernels, because with TwinRead input a batch's H2D time equals its kernel time (PCIe bound) and a
    // one-batch lead left bubbles in both chains.
    constexpr int kSlots = 4;
    Batch *bufs[kSlots] = {&c->work, &c->work2, &c->work3, &c->work4};
    auto upload = [&](size_t bi) -> int {
        const int slot = (int)(bi % kSlots);
        TRACE("upload(%zu) begin", bi);
        CK(cudaEventRecord(c->ev_h0[slot], c->copy_stream));
        TRY(upload_batch(*bufs[slot], cuts[bi], cuts[bi + 1] - cuts[bi], &st.h2d_bytes, c->copy_stream));
        CK(cudaEventRecord(c->ev_up[slot], c->copy_stream));
        TRACE("upload(%zu) issued", bi);
        return 0;
    };
    CK(cudaStreamWaitEvent(c->copy_stream, c->ev[6], 0));   // order after earlier work on the caller's stream
    for (int l = 0; l < 2; ++l) CK(cudaStreamWaitEvent(c->lane_stream[l], c->ev[6], 0));
    struct Pending { Batch *b; size_t n, r0, bi; int slot, lane; bool live; } prev{nullptr, 0, 0, 0, 0, 0, false};
    auto finish_prev = [&]() -> int {
        if (!prev.live) return 0;
        TRACE("finish(%zu) wait", prev.bi);
        {
            LaneGuard g(c, prev.lane);
            TRY(finish_path(c, *prev.b, n_samples, st));
        }
        TRACE("finish(%zu) kernels done: h2d %.3f ms, kernels %.3f ms (device)", prev.bi, ev_ms(c->ev_h0[prev.slot], c->ev_up[prev.slot]),
              ev_ms(c->lane_ev[prev.lane][0], c->lane_ev[prev.lane][4]));
        st.h2d_ms += ev_ms(c->ev_h0[prev.slot], c->ev_up[prev.slot]);
        uint64_t n_cls = 0;
        CK(cudaEventRecord(c->ev_d2h[0], c->d2h_stream));
        TRY(fetch_batch(c, *prev.b, prev.n, c->d2h_stream, status + prev.r0, best_nm + prev.r0, class_offsets + prev.r0, class_base,
                        class_asv, class_cap, &n_cls, &st.d2h_bytes));
        CK(cudaEventRecord(c->ev_d2h[1], c->d2h_stream));
        CK(cudaEventSynchronize(c->ev_d2h[1]));
        st.d2h_ms += ev_ms(c->ev_d2h[0], c->ev_d2h[1]);
        class_base += n_cls;
        prev.live = false;
        TRACE("finish(%zu) fetched", prev.bi);
        if (on_finished) on_finished(prev.bi);
        return 0;
    };
    if (nb) TRY(upload(0));
    if (nb > 1) TRY(upload(1));
    for (size_t bi = 0; bi < nb; ++bi) {
        const int slot = (int)(bi % kSlots), lane = (int)(bi & 1);
        Batch &b = *bufs[slot];
        {
            LaneGuard g(c, lane);
            CK(cudaStreamWaitEvent(c->stream, c->ev_up[slot], 0));
            TRY(enqueue_path(c, b, n_samples));
        }
        TRACE("enqueue(%zu) done", bi);
        // batch bi's kernels are queued: now stage batch bi+2 (its slot held batch bi-2: finished and fetched).  With the
        // pointer entry point this includes the host-side gather of the reads' heap buffers, which so runs under bi's kernels
        if (bi + 2 < nb) TRY(upload(bi + 2));
        TRY(finish_prev());
        prev = Pending{&b, cuts[bi + 1] - cuts[bi], cuts[bi], bi, slot, lane, true};
        st.n_batches += 1;
    }
    TRY(finish_prev());
    TRY(read_counters(c, unambig, ambig, leq10, totals, sample_singleton, true));
    st.d2h_bytes += c->counters_len * 8;
    CK(cudaEventRecord(c->ev[7], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    st.total_ms = ev_ms(c->ev[6], c->ev[7]);
    st.kernel_launches = c->last_stats.kernel_launches;
    c->last_stats = st;
    if (stats) *stats = st;
    return 0;
}

int savont_b200_map_reads(savont_b200_ctx *h, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets,
                          const uint32_t *file_idx, size_t n_reads, int n_samples, uint8_t *status, int32_t *best_nm,
                          uint64_t *class_offsets, uint32_t *class_asv, size_t class_cap, uint64_t *unambig, uint64_t *ambig,
                          uint64_t *leq10, uint64_t *totals, uint64_t *sample_singleton, savont_b200_stats *stats) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    TRY(ensure_index(c));
    if (!offsets || !status || !best_nm || !class_offsets) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    if (n_samples < 1) return fail(SAVONT_B200_ERR_INVALID, "n_samples must be >= 1");
    TRY(validate_file_idx(file_idx, n_reads, n_samples));
    std::vector<size_t> cuts{0};
    for (size_t r0 = 0; r0 < n_reads;) {
        size_t r1 = std::min(n_reads, r0 + ramp_batch(c, cuts.size() - 1, n_reads - r0));
        while (r1 > r0 + 1 && offsets[r1] - offsets[r0] > (3ull << 30)) r1 = r0 + (r1 - r0) / 2;
        cuts.push_back(r1);
        r0 = r1;
    }
    auto up = [&](Batch &b, size_t r0, size_t n, uint64_t *bytes, cudaStream_t st) -> int {
        return batch_upload(c, b, bases, quals, offsets + r0, file_idx ? file_idx + r0 : nullptr, n, bytes, st);
    };
    return map_batches(c, cuts, up, n_samples, status, best_nm, class_offsets, class_asv, class_cap, unambig, ambig, leq10, totals,
                       sample_singleton, stats);
}

static std::vector<size_t> twin_cuts(const Context *c, const TwinInput &in, size_t n_reads) {
    std::vector<size_t> cuts{0};
    for (size_t r0 = 0; r0 < n_reads;) {
        size_t r1 = std::min(n_reads, r0 + ramp_batch(c, cuts.size() - 1, n_reads - r0));
        while (r1 > r0 + 1 && (in.word_offsets[r1] - in.word_offsets[r0]) * 32ull > (3ull << 30)) r1 = r0 + (r1 - r0) / 2;
        cuts.push_back(r1);
        r0 = r1;
    }
    return cuts;
}

int savont_b200_map_twin_reads(savont_b200_ctx *h, const uint64_t *dna_words, const uint64_t *dna_word_offsets,
                               const uint32_t *lengths, const void *snpmer_kmers, int snpmer_kmer_bytes,
                               const uint64_t *snpmer_offsets, const uint32_t *file_idx, size_t n_reads, int n_samples,
                               uint8_t *status, int32_t *best_nm, uint64_t *class_offsets, uint32_t *class_asv, size_t class_cap,
                               uint64_t *unambig, uint64_t *ambig, uint64_t *leq10, uint64_t *totals, uint64_t *sample_singleton,
                               savont_b200_stats *stats) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    TRY(ensure_index(c));
    if (!status || !best_nm || !class_offsets) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    if (n_samples < 1) return fail(SAVONT_B200_ERR_INVALID, "n_samples must be >= 1");
    const TwinInput in{dna_words, dna_word_offsets, lengths, snpmer_kmers, snpmer_kmer_bytes, snpmer_offsets, file_idx};
    TRY(validate_twin(in, n_reads));
    TRY(validate_file_idx(file_idx, n_reads, n_samples));
    const std::vector<size_t> cuts = twin_cuts(c, in, n_reads);
    auto up = [&](Batch &b, size_t r0, size_t n, uint64_t *bytes, cudaStream_t st) -> int {
        return batch_upload_twin(c, b, in, r0, n, bytes, st);
    };
    return map_batches(c, cuts, up, n_samples, status, best_nm, class_offsets, class_asv, class_cap, unambig, ambig, leq10, totals,
                       sample_singleton, stats);
}

// ---- TwinReads as they really lie in the reference's memory: a Vec<TwinRead> is an array of structs whose dna_seq / snpmer_kmers
// are separate heap allocations (src/types.rs:386-412).  The caller passes one pointer per read; per batch the library gathers the
// buffers with host threads into page-locked staging (three slots, one per batch in flight) and uploads from there, so the gather of
// batch i+1 runs while batch i computes.
struct PtrInput {
    const uint64_t *const *dna_ptrs; const uint32_t *lengths; const void *const *snp_ptrs; const uint32_t *snp_counts; int kmer_bytes;
    const uint32_t *file_idx;
};
static int gather_batch(Context *c, int slot, const PtrInput &in, size_t r0, size_t n, TwinInput *out, TwinPrep *prep, double *gather_ms) {
    const auto t0 = std::chrono::steady_clock::now();
    const size_t kb = (size_t)in.kmer_bytes;
    int T = c->host_threads > 0 ? c->host_threads : (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    if ((size_t)T > n / 2048 + 1) T = (int)(n / 2048 + 1);
    // The batch's offset arrays (what batch_upload_twin would compute) are built by the same threads that copy: pass A sums words /
    // SNPmers / bases per chunk of 512 reads, a serial scan over the (few hundred) chunks places the chunks, pass B fills the per-read
    // arrays of a chunk and copies its blocks.  Nothing per read runs on one thread.
    const size_t chunk = 512, n_chunks = (n + chunk - 1) / chunk;
    std::vector<uint64_t> cw(n_chunks + 1, 0), cs(n_chunks + 1, 0), cl(n_chunks + 1, 0);
    std::vector<uint32_t> cmax(n_chunks, 0);
    std::atomic<size_t> next_chunk{0};
    const std::function<void(int)> pass_a = [&](int) {
        while (true) {
            const size_t ch = next_chunk.fetch_add(1);
            if (ch >= n_chunks) break;
            const size_t a = ch * chunk, b = std::min(n, a + chunk);
            uint64_t w = 0, sn = 0, l = 0; uint32_t mx = 0;
            for (size_t i = a; i < b; ++i) {
                const uint32_t len = in.lengths[r0 + i];
                w += ((uint64_t)len + 31u) / 32u; sn += in.snp_counts[r0 + i]; l += len;
                if (len > mx) mx = len;
            }
            cw[ch + 1] = w; cs[ch + 1] = sn; cl[ch + 1] = l; cmax[ch] = mx;
        }
    };
    c->pool.run(T, pass_a);
    uint32_t max_len = 0;
    for (size_t ch = 0; ch < n_chunks; ++ch) { cw[ch + 1] += cw[ch]; cs[ch + 1] += cs[ch]; cl[ch + 1] += cl[ch]; if (cmax[ch] > max_len) max_len = cmax[ch]; }
    const uint64_t n_words = cw[n_chunks], n_snp = cs[n_chunks], total = cl[n_chunks];
    if (total >= (1ull << 32) - 64) return fail(SAVONT_B200_ERR_INVALID, "batch of %llu bases exceeds 4 Gi", (unsigned long long)total);
    if (n_snp >= (1ull << 32)) return fail(SAVONT_B200_ERR_INVALID, "batch of %llu SNPmers is too large", (unsigned long long)n_snp);
    TRY(c->gather_words[slot].ensure((n_words + 2) * 8));
    TRY(c->gather_snp[slot].ensure(n_snp * kb + 16));
    TRY(c->gather_meta[slot].ensure((n + 1) * 16 + n * 8 + 64));
    uint64_t *dw = c->gather_words[slot].as<uint64_t>();
    uint8_t *ds = c->gather_snp[slot].as<uint8_t>();
    uint64_t *rel = c->gather_meta[slot].as<uint64_t>();
    uint64_t *wrel = rel + (n + 1);
    uint32_t *sstart = reinterpret_cast<uint32_t *>(wrel + (n + 1));
    uint32_t *scnt = sstart + n;
    rel[n] = total; wrel[n] = n_words;
    std::atomic<int> bad{0};
    // tuning knobs (measured on the B200 hosts, tools/trace_e2e.py, profiles/r02i_gather_sweep.txt): prefetch distance in reads,
    // prefetch every line of a block or only its first, non-temporal stores for the words
    auto env_int = [](const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; };
    const size_t D = (size_t)std::max(0, env_int("SAVONT_B200_GATHER_PF", 8));
    const bool pf_all = env_int("SAVONT_B200_GATHER_PF_ALL", 1) != 0, nt = env_int("SAVONT_B200_GATHER_NT", 1) != 0;
    next_chunk.store(0);
    const std::function<void(int)> pass_b = [&](int) {
        while (true) {
            const size_t ch = next_chunk.fetch_add(1);
            if (ch >= n_chunks) break;
            const size_t a = ch * chunk, b = std::min(n, a + chunk);
            uint64_t w = cw[ch], sn = cs[ch], l = cl[ch];
            for (size_t i = a; i < b; ++i) {
                // the blocks of read i + D are prefetched while read i is copied (they are scattered over the heap: without the
                // prefetch every copy starts with a DRAM miss)
                if (D && i + D < b) {
                    const char *pw = reinterpret_cast<const char *>(in.dna_ptrs[r0 + i + D]);
                    const char *ps = reinterpret_cast<const char *>(in.snp_ptrs[r0 + i + D]);
                    const size_t bw = pf_all ? (((size_t)in.lengths[r0 + i + D] + 31u) / 32u) * 8 : 1, bs = pf_all ? (size_t)in.snp_counts[r0 + i + D] * kb : 1;
                    if (pw) for (size_t o = 0; o < bw; o += 64) __builtin_prefetch(pw + o, 0, 0);
                    if (ps) for (size_t o = 0; o < bs; o += 64) __builtin_prefetch(ps + o, 0, 0);
                }
                const uint32_t len = in.lengths[r0 + i];
                const uint64_t nw = ((uint64_t)len + 31u) / 32u, ns = in.snp_counts[r0 + i];
                rel[i] = l; wrel[i] = w; sstart[i] = (uint32_t)sn; scnt[i] = (uint32_t)ns;
                if ((nw && !in.dna_ptrs[r0 + i]) || (ns && !in.snp_ptrs[r0 + i])) bad.store(1);
                else {
                    // the staging buffer is read next by the DMA engine, not by a core: non-temporal stores keep it out of the
                    // caches and spare the read-for-ownership of every destination line
                    if (nw && nt) {
                        const long long *src = reinterpret_cast<const long long *>(in.dna_ptrs[r0 + i]);
                        long long *dst = reinterpret_cast<long long *>(dw + w);
                        for (uint64_t q = 0; q < nw; ++q) _mm_stream_si64(dst + q, src[q]);
                    } else if (nw) memcpy(dw + w, in.dna_ptrs[r0 + i], nw * 8);
                    if (ns) memcpy(ds + sn * kb, in.snp_ptrs[r0 + i], ns * kb);
                }
                w += nw; sn += ns; l += len;
            }
            _mm_sfence();
        }
    };
    c->pool.run(T, pass_b);
    if (bad.load()) return fail(SAVONT_B200_ERR_INVALID, "null read buffer in batch starting at read %zu", r0);
    *out = TwinInput{dw, nullptr, in.lengths + r0, ds, in.kmer_bytes, nullptr, in.file_idx ? in.file_idx + r0 : nullptr};
    prep->rel = rel; prep->wrel = wrel; prep->sstart = sstart; prep->scnt = scnt;
    prep->total = total; prep->n_words = n_words; prep->n_snp = n_snp; prep->w0 = 0; prep->s0 = 0; prep->max_len = max_len; prep->n = n;
    *gather_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

int savont_b200_set_host_threads(savont_b200_ctx *h, int n_threads) {
    if (!h) return fail(SAVONT_B200_ERR_INVALID, "null context");
    if (n_threads < 0 || n_threads > 1024) return fail(SAVONT_B200_ERR_INVALID, "n_threads must be in 0..1024");
    h->c.host_threads = n_threads;
    return 0;
}

int savont_b200_map_twin_read_ptrs(savont_b200_ctx *h, const uint64_t *const *dna_ptrs, const uint32_t *lengths,
                                   const void *const *snpmer_ptrs, const uint32_t *snpmer_counts, int snpmer_kmer_bytes,
                                   const uint32_t *file_idx, size_t n_reads, int n_samples, uint8_t *status, int32_t *best_nm,
                                   uint64_t *class_offsets, uint32_t *class_asv, size_t class_cap, uint64_t *unambig, uint64_t *ambig,
                                   uint64_t *leq10, uint64_t *totals, uint64_t *sample_singleton, savont_b200_stats *stats) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    TRY(ensure_index(c));
    if (!status || !best_nm || !class_offsets) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    if (n_reads && (!dna_ptrs || !lengths || !snpmer_ptrs || !snpmer_counts)) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    if (snpmer_kmer_bytes != 6 && snpmer_kmer_bytes != 8) return fail(SAVONT_B200_ERR_INVALID, "snpmer_kmer_bytes must be 6 (Kmer48) or 8 (u64)");
    if (n_samples < 1) return fail(SAVONT_B200_ERR_INVALID, "n_samples must be >= 1");
    TRY(validate_file_idx(file_idx, n_reads, n_samples));
    const PtrInput in{dna_ptrs, lengths, snpmer_ptrs, snpmer_counts, snpmer_kmer_bytes, file_idx};
    std::vector<size_t> cuts{0};
    for (size_t r0 = 0; r0 < n_reads;) {
        size_t r1 = std::min(n_reads, r0 + ramp_batch(c, cuts.size() - 1, n_reads - r0));
        uint64_t bases = 0;
        for (size_t r = r0; r < r1; ++r) { bases += lengths[r]; if (bases > (3ull << 30)) { r1 = std::max(r, r0 + 1); break; } }
        cuts.push_back(r1);
        r0 = r1;
    }
    // A producer thread gathers batch after batch into the four staging slots (its copies run on the context's host pool) and runs
    // ahead of the consumer, which uploads a batch as soon as it is gathered; a slot is handed back when its batch has finished.
    double gather_ms = 0.0;
    struct Pipe {
        std::mutex mu; std::condition_variable cv;
        size_t gathered = 0, released = 0; int rc = 0; std::string err; bool abort = false;
    } pipe;
    const size_t nb = cuts.size() - 1;
    constexpr size_t kRing = 8;            // staging slots: the producer may run four batches ahead of the four the pipeline holds
    std::vector<TwinInput> staged(kRing); std::vector<TwinPrep> preps(kRing);
    std::thread producer([&]() {
        cudaSetDevice(c->device);                            // the staging buffers are page-locked through this device's context
        for (size_t b = 0; b < nb; ++b) {
            {
                std::unique_lock<std::mutex> lk(pipe.mu);
                pipe.cv.wait(lk, [&] { return pipe.abort || b < pipe.released + kRing; });
                if (pipe.abort) return;
            }
            TwinInput ti{}; TwinPrep tp{};
            TRACE("gather(%zu) begin (%zu reads)", b, cuts[b + 1] - cuts[b]);
            const int rc = gather_batch(c, (int)(b % kRing), in, cuts[b], cuts[b + 1] - cuts[b], &ti, &tp, &gather_ms);
            TRACE("gather(%zu) end", b);
            std::unique_lock<std::mutex> lk(pipe.mu);
            if (rc != 0) { pipe.rc = rc; pipe.err = savont_b200_last_error(); pipe.cv.notify_all(); return; }
            staged[b % kRing] = ti; preps[b % kRing] = tp; pipe.gathered = b + 1;
            pipe.cv.notify_all();
        }
    });
    size_t next_up = 0;                    // batches are uploaded in order
    auto up = [&](Batch &b, size_t /*r0*/, size_t /*n*/, uint64_t *bytes, cudaStream_t st) -> int {
        const size_t bi = next_up++;
        {
            std::unique_lock<std::mutex> lk(pipe.mu);
            pipe.cv.wait(lk, [&] { return pipe.rc != 0 || pipe.gathered > bi; });
            if (pipe.rc != 0) return fail(pipe.rc, "%s", pipe.err.c_str());
        }
        return issue_twin(c, b, staged[bi % kRing], 0, preps[bi % kRing], bytes, st);
    };
    auto on_finished = [&](size_t bi) {
        std::unique_lock<std::mutex> lk(pipe.mu);
        pipe.released = bi + 1;
        pipe.cv.notify_all();
    };
    const int rc = map_batches(c, cuts, up, n_samples, status, best_nm, class_offsets, class_asv, class_cap, unambig, ambig, leq10, totals,
                               sample_singleton, stats, on_finished);
    {
        std::unique_lock<std::mutex> lk(pipe.mu);
        pipe.abort = true;
        pipe.cv.notify_all();
    }
    producer.join();
    c->last_stats.gather_ms = gather_ms;
    if (stats) stats->gather_ms = gather_ms;
    return rc;
}

int savont_b200_upload_twin_reads(savont_b200_ctx *h, const uint64_t *dna_words, const uint64_t *dna_word_offsets,
                                  const uint32_t *lengths, const void *snpmer_kmers, int snpmer_kmer_bytes,
                                  const uint64_t *snpmer_offsets, const uint32_t *file_idx, size_t n_reads, int n_samples) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (n_samples < 1) return fail(SAVONT_B200_ERR_INVALID, "n_samples must be >= 1");
    const TwinInput in{dna_words, dna_word_offsets, lengths, snpmer_kmers, snpmer_kmer_bytes, snpmer_offsets, file_idx};
    TRY(validate_twin(in, n_reads));
    TRY(validate_file_idx(file_idx, n_reads, n_samples));
    TRY(batch_upload_twin(c, c->resident, in, 0, n_reads, nullptr));
    CK(cudaStreamSynchronize(c->stream));
    c->resident_samples = n_samples;
    return build_resident_parts(c);
}

int savont_b200_upload_reads(savont_b200_ctx *h, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets,
                             const uint32_t *file_idx, size_t n_reads, int n_samples) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!offsets || (n_reads && !bases)) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    if (n_samples < 1) return fail(SAVONT_B200_ERR_INVALID, "n_samples must be >= 1");
    TRY(validate_file_idx(file_idx, n_reads, n_samples));
    TRY(batch_upload(c, c->resident, bases, quals, offsets, file_idx, n_reads, nullptr));
    CK(cudaStreamSynchronize(c->stream));
    c->resident_samples = n_samples;
    return build_resident_parts(c);
}

int savont_b200_set_resident_parts(savont_b200_ctx *h, int n_parts) {
    if (!h) return fail(SAVONT_B200_ERR_INVALID, "null context");
    if (n_parts < 0 || n_parts > 4096) return fail(SAVONT_B200_ERR_INVALID, "n_parts must be in 0..4096");
    h->c.resident_parts_req = n_parts;
    return 0;
}

int savont_b200_map_resident(savont_b200_ctx *h, int zero_counters, savont_b200_stats *stats) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    TRY(ensure_index(c));
    savont_b200_stats st; memset(&st, 0, sizeof(st));
    memset(&c->last_stats, 0, sizeof(c->last_stats));
    TRY(ensure_counters(c, c->resident_samples, zero_counters != 0));
    CK(cudaEventRecord(c->ev[6], c->stream));
    if (c->res_parts.size() <= 1) {
        TRY(run_path(c, c->resident, c->resident_samples, st));
        st.n_batches = 1;
    } else {
        // part i runs on lane i & 1: its sketch / filter kernels are enqueued while the DP kernel of part i-1 still runs on the other
        // lane (the DP grid leaves room on every SM, launch_align_nm); one host synchronisation per part, behind the enqueue of
        // the next one
        for (int l = 0; l < 2; ++l) CK(cudaStreamWaitEvent(c->lane_stream[l], c->ev[6], 0));
        int prev = -1;
        for (size_t i = 0; i < c->res_parts.size(); ++i) {
            { LaneGuard g(c, (int)(i & 1)); TRY(enqueue_path(c, c->res_parts[i], c->resident_samples)); }
            if (prev >= 0) { LaneGuard g(c, prev & 1); TRY(finish_path(c, c->res_parts[(size_t)prev], c->resident_samples, st)); }
            prev = (int)i;
        }
        { LaneGuard g(c, prev & 1); TRY(finish_path(c, c->res_parts[(size_t)prev], c->resident_samples, st)); }
        st.n_batches = c->res_parts.size();
        c->resident.mapped = true;
    }
    CK(cudaEventRecord(c->ev[7], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    st.total_ms = ev_ms(c->ev[6], c->ev[7]);
    st.kernel_launches = c->last_stats.kernel_launches;
    c->last_stats = st;
    if (stats) *stats = st;
    return 0;
}

int savont_b200_fetch_resident(savont_b200_ctx *h, uint8_t *status, int32_t *best_nm, uint64_t *class_offsets,
                               uint32_t *class_asv, size_t class_cap) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!c->resident.mapped) return fail(SAVONT_B200_ERR_INVALID, "map_resident has not run");
    if (!status || !best_nm || !class_offsets) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    uint64_t n_cls = 0;
    if (c->res_parts.size() <= 1)
        return fetch_batch(c, c->resident, c->resident.n, c->stream, status, best_nm, class_offsets, 0, class_asv, class_cap, &n_cls, nullptr);
    uint64_t base = 0;
    class_offsets[0] = 0;
    for (size_t q = 0; q < c->res_parts.size(); ++q) {
        const size_t r0 = c->res_cuts[q];
        TRY(fetch_batch(c, c->res_parts[q], c->res_parts[q].n, c->stream, status + r0, best_nm + r0, class_offsets + r0, base, class_asv,
                        class_cap, &n_cls, nullptr));
        base += n_cls;
    }
    return 0;
}

// The rows of temp/read_to_asv_mappings.tsv for one mapped batch (src/alignment.rs:1866, :1875-1887): a read's aligned candidates,
// stable-sorted by NM, the first max_rows of them.  Small per-pair arrays come to the host and are grouped there.
static int mapping_rows_batch(Context *c, Batch &b, uint32_t max_rows, uint64_t *row_offsets /* [n + 1], [0] set */, uint32_t *row_asv,
                              int32_t *row_nm, uint32_t *row_mm, size_t row_cap) {
    const size_t n = b.n;
    if (n == 0) return 0;
    std::vector<uint32_t> start(n), cnt(n), low(n);
    std::vector<uint8_t> st(n);
    CK(cudaMemcpyAsync(start.data(), b.pair_start.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(cnt.data(), b.pair_cnt.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(low.data(), b.low_mm.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(st.data(), b.cand_status.p, n, cudaMemcpyDeviceToHost, c->stream));
    uint32_t np = 0;
    CK(cudaMemcpyAsync(&np, b.ctrs.as<uint32_t>() + 1, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::vector<uint32_t> pasv(np);
    std::vector<int32_t> pnm(np);
    if (np) {
        CK(cudaMemcpyAsync(pasv.data(), b.p_asv.p, (size_t)np * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(pnm.data(), b.nm.p, (size_t)np * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    const uint32_t per = c->dp.strand_mode == 1 ? 2u : 1u;
    std::vector<std::pair<int32_t, uint32_t>> rows;
    uint64_t o = row_offsets[0];
    for (size_t r = 0; r < n; ++r) {
        rows.clear();
        if (st[r] == 0) {
            if ((uint64_t)start[r] + cnt[r] > np) return fail(SAVONT_B200_ERR_CUDA, "internal: pair chunk of read %zu outside the pool", r);
            for (uint32_t q = 0; q < cnt[r] / per; ++q) {
                int32_t v = pnm[start[r] + q * per];
                if (per == 2u) v = std::min(v, pnm[start[r] + q * per + 1]);
                rows.emplace_back(v, pasv[start[r] + q * per]);
            }
            std::stable_sort(rows.begin(), rows.end(), [](const std::pair<int32_t, uint32_t> &x, const std::pair<int32_t, uint32_t> &y) { return x.first < y.first; });
        }
        const size_t k = std::min<size_t>(rows.size(), max_rows);
        if (o + k > row_cap) return fail(SAVONT_B200_ERR_CAPACITY, "mapping rows: capacity %zu too small (at read %zu)", row_cap, r);
        for (size_t i = 0; i < k; ++i) { row_asv[o + i] = rows[i].second; row_nm[o + i] = rows[i].first; row_mm[o + i] = low[r]; }
        o += k;
        row_offsets[r + 1] = o;
    }
    return 0;
}

int savont_b200_mapping_rows_resident(savont_b200_ctx *h, uint32_t max_rows_per_read, uint64_t *row_offsets, uint32_t *row_asv, int32_t *row_nm,
                                      uint32_t *row_snp_mismatches, size_t row_cap) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!c->resident.mapped) return fail(SAVONT_B200_ERR_INVALID, "map_resident has not run");
    if (!row_offsets || (row_cap && (!row_asv || !row_nm || !row_snp_mismatches))) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    row_offsets[0] = 0;
    if (c->res_parts.size() <= 1) return mapping_rows_batch(c, c->resident, max_rows_per_read, row_offsets, row_asv, row_nm, row_snp_mismatches, row_cap);
    for (size_t q = 0; q < c->res_parts.size(); ++q)
        TRY(mapping_rows_batch(c, c->res_parts[q], max_rows_per_read, row_offsets + c->res_cuts[q], row_asv, row_nm, row_snp_mismatches, row_cap));
    return 0;
}

// best hit of every read of one mapped batch + its traceback; the CIGAR ops are appended to `cigar` behind *cig_base
static int best_hits_batch(Context *c, Batch &b, uint32_t *best_asv, uint8_t *strand, int32_t *asv_start, int32_t *asv_end,
                           int32_t *read_start, int32_t *read_end, double *identity, uint64_t *cigar_offsets /* [n + 1], [0] set */,
                           uint32_t *cigar, size_t cigar_cap) {
    const size_t n = b.n;
    if (n == 0) return 0;
    DevBuf &d_flag = c->bh_ws[0], &d_off = c->bh_ws[1], &d_tmp = c->bh_ws[2], &d_basv = c->bh_ws[3], &d_bstr = c->bh_ws[4], &d_pr = c->bh_ws[5],
           &d_pa = c->bh_ws[6], &d_ps = c->bh_ws[7];
    auto cleanup = [&]() {};                     // the workspace stays with the context (released by savont_b200_destroy)
    int rc = 0;
    if ((rc = d_flag.ensure(n * 4)) || (rc = d_off.ensure((n + 1) * 4)) || (rc = d_tmp.ensure((n / 2048 + 4) * 4)) ||
        (rc = d_basv.ensure(n * 4)) || (rc = d_bstr.ensure(n + 4)) || (rc = d_pr.ensure(n * 4)) || (rc = d_pa.ensure(n * 4)) ||
        (rc = d_ps.ensure(n + 4))) { cleanup(); return rc; }
    const auto bh_t0 = std::chrono::steady_clock::now();
    launch_best_pairs(filter_view(b), pool_view(b), b.nm.as<int32_t>(), b.rstatus.as<uint8_t>(), b.best_nm.as<int32_t>(), (uint32_t)n,
                      c->dp.strand_mode, d_flag.as<uint32_t>(), d_off.as<uint32_t>(), d_tmp.as<uint32_t>(), d_basv.as<uint32_t>(),
                      d_bstr.as<uint8_t>(), d_pr.as<uint32_t>(), d_pa.as<uint32_t>(), d_ps.as<uint8_t>(), c->stream);
    c->last_stats.kernel_launches += 5;
    std::vector<uint32_t> &off = c->bh_off;
    if (off.size() < n + 1) off.resize(n + 1);
    cudaMemcpyAsync(off.data(), d_off.p, (n + 1) * 4, cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(best_asv, d_basv.p, n * 4, cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(strand, d_bstr.p, n, cudaMemcpyDeviceToHost, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { cleanup(); return fail(SAVONT_B200_ERR_CUDA, "best_pairs: %s", cudaGetErrorString(e)); }
    const size_t np = off[n];
    std::vector<int32_t> &v = c->bh_rec;
    std::vector<uint64_t> &coff = c->bh_coff;
    if (v.size() < np * 5) v.resize(np * 5);
    if (coff.size() < np + 1) coff.resize(np + 1);
    coff[0] = cigar_offsets[0];
    if (np) {
        rc = traceback_pairs(c, b, index_view(c), c->max_asv_len, d_pr.as<uint32_t>(), d_pa.as<uint32_t>(), d_ps.as<uint8_t>(), np, v.data(), v.data() + np,
                             v.data() + 2 * np, v.data() + 3 * np, v.data() + 4 * np, coff.data(), cigar, cigar_cap);
        if (rc) { cleanup(); return rc; }
    }
    for (size_t r = 0; r < n; ++r) {
        if (off[r + 1] == off[r]) {
            asv_start[r] = asv_end[r] = read_start[r] = read_end[r] = -1;
            identity[r] = std::nan("");
        } else {
            const size_t t = off[r];
            asv_end[r] = v[np + t]; read_end[r] = v[2 * np + t]; asv_start[r] = v[3 * np + t]; read_start[r] = v[4 * np + t];
            // the reference's identity with the ASV as the query: 1 - nm / (query_end - query_start), src/chimera.rs:444-452
            const double span = (double)(asv_end[r] - asv_start[r]);
            identity[r] = span > 0.0 ? 1.0 - (double)v[t] / span : 0.0;
        }
        cigar_offsets[r + 1] = coff[off[r + 1]];
    }
    cleanup();
    if (getenv("SAVONT_B200_TRACE")) fprintf(stderr, "[trace] best_hits_batch: %zu reads, %.2f ms in all\n", n, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - bh_t0).count());
    return 0;
}

int savont_b200_best_hits_resident(savont_b200_ctx *h, uint32_t *best_asv, uint8_t *strand, int32_t *asv_start, int32_t *asv_end,
                                   int32_t *read_start, int32_t *read_end, double *identity, uint64_t *cigar_offsets,
                                   uint32_t *cigar, size_t cigar_cap) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!c->resident.mapped) return fail(SAVONT_B200_ERR_INVALID, "map_resident has not run");
    if (!best_asv || !strand || !asv_start || !asv_end || !read_start || !read_end || !identity || !cigar_offsets)
        return fail(SAVONT_B200_ERR_INVALID, "null argument");
    cigar_offsets[0] = 0;
    if (c->res_parts.size() <= 1)
        return best_hits_batch(c, c->resident, best_asv, strand, asv_start, asv_end, read_start, read_end, identity, cigar_offsets, cigar, cigar_cap);
    for (size_t q = 0; q < c->res_parts.size(); ++q) {
        const size_t r0 = c->res_cuts[q];
        TRY(best_hits_batch(c, c->res_parts[q], best_asv + r0, strand + r0, asv_start + r0, asv_end + r0, read_start + r0, read_end + r0,
                            identity + r0, cigar_offsets + r0, cigar, cigar_cap));
    }
    return 0;
}

int savont_b200_counts_dev(savont_b200_ctx *h, void **dev_ptr, size_t *n_u64) {
    if (!h || !dev_ptr || !n_u64) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    if (h->c.counters_len == 0) return fail(SAVONT_B200_ERR_INVALID, "no counters: nothing was mapped yet");
    *dev_ptr = h->c.d_counters.p;
    *n_u64 = h->c.counters_len;
    return 0;
}

int savont_b200_read_counts(savont_b200_ctx *h, uint64_t *unambig, uint64_t *ambig, uint64_t *leq10, uint64_t *totals,
                            uint64_t *sample_singleton) {
    TRY(use_device(h ? &h->c : nullptr));
    return read_counters(&h->c, unambig, ambig, leq10, totals, sample_singleton, false);
}

int savont_b200_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    CK(cudaMallocHost(out, bytes ? bytes : 1));
    return 0;
}
int savont_b200_host_free(void *p) {
    if (p) CK(cudaFreeHost(p));
    return 0;
}

}  // extern "C"
