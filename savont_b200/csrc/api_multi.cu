// api_multi.cu -- one process, several GPUs: the handle SURVEY.md 8(b) asks for (`savont_b200_create(device_ids, n_dev, ..)`).
//
// `savont asv` is ONE process with a rayon pool over the reads (reference src/alignment.rs:1786, src/main.rs:476-480); the drop-in
// for it on an 8-GPU box is therefore one handle that owns a context per device, shards the reads of a call into contiguous
// ranges (read order is preserved, src/main.rs:538), runs every shard on its device from its own host thread, reduces the
// counter vector [unambig | ambig | leq10 | assigned | filtered | singleton S x A] across the devices, and hands back the per-read
// results concatenated in read order.  The ASV set, the SNPmer set and the index are replicated.  There is no exchange step on
// the path; the one collective is the sum of the counter vectors (<= 0.9 MB):
//   mode 0 (default)  one kernel on device 0 that reads every device's vector through peer memory (NVLink / NVSwitch), sums, and
//                     writes the sum back into all of them -- transfer and arithmetic in the same loop, no staging;
//   mode 1            ncclAllReduce (ncclCommInitAll, grouped launch), NCCL loaded with dlopen so that the library has no link-time
//                     dependency on it.
// Built on the public per-context ABI only (savont_b200.h); nothing here touches a context's internals.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "api_internal.cuh"

using namespace sb;

namespace {

constexpr int kMaxDev = 16;
struct PeerVecs { unsigned long long *p[kMaxDev]; int n; };

// every element: sum over the devices' vectors, written back to all of them (device 0 runs it; the peers are mapped)
__global__ void peer_allreduce_u64_kernel(PeerVecs v, size_t len) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long s = 0;
        for (int d = 0; d < v.n; ++d) s += v.p[d][i];
        for (int d = 0; d < v.n; ++d) v.p[d][i] = s;
    }
}

struct Nccl {
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(lib, "ncclAllReduce"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd && GetErrorString;
    }
};

}  // namespace

struct savont_b200_multi {
    int n = 0;
    std::vector<int> dev;
    std::vector<savont_b200_ctx *> ctx;
    std::vector<cudaStream_t> stream;            // one per device, for the collective
    bool distinct = true;                        // no device appears twice (duplicates: several contexts on one GPU, tests)
    int allreduce_mode = 0;
    Nccl nccl; std::vector<ncclComm_t> comms;
    int host_threads = 0;
    // per-device scratch of a map call
    struct Shard { std::vector<uint64_t> class_offsets; std::vector<uint32_t> class_asv; savont_b200_stats st; int rc = 0; std::string err; };
    std::vector<Shard> shard;
};

namespace {

int multi_fail_from(savont_b200_multi *m) {
    for (int d = 0; d < m->n; ++d)
        if (m->shard[d].rc != 0) return fail(m->shard[d].rc, "device %d: %s", m->dev[d], m->shard[d].err.c_str());
    return 0;
}

// runs f(d) for every device on its own host thread; errors (code + thread-local message) are collected per shard
template <typename F>
int for_each_device(savont_b200_multi *m, F f) {
    std::vector<std::thread> th;
    for (int d = 0; d < m->n; ++d) {
        m->shard[d].rc = 0; m->shard[d].err.clear();
        th.emplace_back([m, d, &f]() {
            const int rc = f(d);
            if (rc != 0) { m->shard[d].rc = rc; m->shard[d].err = savont_b200_last_error(); }
        });
    }
    for (auto &t : th) t.join();
    return multi_fail_from(m);
}

int allreduce_counts(savont_b200_multi *m) {
    if (m->n == 1) return 0;
    PeerVecs v{}; v.n = m->n;
    size_t len = 0;
    for (int d = 0; d < m->n; ++d) {
        void *p = nullptr; size_t l = 0;
        TRY(savont_b200_counts_dev(m->ctx[d], &p, &l));
        if (d > 0 && l != len) return fail(SAVONT_B200_ERR_INVALID, "counter vectors differ in length across devices");
        len = l; v.p[d] = static_cast<unsigned long long *>(p);
    }
    if (m->allreduce_mode == 1) {
        if (!m->distinct) return fail(SAVONT_B200_ERR_INVALID, "the NCCL all-reduce needs distinct devices");
        ncclResult_t r = m->nccl.GroupStart();
        for (int d = 0; d < m->n && r == ncclSuccess; ++d) {
            CK(cudaSetDevice(m->dev[d]));
            r = m->nccl.AllReduce(v.p[d], v.p[d], len, ncclUint64, ncclSum, m->comms[d], m->stream[d]);
        }
        if (r == ncclSuccess) r = m->nccl.GroupEnd();
        if (r != ncclSuccess) return fail(SAVONT_B200_ERR_CUDA, "ncclAllReduce: %s", m->nccl.GetErrorString(r));
        for (int d = 0; d < m->n; ++d) { CK(cudaSetDevice(m->dev[d])); CK(cudaStreamSynchronize(m->stream[d])); }
        return 0;
    }
    CK(cudaSetDevice(m->dev[0]));
    const unsigned grid = (unsigned)std::min<size_t>((len + 255) / 256, 148 * 4);
    peer_allreduce_u64_kernel<<<grid ? grid : 1, 256, 0, m->stream[0]>>>(v, len);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(m->stream[0]));
    return 0;
}

void add_stats(savont_b200_stats &a, const savont_b200_stats &b) {
    a.sketch_ms = std::max(a.sketch_ms, b.sketch_ms); a.filter_ms = std::max(a.filter_ms, b.filter_ms);
    a.align_ms = std::max(a.align_ms, b.align_ms); a.assign_ms = std::max(a.assign_ms, b.assign_ms);
    a.h2d_ms = std::max(a.h2d_ms, b.h2d_ms); a.d2h_ms = std::max(a.d2h_ms, b.d2h_ms); a.total_ms = std::max(a.total_ms, b.total_ms);
    a.gather_ms = std::max(a.gather_ms, b.gather_ms);
    a.n_reads += b.n_reads; a.n_bases += b.n_bases; a.n_pairs += b.n_pairs; a.dp_cells += b.dp_cells;
    a.h2d_bytes += b.h2d_bytes; a.d2h_bytes += b.d2h_bytes; a.kernel_launches += b.kernel_launches; a.n_batches += b.n_batches;
}

// shard d of n reads: [n * d / N, n * (d + 1) / N)
inline size_t cut(size_t n, int d, int N) { return n * (size_t)d / (size_t)N; }

// after the per-device maps: counters reduced and read from device 0, classes concatenated in read order
int finish_map(savont_b200_multi *m, size_t n_reads, uint64_t *class_offsets, uint32_t *class_asv, size_t class_cap, uint64_t *unambig,
               uint64_t *ambig, uint64_t *leq10, uint64_t *totals, uint64_t *sample_singleton, savont_b200_stats *stats) {
    TRY(allreduce_counts(m));
    TRY(savont_b200_read_counts(m->ctx[0], unambig, ambig, leq10, totals, sample_singleton));
    uint64_t base = 0;
    class_offsets[0] = 0;
    for (int d = 0; d < m->n; ++d) {
        const size_t r0 = cut(n_reads, d, m->n), cnt = cut(n_reads, d + 1, m->n) - r0;
        const savont_b200_multi::Shard &sh = m->shard[d];
        const uint64_t tot = cnt ? sh.class_offsets[cnt] : 0;
        if (base + tot > class_cap) return fail(SAVONT_B200_ERR_CAPACITY, "class_asv capacity %zu too small (need >= %llu)", class_cap, (unsigned long long)(base + tot));
        for (size_t i = 1; i <= cnt; ++i) class_offsets[r0 + i] = base + sh.class_offsets[i];
        if (tot && class_asv) memcpy(class_asv + base, sh.class_asv.data(), tot * sizeof(uint32_t));
        base += tot;
    }
    if (stats) {
        savont_b200_stats st; memset(&st, 0, sizeof(st));
        for (int d = 0; d < m->n; ++d) add_stats(st, m->shard[d].st);
        *stats = st;
    }
    return 0;
}

}  // namespace

extern "C" {

int savont_b200_create_multi(const int *device_ids, int n_dev, savont_b200_multi **out) {
    if (!out) return fail(SAVONT_B200_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!device_ids || n_dev < 1 || n_dev > kMaxDev) return fail(SAVONT_B200_ERR_INVALID, "n_dev must be in 1..%d", kMaxDev);
    savont_b200_multi *m = new savont_b200_multi();
    m->n = n_dev; m->dev.assign(device_ids, device_ids + n_dev);
    m->ctx.assign(n_dev, nullptr); m->stream.assign(n_dev, nullptr); m->shard.resize(n_dev);
    for (int a = 0; a < n_dev; ++a) for (int b = a + 1; b < n_dev; ++b) if (m->dev[a] == m->dev[b]) m->distinct = false;
    int rc = 0;
    for (int d = 0; d < n_dev && rc == 0; ++d) {
        rc = savont_b200_create(m->dev[d], &m->ctx[d]);
        if (rc == 0 && cudaStreamCreateWithFlags(&m->stream[d], cudaStreamNonBlocking) != cudaSuccess)
            rc = fail(SAVONT_B200_ERR_CUDA, "cudaStreamCreate on device %d failed", m->dev[d]);
    }
    // device 0 reads and writes the peers' counter vectors in the all-reduce kernel
    for (int d = 1; d < n_dev && rc == 0; ++d) {
        if (m->dev[d] == m->dev[0]) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, m->dev[0], m->dev[d]);
        if (!can) { rc = fail(SAVONT_B200_ERR_CUDA, "device %d cannot access device %d as a peer", m->dev[0], m->dev[d]); break; }
        cudaSetDevice(m->dev[0]);
        const cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[d], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = fail(SAVONT_B200_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
    }
    if (rc != 0) { savont_b200_multi_destroy(m); return rc; }
    *out = m;
    return 0;
}

int savont_b200_multi_destroy(savont_b200_multi *m) {
    if (!m) return 0;
    for (size_t d = 0; d < m->comms.size(); ++d) if (m->comms[d]) m->nccl.CommDestroy(m->comms[d]);
    for (int d = 0; d < m->n; ++d) {
        if (m->stream[d]) { cudaSetDevice(m->dev[d]); cudaStreamDestroy(m->stream[d]); }
        if (m->ctx[d]) savont_b200_destroy(m->ctx[d]);
    }
    delete m;
    return 0;
}

int savont_b200_multi_num_devices(savont_b200_multi *m, int *n_dev) {
    if (!m || !n_dev) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    *n_dev = m->n;
    return 0;
}

int savont_b200_multi_ctx(savont_b200_multi *m, int i, savont_b200_ctx **ctx) {
    if (!m || !ctx || i < 0 || i >= m->n) return fail(SAVONT_B200_ERR_INVALID, "bad argument");
    *ctx = m->ctx[i];
    return 0;
}

int savont_b200_multi_set_allreduce(savont_b200_multi *m, int mode) {
    if (!m || (mode != 0 && mode != 1)) return fail(SAVONT_B200_ERR_INVALID, "mode must be 0 (peer-memory kernel) or 1 (NCCL)");
    if (mode == 1 && m->n > 1 && m->comms.empty()) {
        if (!m->distinct) return fail(SAVONT_B200_ERR_INVALID, "the NCCL all-reduce needs distinct devices");
        if (!m->nccl.load()) return fail(SAVONT_B200_ERR_CUDA, "libnccl.so.2 could not be loaded: %s", dlerror());
        m->comms.assign(m->n, nullptr);
        const ncclResult_t r = m->nccl.CommInitAll(m->comms.data(), m->n, m->dev.data());
        if (r != ncclSuccess) { m->comms.clear(); return fail(SAVONT_B200_ERR_CUDA, "ncclCommInitAll: %s", m->nccl.GetErrorString(r)); }
    }
    m->allreduce_mode = mode;
    return 0;
}

int savont_b200_multi_set_host_threads(savont_b200_multi *m, int n_threads) {
    if (!m || n_threads < 0) return fail(SAVONT_B200_ERR_INVALID, "bad argument");
    m->host_threads = n_threads;
    return 0;
}

int savont_b200_multi_set_params(savont_b200_multi *m, const savont_b200_params *p) {
    if (!m) return fail(SAVONT_B200_ERR_INVALID, "null handle");
    for (int d = 0; d < m->n; ++d) TRY(savont_b200_set_params(m->ctx[d], p));
    return 0;
}

int savont_b200_multi_set_batch_reads(savont_b200_multi *m, size_t n) {
    if (!m) return fail(SAVONT_B200_ERR_INVALID, "null handle");
    for (int d = 0; d < m->n; ++d) TRY(savont_b200_set_batch_reads(m->ctx[d], n));
    return 0;
}

int savont_b200_multi_load_snpmers(savont_b200_multi *m, const uint64_t *kmers, size_t n) {
    if (!m) return fail(SAVONT_B200_ERR_INVALID, "null handle");
    return for_each_device(m, [&](int d) { return savont_b200_load_snpmers(m->ctx[d], kmers, n); });
}

int savont_b200_multi_load_asvs(savont_b200_multi *m, const uint8_t *bases, const uint64_t *offsets, size_t n_asv) {
    if (!m) return fail(SAVONT_B200_ERR_INVALID, "null handle");
    return for_each_device(m, [&](int d) { return savont_b200_load_asvs(m->ctx[d], bases, offsets, n_asv); });
}

int savont_b200_multi_map_reads(savont_b200_multi *m, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets,
                                const uint32_t *file_idx, size_t n_reads, int n_samples, uint8_t *status, int32_t *best_nm,
                                uint64_t *class_offsets, uint32_t *class_asv, size_t class_cap, uint64_t *unambig, uint64_t *ambig,
                                uint64_t *leq10, uint64_t *totals, uint64_t *sample_singleton, savont_b200_stats *stats) {
    if (!m || !offsets || !status || !best_nm || !class_offsets) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    TRY(for_each_device(m, [&](int d) {
        const size_t r0 = cut(n_reads, d, m->n), cnt = cut(n_reads, d + 1, m->n) - r0;
        savont_b200_multi::Shard &sh = m->shard[d];
        sh.class_offsets.assign(cnt + 1, 0);
        const size_t cap = std::max<size_t>(class_cap, 16);
        if (sh.class_asv.size() < cap) sh.class_asv.resize(cap);
        return savont_b200_map_reads(m->ctx[d], bases, quals, offsets + r0, file_idx ? file_idx + r0 : nullptr, cnt, n_samples, status + r0,
                                     best_nm + r0, sh.class_offsets.data(), sh.class_asv.data(), cap, nullptr, nullptr, nullptr, nullptr,
                                     nullptr, &sh.st);
    }));
    return finish_map(m, n_reads, class_offsets, class_asv, class_cap, unambig, ambig, leq10, totals, sample_singleton, stats);
}

int savont_b200_multi_map_twin_reads(savont_b200_multi *m, const uint64_t *dna_words, const uint64_t *dna_word_offsets,
                                     const uint32_t *lengths, const void *snpmer_kmers, int snpmer_kmer_bytes,
                                     const uint64_t *snpmer_offsets, const uint32_t *file_idx, size_t n_reads, int n_samples,
                                     uint8_t *status, int32_t *best_nm, uint64_t *class_offsets, uint32_t *class_asv, size_t class_cap,
                                     uint64_t *unambig, uint64_t *ambig, uint64_t *leq10, uint64_t *totals, uint64_t *sample_singleton,
                                     savont_b200_stats *stats) {
    if (!m || !dna_word_offsets || !lengths || !snpmer_offsets || !status || !best_nm || !class_offsets)
        return fail(SAVONT_B200_ERR_INVALID, "null argument");
    TRY(for_each_device(m, [&](int d) {
        const size_t r0 = cut(n_reads, d, m->n), cnt = cut(n_reads, d + 1, m->n) - r0;
        savont_b200_multi::Shard &sh = m->shard[d];
        sh.class_offsets.assign(cnt + 1, 0);
        const size_t cap = std::max<size_t>(class_cap, 16);
        if (sh.class_asv.size() < cap) sh.class_asv.resize(cap);
        // the offset arrays are absolute into dna_words / snpmer_kmers: a shard is the same arrays entered at read r0
        return savont_b200_map_twin_reads(m->ctx[d], dna_words, dna_word_offsets + r0, lengths + r0, snpmer_kmers, snpmer_kmer_bytes,
                                          snpmer_offsets + r0, file_idx ? file_idx + r0 : nullptr, cnt, n_samples, status + r0, best_nm + r0,
                                          sh.class_offsets.data(), sh.class_asv.data(), cap, nullptr, nullptr, nullptr, nullptr, nullptr, &sh.st);
    }));
    return finish_map(m, n_reads, class_offsets, class_asv, class_cap, unambig, ambig, leq10, totals, sample_singleton, stats);
}

int savont_b200_multi_map_twin_read_ptrs(savont_b200_multi *m, const uint64_t *const *dna_ptrs, const uint32_t *lengths,
                                         const void *const *snpmer_ptrs, const uint32_t *snpmer_counts, int snpmer_kmer_bytes,
                                         const uint32_t *file_idx, size_t n_reads, int n_samples, uint8_t *status, int32_t *best_nm,
                                         uint64_t *class_offsets, uint32_t *class_asv, size_t class_cap, uint64_t *unambig,
                                         uint64_t *ambig, uint64_t *leq10, uint64_t *totals, uint64_t *sample_singleton,
                                         savont_b200_stats *stats) {
    if (!m || !status || !best_nm || !class_offsets) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    if (n_reads && (!dna_ptrs || !lengths || !snpmer_ptrs || !snpmer_counts)) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    // the gather threads of all devices share the host's cores
    int per_dev = m->host_threads > 0 ? m->host_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    per_dev = std::max(1, per_dev / m->n);
    for (int d = 0; d < m->n; ++d) TRY(savont_b200_set_host_threads(m->ctx[d], per_dev));
    TRY(for_each_device(m, [&](int d) {
        const size_t r0 = cut(n_reads, d, m->n), cnt = cut(n_reads, d + 1, m->n) - r0;
        savont_b200_multi::Shard &sh = m->shard[d];
        sh.class_offsets.assign(cnt + 1, 0);
        const size_t cap = std::max<size_t>(class_cap, 16);
        if (sh.class_asv.size() < cap) sh.class_asv.resize(cap);
        return savont_b200_map_twin_read_ptrs(m->ctx[d], dna_ptrs + r0, lengths + r0, snpmer_ptrs + r0, snpmer_counts + r0, snpmer_kmer_bytes,
                                              file_idx ? file_idx + r0 : nullptr, cnt, n_samples, status + r0, best_nm + r0,
                                              sh.class_offsets.data(), sh.class_asv.data(), cap, nullptr, nullptr, nullptr, nullptr, nullptr,
                                              &sh.st);
    }));
    return finish_map(m, n_reads, class_offsets, class_asv, class_cap, unambig, ambig, leq10, totals, sample_singleton, stats);
}

}  // extern "C"
