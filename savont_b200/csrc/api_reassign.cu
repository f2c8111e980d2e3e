// api_reassign.cu -- host side of savont_b200_snpmer_compare (Stage 3 reclustering): builds the inverted
// index over the clusters' consensus SNPmers and runs snpmer_compare_kernel over the reads in row batches.
#include "api_internal.cuh"

using namespace sb;

extern "C" int savont_b200_snpmer_compare(savont_b200_ctx *h, const uint64_t *cons_kmers, const uint64_t *cons_offsets, size_t n_cons,
                                          const uint64_t *read_kmers, const uint64_t *read_offsets, size_t n_reads, uint32_t *best_cons,
                                          uint32_t *best_mismatches, uint32_t *best_matches, uint16_t *matrix, savont_b200_stats *stats) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!cons_offsets || !read_offsets || (n_reads && (!best_cons || !best_mismatches || !best_matches)))
        return fail(SAVONT_B200_ERR_INVALID, "null argument");
    if (n_cons >= (1u << 30) || n_reads >= 0xfffffff0ull) return fail(SAVONT_B200_ERR_INVALID, "too many clusters / reads");
    if (n_cons * 4 > 200u * 1024u) return fail(SAVONT_B200_ERR_OVERFLOW, "%zu clusters: at most 51200 fit the shared-memory row", n_cons);
    for (size_t i = 0; i < n_cons; ++i)
        if (cons_offsets[i + 1] < cons_offsets[i]) return fail(SAVONT_B200_ERR_INVALID, "consensus offsets not monotonic at %zu", i);
    for (size_t r = 0; r < n_reads; ++r) {
        if (read_offsets[r + 1] < read_offsets[r]) return fail(SAVONT_B200_ERR_INVALID, "read offsets not monotonic at %zu", r);
        if (read_offsets[r + 1] - read_offsets[r] > 65535) return fail(SAVONT_B200_ERR_OVERFLOW, "read %zu: more than 65535 markers", r);
    }
    c->last_stats = savont_b200_stats{};
    cudaStream_t st = c->stream;
    const int k = c->params.k;
    const uint64_t split_mask = ~(3ull << (k - 1));
    // ---- index: per cluster the LAST entry of a splitmer counts (HashMap::insert, src/asv_cluster.rs:1031-1034) ----
    struct Ent { uint64_t key; uint32_t val; };
    std::vector<Ent> ents;
    ents.reserve((size_t)(cons_offsets[n_cons] - cons_offsets[0]));
    std::vector<std::pair<uint64_t, uint32_t>> one;
    for (size_t ci = 0; ci < n_cons; ++ci) {
        one.clear();
        for (uint64_t i = cons_offsets[ci]; i < cons_offsets[ci + 1]; ++i)
            one.push_back({cons_kmers[i] & split_mask, (uint32_t)((cons_kmers[i] >> (k - 1)) & 3u)});
        std::stable_sort(one.begin(), one.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
        for (size_t i = 0; i < one.size(); ++i)
            if (i + 1 == one.size() || one[i + 1].first != one[i].first) ents.push_back(Ent{one[i].first, (uint32_t)(ci << 2) | one[i].second});
    }
    std::stable_sort(ents.begin(), ents.end(), [](const Ent &x, const Ent &y) { return x.key < y.key; });   // buckets keep cluster order
    size_t n_keys = 0;
    for (size_t i = 0; i < ents.size(); ++i) if (i == 0 || ents[i].key != ents[i - 1].key) ++n_keys;
    const uint32_t cap = next_pow2_u32(std::max<uint64_t>(2 * n_keys + 2, 16));
    const uint32_t shift = shift_for_cap(cap);
    std::vector<uint64_t> keys(cap, kEmpty64); std::vector<uint2> vals(cap, make_uint2(0, 0)); std::vector<uint32_t> flat(ents.size());
    for (size_t i = 0; i < ents.size();) {
        size_t j = i;
        while (j < ents.size() && ents[j].key == ents[i].key) { flat[j] = ents[j].val; ++j; }
        uint32_t slot = hash_slot(ents[i].key, shift);
        while (keys[slot] != kEmpty64) slot = (slot + 1u) & (cap - 1u);
        keys[slot] = ents[i].key; vals[slot] = make_uint2((uint32_t)i, (uint32_t)(j - i));
        i = j;
    }
    DevBuf *w = c->cluster_ws;         // the clustering scratch is free between calls
    const uint64_t m0 = read_offsets[0], n_mark = read_offsets[n_reads] - m0;
    std::vector<uint64_t> rel(n_reads + 1);
    for (size_t r = 0; r <= n_reads; ++r) rel[r] = read_offsets[r] - m0;
    CK(cudaEventRecord(c->ev[0], st));
    TRY(w[0].ensure((size_t)cap * 8)); TRY(w[1].ensure((size_t)cap * 8)); TRY(w[2].ensure(flat.size() * 4 + 16));
    TRY(w[3].ensure(n_mark * 8 + 16)); TRY(w[4].ensure((n_reads + 1) * 8)); TRY(w[5].ensure(n_reads * 12 + 16));
    CK(cudaMemcpyAsync(w[0].p, keys.data(), (size_t)cap * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(w[1].p, vals.data(), (size_t)cap * 8, cudaMemcpyHostToDevice, st));
    if (!flat.empty()) CK(cudaMemcpyAsync(w[2].p, flat.data(), flat.size() * 4, cudaMemcpyHostToDevice, st));
    if (n_mark) CK(cudaMemcpyAsync(w[3].p, read_kmers + m0, n_mark * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(w[4].p, rel.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    c->last_stats.h2d_bytes = (uint64_t)cap * 16 + flat.size() * 4 + n_mark * 8 + (n_reads + 1) * 8;
    SnpCompareArgs a{};
    a.k = k; a.idx_keys = ents.empty() ? nullptr : w[0].as<uint64_t>(); a.idx_vals = w[1].as<uint2>(); a.idx_mask = cap - 1u; a.idx_shift = shift;
    a.idx_entries = w[2].as<uint32_t>(); a.n_cons = (uint32_t)n_cons;
    a.read_kmers = w[3].as<uint64_t>(); a.read_offsets = w[4].as<uint64_t>(); a.n_reads = (uint32_t)n_reads;
    a.best_cons = w[5].as<uint32_t>(); a.best_mismatches = a.best_cons + n_reads; a.best_matches = a.best_cons + 2 * n_reads;
    // row batches: the optional matrix goes through a device buffer of at most 1 GiB
    const size_t rows = matrix ? std::max<size_t>(1, ((size_t)1 << 28) / std::max<size_t>(n_cons, 1)) : std::max<size_t>(n_reads, 1);
    if (matrix) TRY(w[6].ensure(std::min(rows, std::max<size_t>(n_reads, 1)) * std::max<size_t>(n_cons, 1) * 4 + 16));
    CK(cudaEventRecord(c->ev[1], st));
    for (size_t r0 = 0; r0 < n_reads; r0 += rows) {
        const uint32_t nb = (uint32_t)std::min(rows, n_reads - r0);
        a.matrix = matrix ? w[6].as<uint32_t>() : nullptr;
        if (launch_snpmer_compare(a, (uint32_t)r0, nb, c->sm_count, st)) return fail(SAVONT_B200_ERR_OVERFLOW, "too many clusters");
        CK(cudaGetLastError());
        c->last_stats.kernel_launches += 1;
        if (matrix && n_cons) {
            CK(cudaMemcpyAsync(matrix + r0 * n_cons * 2, w[6].p, (size_t)nb * n_cons * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            c->last_stats.d2h_bytes += (uint64_t)nb * n_cons * 4;
        }
    }
    CK(cudaEventRecord(c->ev[2], st));
    if (n_reads) {
        CK(cudaMemcpyAsync(best_cons, a.best_cons, n_reads * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(best_mismatches, a.best_mismatches, n_reads * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(best_matches, a.best_matches, n_reads * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    c->last_stats.d2h_bytes += n_reads * 12;
    c->last_stats.n_reads = n_reads; c->last_stats.n_pairs = (uint64_t)n_reads * n_cons;
    c->last_stats.h2d_ms = ev_ms(c->ev[0], c->ev[1]); c->last_stats.filter_ms = ev_ms(c->ev[1], c->ev[2]);
    c->last_stats.total_ms = ev_ms(c->ev[0], c->ev[2]);
    if (stats) *stats = c->last_stats;
    return 0;
}
