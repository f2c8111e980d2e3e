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

// ---- Stage 3 greedy SNPmer clustering (snpcluster.cu) ----
extern "C" int savont_b200_cluster_reads_by_snpmers(savont_b200_ctx *h, const uint64_t *snp_kmers, const uint64_t *snp_offsets, size_t n_reads,
                                                    const uint32_t *cluster_reads, const uint64_t *cluster_offsets, size_t n_clusters,
                                                    uint32_t iterative_above, uint32_t *rep_of, uint64_t *n_representatives,
                                                    savont_b200_stats *stats) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!snp_offsets || !cluster_offsets || (n_reads && !rep_of)) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    if (n_reads >= 0xfffffff0ull || n_clusters >= 0xfffffff0ull) return fail(SAVONT_B200_ERR_INVALID, "too many reads / clusters");
    const uint64_t m0 = snp_offsets[0], n_mark = snp_offsets[n_reads] - m0;
    const uint64_t c0 = cluster_offsets[0], n_memb = cluster_offsets[n_clusters] - c0;
    if ((n_mark && !snp_kmers) || (n_memb && !cluster_reads)) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    std::vector<uint64_t> rel(n_reads + 1), crel(n_clusters + 1);
    for (size_t r = 0; r <= n_reads; ++r) {
        if (r && snp_offsets[r] < snp_offsets[r - 1]) return fail(SAVONT_B200_ERR_INVALID, "snp_offsets not monotonic at %zu", r);
        rel[r] = snp_offsets[r] - m0;
    }
    for (size_t q = 0; q <= n_clusters; ++q) {
        if (q && cluster_offsets[q] < cluster_offsets[q - 1]) return fail(SAVONT_B200_ERR_INVALID, "cluster_offsets not monotonic at %zu", q);
        crel[q] = cluster_offsets[q] - c0;
    }
    std::vector<uint8_t> seen(n_reads, 0);
    for (uint64_t i = 0; i < n_memb; ++i) {
        const uint32_t r = cluster_reads[c0 + i];
        if (r >= n_reads) return fail(SAVONT_B200_ERR_INVALID, "cluster member %u out of range", r);
        if (seen[r]) return fail(SAVONT_B200_ERR_INVALID, "read %u is a member of two clusters", r);
        seen[r] = 1;
    }
    // largest cluster first: the reads of a cluster are a serial chain, so the long chains must start early
    std::vector<uint32_t> order(n_clusters);
    for (size_t q = 0; q < n_clusters; ++q) order[q] = (uint32_t)q;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return crel[x + 1] - crel[x] > crel[y + 1] - crel[y]; });
    c->last_stats = savont_b200_stats{};
    cudaStream_t st = c->stream;
    DevBuf *w = c->cluster_ws;
    CK(cudaEventRecord(c->ev[0], st));
    TRY(w[0].ensure(n_mark * 8 + 16)); TRY(w[1].ensure((n_reads + 1) * 8)); TRY(w[2].ensure(n_memb * 4 + 1024)); TRY(w[3].ensure((n_clusters + 1) * 8));
    TRY(w[4].ensure(n_clusters * 4 + 16)); TRY(w[5].ensure(n_memb * 4 + 16)); TRY(w[6].ensure(n_memb * 4 + 16)); TRY(w[7].ensure(n_reads * 4 + 16));
    TRY(w[8].ensure(64));
    TRY(upload_host(c, w[0].p, snp_kmers + m0, n_mark * 8, st));
    CK(cudaMemcpyAsync(w[1].p, rel.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_memb) CK(cudaMemcpyAsync(w[2].p, cluster_reads + c0, n_memb * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(w[3].p, crel.data(), (n_clusters + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_clusters) CK(cudaMemcpyAsync(w[4].p, order.data(), n_clusters * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(w[7].p, 0xff, n_reads * 4 + 16, st));                         // reads outside every cluster: UINT32_MAX
    CK(cudaMemsetAsync(w[8].p, 0, 64, st));
    c->last_stats.h2d_bytes = n_mark * 8 + (n_reads + 1) * 8 + n_memb * 4 + (n_clusters + 1) * 8 + n_clusters * 4;
    SnpClusterArgs a{};
    a.k = c->params.k; a.snp_kmers = w[0].as<uint64_t>(); a.snp_offsets = w[1].as<uint64_t>();
    a.cluster_reads = w[2].as<uint32_t>(); a.cluster_offsets = w[3].as<uint64_t>(); a.order = w[4].as<uint32_t>();
    a.n_clusters = (uint32_t)n_clusters; a.iterative_above = iterative_above;
    a.reps = w[5].as<uint32_t>(); a.rep_size = w[6].as<uint32_t>(); a.rep_of = w[7].as<uint32_t>();
    a.work_ctr = w[8].as<uint32_t>(); a.err = w[8].as<uint32_t>() + 1; a.total_reps = reinterpret_cast<unsigned long long *>(w[8].as<uint32_t>() + 2);
    CK(cudaEventRecord(c->ev[1], st));
    // cluster-major copy of the lists (count -> scan -> gather): the cluster kernel then streams consecutive memory
    if (n_mark >= (1ull << 32)) return fail(SAVONT_B200_ERR_OVERFLOW, "more than 4 Gi SNPmers in one call");
    TRY(w[9].ensure(n_memb * 4 + 16)); TRY(w[10].ensure((n_memb + 2) * 4 + 256)); TRY(w[11].ensure((n_memb / 2048 + 4) * 4)); TRY(w[12].ensure(n_mark * 8 + 4096 + 16));
    if (n_memb) {
        launch_snp_cluster_count(a, (uint32_t)n_memb, w[9].as<uint32_t>(), st);
        launch_scan_u32(w[9].as<uint32_t>(), w[10].as<uint32_t>(), (uint32_t)n_memb, w[11].as<uint32_t>(), st);
        a.poff = w[10].as<uint32_t>();
        launch_snp_cluster_gather(a, (uint32_t)n_memb, w[12].as<unsigned long long>(), st);
        a.pkmers = w[12].as<unsigned long long>();
        c->last_stats.kernel_launches += 5;
    }
    // tiles of the round kernels (a tile never spans two clusters), the ring of next representatives, the per-position state
    std::vector<uint32_t> tiles;
    for (size_t q = 0; q < n_clusters; ++q)
        for (uint64_t p0 = crel[q]; p0 < crel[q + 1]; p0 += kSnpTile) { tiles.push_back((uint32_t)q); tiles.push_back((uint32_t)p0); }
    const uint32_t n_tiles = (uint32_t)(tiles.size() / 2);
    TRY(w[13].ensure(3 * n_clusters * 4 + 16)); TRY(w[14].ensure(tiles.size() * 4 + 16)); TRY(w[15].ensure((2 + kSnpTies) * n_memb * 4 + n_memb + 64));
    std::vector<uint32_t> tl(tiles.size());
    for (uint32_t t = 0; t < n_tiles; ++t) { tl[t] = tiles[2 * t]; tl[n_tiles + t] = tiles[2 * t + 1]; }
    if (n_tiles) CK(cudaMemcpyAsync(w[14].p, tl.data(), tl.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(w[13].p, 0xff, 3 * n_clusters * 4 + 16, st));                   // no representative known; last_round = -1
    a.n_memb = (uint32_t)n_memb; a.n_tiles = n_tiles;
    a.tile_cl = w[14].as<uint32_t>(); a.tile_p0 = w[14].as<uint32_t>() + n_tiles;
    a.next3 = w[13].as<uint32_t>(); a.last_round = reinterpret_cast<int *>(w[13].as<uint32_t>() + 3 * n_clusters);
    a.st_mt = w[15].as<uint32_t>(); a.st_first = a.st_mt + n_memb; a.st_tie = a.st_first + n_memb;
    const bool prof = getenv("SAVONT_B200_SNPCLUSTER_PROF") != nullptr;
    cudaEvent_t pe[3];
    a.prof = prof ? 1 : 0;
    if (prof) { for (auto &e : pe) CK(cudaEventCreate(&e)); CK(cudaEventRecord(pe[0], st)); }
    launch_snp_init(a, st);
    // rounds in batches: the number of rounds is the largest number of (non-empty) representatives of a cluster, known only to the device;
    // a round without any representative left is a no-op, so overshooting costs launches only
    uint32_t rounds = 0, batch = 8;
    while (n_tiles) {
        for (uint32_t j = 0; j < batch; ++j) launch_snp_round(a, rounds + j, st);
        rounds += batch;
        int last = -1;
        CK(cudaMemcpyAsync(&last, a.last_round, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (last < (int)rounds - 1) break;
        if (batch < 64) batch *= 2;
    }
    if (prof) CK(cudaEventRecord(pe[1], st));
    launch_snp_resolve(a, c->sm_count, st);
    CK(cudaGetLastError());
    if (prof) {
        CK(cudaEventRecord(pe[2], st)); CK(cudaStreamSynchronize(st));
        uint32_t cnt[16];
        CK(cudaMemcpy(cnt, w[8].p, 64, cudaMemcpyDeviceToHost));
        { unsigned long long pr[4]; memcpy(pr, cnt + 8, 32); fprintf(stderr, "[snpcluster prof] largest cluster %llu reads: head %llu cycles, body %llu cycles, 3+-way ties %llu\n", pr[3], pr[0], pr[1], pr[2]); }
        fprintf(stderr, "[snpcluster prof] %u tiles, %u rounds launched: init + rounds %.3f ms, resolve %.3f ms; tied reads %u, re-scored reads %u\n", n_tiles, rounds,
                ev_ms(pe[0], pe[1]), ev_ms(pe[1], pe[2]), cnt[4], cnt[5]);
        for (auto &e : pe) cudaEventDestroy(e);
    }
    if (n_clusters) c->last_stats.kernel_launches += 2 + rounds;
    c->last_stats.n_batches = rounds;
    CK(cudaEventRecord(c->ev[2], st));
    uint32_t tail[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(tail, w[8].p, sizeof(tail), cudaMemcpyDeviceToHost, st));
    if (n_reads) CK(cudaMemcpyAsync(rep_of, w[7].p, n_reads * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (tail[1] & 1u) return fail(SAVONT_B200_ERR_OVERFLOW, "a read has more than 512 SNPmers");
    if (tail[1] & 4u) return fail(SAVONT_B200_ERR_CUDA, "internal: a read without a compatible representative was not promoted");
    if (tail[1] & 2u) return fail(SAVONT_B200_ERR_INVALID, "a read carries two SNPmers with the same splitmer (TwinRead.snpmer_kmers never does, src/seeding.rs:550-568)");
    unsigned long long nr = 0; memcpy(&nr, tail + 2, 8);
    if (n_representatives) *n_representatives = nr;
    c->last_stats.d2h_bytes = n_reads * 4;
    c->last_stats.n_reads = n_memb; c->last_stats.n_pairs = nr;
    c->last_stats.h2d_ms = ev_ms(c->ev[0], c->ev[1]); c->last_stats.filter_ms = ev_ms(c->ev[1], c->ev[2]);
    c->last_stats.total_ms = ev_ms(c->ev[0], c->ev[2]);
    if (stats) *stats = c->last_stats;
    return 0;
}
