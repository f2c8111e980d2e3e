// api_kmer.cu -- host side of K5 (Stage 1 split-k-mer counting, kernels in kmercount.cu): table / partition sizing, the
// chunked upload of the reads and the two extern "C" entry points savont_b200_count_split_kmers[_resident].
#include "api_internal.cuh"

namespace sb {

struct KcView { SeqBatch in; const uint8_t *rc; };

static uint64_t mod_inverse_pow2(uint64_t a, uint64_t mask) {        // a odd: a^-1 mod 2^64 by Newton iteration, then masked
    uint64_t x = a;                                                   // correct to 3 bits
    for (int i = 0; i < 6; ++i) x *= 2 - a * x;
    return x & mask;
}

// Partitioned counting (kmercount.cu): histogram -> scatter -> per-partition shared-memory counting.  Returns 1 when the
// spill table overflowed (the caller falls back to the single-table algorithm), 0 on success, < 0 on error.
template <typename GetView>
static int kc_run_part(Context *c, size_t n_chunks, uint64_t total_bases, GetView view, int single_strand, uint64_t *out_kmers,
                       uint32_t *out_counts, size_t out_cap, size_t *n_out, savont_b200_kmer_stats &st) {
    const int k = c->dp.k, pbits = kKmerPartBits, rem_bits = 2 * k - pbits;
    const uint32_t P = 1u << pbits;
    KmerPartPlan pl{};
    pl.kmask = (1ull << (2 * k)) - 1ull;
    pl.mul = (kFibMul | 1ull) & pl.kmask; pl.inv = mod_inverse_pow2(pl.mul, pl.kmask);
    pl.pbits = pbits; pl.rem_bits = rem_bits; pl.single_strand = single_strand;
    pl.smem_fill_limit = c->kc_smem_limit;
    // small device block: hist[P] | cursor[P] | part_off[P + 1] (u64) | ctr[16] | out_n | queue | n_distinct (u64)
    const size_t small_bytes = (size_t)P * 8 + (size_t)(P + 1) * 8 + 256 + 128 * 4 + 132 * 4;
    TRY(c->d_kp_small.ensure(small_bytes));
    uint8_t *sm = c->d_kp_small.as<uint8_t>();
    pl.hist = reinterpret_cast<uint32_t *>(sm); pl.cursor = pl.hist + P;
    unsigned long long *d_off = reinterpret_cast<unsigned long long *>(sm + (size_t)P * 8);
    pl.part_off = d_off;
    uint32_t *tail = reinterpret_cast<uint32_t *>(sm + (size_t)P * 8 + (size_t)(P + 1) * 8);
    pl.ctr = tail; pl.out_n = tail + 16; pl.queue = tail + 17; pl.n_distinct = reinterpret_cast<unsigned long long *>(tail + 18);
    uint32_t *cursor1 = tail + 64, *d_tile_start = tail + 64 + 128;
    // spill table
    const uint64_t slots = 1ull << 22;
    TRY(c->d_kc_table.ensure(slots * sizeof(KmerSlot)));
    pl.table = c->d_kc_table.as<KmerSlot>(); pl.tshift = shift_for_cap((uint32_t)slots); pl.tmask = (uint32_t)(slots - 1);
    pl.fill_limit = (uint32_t)(slots / 2);
    CK(cudaEventRecord(c->ev[0], c->stream));
    CK(cudaMemsetAsync(sm, 0, small_bytes, c->stream));
    launch_kmer_table_clear(pl.table, slots, c->stream);
    for (size_t i = 0; i < n_chunks; ++i) {
        KcView v{};
        TRY(view(i, 1, &v));
        launch_kmer_hist(pl, v.in, v.rc, k, c->dp.min_bq, c->sm_count, c->stream);
        CK(cudaGetLastError());
        CK(cudaEventRecord(c->kc_ev_done[i & 1], c->stream));
        st.kernel_launches += 1;
    }
    CK(cudaEventRecord(c->ev[1], c->stream));
    std::vector<uint32_t> hist(P);
    CK(cudaMemcpyAsync(hist.data(), pl.hist, (size_t)P * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::vector<unsigned long long> off(P + 1, 0);
    for (uint32_t i = 0; i < P; ++i) off[i + 1] = off[i] + hist[i];
    const uint64_t total = off[P];
    TRY(c->d_kp_entries.ensure((total + 64) * 4));
    pl.entries = c->d_kp_entries.as<uint32_t>();
    CK(cudaMemcpyAsync(d_off, off.data(), (size_t)(P + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    uint32_t max_len_all = 0;
    for (size_t i = 0; i < n_chunks; ++i) { KcView v{}; TRY(view(i, 0, &v)); if (v.in.max_len > max_len_all) max_len_all = v.in.max_len; }
    const bool two_level = kmer_part_two_level(k, max_len_all);
    std::vector<uint32_t> tile_start(132, 0);
    if (two_level) {
        // level 1 (from the reads, 128 bins) into entries1, level 2 (tiles of entries1) into the final partitions
        const int sub = pbits - kmer_scatter1_bits();
        const uint32_t P1 = 1u << kmer_scatter1_bits(), tile2 = kmer_scatter2_tile();
        for (uint32_t b1 = 0; b1 < P1; ++b1) {
            const unsigned long long nb = off[(size_t)(b1 + 1) << sub] - off[(size_t)b1 << sub];
            tile_start[b1 + 1] = tile_start[b1] + (uint32_t)((nb + tile2 - 1) / tile2);
        }
        TRY(c->d_kp_entries1.ensure((total + 64) * 4));
        CK(cudaMemcpyAsync(d_tile_start, tile_start.data(), 132 * 4, cudaMemcpyHostToDevice, c->stream));
    }
    for (size_t i = 0; i < n_chunks; ++i) {
        KcView v{};
        TRY(view(i, 2, &v));
        if (two_level) launch_kmer_scatter1(pl, c->d_kp_entries1.as<uint32_t>(), cursor1, v.in, v.rc, k, c->dp.min_bq, c->sm_count, c->stream);
        else launch_kmer_scatter(pl, v.in, v.rc, k, c->dp.min_bq, c->sm_count, c->stream);
        CK(cudaGetLastError());
        CK(cudaEventRecord(c->kc_ev_done[i & 1], c->stream));
        st.kernel_launches += 1;
    }
    if (two_level) {
        launch_kmer_scatter2(pl, c->d_kp_entries1.as<uint32_t>(), d_tile_start, tile_start[1u << kmer_scatter1_bits()], k, c->sm_count, c->stream);
        CK(cudaGetLastError());
        st.kernel_launches += 1;
    }
    CK(cudaEventRecord(c->ev[2], c->stream));
    // every survivor has more than two occurrences
    const uint64_t cap64 = std::min<uint64_t>((uint64_t)out_cap, total / 3 + 32);
    const uint32_t cap_dev = (uint32_t)std::min<uint64_t>(cap64, 0xfffffff0ull);
    TRY(c->d_kc_out_k.ensure((size_t)cap_dev * 8 + 64)); TRY(c->d_kc_out_c.ensure((size_t)cap_dev * 8 + 64));
    pl.out_kmers = c->d_kc_out_k.as<uint64_t>(); pl.out_counts = c->d_kc_out_c.as<uint32_t>(); pl.out_cap = cap_dev;
    launch_kmer_count_part(pl, c->sm_count, c->stream);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev[3], c->stream));
    launch_kmer_compact(pl.table, slots, single_strand, pl.out_kmers, pl.out_counts, cap_dev, pl.out_n, c->stream);
    CK(cudaGetLastError());
    st.kernel_launches += 3;
    CK(cudaEventRecord(c->ev[4], c->stream));
    uint32_t t[20];
    CK(cudaMemcpyAsync(t, tail, sizeof(t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    st.hist_ms += ev_ms(c->ev[0], c->ev[1]); st.scatter_ms += ev_ms(c->ev[1], c->ev[2]);
    st.count_ms += ev_ms(c->ev[0], c->ev[3]); st.compact_ms += ev_ms(c->ev[3], c->ev[4]);
    if (t[1] != 0) return 1;                                          // spill table overflow: not a 16S-like input
    const uint32_t n_found = t[16];
    unsigned long long nd = 0; memcpy(&nd, t + 18, 8);
    st.table_slots = slots; st.n_distinct = nd + t[0]; st.partitions = P; st.spilled_distinct = t[0];
    if (n_out) *n_out = n_found;
    if (n_found > out_cap) return fail(SAVONT_B200_ERR_CAPACITY, "k-mer output capacity %zu too small (need >= %u)", out_cap, n_found);
    if (n_found) {
        CK(cudaMemcpyAsync(out_kmers, pl.out_kmers, (size_t)n_found * 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(out_counts, pl.out_counts, (size_t)n_found * 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    st.d2h_bytes += (uint64_t)n_found * 16 + 16;
    return 0;
}

// the two passes + compaction over device-resident chunk views; (re)sizes the Bloom filter and the table, grows the table
// and starts over when it fills up.  `views(i, pass)` returns the i-th chunk after making sure it is on the device.
template <typename GetView>
static int kc_run(Context *c, size_t n_chunks, uint64_t total_bases, GetView view, int single_strand, uint64_t *out_kmers,
                  uint32_t *out_counts, size_t out_cap, size_t *n_out, savont_b200_kmer_stats &st) {
    // sizing: slots = occurrences / 8 rounded up to a power of two (load <= 0.5 is enforced: the table is regrown x4 and the
    // pass re-run when more than half of the slots fill up)
    uint64_t slots = 1ull << 16;
    while (slots * 8ull < total_bases && slots < (1ull << 31)) slots <<= 1;
    TRY(c->d_kc_ctr.ensure(64));
    uint32_t ctr[4] = {0, 0, 0, 0};
    for (int attempt = 0;; ++attempt) {
        TRY(c->d_kc_table.ensure(slots * sizeof(KmerSlot)));
        const uint32_t tshift = shift_for_cap((uint32_t)slots), tmask = (uint32_t)(slots - 1);
        CK(cudaEventRecord(c->ev[0], c->stream));
        CK(cudaMemsetAsync(c->d_kc_ctr.p, 0, 64, c->stream));
        launch_kmer_table_clear(c->d_kc_table.as<KmerSlot>(), slots, c->stream);
        st.kernel_launches += 1;
        for (size_t i = 0; i < n_chunks; ++i) {
            KcView v{};
            TRY(view(i, 1, &v));
            launch_kmer_count(v.in, v.rc, c->dp.k, c->dp.min_bq, c->d_kc_table.as<KmerSlot>(), tshift, tmask, c->d_kc_ctr.as<uint32_t>(),
                              (uint32_t)(slots / 2), c->sm_count, c->stream);
            CK(cudaGetLastError());
            CK(cudaEventRecord(c->kc_ev_done[i & 1], c->stream));
            st.kernel_launches += 1;
        }
        CK(cudaEventRecord(c->ev[1], c->stream));
        CK(cudaMemcpyAsync(ctr, c->d_kc_ctr.p, sizeof(ctr), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        st.count_ms += ev_ms(c->ev[0], c->ev[1]);
        if (ctr[1] == 0) break;
        if (attempt == 4 || slots >= (1ull << 31))
            return fail(SAVONT_B200_ERR_OVERFLOW, "k-mer table sizing did not converge (%llu slots): split the call", (unsigned long long)slots);
        slots <<= 2; st.table_regrows += 1;
    }
    st.table_slots = slots; st.n_distinct = ctr[0];
    CK(cudaEventRecord(c->ev[2], c->stream));
    // compaction: at most table_used survivors
    const uint32_t cap_dev = ctr[0] + 32u;
    TRY(c->d_kc_out_k.ensure((size_t)cap_dev * 8)); TRY(c->d_kc_out_c.ensure((size_t)cap_dev * 8));
    CK(cudaMemsetAsync(c->d_kc_ctr.as<uint32_t>() + 8, 0, 4, c->stream));
    launch_kmer_compact(c->d_kc_table.as<KmerSlot>(), slots, single_strand, c->d_kc_out_k.as<uint64_t>(), c->d_kc_out_c.as<uint32_t>(),
                        cap_dev, c->d_kc_ctr.as<uint32_t>() + 8, c->stream);
    CK(cudaGetLastError());
    st.kernel_launches += 1;
    CK(cudaEventRecord(c->ev[3], c->stream));
    uint32_t n_found = 0;
    CK(cudaMemcpyAsync(&n_found, c->d_kc_ctr.as<uint32_t>() + 8, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    st.compact_ms += ev_ms(c->ev[2], c->ev[3]);
    if (n_out) *n_out = n_found;
    if (n_found > out_cap) return fail(SAVONT_B200_ERR_CAPACITY, "k-mer output capacity %zu too small (need >= %u)", out_cap, n_found);
    if (n_found) {
        CK(cudaMemcpyAsync(out_kmers, c->d_kc_out_k.p, (size_t)n_found * 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(out_counts, c->d_kc_out_c.p, (size_t)n_found * 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    st.d2h_bytes += (uint64_t)n_found * 16 + 16;
    return 0;
}


// ------------------------------------------------------------------------------------------------
// Sharded counting over several GPUs (one context per GPU, one process per GPU): every rank histograms and scatters ITS reads;
// partition p belongs to rank p * n_ranks / P, and the level-2 scatter kernel writes each rank's run of partition p straight into
// the owner's entry buffer (peer memory over NVLink, cudaIpc-mapped by the caller) at the offset the all-gathered histograms
// give it.  After a barrier every owner counts its partitions.  The host-side index arithmetic lives with the caller
// (savont_b200/dist.py: kmer_shard_layout) because it needs all ranks' histograms.
// ------------------------------------------------------------------------------------------------
static int ks_check(Context *c) {
    Batch &b = c->resident;
    if (b.twin) return fail(SAVONT_B200_ERR_INVALID, "the resident reads are TwinReads: k-mer counting needs the FASTQ bytes");
    if (!(kmer_part_supported(c->dp.k, b.max_len) && kmer_part_two_level(c->dp.k, std::max(b.max_len, 1u))))
        return fail(SAVONT_B200_ERR_INVALID, "sharded k-mer counting needs the two-level partitioned algorithm (k <= 19, reads <= 24000 bases)");
    return 0;
}

}  // namespace sb

using namespace sb;

extern "C" {

int savont_b200_kmer_shard_partitions(void) { return 1 << kKmerPartBits; }

int savont_b200_kmer_shard_hist(savont_b200_ctx *h, uint32_t *hist) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!hist) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    TRY(ks_check(c));
    Batch &b = c->resident;
    Context::KmerShard &ks = c->ks;
    ks = Context::KmerShard{};
    const int k = c->dp.k, pbits = kKmerPartBits, rem_bits = 2 * k - pbits;
    const uint32_t P = 1u << pbits;
    KmerPartPlan &pl = ks.pl;
    pl.kmask = (1ull << (2 * k)) - 1ull;
    pl.mul = (kFibMul | 1ull) & pl.kmask; pl.inv = mod_inverse_pow2(pl.mul, pl.kmask);
    pl.pbits = pbits; pl.rem_bits = rem_bits; pl.smem_fill_limit = c->kc_smem_limit;
    const size_t small_bytes = (size_t)P * 8 + (size_t)(P + 1) * 8 + 256 + 128 * 4 + 132 * 4;      // as in kc_run_part
    TRY(c->d_kp_small.ensure(small_bytes));
    uint8_t *sm = c->d_kp_small.as<uint8_t>();
    pl.hist = reinterpret_cast<uint32_t *>(sm); pl.cursor = pl.hist + P;
    unsigned long long *d_off = reinterpret_cast<unsigned long long *>(sm + (size_t)P * 8);
    pl.part_off = d_off;
    uint32_t *tail = reinterpret_cast<uint32_t *>(sm + (size_t)P * 8 + (size_t)(P + 1) * 8);
    pl.ctr = tail; pl.out_n = tail + 16; pl.queue = tail + 17; pl.n_distinct = reinterpret_cast<unsigned long long *>(tail + 18);
    ks.tail = tail; ks.cursor1 = tail + 64; ks.d_tile_start = tail + 64 + 128;
    ks.slots = 1ull << 22;
    TRY(c->d_kc_table.ensure(ks.slots * sizeof(KmerSlot)));
    pl.table = c->d_kc_table.as<KmerSlot>(); pl.tshift = shift_for_cap((uint32_t)ks.slots); pl.tmask = (uint32_t)(ks.slots - 1);
    pl.fill_limit = (uint32_t)(ks.slots / 2);
    CK(cudaEventRecord(c->ev[0], c->stream));
    CK(cudaMemsetAsync(sm, 0, small_bytes, c->stream));
    launch_kmer_table_clear(pl.table, ks.slots, c->stream);
    if (b.n) launch_kmer_hist(pl, seq_view(b), nullptr, k, c->dp.min_bq, c->sm_count, c->stream);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev[1], c->stream));
    CK(cudaMemcpyAsync(hist, pl.hist, (size_t)P * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    ks.off_local.assign(P + 1, 0);
    for (uint32_t i = 0; i < P; ++i) ks.off_local[i + 1] = ks.off_local[i] + hist[i];
    CK(cudaMemcpyAsync(d_off, ks.off_local.data(), (size_t)(P + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    TRY(c->d_kp_entries1.ensure((ks.off_local[P] + 64) * 4));
    ks.active = true;
    return 0;
}

int savont_b200_kmer_shard_alloc(savont_b200_ctx *h, uint64_t n_entries, void **dev_ptr, uint64_t *capacity) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!dev_ptr) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    TRY(c->d_ks_entries.ensure((n_entries + 64) * 4));
    *dev_ptr = c->d_ks_entries.p;
    if (capacity) *capacity = c->d_ks_entries.cap / 4 - 64;
    return 0;
}

int savont_b200_ipc_export(savont_b200_ctx *h, const void *dev_ptr, void *handle64) {
    TRY(use_device(h ? &h->c : nullptr));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t");
    if (!dev_ptr || !handle64) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    cudaIpcMemHandle_t hd;
    CK(cudaIpcGetMemHandle(&hd, const_cast<void *>(dev_ptr)));
    memcpy(handle64, &hd, 64);
    return 0;
}
int savont_b200_ipc_open(savont_b200_ctx *h, const void *handle64, void **dev_ptr) {
    TRY(use_device(h ? &h->c : nullptr));
    if (!dev_ptr || !handle64) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    cudaIpcMemHandle_t hd; memcpy(&hd, handle64, 64);
    CK(cudaIpcOpenMemHandle(dev_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int savont_b200_ipc_close(savont_b200_ctx *h, void *dev_ptr) {
    TRY(use_device(h ? &h->c : nullptr));
    if (dev_ptr) CK(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

int savont_b200_kmer_shard_scatter(savont_b200_ctx *h, int n_ranks, void *const *peer_entries, const uint64_t *dst_off) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    Context::KmerShard &ks = c->ks;
    if (!ks.active) return fail(SAVONT_B200_ERR_INVALID, "savont_b200_kmer_shard_hist has not been called");
    if (n_ranks < 1 || !peer_entries || !dst_off) return fail(SAVONT_B200_ERR_INVALID, "bad argument");
    Batch &b = c->resident;
    KmerPartPlan &pl = ks.pl;
    const int k = c->dp.k;
    const uint32_t P = 1u << kKmerPartBits;
    ks.n_ranks = n_ranks;
    std::vector<uint32_t *> dst(P);
    for (uint32_t p = 0; p < P; ++p) {
        const int owner = (int)(((uint64_t)p * (uint64_t)n_ranks) >> kKmerPartBits);
        if (!peer_entries[owner]) return fail(SAVONT_B200_ERR_INVALID, "null entry buffer of rank %d", owner);
        dst[p] = reinterpret_cast<uint32_t *>(peer_entries[owner]) + dst_off[p];
    }
    TRY(c->d_ks_dstptr.ensure((size_t)P * sizeof(uint32_t *)));
    CK(cudaMemcpyAsync(c->d_ks_dstptr.p, dst.data(), (size_t)P * sizeof(uint32_t *), cudaMemcpyHostToDevice, c->stream));
    pl.dst_ptr = c->d_ks_dstptr.as<uint32_t *>();
    pl.entries = nullptr;
    const int sub = kKmerPartBits - kmer_scatter1_bits();
    const uint32_t P1 = 1u << kmer_scatter1_bits(), tile2 = kmer_scatter2_tile();
    std::vector<uint32_t> tile_start(132, 0);
    for (uint32_t b1 = 0; b1 < P1; ++b1) {
        const unsigned long long nb = ks.off_local[(size_t)(b1 + 1) << sub] - ks.off_local[(size_t)b1 << sub];
        tile_start[b1 + 1] = tile_start[b1] + (uint32_t)((nb + tile2 - 1) / tile2);
    }
    CK(cudaMemcpyAsync(ks.d_tile_start, tile_start.data(), 132 * 4, cudaMemcpyHostToDevice, c->stream));
    if (b.n) launch_kmer_scatter1(pl, c->d_kp_entries1.as<uint32_t>(), ks.cursor1, seq_view(b), nullptr, k, c->dp.min_bq, c->sm_count, c->stream);
    launch_kmer_scatter2(pl, c->d_kp_entries1.as<uint32_t>(), ks.d_tile_start, tile_start[P1], k, c->sm_count, c->stream);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev[2], c->stream));
    CK(cudaStreamSynchronize(c->stream));          // the caller's barrier comes next: the peers' buffers are complete after it
    return 0;
}

int savont_b200_kmer_shard_count(savont_b200_ctx *h, const uint64_t *part_off, int single_strand, uint64_t *out_kmers, uint32_t *out_counts,
                                 size_t out_cap, size_t *n_out, savont_b200_kmer_stats *stats) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    Context::KmerShard &ks = c->ks;
    if (!ks.active || ks.n_ranks == 0) return fail(SAVONT_B200_ERR_INVALID, "savont_b200_kmer_shard_scatter has not been called");
    if (!part_off || ((out_kmers == nullptr) != (out_counts == nullptr))) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    const bool fetch = out_kmers != nullptr;            // NULL outputs: count only (n_out, stats); the survivors stay on the device
    if (!fetch) out_cap = (size_t)1 << 32;
    KmerPartPlan &pl = ks.pl;
    const uint32_t P = 1u << kKmerPartBits;
    const uint64_t total = part_off[P];
    if ((total + 64) * 4 > c->d_ks_entries.cap) return fail(SAVONT_B200_ERR_INVALID, "part_off exceeds the allocated entry buffer");
    savont_b200_kmer_stats st; memset(&st, 0, sizeof(st));
    st.n_reads = c->resident.n; st.n_bases = c->resident.total_bases;
    TRY(c->d_ks_partoff.ensure((size_t)(P + 1) * 8));
    CK(cudaMemcpyAsync(c->d_ks_partoff.p, part_off, (size_t)(P + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    pl.part_off = c->d_ks_partoff.as<unsigned long long>(); pl.entries = c->d_ks_entries.as<uint32_t>(); pl.dst_ptr = nullptr;
    pl.single_strand = single_strand;
    const uint64_t cap64 = std::min<uint64_t>((uint64_t)out_cap, total / 3 + 32);
    const uint32_t cap_dev = (uint32_t)std::min<uint64_t>(cap64, 0xfffffff0ull);
    TRY(c->d_kc_out_k.ensure((size_t)cap_dev * 8 + 64)); TRY(c->d_kc_out_c.ensure((size_t)cap_dev * 8 + 64));
    pl.out_kmers = c->d_kc_out_k.as<uint64_t>(); pl.out_counts = c->d_kc_out_c.as<uint32_t>(); pl.out_cap = cap_dev;
    if (ks.count_runs++ > 0) {
        // a repeated call after ERR_CAPACITY: the entries are intact, the spill table, the counters and the output cursor are
        // not -- reset them so that nothing is counted twice
        CK(cudaMemsetAsync(ks.tail, 0, 64 * 4, c->stream));
        launch_kmer_table_clear(pl.table, ks.slots, c->stream);
    }
    CK(cudaEventRecord(c->ev[5], c->stream));
    launch_kmer_count_part(pl, c->sm_count, c->stream);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev[3], c->stream));
    launch_kmer_compact(pl.table, ks.slots, single_strand, pl.out_kmers, pl.out_counts, cap_dev, pl.out_n, c->stream);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev[4], c->stream));
    uint32_t t[20];
    CK(cudaMemcpyAsync(t, ks.tail, sizeof(t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    st.hist_ms = ev_ms(c->ev[0], c->ev[1]); st.scatter_ms = ev_ms(c->ev[1], c->ev[2]);      // scatter_ms includes the caller's all-gather
    st.count_ms = st.hist_ms + st.scatter_ms + ev_ms(c->ev[5], c->ev[3]); st.compact_ms = ev_ms(c->ev[3], c->ev[4]);
    st.total_ms = st.count_ms + st.compact_ms;
    st.kernel_launches = 6; st.partitions = P; st.table_slots = ks.slots;
    if (t[1] != 0) { ks.active = false; return fail(SAVONT_B200_ERR_OVERFLOW, "spill table overflow in sharded k-mer counting"); }
    const uint32_t n_found = t[16];
    unsigned long long nd = 0; memcpy(&nd, t + 18, 8);
    st.n_distinct = nd + t[0]; st.spilled_distinct = t[0];
    if (n_out) *n_out = n_found;
    if (stats) *stats = st;
    // too small an output: the shard stays active (entries in place), the caller repeats the call with >= *n_out slots
    if (n_found > out_cap) return fail(SAVONT_B200_ERR_CAPACITY, "k-mer output capacity %zu too small (need >= %u)", out_cap, n_found);
    ks.active = false;
    if (n_found && fetch) {
        CK(cudaMemcpyAsync(out_kmers, pl.out_kmers, (size_t)n_found * 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(out_counts, pl.out_counts, (size_t)n_found * 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

}  // extern "C"

namespace sb {
}  // namespace sb

using namespace sb;

extern "C" {

int savont_b200_count_split_kmers(savont_b200_ctx *h, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets,
                                  const uint8_t *rc, size_t n_reads, int single_strand, uint64_t *out_kmers,
                                  uint32_t *out_counts, size_t out_cap, size_t *n_out, savont_b200_kmer_stats *stats) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!offsets || (n_reads && !bases) || (out_cap && (!out_kmers || !out_counts))) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    savont_b200_kmer_stats st; memset(&st, 0, sizeof(st));
    CK(cudaEventRecord(c->ev[6], c->stream));
    // chunks of at most 256 Mi bases / 4 Mi reads (the upload of chunk i+1 overlaps the counting of chunk i); every chunk keeps its own device buffers when the whole input fits in
    // the device memory that is left beside the tables, otherwise two chunk buffers are re-filled in both passes
    std::vector<size_t> cuts{0};
    for (size_t r0 = 0; r0 < n_reads;) {
        size_t r1 = std::min(n_reads, r0 + ((size_t)1 << 22));
        while (r1 > r0 + 1 && offsets[r1] - offsets[r0] > (1ull << 28)) r1 = r0 + (r1 - r0) / 2;
        cuts.push_back(r1); r0 = r1;
    }
    const size_t nch = cuts.size() - 1;
    const uint64_t total = n_reads ? offsets[n_reads] - offsets[0] : 0;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    // the reads (1 or 2 bytes per base) stay on the device next to the partition entries (2 x 4 bytes per k-mer occurrence)
    const uint64_t need_resident = total * (quals ? 2ull : 1ull) + (uint64_t)n_reads * 9ull + total * 8ull + (4ull << 30);
    const bool resident = nch <= 2 || need_resident < (uint64_t)free_b;
    const size_t nbuf = resident ? nch : 2;
    if (c->kc_chunks.size() < nbuf) c->kc_chunks.resize(nbuf);
    std::vector<char> uploaded(nbuf, 0);
    CK(cudaStreamWaitEvent(c->copy_stream, c->ev[6], 0));
    auto upload = [&](size_t i) -> int {
        Context::KcChunk &ch = c->kc_chunks[i % nbuf];
        const size_t r0 = cuts[i], n = cuts[i + 1] - r0;
        const uint64_t base0 = offsets[r0], tot = offsets[r0 + n] - base0;
        TRY(ch.pin.ensure((n + 1) * 8));
        uint64_t *rel = ch.pin.as<uint64_t>();
        uint32_t mx = 0;
        for (size_t j = 0; j <= n; ++j) rel[j] = offsets[r0 + j] - base0;
        for (size_t j = 0; j < n; ++j) { const uint64_t l = rel[j + 1] - rel[j]; if (l >= (1ull << 22)) return fail(SAVONT_B200_ERR_INVALID, "read %zu is too long", r0 + j); if (l > mx) mx = (uint32_t)l; }
        ch.n = (uint32_t)n; ch.max_len = mx; ch.total = tot; ch.has_quals = quals != nullptr; ch.has_rc = rc != nullptr;
        TRY(ch.bases.ensure(tot + 64)); TRY(ch.offsets.ensure((n + 1) * 8));
        if (quals) TRY(ch.quals.ensure(tot + 64));
        if (rc) TRY(ch.rc.ensure(n + 16));
        // buffer reuse (non-resident mode): the kernel that last read this buffer must be done
        CK(cudaStreamWaitEvent(c->copy_stream, c->kc_ev_done[i & 1], 0));
        CK(cudaEventRecord(c->ev_h0[0], c->copy_stream));
        CK(cudaMemcpyAsync(ch.offsets.p, rel, (n + 1) * 8, cudaMemcpyHostToDevice, c->copy_stream));
        if (tot) CK(cudaMemcpyAsync(ch.bases.p, bases + base0, tot, cudaMemcpyHostToDevice, c->copy_stream));
        if (quals && tot) CK(cudaMemcpyAsync(ch.quals.p, quals + base0, tot, cudaMemcpyHostToDevice, c->copy_stream));
        if (rc && n) CK(cudaMemcpyAsync(ch.rc.p, rc + r0, n, cudaMemcpyHostToDevice, c->copy_stream));
        CK(cudaEventRecord(c->kc_ev_up[i & 1], c->copy_stream));
        st.h2d_bytes += (n + 1) * 8 + tot * (quals ? 2ull : 1ull) + (rc ? n : 0);
        return 0;
    };
    auto view = [&](size_t i, int pass, KcView *v) -> int {
        Context::KcChunk &ch = c->kc_chunks[i % nbuf];
        if (pass == 0) {                                              // metadata only (max_len of the chunk)
            v->in = SeqBatch{}; v->rc = nullptr;
            uint32_t mx = 0;
            for (size_t r = cuts[i]; r < cuts[i + 1]; ++r) { const uint64_t l = offsets[r + 1] - offsets[r]; if (l > mx) mx = (uint32_t)std::min<uint64_t>(l, 0xffffffffu); }
            v->in.max_len = mx;
            return 0;
        }
        if (pass == 1 || !resident) {
            if (!(resident && uploaded[i % nbuf])) {
                // non-resident: the pin buffer of this slot is rewritten, so the previous copy out of it must be complete
                if (!resident) CK(cudaStreamSynchronize(c->copy_stream));
                TRY(upload(i));
                uploaded[i % nbuf] = 1;
            }
            CK(cudaStreamWaitEvent(c->stream, c->kc_ev_up[i & 1], 0));
        }
        v->in = SeqBatch{};
        v->in.bases = ch.bases.as<uint8_t>(); v->in.quals = ch.has_quals ? ch.quals.as<uint8_t>() : nullptr;
        v->in.offsets = ch.offsets.as<uint64_t>(); v->in.n = ch.n; v->in.max_len = ch.max_len;
        v->rc = ch.has_rc ? ch.rc.as<uint8_t>() : nullptr;
        return 0;
    };
    st.n_reads = n_reads; st.n_bases = total;
    uint32_t max_len_all = 0;
    for (size_t r = 0; r < n_reads; ++r) { const uint64_t l = offsets[r + 1] - offsets[r]; if (l > max_len_all) max_len_all = (uint32_t)std::min<uint64_t>(l, 0xffffffffu); }
    int rcode = 1;
    if (c->kc_partitioned && kmer_part_supported(c->dp.k, max_len_all))
        rcode = kc_run_part(c, nch, total, view, single_strand, out_kmers, out_counts, out_cap, n_out, st);
    if (rcode == 1) {                                                 // unsupported k, or too many distinct k-mers for the partitions
        std::fill(uploaded.begin(), uploaded.end(), 0);
        rcode = kc_run(c, nch, total, view, single_strand, out_kmers, out_counts, out_cap, n_out, st);
    }
    CK(cudaEventRecord(c->ev[7], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    st.total_ms = ev_ms(c->ev[6], c->ev[7]);
    if (stats) *stats = st;
    return rcode;
}

int savont_b200_count_split_kmers_resident(savont_b200_ctx *h, int single_strand, uint64_t *out_kmers, uint32_t *out_counts,
                                           size_t out_cap, size_t *n_out, savont_b200_kmer_stats *stats) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    Batch &b = c->resident;
    if (b.twin) return fail(SAVONT_B200_ERR_INVALID, "the resident reads are TwinReads: k-mer counting needs the FASTQ bytes");
    savont_b200_kmer_stats st; memset(&st, 0, sizeof(st));
    CK(cudaEventRecord(c->ev[6], c->stream));
    auto view = [&](size_t, int, KcView *v) -> int { v->in = seq_view(b); v->rc = nullptr; return 0; };
    st.n_reads = b.n; st.n_bases = b.total_bases;
    int rcode = 1;
    if (c->kc_partitioned && kmer_part_supported(c->dp.k, b.max_len))
        rcode = kc_run_part(c, b.n ? 1 : 0, b.total_bases, view, single_strand, out_kmers, out_counts, out_cap, n_out, st);
    if (rcode == 1) rcode = kc_run(c, b.n ? 1 : 0, b.total_bases, view, single_strand, out_kmers, out_counts, out_cap, n_out, st);
    CK(cudaEventRecord(c->ev[7], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    st.total_ms = ev_ms(c->ev[6], c->ev[7]);
    if (stats) *stats = st;
    return rcode;
}

}  // extern "C"
