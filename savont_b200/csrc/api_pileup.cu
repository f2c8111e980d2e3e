// api_pileup.cu -- host side of savont_b200_consensus_pileups (Stage 4 pileups).
//
// Stands in for alignment::generate_consensus_pileups (reference src/alignment.rs:415-641).  Per chunk of consensus
// sequences (bounded device workspace): upload the chunk's reads, homopolymer-compress them on the device (pileup.cu),
// align every compressed read to its consensus on both strands (align_nm_kernel), keep the better strand, trace it back
// (align_tb_kernel + device CIGAR compaction), walk the CIGARs into columns and take the per-column median run length.
#include <chrono>

#include "api_internal.cuh"

namespace sb {

namespace {

// device scratch of the call, kept in Context::pileup_ws between calls
struct PileupScratch {
    DevBuf &hcnt, &hoff32, &scan_tmp, &hc, &hq, &hl;
    DevBuf &cplanes, &cplane_off, &clen;
    DevBuf &pr2, &pa2, &ps2, &nm2, &ae2, &re2, &pr, &pa, &ps;
    DevBuf &cstart, &rstart, &rend, &cig_off, &cig, &row_off, &cons_pos_off, &cons_aln_off, &pos_cons, &cnt, &excl, &col_total, &pos_off,
        &entries, &ref_hp, &entries2;
    explicit PileupScratch(DevBuf *w)
        : hcnt(w[0]), hoff32(w[1]), scan_tmp(w[2]), hc(w[3]), hq(w[4]), hl(w[5]), cplanes(w[6]), cplane_off(w[7]), clen(w[8]), pr2(w[9]),
          pa2(w[10]), ps2(w[11]), nm2(w[12]), ae2(w[13]), re2(w[14]), pr(w[15]), pa(w[16]), ps(w[17]), cstart(w[18]), rstart(w[19]),
          rend(w[20]), cig_off(w[21]), cig(w[22]), row_off(w[23]), cons_pos_off(w[24]), cons_aln_off(w[25]), pos_cons(w[26]), cnt(w[27]),
          excl(w[28]), col_total(w[29]), pos_off(w[30]), entries(w[31]), ref_hp(w[32]), entries2(w[33]) {}
};

template <typename T>
int up(DevBuf &d, const std::vector<T> &v, cudaStream_t st) {
    TRY(d.ensure(v.size() * sizeof(T) + 16));
    if (!v.empty()) CK(cudaMemcpyAsync(d.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return 0;
}

// one chunk: consensus sequences [c0, c1), reads [r0, r1) (exactly the reads whose read_cons lies in the chunk)
int pileup_chunk(Context *c, PileupScratch &s, Batch &rb, size_t chunk, int use_hpc, const uint8_t *cons_bases, const uint64_t *cons_offsets, size_t c0, size_t c1,
                 const uint8_t *read_bases, const uint8_t *read_quals, const uint64_t *read_offsets, const uint32_t *read_cons, size_t r0,
                 size_t r1, uint64_t *pos_offsets, savont_b200_pileup_entry *entries, size_t entry_cap, uint64_t *n_entries,
                 uint8_t *ref_hp_length, uint8_t *read_strand, int32_t *read_nm) {
    cudaStream_t st = c->stream;
    cudaEvent_t *ev = c->ev;       // 0 start, 1 reads compressed + packed, 2 both-strand NM, 3 traceback, 4 columns
    const size_t n = r1 - r0, nc = c1 - c0;
    const uint64_t pos0 = cons_offsets[c0] - cons_offsets[0];
    const size_t n_pos = (size_t)(cons_offsets[c1] - cons_offsets[c0]);
    for (size_t r = r0; r < r1; ++r) { read_strand[r] = 0xff; read_nm[r] = -1; }
    auto empty_columns = [&]() {
        for (size_t p = 0; p < n_pos; ++p) { pos_offsets[pos0 + p + 1] = *n_entries; ref_hp_length[pos0 + p] = 1; }
    };
    if (n == 0 || n_pos == 0) { empty_columns(); return 0; }
    if (n_pos >= (1ull << 31)) return fail(SAVONT_B200_ERR_INVALID, "consensus chunk too long");

    // ---- reads: upload, homopolymer compression (src/utils.rs:135-184) ----
    Batch &hb = c->work2;                    // rb: the chunk's raw reads, uploaded by the caller on the copy stream (one chunk ahead)
    static const bool trace = getenv("SAVONT_B200_PILEUP_TRACE") != nullptr;        // host wall clock of the first stage's steps
    auto tnow = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tl = tnow(), tw[6] = {0}; auto lap = [&](int k) { if (trace) { CK(cudaStreamSynchronize(st)); const double x = tnow(); tw[k] += x - tl; tl = x; } return 0; };
    CK(cudaEventRecord(ev[0], st));
    (void)read_bases; (void)read_quals; (void)read_offsets;
    TRY(lap(0));
    TRY(s.hcnt.ensure((n + 1) * 4)); TRY(s.hoff32.ensure((n + 2) * 4)); TRY(s.scan_tmp.ensure((n / 2048 + 4) * 4));
    launch_hpc_count(seq_view(rb), use_hpc, s.hcnt.as<uint32_t>(), st);
    launch_scan_u32(s.hcnt.as<uint32_t>(), s.hoff32.as<uint32_t>(), (uint32_t)n, s.scan_tmp.as<uint32_t>(), st);
    std::vector<uint32_t> hoff32(n + 1);
    CK(cudaMemcpyAsync(hoff32.data(), s.hoff32.p, (n + 1) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint64_t htotal = hoff32[n];
    TRY(lap(1));
    TRY(s.hc.ensure(htotal + 64)); TRY(s.hq.ensure(htotal + 64)); TRY(s.hl.ensure(htotal + 64));
    CK(cudaMemsetAsync(s.hc.as<uint8_t>() + htotal, 0, 64, st));       // the plane packing reads whole words past the last base
    launch_hpc_write(seq_view(rb), use_hpc, s.hoff32.as<uint32_t>(), s.hc.as<uint8_t>(), s.hq.as<uint8_t>(), s.hl.as<uint8_t>(), st);
    c->last_stats.kernel_launches += 5;
    // the compressed reads as a batch of their own (device-resident bases)
    TRY(hb.pin_offsets.ensure((n + 1) * 8));
    uint64_t *hoff = hb.pin_offsets.as<uint64_t>();
    uint32_t hmax = 0;
    for (size_t i = 0; i <= n; ++i) { hoff[i] = hoff32[i]; if (i) hmax = std::max(hmax, hoff32[i] - hoff32[i - 1]); }
    hb.n = (uint32_t)n; hb.total_bases = htotal; hb.max_len = hmax; hb.mapped = false; hb.has_quals = false; hb.has_file_idx = false;
    hb.twin = false; hb.force_generic = false;
    TRY(hb.offsets.ensure((n + 1) * 8));
    CK(cudaMemcpyAsync(hb.offsets.p, hoff, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    std::swap(hb.bases, s.hc);                 // seq_view(hb) reads hb.bases; swapped back before returning
    struct SwapBack { DevBuf &a, &b; ~SwapBack() { std::swap(a, b); } } swap_back{hb.bases, s.hc};
    TRY(hb.planes.ensure(((size_t)plane_word_off(htotal, n) + 4) * sizeof(uint2)));
    launch_pack_planes(seq_view(hb), hb.planes.as<uint2>(), st);
    c->last_stats.kernel_launches += 1;
    c->last_stats.n_reads += n; c->last_stats.n_bases += rb.total_bases;
    TRY(lap(2));

    // ---- consensus sequences: DP planes ----
    std::vector<uint2> planes; std::vector<uint32_t> plane_off, clen;
    uint32_t max_clen = 0;
    std::vector<uint64_t> crel(nc + 1);
    for (size_t i = 0; i <= nc; ++i) crel[i] = cons_offsets[c0 + i] - cons_offsets[c0];
    build_dp_planes(c->dp, cons_bases + cons_offsets[c0], crel.data(), nc, planes, plane_off, clen, &max_clen);
    TRY(up(s.cplanes, planes, st)); TRY(up(s.cplane_off, plane_off, st)); TRY(up(s.clen, clen, st));
    AsvIndex ix{};
    ix.planes = s.cplanes.as<uint2>(); ix.plane_off = s.cplane_off.as<uint32_t>(); ix.len = s.clen.as<uint32_t>(); ix.n_asv = (uint32_t)nc;

    // ---- both strands of every non-empty read against its consensus; minimap2 returns Err for an empty sequence and
    //      the reference skips the read (src/alignment.rs:486) ----
    std::vector<uint32_t> pr2, pa2; std::vector<uint8_t> ps2;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t ci = read_cons[r0 + i] - (uint32_t)c0;
        if (hoff32[i + 1] == hoff32[i] || clen[ci] == 0) continue;
        for (int sd = 0; sd < 2; ++sd) { pr2.push_back((uint32_t)i); pa2.push_back(ci); ps2.push_back((uint8_t)sd); }
    }
    const uint32_t n_aln = (uint32_t)(pr2.size() / 2);
    if (n_aln == 0) { CK(cudaStreamSynchronize(st)); empty_columns(); return 0; }
    TRY(up(s.pr2, pr2, st)); TRY(up(s.pa2, pa2, st)); TRY(up(s.ps2, ps2, st));
    TRY(s.nm2.ensure((size_t)n_aln * 8)); TRY(s.ae2.ensure((size_t)n_aln * 8)); TRY(s.re2.ensure((size_t)n_aln * 8));
    TRY(s.pr.ensure((size_t)n_aln * 4)); TRY(s.pa.ensure((size_t)n_aln * 4)); TRY(s.ps.ensure((size_t)n_aln + 4));
    AlignOut ao{s.nm2.as<int32_t>(), s.ae2.as<int32_t>(), s.re2.as<int32_t>()};
    TRY(lap(3));
    if (trace) fprintf(stderr, "pileup chunk %zu host ms: upload %.2f count+scan %.2f write+pack %.2f planes+pairs %.2f\n", chunk, tw[0], tw[1], tw[2], tw[3]);
    CK(cudaEventRecord(ev[1], st));
    launch_align_nm(seq_view(hb), hb.planes.as<uint2>(), ix, s.pr2.as<uint32_t>(), s.pa2.as<uint32_t>(), s.ps2.as<uint8_t>(), nullptr,
                    2u * n_aln, nullptr, ao, c->dp, c->sm_count, st);
    launch_pick_strand(s.nm2.as<int32_t>(), s.pr2.as<uint32_t>(), s.pa2.as<uint32_t>(), n_aln, s.pr.as<uint32_t>(), s.pa.as<uint32_t>(),
                       s.ps.as<uint8_t>(), st);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev[2], st));
    c->last_stats.kernel_launches += 2;
    c->last_stats.n_pairs += 2ull * n_aln;
    for (uint32_t t = 0; t < n_aln; ++t) c->last_stats.dp_cells += 2ull * clen[pa2[2 * t]] * (2ull * (uint64_t)c->dp.band + 1ull);
    std::vector<uint8_t> strand(n_aln);
    CK(cudaMemcpyAsync(strand.data(), s.ps.p, n_aln, cudaMemcpyDeviceToHost, st));

    // ---- traceback of the chosen strand ----
    std::vector<int32_t> v((size_t)n_aln * 5);
    std::vector<uint64_t> coff((size_t)n_aln + 1, 0);
    std::vector<uint32_t> cigar;
    TbDevice tb;                                      // one sub-batch (the usual case): the ops never leave the device
    const bool host_cigar = getenv("SAVONT_B200_PILEUP_HOST_CIGAR") != nullptr;   // tests: the path of a traceback in several sub-batches
    TRY(traceback_pairs(c, hb, ix, max_clen, s.pr.as<uint32_t>(), s.pa.as<uint32_t>(), s.ps.as<uint8_t>(), n_aln, v.data(),
                        v.data() + n_aln, v.data() + 2 * (size_t)n_aln, v.data() + 3 * (size_t)n_aln, v.data() + 4 * (size_t)n_aln,
                        coff.data(), nullptr, 0, &cigar, host_cigar ? nullptr : &tb));
    if (coff[n_aln] >= (1ull << 32)) return fail(SAVONT_B200_ERR_OVERFLOW, "CIGAR ops of one chunk exceed 4 Gi");
    CK(cudaEventRecord(ev[3], st));
    for (uint32_t t = 0; t < n_aln; ++t) { read_strand[r0 + pr2[2 * t]] = strand[t]; read_nm[r0 + pr2[2 * t]] = v[t]; }

    // ---- columns ----
    std::vector<uint32_t> cig_off32((size_t)n_aln + 1), cons_pos_off(nc + 1), cons_aln_off(nc + 1, 0), pos_cons(n_pos);
    std::vector<uint64_t> row_off(n_aln);
    for (size_t t = 0; t <= n_aln; ++t) cig_off32[t] = (uint32_t)coff[t];
    uint64_t rows = 0;
    for (uint32_t t = 0; t < n_aln; ++t) { row_off[t] = rows; rows += clen[pa2[2 * t]]; cons_aln_off[pa2[2 * t] + 1]++; }
    for (size_t i = 0; i < nc; ++i) {
        cons_aln_off[i + 1] += cons_aln_off[i];
        cons_pos_off[i] = (uint32_t)crel[i];
        for (uint64_t p = crel[i]; p < crel[i + 1]; ++p) pos_cons[p] = (uint32_t)i;
    }
    cons_pos_off[nc] = (uint32_t)crel[nc];
    if (!tb.valid) { cigar.resize(coff[n_aln]); TRY(up(s.cig_off, cig_off32, st)); TRY(up(s.cig, cigar, st)); }
    TRY(up(s.row_off, row_off, st));
    TRY(up(s.cons_pos_off, cons_pos_off, st)); TRY(up(s.cons_aln_off, cons_aln_off, st)); TRY(up(s.pos_cons, pos_cons, st));
    std::vector<int32_t> cs(v.begin() + 3 * (size_t)n_aln, v.begin() + 4 * (size_t)n_aln), rs(v.begin() + 4 * (size_t)n_aln, v.end()),
        re(v.begin() + 2 * (size_t)n_aln, v.begin() + 3 * (size_t)n_aln);
    if (!tb.valid) { TRY(up(s.cstart, cs, st)); TRY(up(s.rstart, rs, st)); TRY(up(s.rend, re, st)); }
    TRY(s.cnt.ensure(rows + 16)); TRY(s.excl.ensure(rows * 2 + 16));
    TRY(s.col_total.ensure((n_pos + 1) * 4)); TRY(s.pos_off.ensure((n_pos + 2) * 4)); TRY(s.scan_tmp.ensure((n_pos / 2048 + 4) * 4));
    TRY(s.ref_hp.ensure(n_pos + 16));
    CK(cudaMemsetAsync(s.cnt.p, 0, rows + 16, st));
    PileupArgs a{};
    a.n_aln = n_aln; a.aln_read = s.pr.as<uint32_t>(); a.aln_cons = s.pa.as<uint32_t>(); a.strand = s.ps.as<uint8_t>();
    a.cons_start = s.cstart.as<int32_t>(); a.read_start = s.rstart.as<int32_t>(); a.read_end = s.rend.as<int32_t>();
    a.cigar_off = s.cig_off.as<uint32_t>(); a.cigar = s.cig.as<uint32_t>(); a.row_off = s.row_off.as<uint64_t>();
    if (tb.valid) {
        a.cons_start = tb.out + 3 * (size_t)n_aln; a.read_start = tb.out + 4 * (size_t)n_aln; a.read_end = tb.out + 2 * (size_t)n_aln;
        a.cigar_off = tb.off; a.cigar = tb.ops;
    }
    a.hpc_offsets = hb.offsets.as<uint64_t>(); a.hc = hb.bases.as<uint8_t>(); a.hq = s.hq.as<uint8_t>(); a.hl = s.hl.as<uint8_t>();
    a.cons_len = s.clen.as<uint32_t>(); a.cons_pos_off = s.cons_pos_off.as<uint32_t>(); a.cons_aln_off = s.cons_aln_off.as<uint32_t>();
    a.pos_cons = s.pos_cons.as<uint32_t>(); a.n_pos = (uint32_t)n_pos; a.read_base = (uint32_t)r0;
    a.cnt = s.cnt.as<uint8_t>(); a.excl = s.excl.as<uint16_t>(); a.col_total = s.col_total.as<uint32_t>();
    a.pos_off = s.pos_off.as<uint32_t>(); a.entries = nullptr;
    launch_pileup_count(a, st);
    launch_pileup_rank(a, st);
    launch_scan_u32(s.col_total.as<uint32_t>(), s.pos_off.as<uint32_t>(), (uint32_t)n_pos, s.scan_tmp.as<uint32_t>(), st);
    std::vector<uint32_t> pos_off(n_pos + 1);
    CK(cudaMemcpyAsync(pos_off.data(), s.pos_off.p, (n_pos + 1) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint64_t ne = pos_off[n_pos];
    if (*n_entries + ne > entry_cap)
        return fail(SAVONT_B200_ERR_CAPACITY, "pileup entry capacity %zu too small (need >= %llu)", entry_cap,
                    (unsigned long long)(*n_entries + ne));
    // two entry buffers: the copy of chunk i to the host (d2h_stream) runs under the kernels of chunk i + 1
    DevBuf &dent = (chunk & 1) ? s.entries2 : s.entries;
    cudaEvent_t ev_free = c->ev_d2h[chunk & 1];               // recorded after the previous copy out of this buffer
    if (chunk >= 2) CK(cudaEventSynchronize(ev_free));
    TRY(dent.ensure(ne * sizeof(PileupEntry) + 16));
    a.entries = dent.as<PileupEntry>();
    launch_pileup_emit(a, st);
    launch_pileup_median(a, s.ref_hp.as<uint8_t>(), st);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev[4], st));
    c->last_stats.kernel_launches += 7;
    CK(cudaMemcpyAsync(ref_hp_length + pos0, s.ref_hp.p, n_pos, cudaMemcpyDeviceToHost, st));
    for (size_t p = 0; p < n_pos; ++p) pos_offsets[pos0 + p + 1] = *n_entries + pos_off[p + 1];
    CK(cudaStreamSynchronize(st));
    if (ne) {
        // page-locked destination: SM copy (see copy_to_host_kernel); pageable: the copy engine's staged path
        cudaPointerAttributes at{};
        const bool pinned = cudaPointerGetAttributes(&at, entries) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer != nullptr;
        if (!pinned) cudaGetLastError();
        if (pinned) {
            launch_copy_to_host(dent.p, reinterpret_cast<uint8_t *>(at.devicePointer) + *n_entries * sizeof(PileupEntry), ne * sizeof(PileupEntry),
                                c->d2h_stream);
            CK(cudaGetLastError());
            c->last_stats.kernel_launches += 1;
        } else {
            CK(cudaMemcpyAsync(entries + *n_entries, dent.p, ne * sizeof(PileupEntry), cudaMemcpyDeviceToHost, c->d2h_stream));
        }
    }
    CK(cudaEventRecord(ev_free, c->d2h_stream));
    c->last_stats.d2h_bytes += ne * sizeof(PileupEntry) + n_pos;
    c->last_stats.sketch_ms += ev_ms(ev[0], ev[1]); c->last_stats.align_ms += ev_ms(ev[1], ev[2]);
    c->last_stats.filter_ms += ev_ms(ev[2], ev[3]); c->last_stats.assign_ms += ev_ms(ev[3], ev[4]);
    c->last_stats.n_batches += 1;
    *n_entries += ne;
    return 0;
}

}  // namespace

}  // namespace sb

using namespace sb;

extern "C" int savont_b200_consensus_pileups(savont_b200_ctx *h, int use_hpc, const uint8_t *cons_bases, const uint64_t *cons_offsets,
                                             size_t n_cons, const uint8_t *read_bases, const uint8_t *read_quals,
                                             const uint64_t *read_offsets, const uint32_t *read_cons, size_t n_reads,
                                             uint64_t *pos_offsets, savont_b200_pileup_entry *entries, size_t entry_cap,
                                             uint64_t *n_entries, uint8_t *ref_hp_length, uint8_t *read_strand, int32_t *read_nm,
                                             savont_b200_stats *stats) {
    static_assert(sizeof(savont_b200_pileup_entry) == sizeof(PileupEntry) && sizeof(PileupEntry) == 12, "pileup entry layout");
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!cons_offsets || !read_offsets || !pos_offsets || !n_entries || (n_reads && (!read_cons || !read_strand || !read_nm)))
        return fail(SAVONT_B200_ERR_INVALID, "null argument");
    const uint64_t total_pos = n_cons ? cons_offsets[n_cons] - cons_offsets[0] : 0;
    if (total_pos && (!cons_bases || !ref_hp_length)) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    for (size_t i = 0; i < n_cons; ++i)
        if (cons_offsets[i + 1] < cons_offsets[i]) return fail(SAVONT_B200_ERR_INVALID, "consensus offsets not monotonic at %zu", i);
    for (size_t r = 0; r < n_reads; ++r) {
        if (read_cons[r] >= n_cons) return fail(SAVONT_B200_ERR_INVALID, "read %zu: consensus %u out of range", r, read_cons[r]);
        if (r && read_cons[r] < read_cons[r - 1]) return fail(SAVONT_B200_ERR_INVALID, "read_cons must be non-decreasing (read %zu)", r);
        if (read_offsets[r + 1] < read_offsets[r]) return fail(SAVONT_B200_ERR_INVALID, "read offsets not monotonic at %zu", r);
    }
    c->last_stats = savont_b200_stats{};
    *n_entries = 0; pos_offsets[0] = 0;
    // chunks of consecutive consensus sequences: <= 2^27 (read, position) cells of column workspace, <= 32768 reads
    // (a single consensus beyond these bounds forms a chunk of its own)
    constexpr uint64_t kMaxCells = 1ull << 27;
    size_t kMaxReads = 32768;
    if (const char *e = getenv("SAVONT_B200_PILEUP_CHUNK_READS")) kMaxReads = std::max<size_t>(1, strtoull(e, nullptr, 10));   // tests
    PileupScratch s(c->pileup_ws);
    int rc = 0;
    struct Chunk { size_t c0, c1, r0, r1; };
    std::vector<Chunk> chunks;
    {
        size_t c0 = 0, r0 = 0;
        while (c0 < n_cons && rc == 0) {
            size_t c1 = c0, r1 = r0;
            uint64_t cells = 0;
            while (c1 < n_cons) {
                size_t re = r1;
                while (re < n_reads && read_cons[re] == c1) ++re;
                if (re - r1 > 32767) { rc = fail(SAVONT_B200_ERR_INVALID, "consensus %zu: %zu reads (at most 32767 per consensus)", c1, re - r1); break; }
                const uint64_t add = (uint64_t)(re - r1) * (cons_offsets[c1 + 1] - cons_offsets[c1]);
                if (c1 > c0 && (cells + add > kMaxCells || re - r0 > kMaxReads)) break;
                cells += add; r1 = re; ++c1;
            }
            if (rc) break;
            chunks.push_back(Chunk{c0, c1, r0, r1});
            c0 = c1; r0 = r1;
        }
    }
    // The raw reads of chunk i + 1 are uploaded (copy stream, second buffer) while chunk i is processed: a chunk's steps are separated
    // by host synchronisations (sizes come back from the device), so without this the link idles during the kernels and vice versa.
    Batch *raw[2] = {&c->work, &c->work3};
    auto upload = [&](size_t ci) -> int {
        const Chunk &k = chunks[ci];
        if (k.r1 > k.r0) TRY(batch_upload(c, *raw[ci & 1], read_bases, read_quals, read_offsets + k.r0, nullptr, k.r1 - k.r0, &c->last_stats.h2d_bytes, c->copy_stream));
        CK(cudaEventRecord(c->ev_up[ci & 1], c->copy_stream));
        return 0;
    };
    if (rc == 0 && !chunks.empty()) rc = upload(0);
    for (size_t ci = 0; ci < chunks.size() && rc == 0; ++ci) {
        const Chunk &k = chunks[ci];
        // chunk ci - 1 (the previous user of the other buffer) has returned: its kernels on the raw reads are complete
        if (ci + 1 < chunks.size()) { rc = upload(ci + 1); if (rc) break; }
        if (cudaStreamWaitEvent(c->stream, c->ev_up[ci & 1], 0) != cudaSuccess) { rc = fail(SAVONT_B200_ERR_CUDA, "pileup: wait for the upload failed"); break; }
        rc = pileup_chunk(c, s, *raw[ci & 1], ci, use_hpc, cons_bases, cons_offsets, k.c0, k.c1, read_bases, read_quals, read_offsets, read_cons, k.r0, k.r1,
                          pos_offsets, entries, entry_cap, n_entries, ref_hp_length, read_strand, read_nm);
    }
    if (rc) cudaStreamSynchronize(c->copy_stream);          // an upload may still be in flight
    {   // the last copies of the entries
        cudaError_t e = cudaStreamSynchronize(c->d2h_stream);
        if (e != cudaSuccess && rc == 0) rc = fail(SAVONT_B200_ERR_CUDA, "pileup entries copy: %s", cudaGetErrorString(e));
    }
    c->last_stats.total_ms = c->last_stats.sketch_ms + c->last_stats.align_ms + c->last_stats.filter_ms + c->last_stats.assign_ms;
    if (stats) *stats = c->last_stats;
    return rc;
}
