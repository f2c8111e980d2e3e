// api_cluster.cu -- host side of savont_b200_lsh_signatures / savont_b200_cluster_reads_by_kmers.
//
// Stands in for asv_cluster::cluster_reads_by_kmers (reference src/asv_cluster.rs:72-190, bucketed branch).  The reference is
// a sequential greedy loop: read i is compared with the representatives created by reads < i and either joins the best one
// or becomes a representative itself.  Here the reads go through in blocks:
//   1. rep_query_kernel compares every read of the block with the representatives known BEFORE the block (bucket hits,
//      candidate window, containment counts of the window);
//   2. the reads that find no match there are the block's potential representatives; pair_stats_kernel gives hits and
//      containment of every (read, earlier potential representative) pair of the block;
//   3. the host walks the block in read order and merges, per read, the pre-block candidates with the representatives that were
//      actually created earlier in the block -- the reference's order (hits, id descending; in-block representatives have the
//      larger ids), window (max-hits group + first 10) and "first strictly greater similarity" rule.  A read that becomes a
//      representative although it had a pre-block match (its match was pushed out of the first 10 by in-block representatives) is
//      not among the pairs of step 2: its pairs are computed on demand.
// The result is identical to the sequential loop; block sizes follow the rate of new representatives.
#include <chrono>

#include "api_internal.cuh"

namespace sb {

namespace {

struct Cand { uint32_t id, hits, cnt, den; };
inline bool ratio_gt(const Cand &a, const Cand &b) { return (uint64_t)a.cnt * b.den > (uint64_t)b.cnt * a.den; }   // a.cnt/a.den > b.cnt/b.den

struct ThresholdTest {       // similarity = ratio.powf(1/k) > threshold (src/asv_cluster.rs:143-152), pow only near the boundary
    double threshold, inv_k, t_pow;
    ThresholdTest(double thr, int k) : threshold(thr), inv_k(1.0 / (double)k), t_pow(thr > 0 ? std::pow(thr, (double)k) : 0.0) {}
    bool pass(uint32_t cnt, uint32_t den) const {
        if (cnt == 0) return false;                          // similarity 0.0 is never > best_similarity = 0.0
        const double ratio = (double)cnt / (double)den;
        if (threshold > 0) {
            if (ratio > t_pow * (1.0 + 1e-9)) return true;
            if (ratio < t_pow * (1.0 - 1e-9)) return false;
        }
        return std::pow(ratio, inv_k) > threshold;
    }
};

template <typename T>
int up(DevBuf &d, const T *v, size_t n, cudaStream_t st) {
    TRY(d.ensure(n * sizeof(T) + 16));
    if (n) CK(cudaMemcpyAsync(d.p, v, n * sizeof(T), cudaMemcpyHostToDevice, st));
    return 0;
}

struct ClusterDev {
    DevBuf &kmers, &offsets, &dk, &nd, &sig, &valid, &tsig, &trep, &toff, &res, &pi, &pj, &ph, &pc, &dsig, &drep;
    explicit ClusterDev(DevBuf *w)
        : kmers(w[0]), offsets(w[1]), dk(w[2]), nd(w[3]), sig(w[4]), valid(w[5]), tsig(w[6]), trep(w[7]), toff(w[8]), res(w[9]), pi(w[10]),
          pj(w[11]), ph(w[12]), pc(w[13]), dsig(w[14]), drep(w[15]) {}
};

// upload + signatures + sorted distinct lists; host copies of nd / sig / valid when asked for
int cluster_prepare(Context *c, ClusterDev &d, const uint64_t *kmers, const uint64_t *offsets, size_t n, std::vector<uint64_t> &rel,
                    std::vector<uint32_t> *nd_h, std::vector<uint64_t> *sig_h, std::vector<uint8_t> *valid_h) {
    cudaStream_t st = c->stream;
    const uint64_t base0 = offsets[0], total = offsets[n] - base0;
    rel.resize(n + 1);
    for (size_t i = 0; i <= n; ++i) {
        if (i && offsets[i] < offsets[i - 1]) return fail(SAVONT_B200_ERR_INVALID, "offsets not monotonic at %zu", i);
        rel[i] = offsets[i] - base0;
    }
    TRY(up(d.kmers, kmers + base0, total, st)); TRY(up(d.offsets, rel.data(), n + 1, st));
    TRY(d.dk.ensure(total * 8 + 16)); TRY(d.nd.ensure(n * 4 + 16)); TRY(d.sig.ensure(n * kLshTables * 8 + 16)); TRY(d.valid.ensure(n + 16));
    launch_lsh_sig(d.kmers.as<uint64_t>(), d.offsets.as<uint64_t>(), (uint32_t)n, d.sig.as<uint64_t>(), d.valid.as<uint8_t>(), st);
    c->last_stats.kernel_launches += 1;
    c->last_stats.h2d_bytes += total * 8 + (n + 1) * 8;
    if (sig_h) {
        sig_h->resize(n * kLshTables);
        if (n) CK(cudaMemcpyAsync(sig_h->data(), d.sig.p, n * kLshTables * 8, cudaMemcpyDeviceToHost, st));
        c->last_stats.d2h_bytes += n * kLshTables * 8;
    }
    if (valid_h) {
        valid_h->resize(n);
        if (n) CK(cudaMemcpyAsync(valid_h->data(), d.valid.p, n, cudaMemcpyDeviceToHost, st));
        c->last_stats.d2h_bytes += n;
    }
    if (nd_h) {
        launch_sort_dedup(d.kmers.as<uint64_t>(), d.offsets.as<uint64_t>(), (uint32_t)n, d.dk.as<uint64_t>(), d.nd.as<uint32_t>(), c->sm_count, st);
        c->last_stats.kernel_launches += 1;
        nd_h->resize(n);
        if (n) CK(cudaMemcpyAsync(nd_h->data(), d.nd.p, n * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        // reads with more minimizers than the shared-memory sort holds: sorted on the host
        std::vector<uint64_t> tmp;
        for (size_t r = 0; r < n; ++r) {
            if ((*nd_h)[r] != 0xffffffffu) continue;
            tmp.assign(kmers + offsets[r], kmers + offsets[r + 1]);
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            (*nd_h)[r] = (uint32_t)tmp.size();
            CK(cudaMemcpy(d.dk.as<uint64_t>() + rel[r], tmp.data(), tmp.size() * 8, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(d.nd.as<uint32_t>() + r, &(*nd_h)[r], 4, cudaMemcpyHostToDevice));
        }
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return 0;
}

}  // namespace

}  // namespace sb

using namespace sb;

extern "C" int savont_b200_lsh_signatures(savont_b200_ctx *h, const uint64_t *kmers, const uint64_t *offsets, size_t n_reads, uint64_t *sig,
                                          uint8_t *valid) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!offsets || (n_reads && (!sig || !valid))) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    if (n_reads >= 0xfffffff0ull) return fail(SAVONT_B200_ERR_INVALID, "too many reads");
    if (offsets[n_reads] > offsets[0] && !kmers) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    c->last_stats = savont_b200_stats{};
    ClusterDev d(c->cluster_ws);
    std::vector<uint64_t> rel, sig_h; std::vector<uint8_t> valid_h;
    TRY(cluster_prepare(c, d, kmers, offsets, n_reads, rel, nullptr, &sig_h, &valid_h));
    if (n_reads) { memcpy(sig, sig_h.data(), n_reads * kLshTables * 8); memcpy(valid, valid_h.data(), n_reads); }
    return 0;
}

extern "C" int savont_b200_cluster_reads_by_kmers(savont_b200_ctx *h, const uint64_t *kmers, const uint64_t *offsets, size_t n_reads,
                                                  double threshold, uint32_t *rep_of, uint64_t *n_reps_out, savont_b200_stats *stats) {
    TRY(use_device(h ? &h->c : nullptr));
    Context *c = &h->c;
    if (!offsets || (n_reads && !rep_of)) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    if (n_reads >= 0xfffffff0ull) return fail(SAVONT_B200_ERR_INVALID, "too many reads");
    if (offsets[n_reads] > offsets[0] && !kmers) return fail(SAVONT_B200_ERR_INVALID, "null argument");
    c->last_stats = savont_b200_stats{};
    cudaStream_t st = c->stream;
    const size_t n = n_reads;
    ClusterDev d(c->cluster_ws);
    std::vector<uint64_t> rel; std::vector<uint8_t> valid; std::vector<uint32_t> nd;
    CK(cudaEventRecord(c->ev[0], st));
    TRY(cluster_prepare(c, d, kmers, offsets, n, rel, &nd, nullptr, &valid));
    CK(cudaEventRecord(c->ev[1], st));
    const ThresholdTest thr(threshold, c->params.k);
    auto tot_of = [&](uint32_t r) { return (uint32_t)(rel[r + 1] - rel[r]); };

    // bucket tables on the host, two levels: `tab` (main, rebuilt when the delta has grown to a quarter of it) and `dlt` (the
    // representatives created since, all with larger ids); per table (signature, representative) sorted.  Flat copies go to
    // the device: the delta after every block that created representatives, the main level only when rebuilt.
    typedef std::pair<uint64_t, uint32_t> Ent;
    std::vector<Ent> tab[kLshTables], dlt[kLshTables], fresh[kLshTables], merged;
    std::vector<uint64_t> flat_sig; std::vector<uint32_t> flat_rep;
    uint32_t toff[2 * (kLshTables + 1)] = {0};             // main offsets, then delta offsets (one device array)
    size_t n_main = 0, n_delta = 0;
    auto flatten = [&](std::vector<Ent> *lvl, uint32_t *off, DevBuf &dsg, DevBuf &drp) -> int {
        size_t e = 0;
        for (int t = 0; t < kLshTables; ++t) e += lvl[t].size();
        flat_sig.resize(e); flat_rep.resize(e);
        e = 0;
        for (int t = 0; t < kLshTables; ++t) {
            off[t] = (uint32_t)e;
            for (const Ent &x : lvl[t]) { flat_sig[e] = x.first; flat_rep[e] = x.second; ++e; }
        }
        off[kLshTables] = (uint32_t)e;
        TRY(up(dsg, flat_sig.data(), e, st)); TRY(up(drp, flat_rep.data(), e, st));
        CK(cudaStreamSynchronize(st));                     // flat_* are reused for the other level
        c->last_stats.h2d_bytes += e * 12;
        return 0;
    };
    bool main_dirty = true;
    bool tables_dirty = true;
    uint64_t n_reps = 0;
    size_t block = 256;
    if (const char *e = getenv("SAVONT_B200_CLUSTER_BLOCK")) block = std::max<size_t>(1, strtoull(e, nullptr, 10));   // tests
    const bool fixed_block = getenv("SAVONT_B200_CLUSTER_BLOCK") != nullptr;
    constexpr size_t kMaxBlock = 16384, kMinBlock = 256, kMaxPairs = 1u << 24;
    size_t kTargetNew = 160;
    if (const char *e = getenv("SAVONT_B200_CLUSTER_TARGET")) kTargetNew = std::max<size_t>(1, strtoull(e, nullptr, 10));   // tuning
    struct PairHit { uint32_t pidx, hits, cnt; };
    std::vector<uint32_t> p0, before, pair_start, hit_off, hit_fill; std::vector<PairHit> hit_list; std::vector<uint8_t> p0_rep;
    std::vector<Cand> inblk, order;
    std::vector<uint32_t> new_ids; std::vector<uint64_t> new_sig;
    double query_ms = 0, pair_ms = 0;
    // host-side wall clock per phase (SAVONT_B200_CLUSTER_TRACE=1 prints it)
    double tw[8] = {0}; auto tnow = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tl = tnow(); auto lap = [&](int k) { const double x = tnow(); tw[k] += x - tl; tl = x; };

    for (size_t b0 = 0; b0 < n;) {
        size_t b1 = std::min(n, b0 + block);
        // ---- 1. the block against the representatives known so far ----
        if (tables_dirty) {
            if (main_dirty) { TRY(flatten(tab, toff, d.tsig, d.trep)); main_dirty = false; }
            TRY(flatten(dlt, toff + kLshTables + 1, d.dsig, d.drep));
            TRY(up(d.toff, toff, (size_t)2 * (kLshTables + 1), st));
            tables_dirty = false;
        }
        lap(0);
        const uint32_t nb = (uint32_t)(b1 - b0);
        TRY(d.res.ensure((size_t)nb * sizeof(ClusterResult)));
        ClusterQuery q{};
        q.offsets = d.offsets.as<uint64_t>(); q.dk = d.dk.as<uint64_t>(); q.nd = d.nd.as<uint32_t>(); q.sig = d.sig.as<uint64_t>();
        q.valid = d.valid.as<uint8_t>(); q.table_sig = d.tsig.as<uint64_t>(); q.table_rep = d.trep.as<uint32_t>();
        q.table_off = d.toff.as<uint32_t>();
        q.delta_sig = d.dsig.as<uint64_t>(); q.delta_rep = d.drep.as<uint32_t>(); q.delta_off = d.toff.as<uint32_t>() + kLshTables + 1;
        q.first = (uint32_t)b0; q.n_block = nb; q.out = d.res.as<ClusterResult>();
        CK(cudaEventRecord(c->ev[2], st));
        launch_rep_query(q, st);
        CK(cudaEventRecord(c->ev[3], st));
        TRY(c->cluster_pin[0].ensure((size_t)nb * sizeof(ClusterResult)));
        const ClusterResult *res = c->cluster_pin[0].as<ClusterResult>();
        CK(cudaMemcpyAsync(c->cluster_pin[0].p, d.res.p, (size_t)nb * sizeof(ClusterResult), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        query_ms += ev_ms(c->ev[2], c->ev[3]);
        c->last_stats.kernel_launches += 1; c->last_stats.d2h_bytes += (size_t)nb * sizeof(ClusterResult);

        lap(1);
        // ---- 2. potential representatives of the block and their pairs ----
        p0.clear(); before.assign(nb, 0); pair_start.assign((size_t)nb + 1, 0);
        for (uint32_t li = 0; li < nb; ++li) {
            before[li] = (uint32_t)p0.size();
            if ((size_t)pair_start[li] + before[li] > kMaxPairs && li > 0) { b1 = b0 + li; break; }   // bounded pair workspace: cut the block here
            pair_start[li + 1] = pair_start[li] + before[li];
            const ClusterResult &r = res[li];
            c->last_stats.dp_cells += r.n_entries;
            const bool match = r.best_rep != 0xffffffffu && thr.pass(r.best_cnt, r.best_den);
            if (!match) p0.push_back((uint32_t)(b0 + li));
        }
        const uint32_t nbk = (uint32_t)(b1 - b0);
        const size_t n_pairs = (size_t)pair_start[nbk];
        lap(2);
        // pairs sharing a bucket, bucketed by read: hit_off[li] .. hit_off[li + 1] into hit_list (pidx, hits, count)
        hit_off.assign((size_t)nbk + 1, 0); hit_list.clear();
        if (n_pairs) {
            q.n_block = nbk;
            TRY(up(d.pi, pair_start.data(), (size_t)nbk + 1, st)); TRY(up(d.pj, p0.data(), p0.size(), st));
            TRY(d.ph.ensure(n_pairs * sizeof(uint4) + 16)); TRY(d.pc.ensure(16));
            CK(cudaMemsetAsync(d.pc.p, 0, 4, st));
            CK(cudaEventRecord(c->ev[2], st));
            launch_block_pair_stats(q, d.pi.as<uint32_t>(), d.pj.as<uint32_t>(), (uint32_t)n_pairs, d.ph.as<uint4>(), d.pc.as<uint32_t>(), st);
            CK(cudaEventRecord(c->ev[3], st));
            uint32_t n_hit = 0;
            CK(cudaMemcpyAsync(&n_hit, d.pc.p, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaGetLastError());
            TRY(c->cluster_pin[1].ensure((size_t)n_hit * sizeof(uint4) + 16));
            if (n_hit) CK(cudaMemcpyAsync(c->cluster_pin[1].p, d.ph.p, (size_t)n_hit * sizeof(uint4), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            const uint4 *raw = c->cluster_pin[1].as<uint4>();
            for (uint32_t x = 0; x < n_hit; ++x) hit_off[raw[x].x + 1]++;
            for (uint32_t li = 0; li < nbk; ++li) hit_off[li + 1] += hit_off[li];
            hit_list.resize(n_hit); hit_fill.assign(hit_off.begin(), hit_off.end() - 1);
            for (uint32_t x = 0; x < n_hit; ++x) hit_list[hit_fill[raw[x].x]++] = PairHit{raw[x].y, raw[x].z, raw[x].w};
            pair_ms += ev_ms(c->ev[2], c->ev[3]);
            c->last_stats.kernel_launches += 1; c->last_stats.n_pairs += n_pairs;
            c->last_stats.h2d_bytes += ((size_t)nbk + 1 + p0.size()) * 4; c->last_stats.d2h_bytes += (size_t)n_hit * sizeof(uint4);
        }
        p0_rep.assign(p0.size(), 0);
        lap(3);

        // ---- 3. the block in read order ----
        size_t n_new = 0;
        struct Extra { uint32_t id; std::vector<uint8_t> hits; std::vector<uint32_t> cnt; };   // a representative outside p0 and the
        std::vector<Extra> extras;                                                              // stats of the reads behind it
        size_t p0_pos = 0;
        for (uint32_t li = 0; li < nbk; ++li) {
            const uint32_t i = (uint32_t)(b0 + li);
            const ClusterResult &r = res[li];
            const bool in_p0 = p0_pos < p0.size() && p0[p0_pos] == i;
            const int32_t my_pidx = in_p0 ? (int32_t)p0_pos : -1;
            if (in_p0) ++p0_pos;
            // representatives created earlier in this block that share a bucket with the read
            inblk.clear();
            uint32_t newmax = r.n_cand ? r.max_hits : 0u;
            for (uint32_t x = hit_off[li]; x < hit_off[li + 1]; ++x) {                 // potential representatives that became one
                const PairHit &ph = hit_list[x];
                if (!p0_rep[ph.pidx]) continue;
                const uint32_t j = p0[ph.pidx];
                inblk.push_back(Cand{j, ph.hits, ph.cnt, std::max(nd[i], tot_of(j))}); newmax = std::max(newmax, ph.hits);
            }
            for (const Extra &ex : extras) {                                            // the rare other ones
                const size_t at = i - ex.id - 1;
                if (ex.hits[at]) { inblk.push_back(Cand{ex.id, ex.hits[at], ex.cnt[at], std::max(nd[i], tot_of(ex.id))}); newmax = std::max<uint32_t>(newmax, ex.hits[at]); }
            }
            bool have = false; Cand best{0, 0, 0, 1};
            if (inblk.empty()) {
                if (r.best_rep != 0xffffffffu) { best = Cand{r.best_rep, r.best_hits, r.best_cnt, r.best_den}; have = true; }
            } else {
                // merged order: hits descending, id descending (in-block ids are larger than every pre-block id)
                std::sort(inblk.begin(), inblk.end(), [](const Cand &a, const Cand &b) { return a.hits != b.hits ? a.hits > b.hits : a.id > b.id; });
                order.clear();
                for (size_t x = 0; x < inblk.size(); ++x) {
                    uint32_t ahead = (uint32_t)x;                                       // in-block candidates in front of it
                    for (uint32_t hh = inblk[x].hits + 1; hh <= (uint32_t)kLshTables; ++hh) ahead += r.n_cand ? r.cnt_h[hh] : 0u;
                    if (inblk[x].hits == newmax || ahead < (uint32_t)kClusterTopN) order.push_back(inblk[x]);
                }
                const uint32_t n_top = std::min<uint32_t>(r.n_cand, (uint32_t)kClusterTopN);
                for (uint32_t qx = 0; qx < n_top; ++qx) {
                    const uint32_t hq = r.top_hits[qx];
                    uint32_t ahead = qx;
                    for (const Cand &e : inblk) if (e.hits >= hq) ++ahead;
                    if (hq == newmax || ahead < (uint32_t)kClusterTopN)
                        order.push_back(Cand{r.top_rep[qx], hq, r.top_cnt[qx], std::max(nd[i], r.top_tot[qx])});
                }
                // a max-hits group larger than the list: its best member stands for the members the list does not hold
                if (r.n_cand && r.n_max > (uint32_t)kClusterTopN && newmax == r.max_hits && r.best_rep != 0xffffffffu)
                    order.push_back(Cand{r.best_rep, r.best_hits, r.best_cnt, r.best_den});
                std::stable_sort(order.begin(), order.end(), [](const Cand &a, const Cand &b) { return a.hits != b.hits ? a.hits > b.hits : a.id > b.id; });
                for (const Cand &e : order)
                    if (e.cnt > 0 && (!have || ratio_gt(e, best))) { best = e; have = true; }   // first strictly greater similarity
            }
            if (have && thr.pass(best.cnt, best.den)) { rep_of[i] = best.id; continue; }
            // new representative (src/asv_cluster.rs:178-187)
            rep_of[i] = i; ++n_reps;
            if (valid[i]) {
                new_ids.push_back(i);
                tables_dirty = true;
                if (my_pidx < 0 && li + 1 < nbk) {
                    lap(4);
                    // not among the potential representatives: hits / containment of the reads behind it, on demand
                    const uint32_t m = nbk - li - 1;
                    std::vector<uint32_t> xi(m), xj(m, i);
                    for (uint32_t x = 0; x < m; ++x) xi[x] = i + 1 + x;
                    extras.push_back(Extra{i, std::vector<uint8_t>(m), std::vector<uint32_t>(m)});
                    Extra &ex = extras.back();
                    TRY(up(d.pi, xi.data(), m, st)); TRY(up(d.pj, xj.data(), m, st));
                    TRY(d.ph.ensure(m + 16)); TRY(d.pc.ensure((size_t)m * 4 + 16));
                    launch_pair_stats(q, d.pi.as<uint32_t>(), d.pj.as<uint32_t>(), m, d.ph.as<uint8_t>(), d.pc.as<uint32_t>(), st);
                    CK(cudaMemcpyAsync(ex.hits.data(), d.ph.p, m, cudaMemcpyDeviceToHost, st));
                    CK(cudaMemcpyAsync(ex.cnt.data(), d.pc.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
                    CK(cudaStreamSynchronize(st));
                    CK(cudaGetLastError());
                    c->last_stats.kernel_launches += 1; c->last_stats.n_pairs += m;
                    lap(5);
                }
                if (my_pidx >= 0) p0_rep[my_pidx] = 1;
                ++n_new;
            }
        }
        lap(4);
        // the block's representatives join the delta level; the delta is folded into the main level when it reaches a quarter of it
        if (n_new) {
            // their signatures (the only ones the host ever needs) come from the device
            const uint32_t nn = (uint32_t)new_ids.size();
            TRY(up(d.pi, new_ids.data(), nn, st)); TRY(d.ph.ensure((size_t)nn * kLshTables * 8 + 16));
            launch_gather_sig(d.sig.as<uint64_t>(), d.pi.as<uint32_t>(), nn, d.ph.as<uint64_t>(), st);
            new_sig.resize((size_t)nn * kLshTables);
            CK(cudaMemcpyAsync(new_sig.data(), d.ph.p, (size_t)nn * kLshTables * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaGetLastError());
            c->last_stats.kernel_launches += 1; c->last_stats.d2h_bytes += (size_t)nn * kLshTables * 8;
            for (uint32_t x = 0; x < nn; ++x)
                for (int t = 0; t < kLshTables; ++t) fresh[t].push_back(Ent(new_sig[(size_t)x * kLshTables + t], new_ids[x]));
            new_ids.clear();
            for (int t = 0; t < kLshTables; ++t) {
                std::sort(fresh[t].begin(), fresh[t].end());
                merged.resize(dlt[t].size() + fresh[t].size());
                std::merge(dlt[t].begin(), dlt[t].end(), fresh[t].begin(), fresh[t].end(), merged.begin());
                dlt[t].swap(merged); fresh[t].clear();
            }
            n_delta += n_new;
            if (n_delta > 1024 && n_delta * 4 > n_main) {
                for (int t = 0; t < kLshTables; ++t) {
                    merged.resize(tab[t].size() + dlt[t].size());
                    std::merge(tab[t].begin(), tab[t].end(), dlt[t].begin(), dlt[t].end(), merged.begin());
                    tab[t].swap(merged); dlt[t].clear();
                }
                n_main += n_delta; n_delta = 0; main_dirty = true;
            }
        }
        // next block size: about kTargetNew new representatives per block
        if (!fixed_block) {
            const double rate = std::max<double>((double)n_new, 0.5) / (double)nbk;
            block = (size_t)std::min<double>((double)kMaxBlock, std::max<double>((double)kMinBlock, (double)kTargetNew / rate));
        }
        c->last_stats.n_batches += 1;
        b0 = b1;
        lap(6);
    }
    if (getenv("SAVONT_B200_CLUSTER_TRACE"))
        fprintf(stderr, "cluster host ms: tables %.1f query %.1f p0 %.1f pairs %.1f resolve %.1f fallback %.1f merge %.1f\n", tw[0], tw[1], tw[2],
                tw[3], tw[4], tw[5], tw[6]);
    if (n_reps_out) *n_reps_out = n_reps;
    CK(cudaStreamSynchronize(st));                 // ev[1] when there was no block
    c->last_stats.n_reads = n; c->last_stats.n_bases = rel[n];
    c->last_stats.sketch_ms = ev_ms(c->ev[0], c->ev[1]); c->last_stats.filter_ms = query_ms; c->last_stats.align_ms = pair_ms;
    c->last_stats.total_ms = c->last_stats.sketch_ms + query_ms + pair_ms;
    if (stats) *stats = c->last_stats;
    return 0;
}
