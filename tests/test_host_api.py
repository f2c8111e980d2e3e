"""Host layer (C++ libsavont_host.so through savont_b200/api.py): EM, writers on CPU; Stage 7 / 7b drivers on GPU."""
import os

import numpy as np
import pytest

from oracle import oracle as O


def _host_available():
    from savont_b200 import api
    return os.path.exists(api._HOST_PATH)


def test_em_matches_oracle_em():
    from savont_b200 import api
    rng = np.random.default_rng(3)
    A = 12
    classes = {}
    for _ in range(40):
        size = int(rng.integers(1, 4))
        key = tuple(sorted(rng.choice(A, size=size, replace=False).tolist()))
        classes[key] = classes.get(key, 0) + int(rng.integers(1, 50))
    keys = sorted(classes)
    offs = np.zeros(len(keys) + 1, np.uint64); offs[1:] = np.cumsum([len(k) for k in keys])
    flat = np.array([a for k in keys for a in k], np.uint32)
    counts = np.array([classes[k] for k in keys], np.uint64)
    total = int(counts.sum())
    ab_h, dp_h, it_h = api.em(A, offs, flat, counts, total)
    ab_o, dp_o, it_o = O.em(A, offs, flat, counts, total)
    assert it_h == it_o
    assert np.array_equal(dp_h, dp_o)
    assert np.array_equal(ab_h.view(np.uint64), ab_o.view(np.uint64))      # same summation order -> same bits
    assert int(dp_h.sum()) in range(total - A, total + A + 1)


def test_writers_follow_reference_formats(tmp_path):
    from savont_b200 import api
    cons = [api.ConsensusSequence(b"NNACGTACGTNN", depth=7, id=3, unambig_best_read_map_count=5, ambig_read_map_count=2, num_map_leq_10nm=6),
            api.ConsensusSequence(b"GGGTTTAAAC", depth=2, id=0)]
    fa, tb = tmp_path / "final_asvs.fasta", tmp_path / "feature-table.tsv"
    api.write_consensus_fasta(cons, fa, "final")
    api.write_feature_table(cons, tb, ["s1"])
    # src/alignment.rs:848-850 header; leading/trailing N trimmed (:842-843)
    assert fa.read_text().splitlines() == [
        ">final_consensus_0_depth_7 debug_id:3 chimera_score:0 unambiguous_read_assignments:5 ambig_read_assignments:2 num_align_leq_10_mismatches:6",
        "ACGTACGT",
        ">final_consensus_1_depth_2 debug_id:0 chimera_score:0 unambiguous_read_assignments:0 ambig_read_assignments:0 num_align_leq_10_mismatches:0",
        "GGGTTTAAAC"]
    # src/main.rs:388-392
    assert tb.read_text().splitlines() == ["#OTU ID\ts1", "final_consensus_0_depth_7\t7", "final_consensus_1_depth_2\t2"]
    for c, d in zip(cons, ([4, 3, 0], [0, 1, 1])):
        c.per_sample_depths = d
    api.write_consensus_fasta(cons, fa, "final")
    api.write_feature_table(cons, tb, ["a", "b", "c"])
    assert fa.read_text().splitlines()[0].startswith(">final_consensus_0_depth_4-3-0 debug_id:3")
    # pooled: dash-joined depths in the id, one column per sample (src/main.rs:394-397)
    assert tb.read_text().splitlines() == ["#OTU ID\ta\tb\tc", "final_consensus_0_depth_4-3-0\t4\t3\t0", "final_consensus_1_depth_0-1-1\t0\t1\t1"]


def test_high_freq_kmers_matches_oracle():
    from savont_b200 import api
    rng = np.random.default_rng(9)
    for n in (1, 7, 99999, 100000, 250003):
        km = rng.permutation(n).astype(np.uint64) + 5
        ct = rng.integers(0, 400, size=(n, 2)).astype(np.uint32)
        ct[rng.integers(0, n, size=max(n // 50, 1))] += 5000
        th, hf = api.high_freq_kmers(km, ct)
        th_o, hf_o = O.high_freq_kmers(km, ct)
        assert th == th_o and np.array_equal(hf, hf_o)
    with pytest.raises(RuntimeError):
        api.high_freq_kmers(np.zeros(0, np.uint64), np.zeros((0, 2), np.uint32))


def test_eq_classes_helper():
    from savont_b200 import api
    status = np.array([0, 1, 0, 0, 2, 0], np.uint8)
    co = np.array([0, 1, 1, 3, 4, 4, 6], np.uint64)
    ca = np.array([2, 0, 1, 2, 0, 1], np.uint32)
    offs, flat, counts, total = api.eq_classes(status, co, ca)
    assert total == 4 and offs.tolist() == [0, 2, 3] and flat.tolist() == [0, 1, 2] and counts.tolist() == [2, 2]
    offs, flat, counts, total = api.eq_classes(status, co, ca, file_idx=np.array([0, 0, 1, 1, 1, 0]), sample=1)
    assert total == 2 and counts.tolist() == [1, 1]


def _oracle_depths(orc, reads, n_asv, n_samples=1, file_idx=None):
    res = orc.map_reads(reads["bases"], reads["quals"], reads["offsets"], file_idx=file_idx, n_samples=n_samples)
    from savont_b200 import api
    out = []
    for s in ([None] if file_idx is None else range(n_samples)):
        offs, flat, counts, total = api.eq_classes(res["status"], res["class_offsets"], res["class_asv"], file_idx, s)
        if total == 0:
            out.append(np.zeros(n_asv, np.uint64)); continue
        _, depths, _ = O.em(n_asv, offs, flat, counts, total)
        out.append(depths)
    return res, out


@pytest.mark.gpu
def test_refine_asv_depths_with_em_matches_oracle_pipeline(tmp_path):
    from savont_b200 import api, synth
    w = synth.Workload("ont16s", n_roots=5, sites_per_genus=6)
    r = w.reads(3000)
    seqs = [bytes(w.asv_bases[int(w.asv_offsets[a]):int(w.asv_offsets[a + 1])]) for a in range(w.n_asv)]
    cons = [api.ConsensusSequence(s, depth=1, id=i) for i, s in enumerate(seqs)]
    args = api.Cli(band=w.band)
    api.refine_asv_depths_with_em(r, cons, w.snpmers, args, temp_dir=tmp_path)
    orc = O.Oracle(O.default_params(band=w.band), w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    res, (depths,) = _oracle_depths(orc, r, w.n_asv)
    kept = [a for a in range(w.n_asv) if depths[a] > 0]
    assert [c.id for c in cons] == kept                                   # zero-depth ASVs dropped (src/alignment.rs:2031)
    assert [c.depth for c in cons] == [int(depths[a]) for a in kept]
    assert [c.unambig_best_read_map_count for c in cons] == [int(res["unambig"][a]) for a in kept]
    assert [c.ambig_read_map_count for c in cons] == [int(res["ambig"][a]) for a in kept]
    assert [c.num_map_leq_10nm for c in cons] == [int(res["leq10"][a]) for a in kept]
    assert (tmp_path / "final_asvs_for_em.fasta").read_text().count(">em_refinement_consensus_") == w.n_asv
    assert sum(c.depth for c in cons) > 0.5 * 3000
    # the same Stage 7 with the reads handed over as TwinRead fields (&[TwinRead], src/alignment.rs:1723-1729)
    from savont_b200 import twin
    sk = O.sketch_batch(O.default_params(band=w.band), w.snpmers, r["bases"], r["quals"], r["offsets"])
    cons2 = [api.ConsensusSequence(s, depth=1, id=i) for i, s in enumerate(seqs)]
    api.refine_asv_depths_with_em(twin.from_sketch(sk, r["offsets"], kmer48=True), cons2, w.snpmers, args)
    assert [(c.id, c.depth, c.unambig_best_read_map_count, c.ambig_read_map_count, c.num_map_leq_10nm) for c in cons2] == \
           [(c.id, c.depth, c.unambig_best_read_map_count, c.ambig_read_map_count, c.num_map_leq_10nm) for c in cons]


@pytest.mark.gpu
def test_compute_per_sample_depths_matches_oracle_pipeline():
    from savont_b200 import api, synth
    w = synth.Workload("pooled", n_roots=4, n_samples=3)
    r = w.reads(2500)
    seqs = [bytes(w.asv_bases[int(w.asv_offsets[a]):int(w.asv_offsets[a + 1])]) for a in range(w.n_asv)]
    cons = [api.ConsensusSequence(s, depth=1, id=i) for i, s in enumerate(seqs)]
    got = api.compute_per_sample_depths(r, 3, cons, w.snpmers, api.Cli(band=w.band))
    orc = O.Oracle(O.default_params(band=w.band), w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    _, per = _oracle_depths(orc, r, w.n_asv, n_samples=3, file_idx=r["file_idx"])
    want = np.stack(per, axis=1).astype(np.int64).tolist()
    assert got == want
    from savont_b200 import twin
    sk = O.sketch_batch(O.default_params(band=w.band), w.snpmers, r["bases"], r["quals"], r["offsets"])
    tw = twin.from_sketch(sk, r["offsets"], file_idx=r["file_idx"])
    assert api.compute_per_sample_depths(tw, 3, cons, w.snpmers, api.Cli(band=w.band)) == want


@pytest.mark.gpu
def test_stage1_counting_feeds_the_read_filter():
    """Stage 1 counting -> high-frequency set -> Stage 1.5 read tests, GPU + host layer vs the oracle chain"""
    from savont_b200 import _lib, api, synth
    w = synth.Workload("ont16s", n_roots=4)
    r = w.reads(3000)
    ids = [f"read_{i:05d}" + (" rc" if i % 3 == 0 else " fw") for i in range(3000)]
    ctx = _lib.Context(0, _lib.default_params(band=w.band))
    ctx.load_snpmers(w.snpmers)
    got = api.read_to_split_kmers(ctx, dict(bases=r["bases"], quals=r["quals"], offsets=r["offsets"], ids=ids))
    rc = np.array([1 if i % 3 == 0 else 0 for i in range(3000)], np.uint8)
    want = O.count_split_kmers(r["bases"], r["quals"], r["offsets"], rc=rc)
    assert np.array_equal(got["kmers"], want["kmers"]) and np.array_equal(got["counts"], want["counts"])
    assert got["n_distinct"] == want["n_distinct"]
    th, hf = api.high_freq_kmers(got["kmers"], got["counts"])
    th_o, hf_o = O.high_freq_kmers(want["kmers"], want["counts"])
    assert th == th_o and np.array_equal(hf, hf_o)
    # lower the bar so that the set is not empty on this small sample, then run the read tests with it
    hf = got["kmers"][got["counts"].sum(1) > np.quantile(got["counts"].sum(1), 0.97)]
    args = api.Cli(band=w.band, min_read_length=1400, max_read_length=1600)
    sel = api.twin_reads_from_snpmers(ctx, dict(bases=r["bases"], quals=r["quals"], offsets=r["offsets"], ids=[i.split()[0] for i in ids]), args,
                                      high_freq_kmers=hf)
    ref = O.twin_reads_filter(O.default_params(band=w.band), O.read_filter(1400, 1600, 98.0), w.snpmers, hf, r["bases"], r["quals"], r["offsets"],
                              [i.split()[0] for i in ids])
    assert np.array_equal(sel["status"], ref["status"]) and np.array_equal(sel["order"], ref["order"])
    assert np.array_equal(sel["n_solid_mini"], ref["n_solid_mini"])
    ctx.close()


@pytest.mark.gpu
def test_python_mirrors_of_sketch_and_candidates():
    from savont_b200 import _lib, api, synth
    w = synth.Workload("ont16s", n_roots=3)
    r = w.reads(50)
    ctx = _lib.Context(0, _lib.default_params(band=w.band))
    ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    o = r["offsets"]
    rd = bytes(r["bases"][int(o[0]):int(o[1])]); ql = bytes(r["quals"][int(o[0]):int(o[1])])
    tw = api.get_twin_read_syncmer(ctx, rd, ql, "r0")
    want = O.sketch_batch(O.default_params(band=w.band), w.snpmers, *O.concat([rd])[:1], np.frombuffer(ql, np.uint8), O.concat([rd])[1])
    assert np.array_equal(tw["minimizer_kmers"], want["mini_kmer"]) and np.array_equal(tw["snpmer_positions"], want["snp_pos"])
    assert tw["est_id"] == want["est_id"][0]
    assert api.get_twin_read_syncmer(ctx, b"ACGT", None) is None           # len < k -> None (src/seeding.rs:339-341)
    asv_sk = api.twin_reads_from_fasta(ctx, [bytes(w.asv_bases[int(w.asv_offsets[a]):int(w.asv_offsets[a + 1])]) for a in range(3)])
    assert len(asv_sk["mini_offsets"]) == 4 and np.all(np.isnan(asv_sk["est_id"]))
    c = api.find_compatible_candidates(ctx, r["bases"], r["quals"], r["offsets"])
    assert len(c["status"]) == 50
    ctx.close()


def test_cluster_lists_from_assignments():
    """host halves of cluster_reads_by_kmers (:192-210) and reassign_reads_to_best_cluster (:1040-1131) on oracle outputs"""
    from savont_b200 import api
    from tests.util import clustering_case, csr_u64, reassign_case
    km, off, _ = clustering_case(31, n_reads=200)
    rep = O.cluster_reads_by_kmers(km, off)
    cl = api.clusters_from_assignment(rep, 1)
    assert sorted(int(x) for c in cl for x in c) == list(range(200))
    assert all(len(a) >= len(b) for a, b in zip(cl, cl[1:]))
    assert all(np.all(np.diff(c) > 0) and int(rep[c[0]]) == int(c[0]) and np.all(rep[c] == c[0]) for c in cl)
    big = api.clusters_from_assignment(rep, 5)
    assert big and all(len(c) >= 5 for c in big) and len(big) < len(cl)
    # reassignment: reads listed cluster by cluster
    cons, reads = reassign_case(5, n_cons=8, n_reads=60)
    ck, co = csr_u64(cons); rk, ro = csr_u64(reads)
    clusters = [np.arange(i * 10, (i + 1) * 10) for i in range(6)] + [np.zeros(0, np.int64)] * 2
    best = O.snpmer_compare(ck, co, rk, ro)["best_cons"]
    new, moved = api.clusters_from_reassignment(clusters, best, 1)
    assert sorted(int(x) for c in new for x in c) == list(range(60)) and moved == int((best != np.repeat(np.arange(6), 10)).sum()) and moved > 0
    assert all(np.all(best[c] == best[c[0]]) for c in new)


def test_pileups_from_columns_shape():
    from savont_b200 import api
    from tests.util import pileup_case
    cons, reads, quals, rcons = pileup_case(9, n_cons=4, reads_per=6)
    cb, co = O.concat(cons); rb, ro = O.concat(reads); qb, _ = O.concat(quals)
    res = O.consensus_pileups(cb, co, rb, qb, ro, rcons, band=16)
    pl = api.pileups_from_columns(res, co)
    assert [len(p) for p in pl] == [len(c) for c in cons]
    assert sum(len(col["bases"]) for p in pl for col in p) == len(res["entries"])
    kinds = {b[0] for p in pl for col in p for b in col["bases"]}
    assert kinds == {"Base", "Deletion", "Insertion"}
    assert all(1 <= len(b[1]) <= 2 for p in pl for col in p for b in col["bases"] if b[0] == "Insertion")


def test_cpp_host_cluster_lists_match_python():
    """savont_host::clusters_from_assignment / clusters_from_reassignment (C++) against the Python halves in api.py"""
    import ctypes as C
    from savont_b200 import api
    from tests.util import clustering_case, csr_u64, reassign_case
    if not _host_available():
        pytest.skip("libsavont_host.so not built")
    H = api.host()
    H.savont_host_clusters_from_assignment.restype = C.c_longlong
    H.savont_host_clusters_from_reassignment.restype = C.c_longlong
    u64p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    km, off, _ = clustering_case(33, n_reads=220)
    rep = np.ascontiguousarray(O.cluster_reads_by_kmers(km, off), np.uint32)
    for min_size in (1, 3):
        offs = np.zeros(len(rep) + 1, np.uint64); mem = np.zeros(len(rep), np.uint64)
        nc = H.savont_host_clusters_from_assignment(rep.ctypes.data_as(u32p), C.c_size_t(len(rep)), C.c_size_t(min_size),
                                                    offs.ctypes.data_as(u64p), mem.ctypes.data_as(u64p))
        want = api.clusters_from_assignment(rep, min_size)
        assert nc == len(want) and nc > 3
        assert all(np.array_equal(mem[int(offs[c]):int(offs[c + 1])], want[c]) for c in range(nc))
    # an out-of-range representative is an error, not a crash
    bad = rep.copy(); bad[5] = 10 ** 6
    assert H.savont_host_clusters_from_assignment(bad.ctypes.data_as(u32p), C.c_size_t(len(bad)), C.c_size_t(1), offs.ctypes.data_as(u64p),
                                                  mem.ctypes.data_as(u64p)) == -1
    cons, reads = reassign_case(6, n_cons=8, n_reads=64)
    ck, co = csr_u64(cons); rk, ro = csr_u64(reads)
    clusters = [np.arange(i * 8, (i + 1) * 8) for i in range(8)]
    best = np.ascontiguousarray(O.snpmer_compare(ck, co, rk, ro)["best_cons"], np.uint32)
    cl_off = np.arange(0, 65, 8, dtype=np.uint64); cl_mem = np.arange(64, dtype=np.uint64)
    offs = np.zeros(9, np.uint64); mem = np.zeros(64, np.uint64); moved = C.c_uint64(0)
    nc = H.savont_host_clusters_from_reassignment(cl_off.ctypes.data_as(u64p), cl_mem.ctypes.data_as(u64p), C.c_size_t(8), best.ctypes.data_as(u32p),
                                                  C.c_size_t(2), offs.ctypes.data_as(u64p), mem.ctypes.data_as(u64p), C.byref(moved))
    want, want_moved = api.clusters_from_reassignment(clusters, best, 2)
    assert nc == len(want) and int(moved.value) == want_moved
    assert all(np.array_equal(mem[int(offs[c]):int(offs[c + 1])], want[c]) for c in range(nc))


@pytest.mark.gpu
def test_python_mirrors_of_stage_2_3_4():
    """api.cluster_reads_by_kmers / reassign_reads_to_best_cluster / generate_consensus_pileups (reference signatures) on the GPU
    against the same host halves fed with the oracle's device-side results"""
    from savont_b200 import _lib, api
    from tests.util import clustering_case, csr_u64, pileup_case, reassign_case
    ctx = _lib.Context(0, _lib.default_params())
    km, off, _ = clustering_case(41, n_reads=180)
    got = api.cluster_reads_by_kmers(ctx, km, off, min_cluster_size=2)
    want = api.clusters_from_assignment(O.cluster_reads_by_kmers(km, off), 2)
    assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))
    cons, reads = reassign_case(8, n_cons=8, n_reads=64)
    ck, co = csr_u64(cons); sk, so = csr_u64(reads)
    clusters = [np.arange(i * 8, (i + 1) * 8)[::-1] for i in range(8)]            # read ids in any order inside a cluster
    got, moved = api.reassign_reads_to_best_cluster(ctx, clusters, ck, co, sk, so, min_cluster_size=1)
    listed = [reads[int(r)] for c in clusters for r in c]
    want, want_moved = api.clusters_from_reassignment(clusters, O.snpmer_compare(ck, co, *csr_u64(listed))["best_cons"], 1)
    assert moved == want_moved and len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))
    pcons, preads, pquals, prcons = pileup_case(12, n_cons=4, reads_per=5)
    cl = [[r for r in range(len(preads)) if prcons[r] == i] for i in range(4)]
    args = api.Cli(); args.band = 16
    pl = api.generate_consensus_pileups(ctx, pcons, cl, preads, pquals, args)
    cb, cof = O.concat(pcons); rb, ro = O.concat(preads); qb, _ = O.concat(pquals)
    want_pl = api.pileups_from_columns(O.consensus_pileups(cb, cof, rb, qb, ro, prcons, band=16), cof)
    assert pl == want_pl and sum(len(col["bases"]) for p in pl for col in p) > 100
    ctx.close()


def test_snpmer_cluster_lists_cpp_matches_python():
    """savont_host::snpmer_clusters_from_assignment (C++) == api.snpmer_clusters_from_assignment (Python) on the oracle's Stage 3
    assignment; the lists partition every k-mer cluster and put the representative's cluster members in ascending order"""
    import ctypes as C
    from savont_b200 import api, build
    from tests.test_gpu_parity import _stage3_case
    km, offs, cr, coffs = _stage3_case(11, 1500, 5, 10)
    kcs = [cr[int(coffs[c]):int(coffs[c + 1])].astype(np.int64) for c in range(5)]
    rep = np.full(1500, 0xffffffff, np.uint32)
    for mem in kcs:
        w_ = O.cluster_reads_by_snpmers(km, offs, mem.astype(np.uint32))
        rep[mem] = w_[mem]
    H = C.CDLL(build.build_host())
    H.savont_host_snpmer_clusters_from_assignment.restype = C.c_longlong
    u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    for min_size in (1, 3):
        want = api.snpmer_clusters_from_assignment(kcs, rep, min_size)
        o = np.zeros(len(cr) + 1, np.uint64); mbuf = np.zeros(max(len(cr), 1), np.uint64); kc_of = np.zeros(max(len(cr), 1), np.uint32)
        nc = H.savont_host_snpmer_clusters_from_assignment(coffs.ctypes.data_as(u64p), cr.ctypes.data_as(u32p), C.c_size_t(5), rep.ctypes.data_as(u32p),
                                                           C.c_size_t(min_size), o.ctypes.data_as(u64p), mbuf.ctypes.data_as(u64p), kc_of.ctypes.data_as(u32p))
        flat = [c for kc in want for c in kc]
        assert nc == len(flat) and nc > 5
        for i, c in enumerate(flat):
            assert np.array_equal(mbuf[int(o[i]):int(o[i + 1])].astype(np.int64), c)
        assert [int(x) for x in kc_of[:nc]] == [i for i, kc in enumerate(want) for _ in kc]
        if min_size == 1:
            for kc, local in zip(kcs, want):
                assert sorted(int(x) for c in local for x in c) == sorted(int(x) for x in kc)
                assert all(len(a) >= len(b) for a, b in zip(local, local[1:])) and all(np.all(np.diff(c) > 0) for c in local)
                assert all(np.all(rep[c] == rep[c[0]]) for c in local)


def test_read_to_asv_mappings_writer_cpp_matches_python(tmp_path):
    """temp/read_to_asv_mappings.tsv (src/alignment.rs:1875-1887): "{read id}\tdebug_id:{consensus id}\t{SNPmer mismatches}\t{NM}",
    C++ writer == Python writer, reads without rows skipped"""
    import ctypes as C
    from savont_b200 import api
    H = api.host()
    rows = dict(row_offsets=np.array([0, 2, 2, 3, 8], np.uint64), asv=np.array([3, 1, 0, 4, 2, 3, 1, 0], np.uint32),
                nm=np.array([4, 9, 0, 7, 7, 8, 12, 40], np.int32), snp_mismatches=np.array([0, 0, 2, 1, 1, 1, 1, 1], np.uint32))
    ids = ["r0 rc", "r1", "r2", "some/longer id"]
    cons = [10, 11, 12, 13, 99]
    p1 = tmp_path / "py.tsv"; p2 = tmp_path / "cpp.tsv"
    api.write_read_to_asv_mappings(p1, ids, cons, rows)
    H.savont_host_write_read_to_asv_mappings.restype = C.c_int
    ca = np.array(cons, np.uint64)
    rc = H.savont_host_write_read_to_asv_mappings(str(p2).encode(), "\n".join(ids).encode(), C.c_size_t(len(ids)),
                                                  ca.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_size_t(len(cons)),
                                                  rows["row_offsets"].ctypes.data_as(C.POINTER(C.c_uint64)), rows["asv"].ctypes.data_as(C.POINTER(C.c_uint32)),
                                                  rows["nm"].ctypes.data_as(C.POINTER(C.c_int32)), rows["snp_mismatches"].ctypes.data_as(C.POINTER(C.c_uint32)))
    assert rc == 0
    txt = open(p1).read()
    assert txt == open(p2).read()
    assert txt.splitlines()[0] == "r0 rc\tdebug_id:13\t0\t4" and txt.splitlines()[2] == "r2\tdebug_id:10\t2\t0" and len(txt.splitlines()) == 8
