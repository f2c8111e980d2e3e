"""GPU parity: libsavont_b200.so (through the C ABI) vs the CPU oracle, bit-exact.
Run on the B200 box: python -m pytest tests -m gpu"""
import random

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import assert_same, mutate, rand_seq, snpmer_set_from, weird_sequences

pytestmark = pytest.mark.gpu

SKETCH_KEYS = ["status", "est_id", "mini_offsets", "mini_pos", "mini_kmer", "mini_fwd", "snp_offsets", "snp_pos", "snp_kmer", "codes"]
CAND_KEYS = ["status", "offsets", "asv", "mm", "mini", "strand"]
MAP_KEYS = ["status", "best_nm", "class_offsets", "class_asv", "unambig", "ambig", "leq10", "totals", "sample_singleton"]


@pytest.fixture(scope="module")
def ctx():
    from savont_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def _params(lib_mod, **kw):
    return lib_mod.default_params(**kw), O.default_params(**kw)


@pytest.mark.parametrize("k,c", [(17, 11), (15, 9), (21, 11), (13, 7), (23, 8), (19, 19), (17, 2), (19, 6)])
def test_sketch_ragged_inputs(ctx, k, c):
    from savont_b200 import _lib
    seqs, quals = weird_sequences(100 + k, k)
    snp = sorted(snpmer_set_from(seqs, k))
    gp, op = _params(_lib, k=k, c=c)
    ctx.set_params(gp)
    ctx.load_snpmers(snp)
    bases, offs = O.concat(seqs)
    qcat, _ = O.concat(quals)
    for q in (qcat, None):
        want = O.sketch_batch(op, snp, bases, q, offs)
        got = ctx.sketch(bases, q, offs)
        assert_same(got, want, SKETCH_KEYS, f"sketch k={k} c={c} quals={'y' if q is not None else 'n'}")


def test_set_params_after_load_snpmers(ctx):
    """the SNPmer Bloom filter is keyed by k: changing k after load_snpmers must rebuild it (no silently dropped SNPmers)"""
    from savont_b200 import _lib
    k, c = 19, 6
    seqs, quals = weird_sequences(100 + k, k)
    snp = sorted(snpmer_set_from(seqs, k))
    assert len(snp) > 0
    ctx.set_params(_lib.default_params())          # k = 17
    ctx.load_snpmers(snp)                          # keys of a k = 19 run
    gp, op = _params(_lib, k=k, c=c)
    ctx.set_params(gp)
    bases, offs = O.concat(seqs)
    qcat, _ = O.concat(quals)
    assert_same(ctx.sketch(bases, qcat, offs), O.sketch_batch(op, snp, bases, qcat, offs), SKETCH_KEYS, "set_params after load_snpmers")


def test_sketch_empty_and_no_snpmers(ctx):
    from savont_b200 import _lib
    gp, op = _params(_lib)
    ctx.set_params(gp)
    ctx.load_snpmers([])
    rng = random.Random(5)
    seqs = [rand_seq(rng, 300) for _ in range(5)]
    bases, offs = O.concat(seqs)
    assert_same(ctx.sketch(bases, None, offs), O.sketch_batch(op, [], bases, None, offs), SKETCH_KEYS, "no snpmers")
    empty = ctx.sketch(np.zeros(0, np.uint8), None, np.zeros(1, np.uint64))
    assert len(empty["status"]) == 0 and len(empty["mini_pos"]) == 0


def test_sketch_synthetic_reads_small_batches(ctx):
    from savont_b200 import _lib, synth
    w = synth.Workload("ont16s", n_roots=6)
    r = w.reads(3000)
    gp, op = _params(_lib, band=w.band)
    ctx.set_params(gp)
    ctx.load_snpmers(w.snpmers)
    ctx.set_batch_reads(700)          # several internal batches
    try:
        got = ctx.sketch(r["bases"], r["quals"], r["offsets"])
    finally:
        ctx.set_batch_reads(0)
    want = O.sketch_batch(op, w.snpmers, r["bases"], r["quals"], r["offsets"])
    assert_same(got, want, SKETCH_KEYS, "synthetic sketch")
    assert np.all(want["est_id"] > 98.0)


def _dp_cases(seed, band):
    rng = random.Random(seed)
    lens = [1, 5, 20, 37, 64, 90, 150, 300] if band < 40 else [40, 90, 300, 700, 1500]
    asvs = [rand_seq(rng, max(L, 17)) for L in lens]
    reads, pr, pa, ps = [], [], [], []
    for a, s in enumerate(asvs):
        for t in range(8):
            rd = mutate(rng, s, rng.randrange(6), rng.randrange(4), rng.randrange(4))
            if t == 4:
                rd = rd[rng.randrange(1, 8):]
            if t == 5:
                rd = rd[:max(1, len(rd) - rng.randrange(1, 8))] + rand_seq(rng, 3)
            if t == 6:
                rd = rand_seq(rng, rng.randrange(1, 2 * band + 40))    # unrelated, odd length
            if t == 7:
                rd = rand_seq(rng, 30) + rd                            # offset beyond small bands
            reads.append(rd); pr.append(len(reads) - 1); pa.append(a); ps.append(t % 2)
    return asvs, reads, pr, pa, ps


# bands with 2B + 1 = 32q + 1 (16, 32, 48, 64, 96, 128, 192) run the q-word + tail-bit kernels, the others whole words
@pytest.mark.parametrize("band", [0, 1, 3, 8, 15, 16, 31, 32, 33, 47, 48, 64, 79, 96, 100, 128, 143, 192, 207])
def test_align_pairs_nm_and_cigar(ctx, band):
    from savont_b200 import _lib
    asvs, reads, pr, pa, ps = _dp_cases(900 + band, band)
    gp, op = _params(_lib, band=band)
    ctx.set_params(gp)
    ctx.load_snpmers([])
    ab, ao = O.concat(asvs)
    ctx.load_asvs(ab, ao)
    orc = O.Oracle(op, [])
    orc.load_asvs(ab, ao)
    rb, ro = O.concat(reads)
    want = orc.align_pairs(rb, ro, pr, pa, ps, want_cigar=True)
    got_nm = ctx.align_pairs(rb, ro, pr, pa, ps, want_cigar=False)
    assert_same(got_nm, want, ["nm", "asv_end", "read_end"], f"align nm band={band}")
    got = ctx.align_pairs(rb, ro, pr, pa, ps, want_cigar=True)
    assert_same(got, want, ["nm", "asv_end", "read_end", "asv_start", "read_start", "cigar_offsets", "cigar"], f"align cigar band={band}")


def test_align_pairs_full_length_reads(ctx):
    """config-2 shaped pairs: 1500 bp, 2% errors, band 64, both strands"""
    from savont_b200 import _lib, synth
    w = synth.Workload("ont16s", n_roots=4)
    r = w.reads(400)
    gp, op = _params(_lib, band=64)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    rng = np.random.default_rng(1)
    pr = np.arange(400, dtype=np.uint32)
    pa = r["truth_asv"].copy()
    pa[::3] = rng.integers(0, w.n_asv, size=len(pa[::3]))        # some wrong ASVs
    ps = r["truth_strand"].copy(); ps[::5] ^= 1                  # some wrong strands
    want = orc.align_pairs(r["bases"], r["offsets"], pr, pa, ps, want_cigar=True)
    got = ctx.align_pairs(r["bases"], r["offsets"], pr, pa, ps, want_cigar=True)
    assert_same(got, want, ["nm", "asv_end", "read_end", "asv_start", "read_start", "cigar_offsets", "cigar"], "full length")


def _make_workload(name, **kw):
    from savont_b200 import synth
    return synth.Workload(name, **kw)


@pytest.mark.parametrize("name,kw,n,min_assigned", [
    ("ont16s", dict(n_roots=8), 4000, 0.9),
    ("ont16s", dict(n_roots=5, sites_per_genus=3), 3000, 0.5),     # near-identical variants -> ambiguous classes
    ("hifi16s", dict(n_roots=8), 3000, 0.9),
    ("operon", dict(n_roots=4), 600, 0.9),
])
def test_candidates_and_map_reads(ctx, name, kw, n, min_assigned):
    from savont_b200 import _lib
    w = _make_workload(name, **kw)
    r = w.reads(n)
    # degrade a few reads: truncate below k, wipe SNPmers with N, flat qualities
    gp, op = _params(_lib, band=w.band)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    assert_same(ctx.find_candidates(r["bases"], r["quals"], r["offsets"]),
                orc.find_candidates(r["bases"], r["quals"], r["offsets"]), CAND_KEYS, f"candidates {name}")
    want = orc.map_reads(r["bases"], r["quals"], r["offsets"])
    ctx.set_batch_reads(1000)
    try:
        got = ctx.map_reads(r["bases"], r["quals"], r["offsets"])
    finally:
        ctx.set_batch_reads(0)
    assert_same(got, want, MAP_KEYS, f"map_reads {name}")
    assert got["stats"]["dp_cells"] == want["dp_cells"]
    assert int(want["totals"][0]) > min_assigned * n
    # device-resident variant gives the same answer
    ctx.upload_reads(r["bases"], r["quals"], r["offsets"])
    ctx.map_resident()
    assert_same(ctx.fetch_resident(), want, MAP_KEYS, f"map_resident {name}")


def test_map_reads_both_strands_mode(ctx):
    from savont_b200 import _lib
    w = _make_workload("ont16s", n_roots=5)
    r = w.reads(1500)
    gp, op = _params(_lib, band=w.band, strand_mode=1)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    want = orc.map_reads(r["bases"], r["quals"], r["offsets"])
    got = ctx.map_reads(r["bases"], r["quals"], r["offsets"])
    assert_same(got, want, MAP_KEYS, "strand_mode=1")
    assert got["stats"]["dp_cells"] == want["dp_cells"]


@pytest.mark.parametrize("strand_mode", [0, 1])
def test_best_hit_alignment_records(ctx, strand_mode):
    """per read: best-hit ASV, strand, coordinates, identity and CIGAR (north star) = the oracle's traceback of the pair
    (read, first ASV of the oracle's class, strand by seed vote / smaller NM)"""
    from savont_b200 import _lib
    w = _make_workload("ont16s", n_roots=5, sites_per_genus=4)
    r = w.reads(1200)
    bases = r["bases"].copy(); offs = r["offsets"]
    bases[int(offs[3]):int(offs[4])] = ord("N")                       # an unassigned read in the middle
    gp, op = _params(_lib, band=w.band, strand_mode=strand_mode)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    want = orc.map_reads(bases, r["quals"], offs)
    cand = orc.find_candidates(bases, r["quals"], offs)
    n = len(offs) - 1
    co = want["class_offsets"].astype(np.int64)
    reads = np.flatnonzero(want["status"] == 0)
    assert 0 < len(reads) < n
    asv = want["class_asv"][co[reads]]
    if strand_mode == 0:
        cofs = cand["offsets"].astype(np.int64)
        strand = np.array([cand["strand"][cofs[rd]:cofs[rd + 1]][list(cand["asv"][cofs[rd]:cofs[rd + 1]]).index(a)]
                           for rd, a in zip(reads, asv)], np.uint8)
    else:
        f = orc.align_pairs(bases, offs, reads, asv, np.zeros(len(reads), np.uint8))["nm"]
        v = orc.align_pairs(bases, offs, reads, asv, np.ones(len(reads), np.uint8))["nm"]
        strand = (v < f).astype(np.uint8)
    ref = orc.align_pairs(bases, offs, reads, asv, strand, want_cigar=True)
    assert np.array_equal(ref["nm"], want["best_nm"][reads])
    ctx.upload_reads(bases, r["quals"], offs)
    ctx.map_resident()
    assert_same(ctx.fetch_resident(), want, MAP_KEYS, "map_resident")
    got = ctx.best_hits_resident()
    hit = got["best_asv"] != 0xffffffff
    assert np.array_equal(np.flatnonzero(hit), reads)
    assert np.array_equal(got["best_asv"][reads], asv) and np.array_equal(got["strand"][reads], strand)
    assert np.all(got["strand"][~hit] == 0xff) and np.all(np.isnan(got["identity"][~hit])) and np.all(got["asv_start"][~hit] == -1)
    for key in ("asv_start", "asv_end", "read_start", "read_end"):
        assert np.array_equal(got[key][reads], ref[key]), key
    gco = got["cigar_offsets"].astype(np.int64); rco = ref["cigar_offsets"].astype(np.int64)
    assert np.array_equal((gco[1:] - gco[:-1])[reads], rco[1:] - rco[:-1]) and np.all((gco[1:] - gco[:-1])[~hit] == 0)
    assert np.array_equal(got["cigar"], ref["cigar"])
    span = (ref["asv_end"] - ref["asv_start"]).astype(np.float64)
    assert np.array_equal(got["identity"][reads], 1.0 - ref["nm"].astype(np.float64) / span)
    # CIGAR consistency: M + D spans the ASV interval, M + I the read interval; the ops are run-length maximal
    ops = got["cigar"] & 15; lens = (got["cigar"] >> 4).astype(np.int64)
    owner = np.repeat(np.arange(n), gco[1:] - gco[:-1])
    asv_span = np.bincount(owner, weights=lens * (ops != 1), minlength=n).astype(np.int64)
    read_span = np.bincount(owner, weights=lens * (ops != 2), minlength=n).astype(np.int64)
    assert np.array_equal(asv_span[reads], got["asv_end"][reads] - got["asv_start"][reads])
    assert np.array_equal(read_span[reads], got["read_end"][reads] - got["read_start"][reads])
    assert ops.max() <= 2 and np.all((ops[1:] != ops[:-1]) | (owner[1:] != owner[:-1]))
    # the same through the parts pipeline (the resident reads cut into 5 parts on two streams): identical results
    ctx.set_resident_parts(5)
    try:
        ctx.upload_reads(bases, r["quals"], offs)
        st = ctx.map_resident()
        assert st["n_batches"] == 5
        assert_same(ctx.fetch_resident(), want, MAP_KEYS, "map_resident in 5 parts")
        got5 = ctx.best_hits_resident()
        for key in ("best_asv", "strand", "asv_start", "asv_end", "read_start", "read_end", "cigar_offsets", "cigar"):
            assert np.array_equal(got5[key], got[key]), key
        assert np.array_equal(np.isnan(got5["identity"]), np.isnan(got["identity"]))
        assert np.array_equal(got5["identity"][reads], got["identity"][reads])
    finally:
        ctx.set_resident_parts(0)


def test_map_reads_pooled_samples_and_degenerate_reads(ctx):
    from savont_b200 import _lib
    w = _make_workload("pooled", n_roots=6, n_samples=5)
    r = w.reads(3000)
    bases = r["bases"].copy(); quals = r["quals"].copy(); offs = r["offsets"]
    # read 0: all N; read 1: flat qualities; read 2: low quality everywhere (SNPmers gated out)
    bases[int(offs[0]):int(offs[1])] = ord("N")
    quals[int(offs[1]):int(offs[2])] = 33 + 40
    quals[int(offs[2]):int(offs[3])] = 33 + 10
    # append reads that are too short / exactly k
    extra = [b"ACGT", b"ACGTACGTACGTACGTA", b""]
    eb, eo = O.concat(extra)
    bases = np.concatenate([bases, eb]); quals = np.concatenate([quals, np.full(len(eb), 33 + 30, np.uint8)])
    offs = np.concatenate([offs, offs[-1] + eo[1:]])
    fidx = np.concatenate([r["file_idx"], np.array([0, 1, 2], np.uint32)])
    gp, op = _params(_lib, band=w.band)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    want = orc.map_reads(bases, quals, offs, file_idx=fidx, n_samples=5)
    got = ctx.map_reads(bases, quals, offs, file_idx=fidx, n_samples=5)
    assert_same(got, want, MAP_KEYS, "pooled")
    assert want["status"][-3] == O.STATUS_TOO_SHORT and want["status"][-1] == O.STATUS_TOO_SHORT
    assert want["sample_singleton"].sum() == want["unambig"].sum()
    assert len(set(fidx.tolist())) == 5


def test_map_reads_properties_at_scale(ctx):
    """size-independent properties on a larger sample than the oracle comfortably checks:
    counters add up, classes are sorted and within range, reruns are identical."""
    from savont_b200 import _lib
    w = _make_workload("ont16s")
    r = w.reads(60000)
    gp, _ = _params(_lib, band=w.band)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    a = ctx.map_reads(r["bases"], r["quals"], r["offsets"])
    b = ctx.map_reads(r["bases"], r["quals"], r["offsets"])
    assert_same(a, b, MAP_KEYS, "rerun")
    co = a["class_offsets"].astype(np.int64); sizes = co[1:] - co[:-1]
    assert np.all((sizes > 0) == (a["status"] == 0))
    assert int(a["totals"][0]) == int((a["status"] == 0).sum())
    assert int(a["totals"][1]) == int(((a["status"] > 0) & (a["status"] < 4)).sum())
    assert int(a["unambig"].sum()) == int((sizes == 1).sum())
    assert int(a["ambig"].sum()) == int(sizes[sizes > 1].sum())
    cls = a["class_asv"]
    assert cls.max() < w.n_asv
    inner = np.ones(len(cls), bool); inner[co[:-1][sizes > 0]] = False
    assert np.all(cls[1:][inner[1:]] > cls[:-1][inner[1:]])        # sorted inside every class
    single = a["status"] == 0
    hit = cls[co[:-1][single & (sizes == 1)]] == r["truth_asv"][single & (sizes == 1)]
    assert hit.mean() > 0.99


def test_errors_are_loud(ctx):
    from savont_b200 import _lib
    with pytest.raises(_lib.SavontB200Error):
        ctx.set_params(_lib.default_params(k=16))
    with pytest.raises(_lib.SavontB200Error):
        ctx.set_params(_lib.default_params(band=500))
    gp = _lib.default_params()
    ctx.set_params(gp); ctx.load_snpmers([])
    rng = random.Random(3)
    asvs = [rand_seq(rng, 200), rand_seq(rng, 10)]
    ab, ao = O.concat(asvs)
    with pytest.raises(_lib.SavontB200Error):
        ctx.load_asvs(ab, ao)                   # ASV shorter than k
    ab, ao = O.concat(asvs[:1])
    ctx.load_asvs(ab, ao)
    rb, ro = O.concat([rand_seq(rng, 100)])
    with pytest.raises(_lib.SavontB200Error):
        ctx.align_pairs(rb, ro, [0], [5], [0])  # ASV index out of range
    with pytest.raises(_lib.SavontB200Error):
        ctx.map_reads(rb, None, ro, class_cap=None, n_samples=0)


def test_map_reads_edge_cases(ctx):
    """empty batch, only-too-short reads, no qualities, out-of-range sample ids"""
    from savont_b200 import _lib
    w = _make_workload("ont16s", n_roots=3)
    gp, op = _params(_lib, band=w.band)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    # empty
    e = ctx.map_reads(np.zeros(0, np.uint8), None, np.zeros(1, np.uint64))
    assert len(e["status"]) == 0 and e["totals"].tolist() == [0, 0] and e["class_offsets"].tolist() == [0]
    # only too-short reads
    sb, so = O.concat([b"ACGT", b"", b"ACGTACGTACGTACGT"])
    t = ctx.map_reads(sb, None, so)
    assert t["status"].tolist() == [4, 4, 4] and t["totals"].tolist() == [0, 0]
    # no qualities at all (FASTA reads): every SNPmer passes the gate (mid-base Q = 60)
    r = w.reads(800)
    assert_same(ctx.map_reads(r["bases"], None, r["offsets"]), orc.map_reads(r["bases"], None, r["offsets"]), MAP_KEYS, "no quals")
    # file_idx beyond n_samples is rejected (it would index outside the S x A matrix); inside the range: per-sample counts
    fidx = np.arange(800, dtype=np.uint32) % 5
    with pytest.raises(_lib.SavontB200Error):
        ctx.map_reads(r["bases"], r["quals"], r["offsets"], file_idx=fidx, n_samples=3)
    bad = fidx.copy(); bad[17] = 0x80000001
    with pytest.raises(_lib.SavontB200Error):
        ctx.map_reads(r["bases"], r["quals"], r["offsets"], file_idx=bad, n_samples=5)
    got = ctx.map_reads(r["bases"], r["quals"], r["offsets"], file_idx=fidx, n_samples=5)
    want = orc.map_reads(r["bases"], r["quals"], r["offsets"], file_idx=fidx, n_samples=5)
    assert_same(got, want, MAP_KEYS, "five samples")
    assert int(got["sample_singleton"].sum()) == int(got["unambig"].sum())
    # lower-case and IUPAC bases follow BYTE_TO_SEQ
    lb = r["bases"].copy(); lb[::7] |= 0x20; lb[5::501] = ord("R")
    assert_same(ctx.map_reads(lb, r["quals"], r["offsets"]), orc.map_reads(lb, r["quals"], r["offsets"]), MAP_KEYS, "lower case / IUPAC")


def test_many_candidates_trigger_pool_regrowth(ctx):
    """near-identical ASVs: many kept candidates per read -> the pair pool overflows its first sizing and the batch is re-run"""
    from savont_b200 import _lib
    w = _make_workload("ont16s", n_roots=2, n_var=40, sites_per_genus=2)
    r = w.reads(1500)
    gp, op = _params(_lib, band=w.band)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    want = orc.map_reads(r["bases"], r["quals"], r["offsets"])
    got = ctx.map_reads(r["bases"], r["quals"], r["offsets"], class_cap=64 * 1500)
    assert_same(got, want, MAP_KEYS, "pool regrowth")
    assert got["stats"]["n_pairs"] > 2 * 1500 + 1024          # more pairs than the first pool sizing
    assert_same(ctx.find_candidates(r["bases"], r["quals"], r["offsets"], cap=64 * 1500),
                orc.find_candidates(r["bases"], r["quals"], r["offsets"]), CAND_KEYS, "candidates with regrowth")


@pytest.mark.parametrize("with_snpmers", [True, False])
def test_low_polymorphism_mode(ctx, with_snpmers):
    """candidate_mode = 1 (refine_asv_depths_with_minimap2 analogue): all ASVs sharing minimizers are candidates;
    works without any SNPmer (the case that auto-enables the mode in the reference, src/main.rs:543)"""
    from savont_b200 import _lib
    w = _make_workload("ont16s", n_roots=6, sites_per_genus=5)
    r = w.reads(2500)
    snp = w.snpmers if with_snpmers else np.zeros(0, np.uint64)
    gp, op = _params(_lib, band=w.band, candidate_mode=1)
    ctx.set_params(gp); ctx.load_snpmers(snp); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc = O.Oracle(op, snp); orc.load_asvs(w.asv_bases, w.asv_offsets)
    assert_same(ctx.find_candidates(r["bases"], r["quals"], r["offsets"], cap=w.n_asv * 2500),
                orc.find_candidates(r["bases"], r["quals"], r["offsets"]), CAND_KEYS, "low-poly candidates")
    want = orc.map_reads(r["bases"], r["quals"], r["offsets"])
    got = ctx.map_reads(r["bases"], r["quals"], r["offsets"], class_cap=w.n_asv * 2500)
    assert_same(got, want, MAP_KEYS, "low-poly map_reads")
    assert int(want["totals"][0]) == 2500                      # every read finds its genus by minimizers alone
    co = want["class_offsets"].astype(np.int64)
    hit = [r["truth_asv"][i] in want["class_asv"][co[i]:co[i + 1]] for i in range(2500)]
    assert np.mean(hit) > 0.97
    ctx.set_params(_lib.default_params())


@pytest.mark.parametrize("name,kw,n,kmer48,k,c", [
    ("ont16s", dict(n_roots=8), 4000, False, 17, 11),
    ("ont16s", dict(n_roots=5, sites_per_genus=3), 2500, True, 17, 11),
    ("operon", dict(n_roots=4), 500, True, 17, 11),
    ("pooled", dict(n_roots=6, n_samples=4), 2500, False, 17, 11),
    ("ont16s", dict(n_roots=6), 1500, True, 15, 9),               # generic-parameter path (unpack + generic sketch kernel)
])
def test_map_twin_reads(ctx, name, kw, n, kmer48, k, c):
    """The TwinRead entry point (Seq<Dna> words + TwinRead.snpmer_kmers, what Stage 7 is handed by the reference,
    src/alignment.rs:1723-1729) gives the same answer as the ASCII entry point and the oracle."""
    from savont_b200 import _lib, twin
    w = _make_workload(name, k=k, **kw)
    r = w.reads(n)
    bases = r["bases"].copy(); offs = r["offsets"]
    bases[5::997] = ord("N")                                         # non-ACGT -> A in dna_seq (src/seeding.rs:604-626)
    gp, op = _params(_lib, band=w.band, k=k, c=c)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    ns = w.n_samples
    fidx = r["file_idx"] if ns > 1 else None
    want = orc.map_reads(bases, r["quals"], offs, file_idx=fidx, n_samples=ns)
    sk = O.sketch_batch(op, w.snpmers, bases, r["quals"], offs)      # what Stage 1.5 stored in the TwinReads
    tw = twin.from_sketch(sk, offs, file_idx=fidx, kmer48=kmer48)
    ctx.set_batch_reads(900)
    try:
        got = ctx.map_twin_reads(tw, n_samples=ns)
    finally:
        ctx.set_batch_reads(0)
    assert_same(got, want, MAP_KEYS, f"map_twin_reads {name} k={k}")
    assert got["stats"]["dp_cells"] == want["dp_cells"]
    assert got["stats"]["h2d_bytes"] < 0.5 * len(bases)               # 2 bits per base + SNPmers instead of two bytes per base
    # the same reads as a Vec<TwinRead> lies in memory: one heap block per read and field, gathered by the library
    heap = twin.scatter_to_heap(tw, seed=n)
    ctx.set_batch_reads(700); ctx.set_host_threads(3)
    try:
        got_p = ctx.map_twin_read_ptrs(heap, n_samples=ns)
    finally:
        ctx.set_batch_reads(0); ctx.set_host_threads(0)
    assert_same(got_p, want, MAP_KEYS, f"map_twin_read_ptrs {name} k={k}")
    assert got_p["stats"]["gather_ms"] > 0.0
    ctx.upload_twin_reads(tw, n_samples=ns)
    ctx.map_resident()
    assert_same(ctx.fetch_resident(), want, MAP_KEYS, f"twin resident {name}")
    ctx.set_resident_parts(3)
    try:
        ctx.upload_twin_reads(tw, n_samples=ns)
        assert ctx.map_resident()["n_batches"] == 3
        assert_same(ctx.fetch_resident(), want, MAP_KEYS, f"twin resident in 3 parts {name}")
    finally:
        ctx.set_resident_parts(0)
    ctx.set_params(_lib.default_params())


def test_map_twin_reads_edge_cases(ctx):
    from savont_b200 import _lib, twin
    w = _make_workload("ont16s", n_roots=3)
    gp, op = _params(_lib, band=w.band)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    # empty call
    e = ctx.map_twin_reads(dict(dna_words=np.zeros(0, np.uint64), dna_word_offsets=np.zeros(1, np.uint64), lengths=np.zeros(0, np.uint32),
                                snpmer_kmers=np.zeros(0, np.uint64), snpmer_offsets=np.zeros(1, np.uint64)))
    assert len(e["status"]) == 0 and e["totals"].tolist() == [0, 0]
    # reads of awkward lengths around the word size, too short reads, a read without SNPmers
    r = w.reads(300)
    seqs = [bytes(r["bases"][int(r["offsets"][i]):int(r["offsets"][i + 1])]) for i in range(300)]
    quals = [bytes(r["quals"][int(r["offsets"][i]):int(r["offsets"][i + 1])]) for i in range(300)]
    for i, cut in enumerate([0, 5, 16, 17, 31, 32, 33, 63, 64, 65, 95, 96, 97, 1023, 1024, 1025]):
        seqs[i] = seqs[i][:cut]; quals[i] = quals[i][:cut]
    bases, offs = O.concat(seqs); qcat, _ = O.concat(quals)
    want = orc.map_reads(bases, qcat, offs)
    sk = O.sketch_batch(op, w.snpmers, bases, qcat, offs)
    got = ctx.map_twin_reads(twin.from_sketch(sk, offs, kmer48=True))
    assert_same(got, want, MAP_KEYS, "twin ragged")
    assert got["status"][0] == 4 and got["status"][2] == 4 and got["status"][3] != 4
    # garbage beyond the last base inside the last word must be ignored
    tw = twin.from_sketch(sk, offs)
    words = tw["dna_words"].copy()
    wo = tw["dna_word_offsets"].astype(np.int64); ln = tw["lengths"].astype(np.int64)
    for i in range(300):
        rem = int(ln[i] % 32)
        if rem and ln[i] > 0:
            words[wo[i + 1] - 1] |= np.uint64(((1 << 64) - 1) ^ ((1 << (2 * rem)) - 1))
    tw["dna_words"] = words
    assert_same(ctx.map_twin_reads(tw), want, MAP_KEYS, "twin with dirty tail bits")
    # invalid arguments are loud
    bad = dict(tw); bad["dna_word_offsets"] = tw["dna_word_offsets"].copy(); bad["dna_word_offsets"][5:] -= np.uint64(3)
    with pytest.raises(_lib.SavontB200Error):
        ctx.map_twin_reads(bad)


@pytest.mark.parametrize("k,c", [(17, 11), (15, 9)])
def test_twin_reads_from_snpmers_ragged(ctx, k, c):
    """a4 (kmer_comp::twin_reads_from_snpmers): per-read tests on the GPU, order on the host, vs the oracle"""
    from savont_b200 import _lib, api
    from tests.test_oracle import _a4_inputs
    reads, quals, ids = _a4_inputs(31 + k, k)
    snp = sorted(snpmer_set_from(reads, k))
    gp, op = _params(_lib, k=k, c=c)
    ctx.set_params(gp); ctx.load_snpmers(snp)
    sk = O.sketch_batch(op, snp, *O.concat(reads)[:1], O.concat(quals)[0], O.concat(reads)[1])
    mo = sk["mini_offsets"].astype(np.int64)
    hf = sorted(set(int(x) for i in (4, 10, 16) for x in sk["mini_kmer"][mo[i]:mo[i + 1]]))
    bases, offs = O.concat(reads); qcat, _ = O.concat(quals)
    for (mn, mx, cut, maxc) in [(0, 0, 0.0, 500), (300, 640, 97.0, 500), (17, 0, 99.5, 3), (400, 500, 90.0, 1)]:
        want = O.twin_reads_filter(op, O.read_filter(mn, mx, cut, maxc), snp, hf, bases, qcat, offs, ids)
        ctx.load_high_freq_kmers(hf)
        ctx.set_read_filter(_lib.read_filter(mn, mx, cut, maxc))
        try:
            got = ctx.twin_reads_from_snpmers(bases, qcat, offs)
        finally:
            ctx.set_read_filter(None)
        assert_same(got, want, ["status", "est_id", "n_solid_mini", "n_solid_snp"], f"a4 k={k} {(mn, mx, cut, maxc)}")
    # host layer: defaults of the Cli + the Stage-7 read order
    args = api.Cli(kmer_size=k, c=c, min_read_length=300, max_read_length=900, quality_value_cutoff=97.0)
    got = api.twin_reads_from_snpmers(ctx, dict(bases=bases, quals=qcat, offsets=offs, ids=ids), args, high_freq_kmers=hf)
    want = O.twin_reads_filter(op, O.read_filter(300, 900, 97.0), snp, hf, bases, qcat, offs, ids)
    assert_same(got, want, ["status", "order"], "a4 host order")
    assert len(want["order"]) > 5 and set(want["status"].tolist()) >= {0, 5, 6, 7}
    ctx.load_high_freq_kmers([])
    ctx.set_params(_lib.default_params())
    with pytest.raises(_lib.SavontB200Error):
        ctx.twin_reads_from_snpmers(bases, qcat, offs)             # no filter set


def test_map_reads_with_read_filter(ctx):
    """FASTQ reads straight into Stage 7: reads failing the a4 tests get status 5 / 6 / 7 and take no part"""
    from savont_b200 import _lib
    w = _make_workload("ont16s", n_roots=6)
    r = w.reads(3000)
    quals = r["quals"].copy(); offs = r["offsets"]
    for i in range(0, 3000, 17):                                     # low quality, not flat
        quals[int(offs[i]):int(offs[i + 1])] = 33 + 9
        quals[int(offs[i])] = 33 + 12
    gp, op = _params(_lib, band=w.band)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    sk = O.sketch_batch(op, w.snpmers, r["bases"], quals, offs)
    mo = sk["mini_offsets"].astype(np.int64)
    genus0 = np.flatnonzero(r["truth_asv"] < 10)[:40]                # the reads of one genus lose (nearly) all minimizers
    hf = sorted(set(int(x) for i in genus0 for x in sk["mini_kmer"][mo[i]:mo[i + 1]]))
    rf_args = (1475, 1525, 98.0)
    orc.set_read_filter(O.read_filter(*rf_args), hf)
    want = orc.map_reads(r["bases"], quals, offs)
    ctx.load_high_freq_kmers(hf); ctx.set_read_filter(_lib.read_filter(*rf_args))
    try:
        ctx.set_batch_reads(800)
        got = ctx.map_reads(r["bases"], quals, offs)
        ctx.set_batch_reads(0)
        assert_same(got, want, MAP_KEYS, "map_reads with read filter")
        assert set(want["status"].tolist()) >= {0, 5, 6, 7}
        ctx.upload_reads(r["bases"], quals, offs); ctx.map_resident()
        assert_same(ctx.fetch_resident(), want, MAP_KEYS, "map_resident with read filter")
        cand = ctx.find_candidates(r["bases"], quals, offs)
        assert np.array_equal(cand["status"][want["status"] >= 4], want["status"][want["status"] >= 4])
    finally:
        ctx.set_batch_reads(0); ctx.set_read_filter(None); ctx.load_high_freq_kmers([])
    orc.set_read_filter(None)
    assert_same(ctx.map_reads(r["bases"], quals, offs), orc.map_reads(r["bases"], quals, offs), MAP_KEYS, "filter off again")


KMER_KEYS = ["kmers", "counts"]


@pytest.mark.parametrize("k,single", [(17, False), (17, True), (15, False), (21, False)])
def test_count_split_kmers_ragged(ctx, k, single):
    """K5 (Stage 1 counting, seq_parse::read_to_split_kmers): rc-tagged reads, N / lower case / U, palindromic splitmers,
    flat and low qualities, reads shorter than k"""
    from savont_b200 import _lib
    from tests.test_oracle import _kmer_count_inputs
    reads, quals, rc = _kmer_count_inputs(40 + k)
    bases, offs = O.concat(reads); qcat, _ = O.concat(quals)
    ctx.set_params(_lib.default_params(k=k, c=min(11, k - 2)))
    for q in (qcat, None):
        for r in (rc, None):
            want = O.count_split_kmers(bases, q, offs, rc=r, k=k, min_bq=25, single_strand=single)
            got = ctx.count_split_kmers(bases, q, offs, rc=r, single_strand=single)
            assert_same(got, want, KMER_KEYS, f"count_split_kmers k={k} single={single} quals={q is not None} rc={r is not None}")
            assert len(want["kmers"]) > 20 and got["stats"]["n_distinct"] == want["n_distinct"]
    ctx.set_params(_lib.default_params(min_bq=30))                 # a different quality gate
    if k == 17:
        assert_same(ctx.count_split_kmers(bases, qcat, offs, rc=rc), O.count_split_kmers(bases, qcat, offs, rc=rc, k=17, min_bq=30),
                    KMER_KEYS, "min_bq 30")
    empty = ctx.count_split_kmers(np.zeros(0, np.uint8), None, np.zeros(1, np.uint64))
    assert len(empty["kmers"]) == 0
    ctx.set_params(_lib.default_params())


def test_count_split_kmers_synthetic_reads(ctx):
    """16S-like redundancy: the counts of 6000 reads, the output capacity regrowth, the resident variant and the
    high-frequency rule on the result"""
    from savont_b200 import _lib
    w = _make_workload("ont16s", n_roots=6)
    r = w.reads(6000)
    rc = (np.arange(6000) % 7 == 0).astype(np.uint8)
    ctx.set_params(_lib.default_params())
    want = O.count_split_kmers(r["bases"], r["quals"], r["offsets"], rc=rc, n_threads=0)
    got = ctx.count_split_kmers(r["bases"], r["quals"], r["offsets"], rc=rc, cap=1000)     # too small on purpose -> regrow
    assert_same(got, want, KMER_KEYS, "synthetic reads")
    assert len(want["kmers"]) > 50000
    st = got["stats"]
    assert st["n_distinct"] == want["n_distinct"] and st["partitions"] == 8192 and st["spilled_distinct"] == 0
    ctx.upload_reads(r["bases"], r["quals"], r["offsets"])
    assert_same(ctx.count_split_kmers_resident(), O.count_split_kmers(r["bases"], r["quals"], r["offsets"]), KMER_KEYS, "resident")
    th, hf = O.high_freq_kmers(want["kmers"], want["counts"])
    assert th >= 100 and len(hf) < len(want["kmers"])


@pytest.mark.parametrize("env,expect", [({"SAVONT_B200_KC_PARTITIONED": "0"}, "single"), ({"SAVONT_B200_KC_SMEM_LIMIT": "40"}, "spill")])
def test_count_split_kmers_other_algorithms(env, expect):
    """the single-table algorithm (also the fallback for k > 21) and the spill path of the partitioned one (shared-memory
    rounds that overflow are counted in the global table) give the same answer"""
    import os
    from savont_b200 import _lib
    w = _make_workload("ont16s", n_roots=5)
    r = w.reads(4000)
    rc = (np.arange(4000) % 4 == 1).astype(np.uint8)
    want = O.count_split_kmers(r["bases"], r["quals"], r["offsets"], rc=rc)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        c2 = _lib.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    try:
        got = c2.count_split_kmers(r["bases"], r["quals"], r["offsets"], rc=rc)
        assert_same(got, want, KMER_KEYS, expect)
        assert got["stats"]["n_distinct"] == want["n_distinct"]
        if expect == "single":
            assert got["stats"]["partitions"] == 0
        else:
            assert got["stats"]["partitions"] == 8192 and got["stats"]["spilled_distinct"] > 1000
        c2.set_params(_lib.default_params(k=23, c=11))                 # k = 23: 46-bit k-mers do not fit the 32-bit entries
        got = c2.count_split_kmers(r["bases"], r["quals"], r["offsets"], rc=rc)
        assert_same(got, O.count_split_kmers(r["bases"], r["quals"], r["offsets"], rc=rc, k=23), KMER_KEYS, "k=23")
        assert got["stats"]["partitions"] == 0
    finally:
        c2.close()


def test_count_split_kmers_long_reads(ctx):
    """reads that fill a level-1 tile on their own, and reads longer than the partitioned kernels support (single table)"""
    from savont_b200 import _lib
    rng = random.Random(77)
    root = rand_seq(rng, 30000)
    ctx.set_params(_lib.default_params())
    for L, expect_parts in ((20000, 8192), (26000, 0)):
        base = (root * 3)[:L]
        reads = [mutate(rng, base, 40, 10, 10) for _ in range(6)] + [rand_seq(rng, 400), b"ACGT"]
        comp = bytes.maketrans(b"ACGT", b"TGCA")
        reads[1] = reads[1].translate(comp)[::-1]; reads[3] = reads[3].translate(comp)[::-1]
        quals = [bytes(33 + rng.choice([20, 30, 40]) for _ in range(len(s))) for s in reads]
        bases, offs = O.concat(reads); qcat, _ = O.concat(quals)
        want = O.count_split_kmers(bases, qcat, offs)
        got = ctx.count_split_kmers(bases, qcat, offs)
        assert_same(got, want, KMER_KEYS, f"long reads {L}")
        assert got["stats"]["partitions"] == expect_parts and got["stats"]["n_distinct"] == want["n_distinct"]
        assert len(want["kmers"]) > 10000


def test_candidates_large_asv_set(ctx):
    """more than 1024 ASVs: the compact candidate table of the filter (dense kernel only for reads that overflow it)"""
    from savont_b200 import _lib
    w = _make_workload("pooled", n_roots=130, n_samples=3)            # 1300 ASVs
    r = w.reads(2500)
    gp, op = _params(_lib, band=w.band)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    assert w.n_asv > 1024
    assert_same(ctx.find_candidates(r["bases"], r["quals"], r["offsets"]), orc.find_candidates(r["bases"], r["quals"], r["offsets"]),
                CAND_KEYS, "candidates A=1300")
    want = orc.map_reads(r["bases"], r["quals"], r["offsets"], file_idx=r["file_idx"], n_samples=3)
    got = ctx.map_reads(r["bases"], r["quals"], r["offsets"], file_idx=r["file_idx"], n_samples=3)
    assert_same(got, want, MAP_KEYS, "map_reads A=1300")
    assert got["stats"]["kernel_launches"] >= 9                       # sketch, compact filter + dense redo, align, assign x5
    gp2, op2 = _params(_lib, band=w.band, strand_mode=1)
    ctx.set_params(gp2); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc2 = O.Oracle(op2, w.snpmers); orc2.load_asvs(w.asv_bases, w.asv_offsets)
    assert_same(ctx.map_reads(r["bases"][:int(r["offsets"][600])], r["quals"][:int(r["offsets"][600])], r["offsets"][:601]),
                orc2.map_reads(r["bases"][:int(r["offsets"][600])], r["quals"][:int(r["offsets"][600])], r["offsets"][:601]), MAP_KEYS,
                "A=1300 both strands")
    ctx.set_params(_lib.default_params())


PILEUP_KEYS = ["pos_offsets", "ref_hp_length", "read_strand", "read_nm"]


def _assert_pileups(got, want, what):
    assert_same(got, want, PILEUP_KEYS, what)
    assert got["entries"].tobytes() == want["entries"].tobytes(), f"{what}: entries differ"


@pytest.mark.parametrize("use_hpc,band,long_runs", [(True, 16, True), (False, 40, False), (True, 3, False), (True, 64, False)])
def test_consensus_pileups_ragged_clusters(ctx, use_hpc, band, long_runs):
    """Stage 4 pileups: empty reads / consensus, consensus without reads, N bases, runs beyond 255, both strands"""
    from savont_b200 import _lib
    from tests.util import pileup_case
    gp, _ = _params(_lib, band=band)
    ctx.set_params(gp)
    cons, reads, quals, rcons = pileup_case(5200 + band, n_cons=7, reads_per=15, long_runs=long_runs)
    cb, co = O.concat(cons); rb, ro = O.concat(reads); qb, _ = O.concat(quals)
    want = O.consensus_pileups(cb, co, rb, qb, ro, rcons, band=band, use_hpc=use_hpc)
    got = ctx.consensus_pileups(cb, co, rb, qb, ro, rcons, use_hpc=use_hpc)
    _assert_pileups(got, want, f"pileups hpc={use_hpc} band={band}")
    assert len(got["entries"]) > 500
    # without qualities (quality 33 everywhere), no reads at all, no consensus at all
    _assert_pileups(ctx.consensus_pileups(cb, co, rb, None, ro, rcons, use_hpc=use_hpc),
                    O.consensus_pileups(cb, co, rb, None, ro, rcons, band=band, use_hpc=use_hpc), "pileups without qualities")
    e = ctx.consensus_pileups(cb, co, b"", None, [0], [], use_hpc=use_hpc)
    assert len(e["entries"]) == 0 and np.all(e["ref_hp_length"] == 1) and np.all(e["pos_offsets"] == 0)
    e = ctx.consensus_pileups(b"", [0], b"", None, [0], [], use_hpc=use_hpc)
    assert len(e["entries"]) == 0 and len(e["pos_offsets"]) == 1
    # a too small entry buffer is an error, not a truncation
    with pytest.raises(_lib.SavontB200Error):
        ctx.consensus_pileups(cb, co, rb, qb, ro, rcons, use_hpc=use_hpc, cap=len(got["entries"]) - 1)


def test_consensus_pileups_many_chunks(ctx, monkeypatch):
    """chunks of <= 20 reads: the double-buffered entry copies and the call-wide read / entry indices"""
    from savont_b200 import _lib
    from tests.util import pileup_case
    gp, _ = _params(_lib, band=24)
    ctx.set_params(gp)
    cons, reads, quals, rcons = pileup_case(77, n_cons=9, reads_per=11)
    cb, co = O.concat(cons); rb, ro = O.concat(reads); qb, _ = O.concat(quals)
    want = O.consensus_pileups(cb, co, rb, qb, ro, rcons, band=24, use_hpc=True)
    monkeypatch.setenv("SAVONT_B200_PILEUP_CHUNK_READS", "20")
    got = ctx.consensus_pileups(cb, co, rb, qb, ro, rcons, use_hpc=True)
    assert got["stats"]["n_batches"] >= 4
    _assert_pileups(got, want, "pileups in chunks")
    monkeypatch.setenv("SAVONT_B200_PILEUP_HOST_CIGAR", "1")      # CIGARs through the host, as when a traceback needs several sub-batches
    _assert_pileups(ctx.consensus_pileups(cb, co, rb, qb, ro, rcons, use_hpc=True), want, "pileups, CIGARs through the host")


def test_consensus_pileups_full_clusters(ctx):
    """16S-sized clusters: 250 reads (MAX_SEQS_CONSENSUS) of ~1500 bp at 2% errors per consensus, band 64"""
    from savont_b200 import _lib, synth
    w = synth.Workload("ont16s", n_roots=2)
    n_cons, per = 6, 250
    gp, _ = _params(_lib, band=64)
    ctx.set_params(gp)
    rng = np.random.default_rng(7)
    asv = [bytes(w.asv_bases[int(w.asv_offsets[a]):int(w.asv_offsets[a + 1])]) for a in range(n_cons)]
    from tests import pyref
    cons = [bytes(b"ACGT"[x] for x in pyref.hpc_with_quality(s, bytes(len(s)))[0]) for s in asv]
    r = w.reads(4000)
    reads, quals, rcons = [], [], []
    for a in range(n_cons):
        idx = np.flatnonzero(r["truth_asv"] % n_cons == a)[:per]
        for i in idx:
            o0, o1 = int(r["offsets"][i]), int(r["offsets"][i + 1])
            # reads drawn from other ASVs of the same genus: realistic mismatches against this consensus
            reads.append(bytes(r["bases"][o0:o1])); quals.append(bytes(r["quals"][o0:o1])); rcons.append(a)
    cb, co = O.concat(cons); rb, ro = O.concat(reads); qb, _ = O.concat(quals)
    want = O.consensus_pileups(cb, co, rb, qb, ro, rcons, band=64, use_hpc=True)
    got = ctx.consensus_pileups(cb, co, rb, qb, ro, rcons, use_hpc=True)
    _assert_pileups(got, want, "full clusters")
    assert len(got["entries"]) > 0.5 * len(reads) * 1000
    del rng


@pytest.mark.parametrize("seed,err,n,block", [(11, 0.02, 400, None), (12, 0.005, 400, "7"), (13, 0.04, 500, "64"), (14, 0.02, 300, "1")])
def test_lsh_signatures_and_kmer_clustering(ctx, monkeypatch, seed, err, n, block):
    """Stage 2: LSH signatures and the greedy k-mer clustering against the sequential oracle, with block sizes that force
    in-block representatives (block=None: the adaptive default)"""
    from savont_b200 import _lib
    from tests.util import clustering_case
    gp, _ = _params(_lib)
    ctx.set_params(gp)
    km, off, _ = clustering_case(800 + seed, n_reads=n, err=err)
    sig, valid = ctx.lsh_signatures(km, off)
    wsig, wvalid = O.lsh_signatures(km, off)
    assert np.array_equal(valid, wvalid) and np.array_equal(sig[valid == 1], wsig[wvalid == 1])
    if block:
        monkeypatch.setenv("SAVONT_B200_CLUSTER_BLOCK", block)
    got = ctx.cluster_reads_by_kmers(km, off)
    want = O.cluster_reads_by_kmers(km, off)
    assert np.array_equal(got["rep_of"], want), np.flatnonzero(got["rep_of"] != want)[:10]
    n_rep = int((want == np.arange(len(want))).sum())
    assert got["n_reps"] == n_rep and 5 < n_rep < n
    # another threshold, empty input, reads without minimizers only
    for thr in (0.9, 0.99):
        assert np.array_equal(ctx.cluster_reads_by_kmers(km, off, threshold=thr)["rep_of"], O.cluster_reads_by_kmers(km, off, threshold=thr))
    e = ctx.cluster_reads_by_kmers(np.zeros(0, np.uint64), [0])
    assert len(e["rep_of"]) == 0 and e["n_reps"] == 0
    e = ctx.cluster_reads_by_kmers(np.array([5, 6], np.uint64), [0, 0, 2, 2])
    assert list(e["rep_of"]) == [0, 1, 2]


def test_kmer_clustering_large_groups(ctx, monkeypatch):
    """many near-identical reads: max-hits groups beyond the first 10 candidates, thousands of bucket entries per read"""
    from savont_b200 import _lib, synth
    gp, op = _params(_lib)
    ctx.set_params(gp)
    w = synth.Workload("ont16s", n_roots=3)
    r = w.reads(6000)
    sk = ctx.sketch(r["bases"], r["quals"], r["offsets"], want_codes=False)
    km, off = np.asarray(sk["mini_kmer"], np.uint64), np.asarray(sk["mini_offsets"], np.uint64)
    want = O.cluster_reads_by_kmers(km, off, threshold=0.97)       # a high threshold: many representatives per species
    n_rep = int((want == np.arange(len(want))).sum())
    assert n_rep > 300
    got = ctx.cluster_reads_by_kmers(km, off, threshold=0.97)
    assert np.array_equal(got["rep_of"], want), np.flatnonzero(got["rep_of"] != want)[:10]
    monkeypatch.setenv("SAVONT_B200_CLUSTER_BLOCK", "1500")
    got = ctx.cluster_reads_by_kmers(km, off, threshold=0.97)
    assert np.array_equal(got["rep_of"], want), np.flatnonzero(got["rep_of"] != want)[:10]
    got = ctx.cluster_reads_by_kmers(km, off)
    assert np.array_equal(got["rep_of"], O.cluster_reads_by_kmers(km, off))


@pytest.mark.parametrize("world", [1, 2, 3])
def test_count_split_kmers_sharded_ranks_on_one_gpu(world):
    """the multi-GPU k-mer counting with its ranks emulated by `world` contexts on one device: the peers' entry buffers are plain
    device pointers here (cudaIpc-mapped under torchrun); the union of the shards must be the single-GPU / oracle result"""
    from savont_b200 import _lib, dist as sd, synth
    w = synth.Workload("ont16s", n_roots=3)
    r = w.reads(3000)
    ctxs = [_lib.Context(0, _lib.default_params(band=w.band)) for _ in range(world)]
    try:
        for rk, c in enumerate(ctxs):
            lo, hi = sd.shard_bounds(3000, world, rk)
            sh = sd.slice_reads(r, lo, hi)
            c.upload_reads(sh["bases"], sh["quals"], sh["offsets"])
        H = np.stack([c.kmer_shard_hist() for c in ctxs]).astype(np.uint64)
        owner, part_off, dst_off = sd.kmer_shard_layout(H)
        ptrs = [c.kmer_shard_alloc(int(part_off[rk, -1])) for rk, c in enumerate(ctxs)]
        for rk, c in enumerate(ctxs):
            c.kmer_shard_scatter(ptrs, dst_off[rk])
        # rank 0 starts with far too small an output: ERR_CAPACITY must leave the shard recoverable (second call with the size the
        # first one reported, nothing counted twice)
        outs = [c.kmer_shard_count(part_off[rk], cap=(64 if rk == 0 else 1 << 20)) for rk, c in enumerate(ctxs)]
        km = np.concatenate([o["kmers"] for o in outs]); ct = np.concatenate([o["counts"] for o in outs])
        o = np.argsort(km, kind="stable"); km, ct = km[o], ct[o]
        want = O.count_split_kmers(r["bases"], r["quals"], r["offsets"])
        assert len(want["kmers"]) > 1000
        assert np.array_equal(km, want["kmers"]) and np.array_equal(ct, want["counts"])
        if world > 1:
            assert all(len(o_["kmers"]) > 0 for o_ in outs)            # every rank owns part of the result
        # the phases must be called in order
        with pytest.raises(_lib.SavontB200Error):
            ctxs[0].kmer_shard_count(part_off[0])
    finally:
        for c in ctxs:
            c.close()


SNPCMP_KEYS = ["best_cons", "best_mismatches", "best_matches"]


@pytest.mark.parametrize("seed,k", [(1, 17), (2, 17), (3, 15), (4, 21)])
def test_snpmer_compare_small_cases(ctx, seed, k):
    """Stage 3 reclustering: reads and consensus lists against cluster consensus SNPmers (empty reads / lists, unknown sites,
    duplicated splitmers, identical lists)"""
    from savont_b200 import _lib
    from tests.util import csr_u64, reassign_case
    gp, _ = _params(_lib, k=k, c=k - 6)
    ctx.set_params(gp)
    cons, reads = reassign_case(90 + seed, k=k)
    ck, co = csr_u64(cons); rk, ro = csr_u64(reads)
    want = O.snpmer_compare(ck, co, rk, ro, k=k, want_matrix=True)
    got = ctx.snpmer_compare(ck, co, rk, ro, want_matrix=True)
    assert_same(got, want, SNPCMP_KEYS + ["matrix"], f"snpmer compare k={k}")
    assert_same(ctx.snpmer_compare(ck, co, rk, ro), want, SNPCMP_KEYS, "without matrix")
    # consensus lists against each other, no clusters, no reads
    assert_same(ctx.snpmer_compare(ck, co, ck, co, want_matrix=True), O.snpmer_compare(ck, co, ck, co, k=k, want_matrix=True),
                SNPCMP_KEYS + ["matrix"], "consensus vs consensus")
    e = ctx.snpmer_compare(np.zeros(0, np.uint64), [0], rk, ro)
    assert np.all(e["best_cons"] == 0xffffffff)
    assert len(ctx.snpmer_compare(ck, co, np.zeros(0, np.uint64), [0])["best_cons"]) == 0


def test_snpmer_compare_16s_clusters(ctx):
    """500 'clusters' (the ASVs' SNPmers as consensus lists) against the SNPmers of 20 k ONT reads"""
    from savont_b200 import _lib, synth
    w = synth.Workload("ont16s")
    gp, _ = _params(_lib)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers)
    sa = ctx.sketch(w.asv_bases, None, w.asv_offsets, want_codes=False)
    r = w.reads(20000)
    sr = ctx.sketch(r["bases"], r["quals"], r["offsets"], want_codes=False)
    ck, co = np.asarray(sa["snp_kmer"], np.uint64), np.asarray(sa["snp_offsets"], np.uint64)
    rk, ro = np.asarray(sr["snp_kmer"], np.uint64), np.asarray(sr["snp_offsets"], np.uint64)
    want = O.snpmer_compare(ck, co, rk[:int(ro[3000])], ro[:3001], want_matrix=True)
    got = ctx.snpmer_compare(ck, co, rk, ro)
    assert_same({k_: got[k_][:3000] for k_ in SNPCMP_KEYS}, want, SNPCMP_KEYS, "16S clusters")
    gm = ctx.snpmer_compare(ck, co, rk[:int(ro[3000])], ro[:3001], want_matrix=True)
    assert_same(gm, want, SNPCMP_KEYS + ["matrix"], "16S clusters, matrix")
    # most reads are reassigned to an ASV of their own genus with several matching SNPmers
    assert float((got["best_matches"] > 2).mean()) > 0.8


def test_count_split_kmers_sharded_processes_on_one_gpu():
    """the real multi-process driver (savont_b200/dist.py: one process per rank, cudaIpc-mapped peer buffers, buffer growth between
    calls) with both ranks on this GPU and gloo collectives; tools/shard_check.py compares the union of the shards with the
    single-GPU count and exits non-zero on a mismatch"""
    import os
    import socket
    import subprocess
    import sys
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SHARD_CHECK_BACKEND="gloo")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "tools", "shard_check.py"), "8000"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "world=2" in r.stdout and " OK" in r.stdout


@pytest.mark.parametrize("n_ctx,allreduce", [(1, "peer"), (3, "peer"), (2, "nccl")])
def test_multi_handle_matches_single_context_and_oracle(n_ctx, allreduce):
    """savont_b200_create_multi: one process, a context per device (here: per entry of the device list -- on a one-GPU box several
    contexts share device 0; with >= n_ctx GPUs they are distinct devices), reads sharded in read order, counter vectors summed by
    the peer-memory kernel or NCCL, results concatenated: every entry point must reproduce the oracle for the WHOLE read set"""
    import torch
    from savont_b200 import _lib, twin
    n_gpu = torch.cuda.device_count()
    devs = list(range(n_ctx)) if n_gpu >= n_ctx else [0] * n_ctx
    if allreduce == "nccl" and len(set(devs)) < n_ctx:
        pytest.skip("the NCCL all-reduce needs distinct devices")
    w = _make_workload("pooled", n_roots=5, n_samples=3)
    r = w.reads(2600)
    gp, op = _params(_lib, band=w.band)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    want = orc.map_reads(r["bases"], r["quals"], r["offsets"], file_idx=r["file_idx"], n_samples=3)
    m = _lib.MultiContext(devs, gp, allreduce=allreduce)
    try:
        m.load_snpmers(w.snpmers); m.load_asvs(w.asv_bases, w.asv_offsets)
        m.set_batch_reads(500)
        got = m.map_reads(r["bases"], r["quals"], r["offsets"], file_idx=r["file_idx"], n_samples=3)
        assert_same(got, want, MAP_KEYS, f"multi map_reads x{n_ctx}")
        assert got["stats"]["n_reads"] == 2600 and got["stats"]["n_batches"] >= n_ctx
        sk = O.sketch_batch(op, w.snpmers, r["bases"], r["quals"], r["offsets"])
        tw = twin.from_sketch(sk, r["offsets"], file_idx=r["file_idx"], kmer48=True)
        assert_same(m.map_twin_reads(tw, n_samples=3), want, MAP_KEYS, f"multi map_twin_reads x{n_ctx}")
        m.set_host_threads(4)
        assert_same(m.map_twin_read_ptrs(twin.scatter_to_heap(tw), n_samples=3), want, MAP_KEYS, f"multi map_twin_read_ptrs x{n_ctx}")
        # fewer reads than devices, and none at all
        few = {k_: (v[:2] if k_ in ("lengths", "file_idx") else v) for k_, v in tw.items()}
        few["dna_word_offsets"] = tw["dna_word_offsets"][:3]; few["snpmer_offsets"] = tw["snpmer_offsets"][:3]
        few["snpmer_kmers"] = tw["snpmer_kmers"][:6 * int(tw["snpmer_offsets"][2])]
        g2 = m.map_twin_reads(few, n_samples=3)
        assert np.array_equal(g2["status"], want["status"][:2]) and np.array_equal(g2["best_nm"], want["best_nm"][:2])
        e = m.map_reads(np.zeros(0, np.uint8), None, np.zeros(1, np.uint64))
        assert len(e["status"]) == 0 and e["totals"].tolist() == [0, 0]
        # an error on one device surfaces with its message
        with pytest.raises(_lib.SavontB200Error):
            m.map_reads(r["bases"], r["quals"], r["offsets"], file_idx=np.full(2600, 7, np.uint32), n_samples=3)
    finally:
        m.close()


def _stage3_case(seed, n_reads, n_clusters, n_sites, dup_frac=0.3):
    """reads = SNPmer lists over a few biallelic sites per k-mer cluster: haplotypes + random allele flips / drop-outs, many exact
    duplicates (ties in matches and cluster sizes), reads without SNPmers, reads outside every cluster"""
    rng = np.random.default_rng(seed)
    k = 17
    mid = k // 2
    lists, members = [], [[] for _ in range(n_clusters)]
    site_sp = rng.integers(0, 1 << 30, size=(n_clusters, n_sites)).astype(np.uint64)
    def kmer(sp, allele):
        lo = int(sp) & ((1 << (2 * mid)) - 1); hi = int(sp) >> (2 * mid)
        return (hi << (2 * mid + 2)) | (int(allele) << (2 * mid)) | lo
    haps = rng.integers(0, 2, size=(n_clusters, 6, n_sites))
    for r in range(n_reads):
        c = int(rng.integers(0, n_clusters + 1))
        if lists and rng.random() < dup_frac:
            lists.append(list(lists[int(rng.integers(0, len(lists)))]))
        else:
            cc = min(c, n_clusters - 1)
            h = haps[cc, int(rng.integers(0, 6))].copy()
            flip = rng.random(n_sites) < 0.05
            keep = rng.random(n_sites) < 0.8
            alle = np.where(flip, 2 + h, h)                                    # a third / fourth allele now and then
            lists.append([kmer(site_sp[cc, s], alle[s]) for s in range(n_sites) if keep[s]] if rng.random() > 0.03 else [])
        if c < n_clusters:
            members[c].append(r)
    for m_ in members:
        rng.shuffle(m_)
    offs = np.zeros(n_reads + 1, np.uint64); offs[1:] = np.cumsum([len(x) for x in lists])
    km = np.array([v for x in lists for v in x], np.uint64)
    coffs = np.zeros(n_clusters + 1, np.uint64); coffs[1:] = np.cumsum([len(m_) for m_ in members])
    cr = np.array([v for m_ in members for v in m_], np.uint32)
    return km, offs, cr, coffs


@pytest.mark.parametrize("seed,n_reads,n_clusters,n_sites,above", [(1, 3000, 7, 12, 1000), (2, 6000, 3, 20, 40), (3, 500, 40, 5, 1000), (4, 2500, 1, 30, 5),
                                                               (5, 3000, 2, 80, 1000), (6, 20000, 3, 3, 1000)])
def test_stage3_greedy_snpmer_clustering(ctx, seed, n_reads, n_clusters, n_sites, above):
    """savont_b200_cluster_reads_by_snpmers (one CTA per k-mer cluster, reads in sequence) == the oracle's restatement of
    asv_cluster::cluster_reads_by_snpmers, cluster by cluster, in the indexed branch and beyond `above` representatives"""
    from savont_b200 import _lib
    ctx.set_params(_lib.default_params())
    km, offs, cr, coffs = _stage3_case(seed, n_reads, n_clusters, n_sites)
    got = ctx.cluster_reads_by_snpmers(km, offs, cr, coffs, iterative_above=above)
    want = np.full(n_reads, 0xffffffff, np.uint32)
    for c in range(n_clusters):
        mem = cr[int(coffs[c]):int(coffs[c + 1])]
        w_ = O.cluster_reads_by_snpmers(km, offs, mem, iterative_above=above)
        want[mem] = w_[mem]
    assert np.array_equal(got["rep_of"], want)
    reps = np.flatnonzero(got["rep_of"] == np.arange(n_reads))
    assert got["n_representatives"] == len(reps) and len(reps) > n_clusters
    # no clusters / empty clusters
    e = ctx.cluster_reads_by_snpmers(km, offs, np.zeros(0, np.uint32), np.zeros(1, np.uint64))
    assert np.all(e["rep_of"] == 0xffffffff) and e["n_representatives"] == 0
    # a read with two SNPmers on one splitmer is rejected
    bad = km.copy(); first = int(np.flatnonzero((offs[1:] - offs[:-1]) >= 2)[0]); bad[int(offs[first]) + 1] = bad[int(offs[first])] ^ np.uint64(1 << 16)
    if first in set(cr.tolist()):
        with pytest.raises(_lib.SavontB200Error):
            ctx.cluster_reads_by_snpmers(bad, offs, cr, coffs)


def test_stage3_many_tied_representatives(ctx):
    """16 mutually incompatible representatives that all contain one shared SNPmer; reads carrying only that SNPmer tie on all of
    them (more than the per-read state holds: the warp re-scores the read), the live cluster sizes and then the read ids decide"""
    from savont_b200 import _lib
    ctx.set_params(_lib.default_params())
    rng = np.random.default_rng(7)
    k = 17; mid = k // 2
    def kmer(sp, allele):
        lo = int(sp) & ((1 << (2 * mid)) - 1); hi = int(sp) >> (2 * mid)
        return (hi << (2 * mid + 2)) | (int(allele) << (2 * mid)) | lo
    s0, s1, s2, s3 = (int(x) for x in rng.integers(0, 1 << 30, 4))
    lists = [[kmer(s0, 0), kmer(s1, x), kmer(s2, y)] for x in range(4) for y in range(4)]
    for _ in range(3000):
        u = rng.random()
        if u < 0.5: lists.append([kmer(s0, 0)])                                               # ties on all 16
        elif u < 0.7: lists.append([kmer(s0, 0), kmer(s1, int(rng.integers(0, 4)))])          # ties on 4
        elif u < 0.8: lists.append([kmer(s1, int(rng.integers(0, 4))), kmer(s2, int(rng.integers(0, 4))), kmer(s3, 1)])
        elif u < 0.9: lists.append([kmer(s0, 0), kmer(s1, int(rng.integers(0, 4))), kmer(s2, int(rng.integers(0, 4)))])
        else: lists.append([kmer(s0, 1)] if rng.random() < 0.5 else [])
    n = len(lists)
    offs = np.zeros(n + 1, np.uint64); offs[1:] = np.cumsum([len(x) for x in lists])
    km = np.array([v for x in lists for v in x], np.uint64)
    mem = np.concatenate([np.arange(16), 16 + rng.permutation(n - 16)]).astype(np.uint32)
    coffs = np.array([0, n], np.uint64)
    for above in (1000, 10):
        got = ctx.cluster_reads_by_snpmers(km, offs, mem, coffs, iterative_above=above)
        want = O.cluster_reads_by_snpmers(km, offs, mem, iterative_above=above)
        assert np.array_equal(got["rep_of"], want)
    assert len(np.unique(want)) >= 17


@pytest.mark.parametrize("name,n,ns", [("operon", 5000, 1), ("pooled", 6000, 20)])
def test_named_configs_at_full_asv_sets(ctx, name, n, ns):
    """configs 4 and 5 of BASELINE.json at their full ASV sets (2000 ASVs / 4.5 kb reads / band 192; 5000 ASVs / 20 samples): the
    whole path (FASTQ entry, resident pass and TwinRead pointer entry) against the oracle -- the shapes bench.py --total-reads streams"""
    from savont_b200 import _lib, twin
    w = _make_workload(name)
    assert w.n_asv == (2000 if name == "operon" else 5000) and w.n_samples == ns
    r = w.reads(n)
    fidx = r["file_idx"] if ns > 1 else None
    gp, op = _params(_lib, band=w.band)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    want = orc.map_reads(r["bases"], r["quals"], r["offsets"], file_idx=fidx, n_samples=ns)
    assert int(want["totals"][0]) > 0.8 * n
    ctx.set_batch_reads(1500)
    try:
        assert_same(ctx.map_reads(r["bases"], r["quals"], r["offsets"], file_idx=fidx, n_samples=ns), want, MAP_KEYS, f"map_reads {name}")
        sk = O.sketch_batch(op, w.snpmers, r["bases"], r["quals"], r["offsets"])
        heap = twin.scatter_to_heap(twin.from_sketch(sk, r["offsets"], file_idx=fidx, kmer48=True))
        assert_same(ctx.map_twin_read_ptrs(heap, n_samples=ns), want, MAP_KEYS, f"map_twin_read_ptrs {name}")
    finally:
        ctx.set_batch_reads(0)
    ctx.upload_reads(r["bases"], r["quals"], r["offsets"], file_idx=fidx, n_samples=ns)
    ctx.map_resident()
    assert_same(ctx.fetch_resident(), want, MAP_KEYS, f"map_resident {name}")
    ctx.set_params(_lib.default_params())


def test_qual_seq(ctx):
    """savont_b200_qual_seq == oracle.qual_seq: bins of 4, QualCompact3 rounding, word-aligned reads, empty reads, capacity error"""
    from savont_b200 import _lib
    rng = np.random.default_rng(9)
    lens = np.concatenate([np.arange(0, 140), rng.integers(1000, 1700, 400)])
    flat = rng.integers(0, 127, size=int(lens.sum()), dtype=np.uint8)
    offs = np.zeros(len(lens) + 1, np.uint64); offs[1:] = np.cumsum(lens)
    want_w, want_o = O.qual_seq(flat, offs)
    got_w, got_o = ctx.qual_seq(flat, offs)
    assert np.array_equal(got_o, want_o) and np.array_equal(got_w, want_w)
    import ctypes as C
    wo = np.zeros(len(lens) + 1, np.uint64); small = np.zeros(4, np.uint64)
    rc = ctx.L.savont_b200_qual_seq(ctx.h, _lib._p(flat, C.c_uint8), _lib._p(offs, C.c_uint64), C.c_size_t(len(lens)), _lib._p(wo, C.c_uint64),
                                    _lib._p(small, C.c_uint64), C.c_size_t(4))
    assert rc == _lib.ERR_CAPACITY


@pytest.mark.parametrize("strand_mode", [0, 1])
def test_mapping_rows_resident(ctx, strand_mode, tmp_path):
    """a12's debug rows (temp/read_to_asv_mappings.tsv, src/alignment.rs:1866, :1875-1887): per read its aligned candidates stable-sorted
    by NM, at most 5, with the SNPmer mismatch count -- composed from the oracle's candidates + pair alignments; near-identical variants
    so that reads have several candidates; also through the 4-part resident pipeline, and the text writer"""
    from savont_b200 import _lib, api
    w = _make_workload("ont16s", n_roots=5, sites_per_genus=3)
    r = w.reads(1500)
    bases = r["bases"].copy(); offs = r["offsets"]
    bases[int(offs[7]):int(offs[8])] = ord("N")
    gp, op = _params(_lib, band=w.band, strand_mode=strand_mode)
    ctx.set_params(gp); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    cand = orc.find_candidates(bases, r["quals"], offs)
    cofs = cand["offsets"].astype(np.int64)
    n = len(offs) - 1
    pr = np.repeat(np.arange(n, dtype=np.uint32), cofs[1:] - cofs[:-1])
    if strand_mode == 0:
        nm = orc.align_pairs(bases, offs, pr, cand["asv"], cand["strand"])["nm"]
    else:
        f = orc.align_pairs(bases, offs, pr, cand["asv"], np.zeros(len(pr), np.uint8))["nm"]
        v = orc.align_pairs(bases, offs, pr, cand["asv"], np.ones(len(pr), np.uint8))["nm"]
        nm = np.minimum(f, v)
    want_off = [0]; want_asv = []; want_nm = []; want_mm = []
    for rd in range(n):
        rows = sorted(range(cofs[rd], cofs[rd + 1]), key=lambda i: int(nm[i]))[:5]        # sorted() is stable; candidates ascend by ASV
        want_asv += [int(cand["asv"][i]) for i in rows]; want_nm += [int(nm[i]) for i in rows]; want_mm += [int(cand["mm"][i]) for i in rows]
        want_off.append(len(want_asv))
    assert max(np.diff(want_off)) >= 2 and min(np.diff(want_off)) == 0
    for parts in (1, 4):
        ctx.set_resident_parts(parts)
        try:
            ctx.upload_reads(bases, r["quals"], offs)
            ctx.map_resident()
            got = ctx.mapping_rows_resident()
        finally:
            ctx.set_resident_parts(1)
        assert np.array_equal(got["row_offsets"], np.array(want_off, np.uint64))
        assert got["asv"].tolist() == want_asv and got["nm"].tolist() == want_nm and got["snp_mismatches"].tolist() == want_mm
    one = ctx.mapping_rows_resident(max_rows=1)
    res = ctx.fetch_resident()
    assigned = res["status"] == 0
    assert np.array_equal((one["row_offsets"][1:] - one["row_offsets"][:-1]) == 1, assigned)
    assert np.array_equal(one["nm"], res["best_nm"][assigned])                    # the first row is a best hit
    ids = [f"read{i} rc" if i % 3 == 0 else f"read{i}" for i in range(n)]
    cons_ids = [1000 + a for a in range(w.n_asv)]
    path = tmp_path / "read_to_asv_mappings.tsv"
    api.write_read_to_asv_mappings(path, ids, cons_ids, got)
    lines = open(path).read().splitlines()
    assert len(lines) == len(want_asv)
    first = int(np.flatnonzero(np.diff(want_off) > 0)[0])
    assert lines[0] == f"{ids[first]}\tdebug_id:{1000 + want_asv[0]}\t{want_mm[0]}\t{want_nm[0]}"
    ctx.set_params(_lib.default_params())


def test_map_reads_from_pageable_memory_uses_the_staging_ring(ctx):
    """batch_upload sends page-locked sources straight to the DMA engine and pageable ones (>= 16 MB) through the ring of pinned chunks
    filled by the host pool (upload_host): same results either way, for every batch of the call"""
    from savont_b200 import _lib
    w = _make_workload("ont16s", n_roots=6)
    n = 50000
    r = w.reads(n)                                                   # numpy arrays: pageable, ~75 MB of bases and of qualities
    assert r["bases"].nbytes > 2 * (16 << 20)
    ctx.set_params(_lib.default_params(band=w.band)); ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
    got = ctx.map_reads(r["bases"], r["quals"], r["offsets"])
    pb = _lib.pinned_array(r["bases"].nbytes); pb[:] = r["bases"]
    pq = _lib.pinned_array(r["quals"].nbytes); pq[:] = r["quals"]
    ref = ctx.map_reads(pb, pq, r["offsets"])
    assert_same(got, ref, MAP_KEYS, "pageable vs page-locked input")
    assert int(got["totals"][0]) > 0.8 * n
    # the TwinRead fields from pageable memory (dna words of the one batch: 19 MB)
    from savont_b200 import twin
    tw = twin.twin_reads_from_snpmers(ctx, r["bases"], r["quals"], r["offsets"], kmer48=True)
    assert tw["dna_words"].nbytes > (16 << 20)
    assert_same(ctx.map_twin_reads(tw), ref, MAP_KEYS, "TwinRead fields from pageable memory")
    _, op = _params(_lib, band=w.band)
    orc = O.Oracle(op, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    m = 1500
    want = orc.map_reads(r["bases"][:int(r["offsets"][m])], r["quals"][:int(r["offsets"][m])], r["offsets"][:m + 1])
    assert np.array_equal(got["status"][:m], want["status"]) and np.array_equal(got["best_nm"][:m], want["best_nm"])
    ctx.set_params(_lib.default_params())
