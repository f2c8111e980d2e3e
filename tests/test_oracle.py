"""Oracle pinning: the reference's own known-answer tests plus an independent pure-Python
restatement.  CPU only."""
import random
import struct

import numpy as np
import pytest

from oracle import oracle as O
from tests import pyref


def test_kmer48_from_u64_layout():
    # src/types.rs:1112-1116  Kmer48::from_u64(1).data == [1,0,0,0,0,0]
    import ctypes as C
    out = (C.c_uint8 * 6)()
    O.lib().so_kmer48_from_u64(C.c_uint64(1), out)
    assert list(out) == [1, 0, 0, 0, 0, 0]
    O.lib().so_kmer48_from_u64(C.c_uint64(0x0000a1b2c3d4e5f6), out)
    assert list(out) == [0xf6, 0xe5, 0xd4, 0xc3, 0xb2, 0xa1]
    assert O.lib().so_kmer48_to_u64(out) == 0x0000a1b2c3d4e5f6


def test_two_bit_encoding_acgtg():
    # src/types.rs:1118-1128  "ACGTG" -> 0b00_01_10_11_10
    assert O.lib().so_kmer_from_ascii(b"ACGTG", 5) == 0b0001101110


def test_kmer_reverse_complement():
    # src/types.rs:1130-1139  revcomp(ACGTG) = CACGT -> 0b01_00_01_10_11
    x = O.lib().so_kmer_from_ascii(b"ACGTG", 5)
    assert O.lib().so_kmer_revcomp(x, 5) == 0b0100011011
    assert O.lib().so_kmer_revcomp(x, 5) == O.lib().so_kmer_from_ascii(b"CACGT", 5)


def test_byte_to_seq_table():
    # src/types.rs:92-101
    L = O.lib()
    for ch, v in [("A", 0), ("a", 0), ("C", 1), ("c", 1), ("G", 2), ("g", 2), ("T", 3), ("t", 3), ("U", 3), ("u", 3),
                  ("N", 0), ("n", 0), ("R", 0), ("-", 0)]:
        assert L.so_encode_base(ord(ch)) == v
    table = [L.so_encode_base(b) for b in range(256)]
    assert table[:4] == [0, 1, 2, 3]
    assert sum(1 for t in table if t) == 8 + 3  # C c G g T t U u + raw bytes 1,2,3


@pytest.mark.parametrize("x,h", [
    (0x0, 0x77cfa1eef01bca90), (0x1, 0x5bca7c69b794f8ce), (0x2, 0xb795033f6f2a0674),
    (0x3fff, 0xfd1fa399fa3deb64), (0x123456789, 0xfd3d49f2cb78951a), (0x3ffffffff, 0xb0cddf11a4f33e9c)])
def test_mm_hash64_kats(x, h):
    # values computed from the formula at src/seeding.rs:18-28
    assert O.mm_hash64(x) == h
    assert pyref.mm_hash64(x) == h


def test_powi_and_ratio_constants():
    # 0.95f64.powi(17) by repeated squaring, bits 0x3fdac27bcb63c568
    v = O.lib().so_powi(0.95, 17)
    assert struct.unpack("<Q", struct.pack("<d", v))[0] == 0x3fdac27bcb63c568
    assert v == 0.41812033521917735
    assert (11.0 / 200.0) / 11.0 == 0.005
    assert O.lib().so_split_mask(17) == (~(3 << 16)) & ((1 << 64) - 1)


def _rand_seq(rng, n, alphabet=b"ACGT"):
    return bytes(rng.choice(alphabet) for _ in range(n))


def _snpmer_set_from(seqs, k, every=7):
    """take some canonical k-mers (both alleles of their middle base) as a SNPmer set"""
    s = set()
    mask = ((1 << 64) - 1) ^ (3 << (k - 1))
    for sq in seqs:
        for p in range(0, len(sq) - k + 1, every):
            f = 0
            for b in sq[p:p + k]:
                f = (f << 2) | pyref.enc(b)
            r = pyref.revcomp_kmer(f, k)
            canon = f if (f & mask) < (r & mask) else r
            s.add(canon)
            s.add((canon & mask) | (((canon >> (k - 1)) + 1) & 3) << (k - 1))
    return s


@pytest.mark.parametrize("k,c", [(17, 11), (15, 9), (21, 11), (13, 7)])
def test_sketch_matches_python_restatement(k, c):
    rng = random.Random(1234 + k)
    seqs, quals = [], []
    for t in range(24):
        L = rng.choice([k - 1, k, k + 1, 30, 64, 200, 333, 800])
        alpha = b"ACGT" if t % 3 else b"ACGTNacgtRY"
        sq = _rand_seq(rng, L, alpha)
        if t % 5 == 0 and L > 60:   # low-complexity stretch -> duplicate k-mers / splitmers
            sq = sq[:20] + b"ACACACACACACACACACACACACACACAC" + sq[50:]
        if t % 7 == 0 and L > 100:  # exact repeat -> duplicated SNPmers
            sq = sq[:50] + sq[:50] + sq[100:]
        seqs.append(sq)
        if t % 4 == 3:
            quals.append(bytes([40 + 33] * L))            # flat qualities
        else:
            quals.append(bytes(33 + rng.choice([5, 12, 25, 26, 30, 40, 50]) for _ in range(L)))
    snp = _snpmer_set_from(seqs, k)
    p = O.default_params(k=k, c=c)
    bases, offs = O.concat(seqs)
    qcat, _ = O.concat(quals)
    for use_q in (True, False):
        got = O.sketch_batch(p, sorted(snp), bases, qcat if use_q else None, offs)
        for i, sq in enumerate(seqs):
            ref = pyref.sketch(sq, quals[i] if use_q else None, k, c, 25, snp)
            if ref is None:
                assert got["status"][i] == O.STATUS_TOO_SHORT
                continue
            assert got["status"][i] == 0
            a, b = int(got["mini_offsets"][i]), int(got["mini_offsets"][i + 1])
            assert [(int(x), int(y), int(z)) for x, y, z in
                    zip(got["mini_pos"][a:b], got["mini_kmer"][a:b], got["mini_fwd"][a:b])] == ref["mini"]
            a, b = int(got["snp_offsets"][i]), int(got["snp_offsets"][i + 1])
            assert [(int(x), int(y)) for x, y in zip(got["snp_pos"][a:b], got["snp_kmer"][a:b])] == ref["snp"]
            o0, o1 = int(offs[i]), int(offs[i + 1])
            assert list(got["codes"][o0:o1]) == ref["codes"]
            if ref["est_id"] is None:
                assert np.isnan(got["est_id"][i])
            else:
                assert got["est_id"][i] == ref["est_id"]


def test_sketch_warmup_quirk_positions():
    # k-mer start positions 0..9 are never syncmer-tested (k=17,c=11)
    rng = random.Random(7)
    seqs = [_rand_seq(rng, 400) for _ in range(20)]
    bases, offs = O.concat(seqs)
    got = O.sketch_batch(O.default_params(), [], bases, None, offs)
    assert got["mini_pos"].min() >= 10
    dens = len(got["mini_pos"]) / sum(len(s) for s in seqs)
    assert 0.06 < dens < 0.12     # open-syncmer density ~ 1/11


def _mutate(rng, s, nsub, nins, ndel):
    s = bytearray(s)
    for _ in range(nsub):
        i = rng.randrange(len(s)); s[i] = rng.choice(b"ACGT")
    for _ in range(nins):
        i = rng.randrange(len(s)); s.insert(i, rng.choice(b"ACGT"))
    for _ in range(ndel):
        i = rng.randrange(len(s)); del s[i]
    return bytes(s)


@pytest.mark.parametrize("band", [0, 1, 3, 8, 33])
def test_banded_dp_matches_full_matrix(band):
    rng = random.Random(99 + band)
    asvs = [_rand_seq(rng, rng.choice([20, 37, 64, 90])) for _ in range(6)]
    reads, pr, pa, ps = [], [], [], []
    for a, s in enumerate(asvs):
        for t in range(6):
            rd = _mutate(rng, s, rng.randrange(4), rng.randrange(3), rng.randrange(3))
            if t == 4:
                rd = rd[rng.randrange(1, 6):]
            if t == 5:
                rd = rd[:-rng.randrange(1, 6)] + _rand_seq(rng, 3)
            reads.append(rd); pr.append(len(reads) - 1); pa.append(a); ps.append(t % 2)
    p = O.default_params(band=band)
    orc = O.Oracle(p, [])
    ab, ao = O.concat(asvs)
    orc.load_asvs(ab, ao)
    rb, ro = O.concat(reads)
    got = orc.align_pairs(rb, ro, pr, pa, ps, want_cigar=True)
    comp = {0: 3, 1: 2, 2: 1, 3: 0}
    for i in range(len(pr)):
        a = [pyref.enc(b) for b in asvs[pa[i]]]
        r = [pyref.enc(b) for b in reads[pr[i]]]
        if ps[i]:
            r = [comp[x] for x in reversed(r)]
        nm, ie, je = pyref.banded_nm(a, r, band)
        assert (int(got["nm"][i]), int(got["asv_end"][i]), int(got["read_end"][i])) == (nm, ie, je)
        # CIGAR consistency: spans and edit count
        c0, c1 = int(got["cigar_offsets"][i]), int(got["cigar_offsets"][i + 1])
        ai, rj, edits = int(got["asv_start"][i]), int(got["read_start"][i]), 0
        assert ai == 0 or rj == 0
        for e in got["cigar"][c0:c1]:
            ln, op = int(e) >> 4, int(e) & 15
            if op == 0:
                edits += sum(1 for t in range(ln) if a[ai + t] != r[rj + t]); ai += ln; rj += ln
            elif op == 1:
                edits += ln; rj += ln
            else:
                edits += ln; ai += ln
        assert (ai, rj) == (ie, je)
        assert edits == nm


def test_em_simple():
    # two ASVs; 10 reads unambiguous to 0, 10 ambiguous {0,1}: EM pushes everything to ASV 0
    ab, depths, it = O.em(2, [0, 1, 3], [0, 0, 1], [10, 10], 20)
    assert depths.tolist() == [20, 0]
    assert abs(ab.sum() - 1.0) < 1e-9 and it > 1
    ab, depths, it = O.em(3, [0, 1, 2, 3], [0, 1, 2], [5, 3, 2], 10)
    assert depths.tolist() == [5, 3, 2]


def test_low_polymorphism_mode_needs_no_snpmers():
    """candidate_mode = 1: with an empty SNPmer set the default mode assigns nothing, the low-polymorphism mode
    (src/alignment.rs:1527-1719 analogue) assigns every read to its genus by minimizers + NM"""
    from savont_b200 import synth
    w = synth.Workload("ont16s", n_roots=3, sites_per_genus=5)
    r = w.reads(120)
    for mode, expect_all in ((0, False), (1, True)):
        orc = O.Oracle(O.default_params(band=w.band, candidate_mode=mode), [])
        orc.load_asvs(w.asv_bases, w.asv_offsets)
        res = orc.map_reads(r["bases"], r["quals"], r["offsets"])
        if expect_all:
            assert int(res["totals"][0]) == 120
            co = res["class_offsets"].astype(int)
            assert all(r["truth_asv"][i] // 10 == res["class_asv"][co[i]] // 10 for i in range(120))   # right genus
        else:
            assert int(res["totals"][0]) == 0 and np.all(res["status"] == O.STATUS_NO_CANDIDATE)


def _a4_inputs(seed=11, k=17):
    """reads that trip every a4 test: too short / too long, below k, repetitive (tandem repeats and high-frequency
    k-mers), low quality, flat quality (est_id None), duplicate ids"""
    import random
    from tests.util import rand_seq
    rng = random.Random(seed)
    reads, quals, ids = [], [], []
    for t in range(60):
        L = rng.choice([10, 16, 300, 420, 500, 640, 900])
        sq = rand_seq(rng, L)
        if t % 6 == 1 and L >= 300:
            unit = rand_seq(rng, rng.choice([3, 7, 23]))
            sq = (unit * (L // len(unit) + 1))[:L]                   # tandem repeat: few distinct minimizers
        q = [33 + rng.choice([12, 20, 30, 40]) for _ in range(L)]
        if t % 5 == 2:
            q = [33 + rng.choice([8, 10, 14]) for _ in range(L)]     # est_id well below any sensible cutoff
        if t % 7 == 3:
            q = [33 + 30] * L                                        # flat: est_id None, never dropped by quality
        reads.append(sq); quals.append(bytes(q)); ids.append(f"r{rng.randrange(40):03d}" if t % 9 else "r007")
    return reads, quals, ids


@pytest.mark.parametrize("k,c", [(17, 11), (15, 9)])
def test_twin_reads_filter_matches_python_restatement(k, c):
    from tests.util import snpmer_set_from
    reads, quals, ids = _a4_inputs(11 + k, k)
    snp = sorted(snpmer_set_from(reads, k))
    p = O.default_params(k=k, c=c)
    # high-frequency k-mers: every minimizer of three of the reads -> those reads lose all their solid minimizers
    hf = set()
    for i in (4, 10, 16):
        sk = pyref.sketch(reads[i], quals[i], k, c, 25, set(snp))
        if sk:
            hf.update(m[1] for m in sk["mini"])
    bases, offs = O.concat(reads); qcat, _ = O.concat(quals)
    for (mn, mx, cut, maxc) in [(0, 0, 0.0, 500), (300, 640, 97.0, 500), (17, 0, 99.5, 3), (400, 500, 90.0, 1)]:
        rf = O.read_filter(mn, mx, cut, maxc)
        got = O.twin_reads_filter(p, rf, snp, sorted(hf), bases, qcat, offs, ids)
        want_status, want_order = pyref.twin_reads_filter(reads, quals, ids, k, c, 25, set(snp), hf, mn, mx, cut, maxc)
        assert got["status"].tolist() == want_status, (mn, mx, cut, maxc)
        assert got["order"].tolist() == want_order, (mn, mx, cut, maxc)
    # every status occurs in at least one of the settings above
    rf = O.read_filter(300, 640, 97.0, 500)
    st = O.twin_reads_filter(p, rf, snp, sorted(hf), bases, qcat, offs, ids)["status"]
    assert set(st.tolist()) >= {0, 5, 6, 7}


def test_map_reads_with_read_filter_skips_filtered_reads():
    """reads failing a4 never reach Stage 7: not assigned, not counted as filtered"""
    from savont_b200 import synth
    w = synth.Workload("ont16s", n_roots=3)
    r = w.reads(200)
    quals = r["quals"].copy()
    quals[int(r["offsets"][3]):int(r["offsets"][4])] = 33 + 9          # low quality read ...
    quals[int(r["offsets"][3])] = 33 + 11                              # ... but not flat (flat qualities -> est_id None, src/seeding.rs:570-573)
    p = O.default_params(band=w.band)
    orc = O.Oracle(p, w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    base = orc.map_reads(r["bases"], quals, r["offsets"])
    rf = O.read_filter(1490, 1505, 98.0)
    orc.set_read_filter(rf, [])
    got = orc.map_reads(r["bases"], quals, r["offsets"])
    flt = O.twin_reads_filter(p, rf, w.snpmers, [], r["bases"], quals, r["offsets"])
    assert got["status"][3] == O.STATUS_LOW_QUALITY
    drop = flt["status"] != 0
    assert drop.any() and (~drop).any()
    assert np.array_equal(got["status"][drop], flt["status"][drop])
    assert np.array_equal(got["status"][~drop], base["status"][~drop])
    assert int(got["totals"][0]) == int((got["status"] == 0).sum())
    assert int(got["totals"][1]) == int(((got["status"] > 0) & (got["status"] < 4)).sum())
    orc.set_read_filter(None)
    assert np.array_equal(orc.map_reads(r["bases"], quals, r["offsets"])["status"], base["status"])


def _kmer_count_inputs(seed=5):
    """a small read set with shared k-mers on both strands, rc-tagged reads, N / lower case / U bases, palindromic splitmers,
    low and flat qualities, reads shorter than k"""
    import random
    from tests.util import mutate, rand_seq
    rng = random.Random(seed)
    roots = [rand_seq(rng, 120) for _ in range(3)]
    roots[0] = roots[0][:40] + b"ACGTACGTACGTACGTACGTACGT" + roots[0][64:]          # palindromic stretches
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    reads, quals, rc = [], [], []
    for t in range(70):
        s = mutate(rng, roots[t % 3], rng.randrange(3), rng.randrange(2), rng.randrange(2))
        if t % 2:
            s = s.translate(comp)[::-1]
        if t % 11 == 0:
            s = s[:30] + b"N" + s[31:60] + b"u" + s[61:90].lower() + s[90:]
        if t % 13 == 5:
            s = s[:rng.randrange(0, 17)]
        q = bytes(33 + rng.choice([10, 24, 25, 26, 35]) for _ in range(len(s)))
        if t % 9 == 4:
            q = bytes([33 + 12] * len(s))                                             # flat: no quality skip at all
        reads.append(s); quals.append(q); rc.append(1 if t % 5 == 0 else 0)
    return reads, quals, rc


@pytest.mark.parametrize("k,single", [(17, False), (17, True), (15, False), (21, False)])
def test_count_split_kmers_matches_python_restatement(k, single):
    reads, quals, rc = _kmer_count_inputs(5 + k)
    bases, offs = O.concat(reads); qcat, _ = O.concat(quals)
    for q in (qcat, None):
        got = O.count_split_kmers(bases, q, offs, rc=rc, k=k, min_bq=25, single_strand=single)
        want, nd = pyref.split_kmer_counts(reads, quals if q is not None else [None] * len(reads), rc, k, 25, single)
        assert got["n_distinct"] == nd
        assert [(int(a), int(b[0]), int(b[1])) for a, b in zip(got["kmers"], got["counts"])] == want
        assert len(want) > 20
    # rc = None is all zeros; multi-threaded extraction gives the same answer
    a = O.count_split_kmers(bases, qcat, offs, rc=None, k=k)
    b = O.count_split_kmers(bases, qcat, offs, rc=np.zeros(len(reads), np.uint8), k=k, n_threads=3)
    assert np.array_equal(a["kmers"], b["kmers"]) and np.array_equal(a["counts"], b["counts"])


def test_high_freq_threshold_rule():
    # src/kmer_comp.rs:474: sorted[len - len/100000 - 1].max(100); members are strictly above the threshold
    km = np.arange(10, dtype=np.uint64) + 7
    ct = np.array([[1, 2], [50, 60], [100, 1], [0, 100], [99, 1], [300, 1], [2, 2], [120, 0], [3, 3], [1, 1]], np.uint32)
    th, hf = O.high_freq_kmers(km, ct)
    assert th == 301 and len(hf) == 0                                                 # n < 100000: the maximum itself
    n = 200001
    km = np.arange(n, dtype=np.uint64); ct = np.zeros((n, 2), np.uint32); ct[:, 0] = np.arange(n) % 1000; ct[:, 1] = 1
    th, hf = O.high_freq_kmers(km, ct)
    s = np.sort(ct.sum(1)); assert th == max(100, int(s[n - n // 100000 - 1]))
    assert np.array_equal(hf, km[ct.sum(1) > th])


@pytest.mark.parametrize("use_hpc,band", [(True, 16), (False, 40), (True, 3)])
def test_consensus_pileups_match_python_restatement(use_hpc, band):
    """so_consensus_pileups vs tests/pyref.py (compression, strand choice over the oracle's DP, CIGAR walk, median)"""
    from tests.util import pileup_case
    cons, reads, quals, rcons = pileup_case(4100 + band, long_runs=use_hpc)
    cb, co = O.concat(cons); rb, ro = O.concat(reads); qb, _ = O.concat(quals)
    got = O.consensus_pileups(cb, co, rb, qb, ro, rcons, band=band, use_hpc=use_hpc)
    # independent walk; the alignments themselves come from so_align_pairs on the compressed reads
    comp = [pyref.hpc_with_quality(r, q, use_hpc) for r, q in zip(reads, quals)]
    hreads = [bytes(b"ACGT"[x] for x in hc) for hc, _, _ in comp]
    orc = O.Oracle(O.default_params(band=band), [])
    ab, ao = O.concat([c_ or b"A" * 30 for c_ in cons])     # the ASV loader rejects empty sequences; never paired below
    orc.load_asvs(ab, ao)
    keep = [r for r in range(len(reads)) if len(hreads[r]) and len(cons[rcons[r]])]
    pr = np.repeat(np.array(keep, np.uint32), 2); pa = np.array([rcons[r] for r in pr], np.uint32); ps = np.tile(np.array([0, 1], np.uint8), len(keep))
    hb, ho = O.concat(hreads)
    al = orc.align_pairs(hb, ho, pr, pa, ps, want_cigar=True)
    aln = [None] * len(reads)
    for t, r in enumerate(keep):
        sd = 1 if al["nm"][2 * t + 1] < al["nm"][2 * t] else 0
        i = 2 * t + sd
        cig = [(int(e) >> 4, int(e) & 15) for e in al["cigar"][int(al["cigar_offsets"][i]):int(al["cigar_offsets"][i + 1])]]
        aln[r] = (sd, int(al["asv_start"][i]), int(al["read_start"][i]), int(al["read_end"][i]), cig)
        assert got["read_strand"][r] == sd and got["read_nm"][r] == al["nm"][i]
    for r in range(len(reads)):
        if aln[r] is None:
            assert got["read_strand"][r] == 0xff and got["read_nm"][r] == -1
    cols, hp = pyref.pileup_columns([len(c_) for c_ in cons], [(rcons[r],) + comp[r] for r in range(len(reads))], aln)
    flat = [e for c_ in cols for col in c_ for e in col]
    assert len(flat) == len(got["entries"]) and len(flat) > 100
    pos = 0; k = 0
    for c_ in cols:
        for col in c_:
            assert int(got["pos_offsets"][pos + 1]) - int(got["pos_offsets"][pos]) == len(col)
            for (r, kind, bs, qs, hs) in col:
                e = got["entries"][k]; k += 1
                assert (int(e["read"]), int(e["kind"]), int(e["n"])) == (r, kind, len(bs))
                assert list(e["base"][:len(bs)]) == bs and list(e["qual"][:len(bs)]) == list(qs) and list(e["hp"][:len(bs)]) == list(hs)
            pos += 1
    assert list(got["ref_hp_length"]) == [x for h in hp for x in h]
    if use_hpc:
        assert max(max(h) for h in hp if h) > 1


def test_fx_hash_known_answers():
    """FxHasher64 (fxhash 0.2.1) from its published definition: hash = (rotl(hash, 5) ^ word) * 0x517cc1b727220a95"""
    K = 0x517cc1b727220a95
    assert pyref.fx_hash2(0, 0) == 0
    assert pyref.fx_hash2(0, 1) == K
    assert pyref.fx_hash2(1, 0) == (((K << 5) | (K >> 59)) & (2 ** 64 - 1)) * K % 2 ** 64
    sig = pyref.lsh_signatures([5, 9, 5, 1 << 40, 77])
    assert len(sig) == 20 and all(s is not None for s in sig) and pyref.lsh_signatures([1, 2]) == [None] * 20


@pytest.mark.parametrize("seed,err", [(1, 0.02), (2, 0.005), (3, 0.04)])
def test_lsh_and_kmer_clustering_match_python_restatement(seed, err):
    from tests.util import clustering_case
    km, off, _ = clustering_case(600 + seed, n_reads=260, err=err)
    minis = [[int(x) for x in km[int(off[i]):int(off[i + 1])]] for i in range(len(off) - 1)]
    sig, valid = O.lsh_signatures(km, off)
    for i, m in enumerate(minis):
        want = pyref.lsh_signatures(m)
        assert bool(valid[i]) == (want[0] is not None)
        if valid[i]:
            assert [int(x) for x in sig[i]] == want
    rep = O.cluster_reads_by_kmers(km, off)
    want = pyref.cluster_reads_by_kmers(minis)
    assert [int(x) for x in rep] == want
    n_rep = sum(1 for i, r in enumerate(want) if r == i)
    assert 6 <= n_rep < len(want) and any(len(m) != len(set(m)) for m in minis) and any(len(m) < 3 for m in minis)


@pytest.mark.parametrize("seed", [1, 2])
def test_snpmer_compare_matches_python_restatement(seed):
    from tests.util import csr_u64, reassign_case
    cons, reads = reassign_case(70 + seed)
    ck, co = csr_u64(cons); rk, ro = csr_u64(reads)
    got = O.snpmer_compare(ck, co, rk, ro, want_matrix=True)
    want, mat = pyref.reassign_reads(cons, reads)
    assert [(int(a), int(b), int(c_)) for a, b, c_ in zip(got["best_cons"], got["best_mismatches"], got["best_matches"])] == want
    assert got["matrix"].tolist() == [[list(x) for x in row] for row in mat]
    assert len(set(int(x) for x in got["best_cons"])) > 5 and int(got["best_mismatches"].max()) == 0 and int(got["matrix"][:, :, 1].max()) > 0
    # consensus lists against each other (compare_consensus_snpmers): the diagonal has no mismatches
    cc = O.snpmer_compare(ck, co, ck, co, want_matrix=True)["matrix"]
    assert all(int(cc[i, i, 1]) == 0 or i == 5 for i in range(len(cons)))


@pytest.mark.parametrize("seed,iter_above", [(1, 1000), (2, 1000), (3, 4)])
def test_snpmer_greedy_clustering_matches_python_restatement(seed, iter_above):
    """Stage 3 greedy loop (oracle only so far): C vs Python, indexed branch and the branch above `iterative_above` representatives"""
    from tests.util import csr_u64, reassign_case
    _, reads = reassign_case(300 + seed, n_cons=10, n_sites=40, n_reads=220)
    rng = random.Random(seed)
    cluster = [r for r in range(len(reads)) if r % 9 != 4]
    rng.shuffle(cluster)                                             # any order inside the cluster; reads outside stay untouched
    sk, so = csr_u64(reads)
    got = O.cluster_reads_by_snpmers(sk, so, cluster, iterative_above=iter_above)
    want = pyref.cluster_reads_by_snpmers(reads, cluster, iterative_above=iter_above)
    assert {r: int(got[r]) for r in cluster} == want
    assert all(int(got[r]) == 0xffffffff for r in range(len(reads)) if r not in want)
    n_rep = sum(1 for r, p in want.items() if r == p)
    assert 8 <= n_rep < len(cluster)


def test_fuzz_stage2_stage3_restatements_agree():
    """random small inputs with many ties (few distinct k-mers, equal bucket hits, equal containment ratios, duplicated splitmers): the C
    oracle and the Python restatement of Stage 2 clustering, the SNPmer comparison and the Stage 3 greedy loop must agree"""
    from hypothesis import given, settings, strategies as st
    from tests.util import csr_u64
    pool = [((i * 0x9E3779B97F4A7C15) >> 20) & ((1 << 34) - 1) for i in range(1, 25)]
    lists = st.lists(st.lists(st.sampled_from(pool), min_size=0, max_size=10), min_size=1, max_size=30)

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(lists, st.sampled_from([0.3, 0.8, 0.95]))
    def stage2(minis, thr):
        km, off = csr_u64(minis)
        assert [int(x) for x in O.cluster_reads_by_kmers(km, off, threshold=thr)] == pyref.cluster_reads_by_kmers(minis, threshold=thr)

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(lists, lists)
    def compare(cons, reads):
        ck, co = csr_u64(cons); rk, ro = csr_u64(reads)
        got = O.snpmer_compare(ck, co, rk, ro, want_matrix=True)
        want, mat = pyref.reassign_reads(cons, reads)
        assert [(int(a), int(b), int(c_)) for a, b, c_ in zip(got["best_cons"], got["best_mismatches"], got["best_matches"])] == want
        assert got["matrix"].tolist() == [[list(x) for x in row] for row in mat]

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(lists, st.sampled_from([1000, 2]))
    def stage3(reads, it):
        sk, so = csr_u64(reads)
        cluster = list(range(len(reads)))
        got = O.cluster_reads_by_snpmers(sk, so, cluster, iterative_above=it)
        assert {r: int(got[r]) for r in cluster} == pyref.cluster_reads_by_snpmers(reads, cluster, iterative_above=it)

    stage2(); compare(); stage3()


def test_fuzz_pileup_walk_restatements_agree():
    """tiny consensus / read pairs, related or not (alignments that start with an insertion, end in a deletion, clip either end):
    strand choice / NM of so_consensus_pileups against the full-matrix DP of tests/pyref.py, provenance of every entry, the median"""
    from hypothesis import given, settings, strategies as st
    seqs = st.text(alphabet="ACGT", min_size=1, max_size=24).map(lambda s_: s_.encode())

    @settings(max_examples=80, deadline=None, derandomize=True)
    @given(st.lists(seqs, min_size=1, max_size=3), st.lists(st.tuples(seqs, st.integers(0, 2)), min_size=1, max_size=8), st.sampled_from([2, 5, 30]),
           st.booleans())
    def run(cons, reads_c, band, use_hpc):
        reads = [r for r, _ in reads_c]
        rcons = sorted(c_ % len(cons) for _, c_ in reads_c)
        quals = [bytes(33 + (7 * i + j) % 40 for j in range(len(r))) for i, r in enumerate(reads)]
        cb, co = O.concat(cons); rb, ro = O.concat(reads); qb, _ = O.concat(quals)
        got = O.consensus_pileups(cb, co, rb, qb, ro, rcons, band=band, use_hpc=use_hpc)
        comp = [pyref.hpc_with_quality(r, q, use_hpc) for r, q in zip(reads, quals)]
        hreads = [bytes(b"ACGT"[x] for x in hc) for hc, _, _ in comp]
        for r in range(len(reads)):                                   # strand choice and NM against the full-matrix DP of pyref
            a = [pyref.enc(b) for b in cons[rcons[r]]]
            best = None
            for sd in (0, 1):
                rd = [pyref.enc(b) for b in hreads[r]]
                if sd:
                    rd = [3 - x for x in reversed(rd)]
                nm, ie, je = pyref.banded_nm(a, rd, band)
                if best is None or nm < best[0]:
                    best = (nm, sd)
            assert int(got["read_nm"][r]) == best[0] and int(got["read_strand"][r]) == best[1]
        # entries: every Base entry's column base / quality / run length must come from the compressed read in its orientation
        po, ent = got["pos_offsets"], got["entries"]
        assert int(po[-1]) == len(ent) and np.all(np.diff(po.astype(np.int64)) >= 0)
        for e in ent:
            r = int(e["read"]); hc, hq, hl = comp[r]
            sd = int(got["read_strand"][r])
            alphabet = [b"ACGT"[3 - x] if sd else b"ACGT"[x] for x in hc]
            for j in range(int(e["n"])):
                assert int(e["base"][j]) in alphabet and int(e["qual"][j]) in hq and int(e["hp"][j]) in hl
        hp = got["ref_hp_length"]
        for p in range(len(hp)):
            v = sorted(int(e["hp"][0]) for e in ent[int(po[p]):int(po[p + 1])] if int(e["kind"]) == 0)
            want = 1 if not v else ((v[len(v) // 2 - 1] + v[len(v) // 2]) // 2 if len(v) % 2 == 0 else v[len(v) // 2])
            assert int(hp[p]) == want
    run()


def test_qual_seq_matches_elementwise_restatement():
    """oracle.qual_seq (numpy) == the loop of get_twin_read_syncmer + QualCompact3::try_from_bits, incl. partial last bins, one-base
    reads, bytes below '!' and above 'N' + 33"""
    from tests import pyref
    rng = np.random.default_rng(5)
    quals = [bytes(rng.integers(0, 127, size=int(L), dtype=np.uint8)) for L in list(range(1, 70)) + [128, 129, 1500, 1503]]
    quals.append(bytes([33] * 64)); quals.append(bytes([126] * 65)); quals.append(bytes([34, 35, 37, 38, 76, 77, 78, 32]))
    flat = np.frombuffer(b"".join(quals), np.uint8)
    offs = np.zeros(len(quals) + 1, np.uint64); offs[1:] = np.cumsum([len(x) for x in quals])
    words, wo = O.qual_seq(flat, offs)
    for r, qs in enumerate(quals):
        want = pyref.qual_seq(qs)
        w = words[int(wo[r]):int(wo[r + 1])]
        assert len(w) == (len(want) + 15) // 16
        got = [int((int(w[i // 16]) >> (4 * (i % 16))) & 15) for i in range(len(want))]
        assert got == want
        tail = int(w[-1]) >> (4 * ((len(want) - 1) % 16 + 1)) if len(want) % 16 else 0
        assert tail == 0
