"""Config 1 of BASELINE.json (real ONT reads vs the zymo reference 16S copies): committed golden vectors
(tests/golden/zymo_subset.npz, made by tests/golden/make_golden.py from the reference's own test data).
CPU: the oracle still reproduces them.  GPU: the library reproduces them through the C ABI."""
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as O

G = os.path.join(os.path.dirname(__file__), "golden", "zymo_subset.npz")


def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


@pytest.fixture(scope="module")
def g():
    return np.load(G)


def test_fixture_shape(g):
    assert len(g["asv_offsets"]) - 1 == 63 and len(g["read_offsets"]) - 1 == 260
    lens = (g["read_offsets"][1:] - g["read_offsets"][:-1]).astype(int)
    assert lens.min() >= 1100 and lens.max() <= 2000                  # src/cli.rs:87-92
    assert int(g["totals"][0]) == int((g["status"] == 0).sum()) == 249


def test_oracle_reproduces_golden(g):
    p = O.default_params(band=64)
    sk = O.sketch_batch(p, g["snpmers"], g["read_bases"], g["read_quals"], g["read_offsets"])
    assert _digest(sk["mini_pos"], sk["mini_kmer"], sk["mini_fwd"], sk["snp_pos"], sk["snp_kmer"]) == str(g["sk_digest"])
    assert np.array_equal(sk["mini_offsets"], g["sk_mini_offsets"]) and np.array_equal(sk["est_id"], g["sk_est_id"])
    orc = O.Oracle(p, g["snpmers"]); orc.load_asvs(g["asv_bases"], g["asv_offsets"])
    res = orc.map_reads(g["read_bases"], g["read_quals"], g["read_offsets"])
    for key in ("status", "best_nm", "class_offsets", "class_asv", "unambig", "ambig", "leq10", "totals"):
        assert np.array_equal(res[key], g[key]), key


@pytest.mark.gpu
def test_gpu_reproduces_golden(g):
    from savont_b200 import _lib
    ctx = _lib.Context(0, _lib.default_params(band=64))
    ctx.load_snpmers(g["snpmers"]); ctx.load_asvs(g["asv_bases"], g["asv_offsets"])
    sk = ctx.sketch(g["read_bases"], g["read_quals"], g["read_offsets"])
    assert _digest(sk["mini_pos"], sk["mini_kmer"], sk["mini_fwd"], sk["snp_pos"], sk["snp_kmer"]) == str(g["sk_digest"])
    assert np.array_equal(sk["snp_offsets"], g["sk_snp_offsets"]) and np.array_equal(sk["est_id"], g["sk_est_id"])
    ask = ctx.sketch(g["asv_bases"], None, g["asv_offsets"])
    assert _digest(ask["mini_pos"], ask["mini_kmer"], ask["mini_fwd"], ask["snp_pos"], ask["snp_kmer"]) == str(g["asv_sk_digest"])
    cand = ctx.find_candidates(g["read_bases"], g["read_quals"], g["read_offsets"])
    for key in ("status", "offsets", "asv", "mm", "mini", "strand"):
        assert np.array_equal(cand[key], g["cand_" + key]), key
    res = ctx.map_reads(g["read_bases"], g["read_quals"], g["read_offsets"])
    for key in ("status", "best_nm", "class_offsets", "class_asv", "unambig", "ambig", "leq10", "totals"):
        assert np.array_equal(res[key], g[key]), key
    # reverse-complementing every read must give the same assignments with the opposite strand
    comp = np.zeros(256, np.uint8); comp[:] = ord("N")
    for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
        comp[a] = b
    o = g["read_offsets"].astype(np.int64)
    rb = np.concatenate([comp[g["read_bases"][o[i]:o[i + 1]][::-1]] for i in range(len(o) - 1)])
    rq = np.concatenate([g["read_quals"][o[i]:o[i + 1]][::-1] for i in range(len(o) - 1)])
    rc = ctx.map_reads(rb, rq, g["read_offsets"])
    crc = ctx.find_candidates(rb, rq, g["read_offsets"])
    assert np.array_equal(crc["asv"], cand["asv"]) and np.array_equal(crc["strand"], 1 - cand["strand"])
    for key in ("status", "best_nm", "class_offsets", "class_asv", "totals"):
        assert np.array_equal(rc[key], g[key]), key
    ctx.close()


# ---- the two stages in front of Stage 7 on the same real reads (tests/golden/zymo_stage1.npz, make_golden_stage1.py) ----
G1 = os.path.join(os.path.dirname(__file__), "golden", "zymo_stage1.npz")


@pytest.fixture(scope="module")
def g1():
    return np.load(G1)


def _ids(g1):
    o = g1["id_offsets"].astype(np.int64)
    return [bytes(g1["id_bytes"][o[i]:o[i + 1]]).decode() for i in range(len(o) - 1)]


def test_stage1_fixture_shape(g1):
    ids = _ids(g1)
    assert len(ids) == 400 and int(g1["rc"].sum()) == 205
    assert all((i.split()[-1] == "rc") == bool(r) for i, r in zip(ids, g1["rc"]))     # src/seq_parse.rs:361-369
    assert set(g1["flt_status"].tolist()) == {0, 5, 7}


def test_oracle_reproduces_stage1_golden(g1):
    kc = O.count_split_kmers(g1["read_bases"], g1["read_quals"], g1["read_offsets"], rc=g1["rc"])
    assert len(kc["kmers"]) == int(g1["kc_n"]) and kc["n_distinct"] == int(g1["kc_n_distinct"])
    assert _digest(kc["kmers"], kc["counts"]) == str(g1["kc_digest"])
    th, hf = O.high_freq_kmers(kc["kmers"], kc["counts"])
    assert th == int(g1["hf_thresh"]) and len(hf) == int(g1["hf_n"])
    flt = O.twin_reads_filter(O.default_params(band=64), O.read_filter(1100, 2000, 98.0), g1["snpmers"], g1["hf_set"],
                              g1["read_bases"], g1["read_quals"], g1["read_offsets"], _ids(g1))
    for key in ("status", "order", "est_id", "n_solid_mini", "n_solid_snp"):
        assert np.array_equal(flt[key], g1["flt_" + {"n_solid_mini": "solid_mini", "n_solid_snp": "solid_snp"}.get(key, key)]), key


@pytest.mark.gpu
def test_gpu_reproduces_stage1_golden(g1):
    from savont_b200 import _lib, api
    ctx = _lib.Context(0, _lib.default_params(band=64))
    ctx.load_snpmers(g1["snpmers"])
    reads = dict(bases=g1["read_bases"], quals=g1["read_quals"], offsets=g1["read_offsets"], ids=_ids(g1))
    kc = api.read_to_split_kmers(ctx, reads)                                         # rc flags from the ids
    assert len(kc["kmers"]) == int(g1["kc_n"]) and kc["n_distinct"] == int(g1["kc_n_distinct"])
    assert _digest(kc["kmers"], kc["counts"]) == str(g1["kc_digest"])
    assert np.array_equal(kc["kmers"][:16], g1["kc_head"]) and np.array_equal(kc["counts"][:16], g1["kc_head_counts"])
    th, hf = api.high_freq_kmers(kc["kmers"], kc["counts"])
    assert th == int(g1["hf_thresh"]) and len(hf) == int(g1["hf_n"])
    sel = api.twin_reads_from_snpmers(ctx, reads, api.Cli(band=64), high_freq_kmers=g1["hf_set"])    # Cli defaults 1100 / 2000 / 98.0
    assert np.array_equal(sel["status"], g1["flt_status"]) and np.array_equal(sel["order"], g1["flt_order"])
    assert np.array_equal(sel["est_id"].view(np.uint64), g1["flt_est_id"].view(np.uint64))
    assert np.array_equal(sel["n_solid_mini"], g1["flt_solid_mini"]) and np.array_equal(sel["n_solid_snp"], g1["flt_solid_snp"])
    ctx.close()


# ---- the neighbouring stages on the same reads (tests/golden/zymo_stages.npz, made by tests/golden/make_golden_stages.py) ----
GS = os.path.join(os.path.dirname(__file__), "golden", "zymo_stages.npz")
STAGE_KEYS = ("lsh_digest", "rep_of", "best_cons", "best_mismatches", "best_matches", "pile_digest", "pile_read_nm", "pile_n_entries")


def _stages_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_stages", os.path.join(os.path.dirname(__file__), "golden", "make_golden_stages.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def test_oracle_reproduces_golden_stages(g):
    gs = np.load(GS)
    m = _stages_module()
    out = m.outputs(g, O.lsh_signatures, O.cluster_reads_by_kmers, O.snpmer_compare,
                    lambda cb, co, rb, rq, ro, rc: O.consensus_pileups(cb, co, rb, rq, ro, rc, band=64, use_hpc=True))
    for key in STAGE_KEYS:
        assert np.array_equal(out[key], gs[key]), key
    rep = gs["rep_of"]
    assert 10 < int((rep == np.arange(len(rep))).sum()) < 100 and int(gs["pile_n_entries"]) > 200000


@pytest.mark.gpu
def test_gpu_reproduces_golden_stages(g):
    from savont_b200 import _lib
    gs = np.load(GS)
    m = _stages_module()
    ctx = _lib.Context(0, _lib.default_params(band=64))
    out = m.outputs(g, ctx.lsh_signatures, lambda km, off: ctx.cluster_reads_by_kmers(km, off)["rep_of"], ctx.snpmer_compare,
                    lambda cb, co, rb, rq, ro, rc: ctx.consensus_pileups(cb, co, rb, rq, ro, rc, use_hpc=True))
    for key in STAGE_KEYS:
        assert np.array_equal(out[key], gs[key]), key
    ctx.close()
