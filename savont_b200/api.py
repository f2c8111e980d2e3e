"""Python mirror of the reference's Rust interface for the read-vs-ASV path (same function names and
argument meaning as savont v0.6.4), implemented over libsavont_b200.so (C ABI) and libsavont_host.so (the
C++ host layer: equivalence classes, EM, writers).  No CPU fallback: without the CUDA library every
function raises.

  get_twin_read_syncmer      src/seeding.rs:317-327
  twin_reads_from_fasta      src/kmer_comp.rs:39-66
  find_compatible_candidates src/asv_cluster.rs:356-360 (+ the selection of src/alignment.rs:1796-1843)
  refine_asv_depths_with_em  src/alignment.rs:1723-1729
  compute_per_sample_depths  src/alignment.rs:2044-2051
  write_consensus_fasta      src/alignment.rs:830-860
  write_feature_table        src/main.rs:381-400
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _lib

_HOST_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libsavont_host.so")
_host = None


def host():
    global _host
    if _host is None:
        _lib.lib()                      # libsavont_b200.so first (rpath $ORIGIN also finds it)
        H = C.CDLL(_HOST_PATH)
        H.savont_host_last_error.restype = C.c_char_p
        H.savont_host_em.restype = C.c_int
        H.savont_host_refine.restype = C.c_longlong
        H.savont_host_per_sample.restype = C.c_int
        H.savont_host_write_outputs.restype = C.c_int
        H.savont_host_twin_read_order.restype = C.c_longlong
        H.savont_host_high_freq_kmers.restype = C.c_longlong
        H.savont_host_refine_twin.restype = C.c_longlong
        H.savont_host_per_sample_twin.restype = C.c_int
        _host = H
    return _host


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _fail(what):
    raise RuntimeError(f"{what}: {host().savont_host_last_error().decode()}")


@dataclass
class Cli:
    """the Cli fields the path reads (src/cli.rs)"""
    kmer_size: int = 17
    c: int = 11
    blockmer_length: int = 3
    minimum_base_quality: int = 25
    band: int = 64
    strand_mode: int = 0
    low_polymorphism: bool = False     # src/alignment.rs:1730: refine_asv_depths_with_minimap2 instead of the SNPmer index
    min_read_length: int = 1100        # src/cli.rs:87-88
    max_read_length: int = 2000        # src/cli.rs:91-92
    quality_value_cutoff: float = 98.0  # src/cli.rs:95-96
    primary_clustering_threshold: float = 0.95   # src/cli.rs:185-186 (Stage 2)
    min_cluster_size: int = 12          # src/cli.rs (Stage 2 / 3 size filter; --hifi lowers it, src/main.rs:470-473)
    use_hpc: bool = True                # homopolymer compression in Stage 4 (src/alignment.rs:481)

    def params(self) -> _lib.Params:
        return _lib.default_params(k=self.kmer_size, c=self.c, l=self.blockmer_length, min_bq=self.minimum_base_quality,
                                   band=self.band, strand_mode=self.strand_mode, candidate_mode=1 if self.low_polymorphism else 0)


@dataclass
class ConsensusSequence:
    """src/types.rs:162-190, the fields Stage 7 / 7b and the writers touch"""
    sequence: bytes
    depth: int = 0
    id: int = 0
    unambig_best_read_map_count: int = 0
    ambig_read_map_count: int = 0
    num_map_leq_10nm: int = 0
    per_sample_depths: list = field(default_factory=list)


def _concat(seqs):
    offs = np.zeros(len(seqs) + 1, np.uint64)
    if len(seqs):
        offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(bytes(s) for s in seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
    return buf, offs


def get_twin_read_syncmer(ctx: _lib.Context, string: bytes, qualities, id: str = ""):
    """one sequence -> dict(minimizer_positions, minimizer_kmers, snpmer_positions, snpmer_kmers, est_id) or None"""
    b, o = _concat([string])
    q = None if qualities is None else np.frombuffer(bytes(qualities), np.uint8)
    r = ctx.sketch(b, q, o, want_codes=False)
    if r["status"][0] != 0:
        return None
    est = float(r["est_id"][0])
    return dict(id=id, minimizer_positions=r["mini_pos"], minimizer_kmers=r["mini_kmer"], snpmer_positions=r["snp_pos"],
                snpmer_kmers=r["snp_kmer"], est_id=None if np.isnan(est) else est, base_length=len(string))


def twin_reads_from_fasta(ctx: _lib.Context, seqs):
    """ASV sequences -> sketches (qualities None, no high-frequency pruning)"""
    b, o = _concat(seqs)
    return ctx.sketch(b, None, o, want_codes=False)


def read_to_split_kmers(ctx: _lib.Context, reads, single_strand=False):
    """Stage 1 counting (src/seq_parse.rs:12-83, bloom_filter_size 0): reads = dict(bases, quals, offsets, ids [optional]);
    a read whose id ends in the token "rc" is reverse-complemented first (:361-369).
    -> dict(kmers, counts [n, 2], n_distinct), sorted by k-mer"""
    rc = None
    if reads.get("ids") is not None:
        rc = np.array([1 if (i.decode() if isinstance(i, (bytes, bytearray)) else i).split()[-1:] == ["rc"] else 0 for i in reads["ids"]], np.uint8)
    out = ctx.count_split_kmers(reads["bases"], reads.get("quals"), reads["offsets"], rc=rc, single_strand=single_strand)
    return dict(kmers=out["kmers"], counts=out["counts"], n_distinct=out["stats"]["n_distinct"])


def high_freq_kmers(kmers, counts):
    """src/kmer_comp.rs:459-500 -> (high_freq_thresh, high-frequency k-mers)"""
    H = host()
    km = np.ascontiguousarray(kmers, np.uint64); ct = np.ascontiguousarray(counts, np.uint32).reshape(-1)
    out = np.zeros(max(len(km), 1), np.uint64); nh = C.c_size_t(0)
    th = H.savont_host_high_freq_kmers(_p(km, C.c_uint64), _p(ct, C.c_uint32), C.c_size_t(len(km)), _p(out, C.c_uint64), C.byref(nh))
    if th < 0:
        _fail("high_freq_kmers")
    return int(th), out[:int(nh.value)]


def twin_reads_from_snpmers(ctx: _lib.Context, reads, args: Cli, high_freq_kmers=None):
    """Stage 1.5, the part Stage 7 depends on (src/kmer_comp.rs:68-258 + src/main.rs:538): which FASTQ reads become
    TwinReads and in which order.  reads: dict(bases, quals, offsets, ids [optional list of str/bytes]).
    Returns dict(status, est_id, n_solid_mini, n_solid_snp, order, auto_low_polymorphism)."""
    H = host()
    ctx.load_high_freq_kmers(high_freq_kmers)
    ctx.set_read_filter(_lib.read_filter(args.min_read_length, args.max_read_length, args.quality_value_cutoff))
    try:
        out = ctx.twin_reads_from_snpmers(reads["bases"], reads.get("quals"), reads["offsets"])
    finally:
        ctx.set_read_filter(None)
    n = len(out["status"])
    ids = reads.get("ids")
    idb = ido = None
    if ids is not None:
        idb, ido = _concat([i.encode() if isinstance(i, str) else bytes(i) for i in ids])
    order = np.zeros(max(n, 1), np.uint32)
    st = np.ascontiguousarray(out["status"]); est = np.ascontiguousarray(out["est_id"])
    nk = H.savont_host_twin_read_order(_p(st, C.c_uint8), _p(est, C.c_double), C.c_size_t(n), _p(idb, C.c_uint8), _p(ido, C.c_uint64),
                                       _p(order, C.c_uint32))
    if nk < 0:
        _fail("twin_read_order")
    out["order"] = order[:int(nk)]
    kept = out["order"]
    out["auto_low_polymorphism"] = bool(len(kept) and float((out["n_solid_snp"][kept] == 0).mean()) > 0.75)
    return out


def find_compatible_candidates(ctx: _lib.Context, bases, quals, offsets):
    return ctx.find_candidates(bases, quals, offsets)


def em(n_asv, cls_offsets, cls_asv, counts, total_assigned):
    H = host()
    co = np.ascontiguousarray(cls_offsets, np.uint64); ca = np.ascontiguousarray(cls_asv, np.uint32); cn = np.ascontiguousarray(counts, np.uint64)
    ab = np.zeros(max(n_asv, 1), np.float64); dp = np.zeros(max(n_asv, 1), np.uint64)
    it = H.savont_host_em(C.c_size_t(n_asv), _p(co, C.c_uint64), _p(ca, C.c_uint32), _p(cn, C.c_uint64), C.c_size_t(len(cn)),
                          C.c_uint64(int(total_assigned)), _p(ab, C.c_double), _p(dp, C.c_uint64))
    if it < 0:
        _fail("savont_host_em")
    return ab[:n_asv], dp[:n_asv], it


def eq_classes(status, class_offsets, class_asv, file_idx=None, sample=None):
    """sorted equivalence classes (CSR) with counts from per-read classes"""
    hist = {}
    co = np.asarray(class_offsets, np.int64)
    sel = np.flatnonzero(np.asarray(status) == 0)
    if sample is not None and file_idx is not None:
        sel = sel[np.asarray(file_idx)[sel] == sample]
    for i in sel:
        key = tuple(int(x) for x in class_asv[co[i]:co[i + 1]])
        hist[key] = hist.get(key, 0) + 1
    keys = sorted(hist)
    offs = np.zeros(len(keys) + 1, np.uint64)
    if keys:
        offs[1:] = np.cumsum([len(k) for k in keys])
    flat = np.array([a for k in keys for a in k], np.uint32)
    return offs, flat, np.array([hist[k] for k in keys], np.uint64), int(len(sel))


def _twin_ptrs(tw):
    """TwinRead-field dict (savont_b200/twin.py) -> ctypes arguments of the *_twin host entry points"""
    w, wo, ln, sk, kb, so, fi = _lib.Context._twin_args(tw)
    keep = (w, wo, ln, sk, so, fi)
    return keep, [_p(w, C.c_uint64), _p(wo, C.c_uint64), _p(ln, C.c_uint32), C.c_void_p(sk.ctypes.data), C.c_int(kb), _p(so, C.c_uint64)], fi, len(ln)


def refine_asv_depths_with_em(twin_reads, consensuses, kmer_info, args: Cli, temp_dir=None, device=0):
    """Stage 7.  twin_reads: dict(bases, quals, offsets) (raw reads) or the TwinRead fields dict(dna_words, dna_word_offsets,
    lengths, snpmer_kmers, snpmer_offsets) of savont_b200/twin.py (what the reference passes: &[TwinRead]);
    kmer_info: u64 SNPmer k-mers (both alleles); consensuses: list[ConsensusSequence], updated and filtered in place like
    the reference."""
    H = host()
    ab, ao = _concat([c.sequence for c in consensuses])
    snp = np.ascontiguousarray(kmer_info, np.uint64)
    A = len(consensuses)
    out = [np.zeros(max(A, 1), np.uint64) for _ in range(5)]
    p = args.params()
    td = None if temp_dir is None else str(temp_dir).encode()
    if "dna_words" in twin_reads:
        keep, tp, _, n_reads = _twin_ptrs(twin_reads)
        n = H.savont_host_refine_twin(C.c_int(device), C.byref(p), _p(snp, C.c_uint64), C.c_size_t(len(snp)), _p(ab, C.c_uint8),
                                      _p(ao, C.c_uint64), C.c_size_t(A), *tp, C.c_size_t(n_reads), td, *[_p(x, C.c_uint64) for x in out])
    else:
        b = np.ascontiguousarray(twin_reads["bases"], np.uint8)
        q = None if twin_reads.get("quals") is None else np.ascontiguousarray(twin_reads["quals"], np.uint8)
        o = np.ascontiguousarray(twin_reads["offsets"], np.uint64)
        n = H.savont_host_refine(C.c_int(device), C.byref(p), _p(snp, C.c_uint64), C.c_size_t(len(snp)), _p(ab, C.c_uint8), _p(ao, C.c_uint64),
                                 C.c_size_t(A), _p(b, C.c_uint8), _p(q, C.c_uint8), _p(o, C.c_uint64), C.c_size_t(len(o) - 1),
                                 td, *[_p(x, C.c_uint64) for x in out])
    if n < 0:
        _fail("refine_asv_depths_with_em")
    kept = []
    for i in range(int(n)):
        c = consensuses[int(out[4][i])]
        c.depth = int(out[0][i]); c.unambig_best_read_map_count = int(out[1][i])
        c.ambig_read_map_count = int(out[2][i]); c.num_map_leq_10nm = int(out[3][i])
        kept.append(c)
    consensuses[:] = kept


def compute_per_sample_depths(twin_reads, n_samples, consensuses, kmer_info, args: Cli, device=0):
    """Stage 7b -> list[asv][sample]; twin_reads as in refine_asv_depths_with_em, with file_idx"""
    H = host()
    ab, ao = _concat([c.sequence for c in consensuses])
    snp = np.ascontiguousarray(kmer_info, np.uint64)
    A = len(consensuses)
    out = np.zeros(max(A * n_samples, 1), np.uint64)
    p = args.params()
    if "dna_words" in twin_reads:
        keep, tp, fi, n_reads = _twin_ptrs(twin_reads)
        rc = H.savont_host_per_sample_twin(C.c_int(device), C.byref(p), _p(snp, C.c_uint64), C.c_size_t(len(snp)), _p(ab, C.c_uint8),
                                           _p(ao, C.c_uint64), C.c_size_t(A), *tp, _p(fi, C.c_uint32), C.c_size_t(n_reads),
                                           C.c_size_t(n_samples), _p(out, C.c_uint64))
    else:
        b = np.ascontiguousarray(twin_reads["bases"], np.uint8)
        q = None if twin_reads.get("quals") is None else np.ascontiguousarray(twin_reads["quals"], np.uint8)
        o = np.ascontiguousarray(twin_reads["offsets"], np.uint64)
        fi = np.ascontiguousarray(twin_reads["file_idx"], np.uint32)
        rc = H.savont_host_per_sample(C.c_int(device), C.byref(p), _p(snp, C.c_uint64), C.c_size_t(len(snp)), _p(ab, C.c_uint8), _p(ao, C.c_uint64),
                                      C.c_size_t(A), _p(b, C.c_uint8), _p(q, C.c_uint8), _p(o, C.c_uint64), _p(fi, C.c_uint32),
                                      C.c_size_t(len(o) - 1), C.c_size_t(n_samples), _p(out, C.c_uint64))
    if rc != 0:
        _fail("compute_per_sample_depths")
    return out[:A * n_samples].reshape(A, n_samples).astype(np.int64).tolist()


def _write(consensuses, fasta_path, prefix, table_path, sample_names):
    H = host()
    ab, ao = _concat([c.sequence for c in consensuses])
    A = len(consensuses)
    arr = lambda f: np.array([f(c) for c in consensuses] or [0], np.uint64)
    ns = len(consensuses[0].per_sample_depths) if A and consensuses[0].per_sample_depths else 0
    ps = np.array([d for c in consensuses for d in c.per_sample_depths] or [0], np.uint64) if ns else None
    rc = H.savont_host_write_outputs(_p(ab, C.c_uint8), _p(ao, C.c_uint64), C.c_size_t(A), _p(arr(lambda c: c.depth), C.c_uint64),
                                     _p(arr(lambda c: c.id), C.c_uint64), _p(arr(lambda c: c.unambig_best_read_map_count), C.c_uint64),
                                     _p(arr(lambda c: c.ambig_read_map_count), C.c_uint64), _p(arr(lambda c: c.num_map_leq_10nm), C.c_uint64),
                                     _p(ps, C.c_uint64), C.c_size_t(ns), None if fasta_path is None else str(fasta_path).encode(),
                                     prefix.encode(), None if table_path is None else str(table_path).encode(),
                                     "\t".join(sample_names or ["sample"]).encode())
    if rc != 0:
        _fail("write outputs")


def write_consensus_fasta(consensuses, output_path, prefix):
    _write(consensuses, output_path, prefix, None, None)


def write_feature_table(consensuses, path, sample_names):
    _write(consensuses, None, "final", path, sample_names)


def write_read_to_asv_mappings(path, read_ids, consensus_ids, rows):
    """temp/read_to_asv_mappings.tsv (src/alignment.rs:1766-1768, :1875-1887) from Context.mapping_rows_resident():
    "{read id}\tdebug_id:{consensus id}\t{SNPmer mismatches}\t{NM}" per row, reads in index order"""
    ro = rows["row_offsets"]
    with open(path, "w") as f:
        for r, rid in enumerate(read_ids):
            for i in range(int(ro[r]), int(ro[r + 1])):
                f.write(f"{rid}\tdebug_id:{consensus_ids[int(rows['asv'][i])]}\t{int(rows['snp_mismatches'][i])}\t{int(rows['nm'][i])}\n")


# ---------------------------------------------------------------------------------------------------------
# Stage 2 / 3 / 4 neighbours of the path: host halves of the reference functions
# ---------------------------------------------------------------------------------------------------------
def clusters_from_assignment(rep_of, min_cluster_size: int = 1):
    """the tail of cluster_reads_by_kmers (src/asv_cluster.rs:192-210): group by representative, order by size descending,
    members ascending, drop clusters below min_cluster_size.  Equal sizes: the reference orders by `a.first()` of a member list
    filled in std::HashMap iteration order (random per process); here by the smallest member (= the representative)."""
    rep_of = np.asarray(rep_of, np.int64)
    order = np.argsort(rep_of, kind="stable")
    reps, start = np.unique(rep_of[order], return_index=True)
    groups = np.split(order, start[1:]) if len(order) else []
    clusters = [np.sort(g) for g in groups]
    clusters.sort(key=lambda c: (-len(c), int(c[0])))
    return [c for c in clusters if len(c) >= min_cluster_size]


def cluster_reads_by_kmers(ctx: _lib.Context, minimizer_kmers, minimizer_offsets, min_cluster_size: int = 1, threshold: float = 0.95):
    """asv_cluster::cluster_reads_by_kmers (src/asv_cluster.rs:72-249) on `TwinRead::minimizer_kmers()` in CSR"""
    out = ctx.cluster_reads_by_kmers(minimizer_kmers, minimizer_offsets, threshold=threshold)
    return clusters_from_assignment(out["rep_of"], min_cluster_size)


def snpmer_clusters_from_assignment(kmer_clusters, rep_of, min_cluster_size: int = 1):
    """the tail of the per-cluster loop of cluster_reads_by_snpmers (src/asv_cluster.rs:660-690): per k-mer cluster its SNPmer
    clusters -- grouped by representative, members ascending, larger first (equal sizes: smaller first member), small ones dropped"""
    rep_of = np.asarray(rep_of, np.int64)
    out = []
    for kc in kmer_clusters:
        kc = np.asarray(kc, np.int64)
        local = clusters_from_assignment_subset(kc, rep_of[kc])
        out.append([c for c in local if len(c) >= min_cluster_size])
    return out


def clusters_from_assignment_subset(reads, reps):
    order = np.lexsort((reads, reps))
    reads, reps = reads[order], reps[order]
    cuts = np.flatnonzero(np.diff(reps)) + 1 if len(reps) else np.zeros(0, np.int64)
    groups = np.split(reads, cuts) if len(reads) else []
    groups.sort(key=lambda c: (-len(c), int(c[0])))
    return groups


def cluster_reads_by_snpmers(ctx: _lib.Context, snpmer_kmers, snpmer_offsets, kmer_clusters, min_cluster_size: int = 1, iterative_above: int = 1000):
    """asv_cluster::cluster_reads_by_snpmers up to the reclustering rounds (src/asv_cluster.rs:561-700) on `TwinRead::snpmer_kmers()` in
    CSR and the k-mer clusters of Stage 2 -> per k-mer cluster the list of its SNPmer clusters"""
    cr = np.concatenate([np.asarray(c, np.uint32) for c in kmer_clusters]) if len(kmer_clusters) else np.zeros(0, np.uint32)
    co = np.zeros(len(kmer_clusters) + 1, np.uint64)
    co[1:] = np.cumsum([len(c) for c in kmer_clusters])
    out = ctx.cluster_reads_by_snpmers(snpmer_kmers, snpmer_offsets, cr, co, iterative_above=iterative_above)
    return snpmer_clusters_from_assignment(kmer_clusters, out["rep_of"], min_cluster_size)


def clusters_from_reassignment(clusters, best_cons, min_cluster_size: int = 1):
    """the tail of reassign_reads_to_best_cluster (src/asv_cluster.rs:1040-1131): reads listed cluster by cluster (`clusters`),
    best_cons[i] = new cluster of the i-th listed read -> (new clusters: members ascending, empty / small ones dropped,
    number of reads that changed cluster)"""
    flat = np.concatenate([np.asarray(c, np.int64) for c in clusters]) if clusters else np.zeros(0, np.int64)
    old = np.concatenate([np.full(len(c), i, np.int64) for i, c in enumerate(clusters)]) if clusters else np.zeros(0, np.int64)
    best = np.asarray(best_cons, np.int64)
    new = [np.sort(flat[best == i]) for i in range(len(clusters))]
    return [c for c in new if len(c) and len(c) >= min_cluster_size], int((best != old).sum())


def reassign_reads_to_best_cluster(ctx: _lib.Context, clusters, cons_kmers, cons_offsets, snpmer_kmers, snpmer_offsets, min_cluster_size=1):
    """asv_cluster::reassign_reads_to_best_cluster (src/asv_cluster.rs:1007-1131, SNPmer markers): `cons_*` = the ConsensusSnpmer
    k-mers of the current clusters, `snpmer_*` = snpmers_vec() k-mers of ALL reads (CSR indexed by read id)"""
    so = np.asarray(snpmer_offsets, np.uint64); sk = np.asarray(snpmer_kmers, np.uint64)
    ids = np.concatenate([np.asarray(c, np.int64) for c in clusters]) if clusters else np.zeros(0, np.int64)
    lens = (so[ids + 1] - so[ids]).astype(np.int64)
    ro = np.zeros(len(ids) + 1, np.uint64); ro[1:] = np.cumsum(lens)
    rk = np.concatenate([sk[int(so[i]):int(so[i + 1])] for i in ids]) if len(ids) else np.zeros(0, np.uint64)
    out = ctx.snpmer_compare(cons_kmers, cons_offsets, rk, ro)
    return clusters_from_reassignment(clusters, out["best_cons"], min_cluster_size)


def pileups_from_columns(res, cons_offsets):
    """savont_b200_consensus_pileups output -> the reference's Vec<Vec<Pileup>> shape: per consensus a list of columns
    dict(ref_hp_length, bases=[("Base", base, qual, hp) | ("Deletion",) | ("Insertion", [(base, qual, hp), ..])]) (src/alignment.rs:17-60)"""
    co = np.asarray(cons_offsets, np.int64) - int(cons_offsets[0])
    po, ent, hp = res["pos_offsets"], res["entries"], res["ref_hp_length"]
    out = []
    for ci in range(len(co) - 1):
        cols = []
        for p in range(int(co[ci]), int(co[ci + 1])):
            col = []
            for e in ent[int(po[p]):int(po[p + 1])]:
                kind = int(e["kind"])
                if kind == 0:
                    col.append(("Base", chr(e["base"][0]), int(e["qual"][0]), int(e["hp"][0])))
                elif kind == 1:
                    col.append(("Deletion",))
                else:
                    col.append(("Insertion", [(chr(e["base"][j]), int(e["qual"][j]), int(e["hp"][j])) for j in range(int(e["n"]))]))
            cols.append(dict(ref_hp_length=int(hp[p]), bases=col))
        out.append(cols)
    return out


def generate_consensus_pileups(ctx: _lib.Context, consensus_seqs, clusters, read_seqs, read_quals, args: Cli):
    """alignment::generate_consensus_pileups (src/alignment.rs:415-641): consensus_seqs[i] = the (compressed) consensus of cluster i,
    clusters[i] = its read ids (the first MAX_SEQS_CONSENSUS = 250 are used, :446-450), read_seqs / read_quals = per read id the
    bases decoded from dna_seq and the per-base qualities.  Returns the reference's Vec<Vec<Pileup>> shape (pileups_from_columns)."""
    MAX_SEQS_CONSENSUS = 250                                  # src/constants.rs:60
    cb, co = _concat(consensus_seqs)
    ids = [int(r) for c in clusters for r in list(c)[:MAX_SEQS_CONSENSUS]]
    rcons = np.array([i for i, c in enumerate(clusters) for _ in list(c)[:MAX_SEQS_CONSENSUS]], np.uint32)
    rb, ro = _concat([read_seqs[r] for r in ids])
    rq = None if read_quals is None else _concat([read_quals[r] for r in ids])[0]
    ctx.set_params(args.params())
    res = ctx.consensus_pileups(cb, co, rb, rq, ro, rcons, use_hpc=args.use_hpc)
    return pileups_from_columns(res, co)
