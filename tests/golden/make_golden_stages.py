#!/usr/bin/env python
"""Generates tests/golden/zymo_stages.npz from tests/golden/zymo_subset.npz (the reference's own zymo test data, see
make_golden.py): the CPU oracle's outputs for the neighbouring stages on the same 260 real ONT reads and 63 reference copies --
Stage 2 (LSH signatures + greedy k-mer clustering of the reads), Stage 3 reclustering comparison (the reads' SNPmers against the
reference copies' SNPmers as cluster consensus lists) and Stage 4 pileups (every assigned read against the homopolymer-compressed
copy it was assigned to).  Only small vectors and digests are stored."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def inputs(g):
    """the derived inputs of the three stages (also used by tests/test_golden.py)"""
    p = O.default_params(band=64)
    sk = O.sketch_batch(p, g["snpmers"], g["read_bases"], g["read_quals"], g["read_offsets"])
    ask = O.sketch_batch(p, g["snpmers"], g["asv_bases"], None, g["asv_offsets"])
    # Stage 4: reads with a class, grouped by their first class member; consensus = the run-length compressed copy
    co, ca = g["class_offsets"].astype(np.int64), g["class_asv"]
    first = np.array([int(ca[co[i]]) if co[i + 1] > co[i] else -1 for i in range(len(co) - 1)])
    order = np.argsort(np.where(first < 0, 1 << 30, first), kind="stable")
    order = order[first[order] >= 0]
    used = np.unique(first[order])
    ro = g["read_offsets"].astype(np.int64)
    reads = [bytes(g["read_bases"][ro[i]:ro[i + 1]]) for i in order]
    quals = [bytes(g["read_quals"][ro[i]:ro[i + 1]]) for i in order]
    ao = g["asv_offsets"].astype(np.int64)
    cons = []
    for a in used:
        s = np.asarray(g["asv_bases"][ao[a]:ao[a + 1]])
        cons.append(s[np.concatenate([[True], s[1:] != s[:-1]])].tobytes())
    rcons = np.searchsorted(used, first[order]).astype(np.uint32)
    return dict(mini=(np.asarray(sk["mini_kmer"], np.uint64), np.asarray(sk["mini_offsets"], np.uint64)),
                snp=(np.asarray(sk["snp_kmer"], np.uint64), np.asarray(sk["snp_offsets"], np.uint64)),
                asv_snp=(np.asarray(ask["snp_kmer"], np.uint64), np.asarray(ask["snp_offsets"], np.uint64)),
                pile=(O.concat(cons), O.concat(reads), O.concat(quals)[0], rcons))


def outputs(g, lsh, cluster, compare, pileups):
    """runs the three stages through the given callables and reduces them to what the fixture stores"""
    x = inputs(g)
    sig, valid = lsh(*x["mini"])
    rep = cluster(*x["mini"])
    cmp_ = compare(*x["asv_snp"], *x["snp"])
    (cb, co), (rb, ro), rq, rc = x["pile"]
    pl = pileups(cb, co, rb, rq, ro, rc)
    return dict(lsh_digest=np.array(digest(sig, valid)), rep_of=np.asarray(rep, np.uint32), best_cons=cmp_["best_cons"],
                best_mismatches=cmp_["best_mismatches"], best_matches=cmp_["best_matches"],
                pile_digest=np.array(digest(pl["pos_offsets"], pl["entries"], pl["ref_hp_length"], pl["read_strand"], pl["read_nm"])),
                pile_read_nm=pl["read_nm"], pile_n_entries=np.array(len(pl["entries"])))


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "zymo_subset.npz"))
    out = outputs(g, O.lsh_signatures, O.cluster_reads_by_kmers, O.snpmer_compare,
                  lambda cb, co, rb, rq, ro, rc: O.consensus_pileups(cb, co, rb, rq, ro, rc, band=64, use_hpc=True))
    path = os.path.join(ROOT, "tests", "golden", "zymo_stages.npz")
    np.savez_compressed(path, **out)
    rep = out["rep_of"]
    print(path, os.path.getsize(path), "bytes; representatives", int((rep == np.arange(len(rep))).sum()), "of", len(rep),
          "reads; reassigned to", len(np.unique(out["best_cons"])), "copies; pileup entries", int(out["pile_n_entries"]),
          "median NM of the compressed reads", float(np.median(out["pile_read_nm"])))


if __name__ == "__main__":
    main()
