#!/usr/bin/env python
"""Generates tests/golden/zymo_subset.npz (config 1 of BASELINE.json: real ONT R10.4 16S reads against the 63
zymo reference 16S copies) from the reference's own test data.  Runs ONLY in the build container, where
/root/reference exists; the GPU box uses the committed .npz.

Contents: the 63 reference sequences (as the ASV set), the first 260 reads of ont_zymo_1000.trimmed.fq.gz that
pass the reference's length window [1100, 2000] (src/cli.rs:87-92), the SNPmer set derived from the ASVs the way
Stage 1 would (savont_b200.synth.snpmers_from_sequences), and the CPU oracle's outputs for them (sketch
digests, candidates, per-read classes, counters).  NM follows the north-star definition; parity with the
reference's minimap2 NM is unpinned."""
import gzip
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402
from savont_b200 import synth           # noqa: E402

REF = "/root/reference/tests/data"


def read_fasta(path):
    seqs, cur = [], []
    for line in gzip.open(path, "rt"):
        if line.startswith(">"):
            if cur: seqs.append("".join(cur).encode())
            cur = []
        else:
            cur.append(line.strip())
    if cur: seqs.append("".join(cur).encode())
    return seqs


def read_fastq(path, limit):
    out = []
    with gzip.open(path, "rt") as f:
        while len(out) < limit:
            h = f.readline()
            if not h: break
            s = f.readline().strip(); f.readline(); q = f.readline().strip()
            if 1100 <= len(s) <= 2000:
                out.append((s.encode(), q.encode()))
    return out


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    asvs = read_fasta(os.path.join(REF, "zymo_ref_asvs.fa.gz"))
    reads = read_fastq(os.path.join(REF, "ont_zymo_1000.trimmed.fq.gz"), 260)
    ab, ao = O.concat(asvs)
    rb, ro = O.concat([r[0] for r in reads]); rq, _ = O.concat([r[1] for r in reads])
    snp = synth.snpmers_from_sequences(ab, ao, 17)
    p = O.default_params(band=64)
    orc = O.Oracle(p, snp); orc.load_asvs(ab, ao)
    sk = O.sketch_batch(p, snp, rb, rq, ro)
    ask = O.sketch_batch(p, snp, ab, None, ao)
    cand = orc.find_candidates(rb, rq, ro)
    res = orc.map_reads(rb, rq, ro)
    out = os.path.join(ROOT, "tests", "golden", "zymo_subset.npz")
    np.savez_compressed(out, asv_bases=ab, asv_offsets=ao, read_bases=rb, read_quals=rq, read_offsets=ro, snpmers=snp,
                        sk_mini_offsets=sk["mini_offsets"], sk_snp_offsets=sk["snp_offsets"], sk_est_id=sk["est_id"],
                        sk_digest=np.array(digest(sk["mini_pos"], sk["mini_kmer"], sk["mini_fwd"], sk["snp_pos"], sk["snp_kmer"])),
                        asv_sk_digest=np.array(digest(ask["mini_pos"], ask["mini_kmer"], ask["mini_fwd"], ask["snp_pos"], ask["snp_kmer"])),
                        cand_status=cand["status"], cand_offsets=cand["offsets"], cand_asv=cand["asv"], cand_mm=cand["mm"],
                        cand_mini=cand["mini"], cand_strand=cand["strand"],
                        status=res["status"], best_nm=res["best_nm"], class_offsets=res["class_offsets"], class_asv=res["class_asv"],
                        unambig=res["unambig"], ambig=res["ambig"], leq10=res["leq10"], totals=res["totals"])
    print(out, os.path.getsize(out), "bytes;", len(asvs), "ASVs,", len(reads), "reads,", len(snp), "SNPmers; assigned", res["totals"])
    print("status histogram", np.bincount(res["status"], minlength=5), "class sizes", np.bincount((res["class_offsets"][1:] - res["class_offsets"][:-1]).astype(int)))
    print("best_nm median", np.median(res["best_nm"][res["status"] == 0]), "strand rev frac", cand["strand"].mean())


if __name__ == "__main__":
    main()
