#!/usr/bin/env python
"""Generates tests/golden/zymo_stage1.npz: the first 400 records of the reference's ont_zymo_1000.trimmed.fq.gz as they
are (ids with their trailing "rc" tokens, reads outside the length window included), and the CPU oracle's outputs for the
two stages in front of Stage 7: Stage 1 counting (seq_parse::read_to_split_kmers, src/seq_parse.rs:12-83) with the
high-frequency rule (src/kmer_comp.rs:459-500), and the read-side wrapper (kmer_comp::twin_reads_from_snpmers,
src/kmer_comp.rs:68-258, Cli defaults 1100 / 2000 / 98.0) with the Stage-7 read order (src/main.rs:538).
Runs ONLY in the build container, where /root/reference exists; the GPU box uses the committed .npz."""
import gzip
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402

REF = "/root/reference/tests/data"


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    recs = []
    with gzip.open(os.path.join(REF, "ont_zymo_1000.trimmed.fq.gz"), "rt") as f:
        while len(recs) < 400:
            h = f.readline()
            if not h:
                break
            s = f.readline().strip(); f.readline(); q = f.readline().strip()
            recs.append((h[1:].rstrip("\n"), s.encode(), q.encode()))
    ids = [r[0] for r in recs]
    rc = np.array([1 if i.split()[-1:] == ["rc"] else 0 for i in ids], np.uint8)       # src/seq_parse.rs:361-369
    rb, ro = O.concat([r[1] for r in recs]); rq, _ = O.concat([r[2] for r in recs])
    idb, ido = O.concat_ids(ids)
    g = np.load(os.path.join(ROOT, "tests", "golden", "zymo_subset.npz"))
    snp = g["snpmers"]
    kc = O.count_split_kmers(rb, rq, ro, rc=rc)
    thresh, hf = O.high_freq_kmers(kc["kmers"], kc["counts"])
    p = O.default_params(band=64)
    # the sample is too small for a k-mer to exceed the threshold of 100: take the 200 most frequent k-mers as the
    # high-frequency set so that the repetitive-read test has something to look up
    top = kc["kmers"][np.argsort(-kc["counts"].sum(1).astype(np.int64), kind="stable")[:200]]
    flt = O.twin_reads_filter(p, O.read_filter(1100, 2000, 98.0), snp, top, rb, rq, ro, ids)
    out = os.path.join(ROOT, "tests", "golden", "zymo_stage1.npz")
    np.savez_compressed(out, read_bases=rb, read_quals=rq, read_offsets=ro, id_bytes=idb, id_offsets=ido, rc=rc, snpmers=snp,
                        kc_n=np.array(len(kc["kmers"])), kc_n_distinct=np.array(kc["n_distinct"]),
                        kc_digest=np.array(digest(kc["kmers"], kc["counts"])), kc_head=kc["kmers"][:16], kc_head_counts=kc["counts"][:16],
                        hf_thresh=np.array(thresh), hf_n=np.array(len(hf)), hf_set=top,
                        flt_status=flt["status"], flt_order=flt["order"], flt_est_id=flt["est_id"],
                        flt_solid_mini=flt["n_solid_mini"], flt_solid_snp=flt["n_solid_snp"])
    print(out, os.path.getsize(out), "bytes;", len(recs), "reads,", int(rc.sum()), "rc-tagged;", len(kc["kmers"]), "k-mers of",
          kc["n_distinct"], "distinct; thresh", thresh, "hf", len(hf))
    print("a4 status histogram", np.bincount(flt["status"], minlength=8), "kept", len(flt["order"]))


if __name__ == "__main__":
    main()
