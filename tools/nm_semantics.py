#!/usr/bin/env python
"""How far is the product's NM definition from the reference's?  

The reference takes `alignment.nm` of minimap2 `lr:hq` (src/alignment.rs:1848-1863); best-NM ties decide the equivalence
classes (:1865-1870) and with them the EM depths.  minimap2 cannot be built here, so this is a SENSITIVITY study with the
CPU oracle (test infrastructure), which can compute NM three ways (so_params.nm_mode):
  0  banded S2 (the product's definition)
  1  S2 with a four times wider band (without any band S2 degenerates: a one-base corner alignment scores 0)
  2  local dual-affine alignment with lr:hq's scoring (match 2, mismatch 4, gaps min(4+2l, 24+l)), NM on its traceback
Everything in front of the NM call (sketch, candidates, selection) is identical in all three runs, so a read can only
differ in its class when it has >= 2 aligned candidates, and in `leq10` through best_nm itself.

usage: python tools/nm_semantics.py [--zymo-all] [--synthetic N] [--out tests/golden/nm_semantics.json]
  --zymo-all   all reads of the reference's tests/data/ont_zymo_1000.trimmed.fq.gz (needs /root/reference; the default is
               the committed 260-read subset tests/golden/zymo_subset.npz)"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402

MODES = {0: "S2 banded (product)", 1: "S2, band 4B", 2: "affine local (lr:hq scoring)"}


def classes(res):
    off = res["class_offsets"]
    return [tuple(res["class_asv"][int(off[i]):int(off[i + 1])]) for i in range(len(off) - 1)]


def depths(res, n_asv):
    """integer depths as the reference derives them (EM over the class histogram, src/alignment.rs:1957-2038)"""
    hist = {}
    for c in classes(res):
        if c:
            hist[c] = hist.get(c, 0) + 1
    keys = sorted(hist)
    offs = np.zeros(len(keys) + 1, np.uint64); flat = []
    for i, k in enumerate(keys):
        flat.extend(k); offs[i + 1] = len(flat)
    cnt = np.array([hist[k] for k in keys], np.uint64)
    total = int(res["totals"][0])
    if total == 0:
        return np.zeros(n_asv, np.int64)
    _, dp, _ = O.em(n_asv, offs, np.array(flat, np.uint32), cnt, total)
    return np.asarray(dp).astype(np.int64)


def study(name, asv_b, asv_o, rb, rq, ro, snp, band, threads):
    n_asv = len(asv_o) - 1
    out = {"dataset": name, "n_reads": int(len(ro) - 1), "n_asv": int(n_asv), "band": band, "modes": {}}
    base = None
    for mode in (0, 1, 2):
        p = O.default_params(band=band, nm_mode=mode)
        orc = O.Oracle(p, snp); orc.load_asvs(asv_b, asv_o)
        t0 = time.perf_counter()
        res = orc.map_reads(rb, rq, ro, n_threads=threads)
        dt = time.perf_counter() - t0
        cl = classes(res); dep = depths(res, n_asv)
        row = {"definition": MODES[mode], "assigned": int(res["totals"][0]), "seconds": round(dt, 1),
               "multi_candidate_reads": None, "leq10_total": int(np.sum(res["leq10"])),
               "median_best_nm": float(np.median(res["best_nm"][res["status"] == 0])) if res["totals"][0] else None}
        if mode == 0:
            base = (res, cl, dep)
            cand = orc.find_candidates(rb, rq, ro)
            row["multi_candidate_reads"] = int(np.sum((cand["offsets"][1:] - cand["offsets"][:-1]) >= 2))
            row["multi_asv_classes"] = int(sum(1 for c in cl if len(c) > 1))
        else:
            r0, c0, d0 = base
            both = (res["status"] == 0) & (r0["status"] == 0)
            row["status_changed_reads"] = int(np.sum(res["status"] != r0["status"]))
            row["class_changed_reads"] = int(sum(1 for i in np.flatnonzero(both) if cl[i] != c0[i]))
            row["class_changed_frac"] = row["class_changed_reads"] / max(int(np.sum(both)), 1)
            row["best_nm_changed_reads"] = int(np.sum(res["best_nm"][both] != r0["best_nm"][both]))
            row["best_nm_mean_abs_diff"] = float(np.mean(np.abs(res["best_nm"][both].astype(np.int64) - r0["best_nm"][both])))
            row["asvs_with_changed_depth"] = int(np.sum(dep != d0))
            row["max_abs_depth_change"] = int(np.max(np.abs(dep - d0))) if n_asv else 0
            row["sum_abs_depth_change"] = int(np.sum(np.abs(dep - d0)))
            row["asvs_with_changed_leq10"] = int(np.sum(res["leq10"] != r0["leq10"]))
        out["modes"][str(mode)] = row
        print(name, row, flush=True)
        orc.close()
    return out


def lowpoly_study(name, asv_b, asv_o, rb, rq, ro, snp, band, threads):
    """candidate_mode = 1 (refine_asv_depths_with_minimap2, src/alignment.rs:1527-1719).  The reference keeps minimap2's hits with
    mapq > 0 (:1579-1581) -- in practice the primary hit of a read that has a clearly better target than any other -- and DROPS a read
    whose best targets cannot be told apart (mapq 0).  The product keeps the minimum-NM ties as a class.  The two agree on every read
    with a unique best ASV; this counts the others and what dropping them would do to the depths (an upper bound of the divergence)."""
    n_asv = len(asv_o) - 1
    out = {"dataset": name, "n_reads": int(len(ro) - 1), "n_asv": int(n_asv), "band": band, "modes": {}}
    for mode in (0, 2):
        p = O.default_params(band=band, nm_mode=mode, candidate_mode=1)
        orc = O.Oracle(p, snp); orc.load_asvs(asv_b, asv_o)
        res = orc.map_reads(rb, rq, ro, n_threads=threads)
        cl = classes(res)
        dep = depths(res, n_asv)
        tied = [i for i, c in enumerate(cl) if len(c) > 1]
        # the same result with the tied reads dropped (what mapq > 0 does to them)
        keep = dict(res)
        off = res["class_offsets"]; flat = []; noff = np.zeros(len(cl) + 1, np.uint64)
        for i, c in enumerate(cl):
            if len(c) == 1:
                flat.extend(c)
            noff[i + 1] = len(flat)
        keep["class_offsets"] = noff; keep["class_asv"] = np.array(flat, np.uint32)
        keep["totals"] = np.array([int(res["totals"][0]) - len(tied), int(res["totals"][1]) + len(tied)], np.uint64)
        dep2 = depths(keep, n_asv)
        row = {"definition": MODES[mode], "assigned": int(res["totals"][0]), "reads_with_tied_best_asvs": len(tied),
               "tied_frac_of_assigned": len(tied) / max(int(res["totals"][0]), 1),
               "asvs_with_changed_depth_if_ties_dropped": int(np.sum(dep != dep2)),
               "max_abs_depth_change_if_ties_dropped": int(np.max(np.abs(dep - dep2))) if n_asv else 0,
               "sum_abs_depth_change_if_ties_dropped": int(np.sum(np.abs(dep - dep2))), "sum_depth": int(np.sum(dep))}
        out["modes"][str(mode)] = row
        print("low-polymorphism mode,", name, row, flush=True)
        orc.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--zymo-all", action="store_true")
    ap.add_argument("--synthetic", type=int, default=2000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=None)
    ap.add_argument("--lowpoly", action="store_true", help="the candidate_mode = 1 study (ties vs the reference's mapq > 0 rule) instead of the NM study")
    a = ap.parse_args()
    fn = lowpoly_study if a.lowpoly else study
    results = []
    from savont_b200 import synth
    if a.zymo_all:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import make_golden as G
        asvs = G.read_fasta(os.path.join(G.REF, "zymo_ref_asvs.fa.gz"))
        reads = G.read_fastq(os.path.join(G.REF, "ont_zymo_1000.trimmed.fq.gz"), 10 ** 9)
        ab, ao = O.concat(asvs)
        rb, ro = O.concat([r[0] for r in reads]); rq, _ = O.concat([r[1] for r in reads])
        snp = synth.snpmers_from_sequences(ab, ao, 17)
        results.append(fn("zymo: all reads of ont_zymo_1000 in [1100, 2000] vs the 63 zymo reference copies", ab, ao, rb, rq, ro, snp, 64, a.threads))
    else:
        z = np.load(os.path.join(ROOT, "tests", "golden", "zymo_subset.npz"))
        results.append(fn("zymo subset: 260 reads (tests/golden/zymo_subset.npz) vs the 63 zymo reference copies", z["asv_bases"], z["asv_offsets"],
                             z["read_bases"], z["read_quals"], z["read_offsets"], z["snpmers"], 64, a.threads))
    if a.synthetic > 0:
        w = synth.Workload("ont16s")
        r = w.reads(a.synthetic)
        results.append(fn(f"config 2 (ont16s): first {a.synthetic} synthetic reads vs 500 ASVs", w.asv_bases, w.asv_offsets, r["bases"], r["quals"],
                             r["offsets"], w.snpmers, w.band, a.threads))
    if a.out:
        json.dump(results, open(a.out, "w"), indent=1)
        print("wrote", a.out)


if __name__ == "__main__":
    main()
