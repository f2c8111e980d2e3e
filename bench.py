#!/usr/bin/env python
"""bench.py -- reads aligned/s (and GCUPS of the banded DP) of the read-vs-ASV assignment path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ont16s] [--reads R]

A "step" = one pass of the hot path (sketch -> candidate filter -> banded DP -> tie/eq-class
counters) over one batch of synthetic reads (BASELINE.json configs[1]: 1M ONT 16S reads vs 500
ASVs by default).  N>1 (under torchrun): reads are sharded, R reads per rank (weak scaling), ASV
index replicated, one all-reduce of the per-ASV count vector per step (NCCL, in place on the
library's device counter vector).

  value     whole-job reads/s with the reads (FASTQ bytes) already resident in HBM (kernel pipeline only)
  e2e       the same metric through the drop-in call of the reference's Stage 7 seam: `refine_asv_depths_with_em(twin_reads:
            &[TwinRead], ..)` hands over TwinReads as they lie in the reference's memory -- one pageable heap block per read for
            dna_seq and for snpmer_kmers, one pointer per read -- through savont_b200_map_twin_read_ptrs: the library's host threads
            gather the blocks into page-locked staging INSIDE the timed region, as are the H2D of them and the D2H of the per-read
            results.  With --no-twin: the FASTQ-byte call below.
  e2e_twin_contiguous   the TwinRead fields already concatenated in pinned memory (savont_b200_map_twin_reads): no gather
  e2e_fastq the same through savont_b200_map_reads: ASCII bases + Phred bytes (the form the reference arm / CPU oracle
            is fed with), 5.5x more H2D bytes -> PCIe bound
  value_twin_resident   `value` with the TwinRead fields resident instead of FASTQ bytes (no ASCII encode, no SNPmer detection)
  value_with_cigar      the resident pass followed by savont_b200_best_hits_resident: strand, coordinates, identity and CIGAR of
            every read's best hit fetched as well (rank 0, one pass, host wall clock; --no-cigar skips it)
  e2e_single_process    N > 1: rank 0 alone drives all N GPUs through savont_b200_create_multi (the other ranks idle)
  multi_gpu_check       N > 1: all-reduced counters == sum over ranks == one GPU over the concatenated samples == oracle
  roofline  the dominant kernel (banded DP): achieved GCUPS vs the INT32 ALU-pipe peak; traffic from the committed ncu capture
            of the same kernel (profiles/<PROFILE_TAG>_*_full.txt) while its duration is within 10 % of this run's
  cpu_baseline / --impl reference: the CPU oracle (C port of the reference path; the Rust
            reference itself cannot be built here) on all host cores.
  --path kmercount | cluster | reassign | snpcluster | pileup: the neighbouring stages (one JSON line each, same keys)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=float(d.get("hbm_gbs", 6650.0)), source="measured (MEASURED_PEAKS.json)",
                    sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)))
    return dict(hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)", sm_max_mhz=1965.0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        busy = [x for x in sm if x > 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def _workload(args):
    from savont_b200 import synth
    w = synth.Workload(args.workload)
    return w


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference path on all host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import oracle as O
    w = _workload(args)
    cores = os.cpu_count() or 1
    sample = args.cpu_reads
    r = w.reads(sample)
    orc = O.Oracle(O.default_params(band=w.band), w.snpmers)
    orc.load_asvs(w.asv_bases, w.asv_offsets)
    times, cells = [], 0
    for it in range(args.warmup + args.steps):
        if it >= args.warmup and it > args.warmup and sum(times) > 150:
            break
        t0 = time.perf_counter()
        res = orc.map_reads(r["bases"], r["quals"], r["offsets"], n_threads=cores)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt); cells = res["dp_cells"]
    steps = len(times)
    tot = sum(times)
    value = sample * steps / tot
    line = {
        "impl": "reference", "metric": "reads aligned/sec (read-vs-ASV assignment path)", "value": value, "unit": "reads/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 (bit-parallel INT32 DP) / f64 thresholds", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w.cfg['desc']}", "n_asv": w.n_asv, "band": w.band, "reads_per_step": sample},
        "gcups": cells * steps / tot / 1e9,
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} reads of the workload per step, OpenMP over reads, scalar banded DP "
                                   "(oracle/savont_oracle.c; the Rust reference + minimap2 cannot be built in this image)"},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_kmercount_sharded(args, rank, world, local_rank):
    """--path kmercount --gpus N: every rank counts its own shard of reads; partition p of the 8192 hash partitions belongs to rank
    p * N / 8192 and the scatter kernel writes the runs straight into the owner's buffer over NVLink (savont_b200/dist.py:
    count_split_kmers_sharded).  Weak scaling: --reads per GPU."""
    import torch
    import torch.distributed as dist
    from savont_b200 import _lib, dist as sd
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    _bind_to_gpu_numa_node(local_rank)
    w = _workload(args)
    n_reads = args.reads
    ctx = _lib.Context(local_rank, _lib.default_params(band=w.band))
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    r = w.reads(n_reads, first=rank * n_reads, alloc=_lib.pinned_array)
    ctx.upload_reads(r["bases"], r["quals"], r["offsets"])
    peers = sd.KmerShardPeers(ctx)
    cap = max(1 << 24, n_reads * 16)

    def barrier():
        dist.barrier(); torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank); sampler.start()
    for _ in range(args.warmup):
        out = sd.count_split_kmers_sharded(ctx, peers, cap=cap, sort=False, fetch=False)
    barrier()
    ms, stats = [], []
    for _ in range(args.steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = sd.count_split_kmers_sharded(ctx, peers, cap=cap, sort=False, fetch=False)   # kernel-only, like the single-GPU line:
        e1.record(stream)                                                                  # the survivors stay on the device
        barrier()
        ms.append(e0.elapsed_time(e1)); stats.append(out["stats"])
    e2e_ms = []
    out_bufs = None
    for _ in range(max(1, min(args.steps, args.e2e_steps))):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.upload_reads(r["bases"], r["quals"], r["offsets"])
        res = sd.count_split_kmers_sharded(ctx, peers, cap=cap, sort=False)
        e1.record(stream)
        barrier()
        e2e_ms.append(e0.elapsed_time(e1))
    clocks = sampler.stop()
    t = torch.tensor([sum(ms), sum(e2e_ms), float(len(res["kmers"]))], dtype=torch.float64, device=dev)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone(); dist.all_reduce(tsum)
    peers.close()
    if rank == 0:
        peaks = _peaks()
        st = stats[-1]
        total_ms, e2e_total = float(tmax[0]), float(tmax[1])
        bases = st["n_bases"]; n_kmers = bases - 16 * n_reads
        kern_ms = float(np.mean([s_["hist_ms"] + (s_["count_ms"] - s_["hist_ms"] - s_["scatter_ms"]) + s_["compact_ms"] for s_ in stats]))
        alg = 2 * 2 * bases + 4 * 4 * n_kmers + 16 * len(res["kmers"])          # per GPU, as in the single-GPU line
        line = {
            "metric": "reads counted/sec (Stage 1 split k-mer counting)", "value": world * n_reads * args.steps / (total_ms / 1e3), "unit": "reads/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 k-mers / u32 counts", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w.cfg['desc']} (k-mer counting of the reads)", "reads_per_gpu_per_step": n_reads,
                       "parallelism": f"reads sharded x{world}; hash partitions owned by rank p*{world}/8192; scatter into peer HBM over NVLink",
                       "l2": "inputs and partition entries larger than L2"},
            "stage_ms_rank0": {"hist": float(np.mean([s_["hist_ms"] for s_ in stats])),
                               "scatter+allgather+peer stores": float(np.mean([s_["scatter_ms"] for s_ in stats])),
                               "count": float(np.mean([s_["count_ms"] - s_["hist_ms"] - s_["scatter_ms"] for s_ in stats])),
                               "compact": float(np.mean([s_["compact_ms"] for s_ in stats]))},
            "survivors_total": int(tsum[2]),
            "roofline": {"bound": "hbm", "kernel": "kmer_hist + kmer_scatter1/2 + kmer_count_part (per GPU)",
                         "achieved": alg / (total_ms / args.steps / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "peak_source": peaks["source"],
                         "frac": alg / (total_ms / args.steps / 1e3) / 1e9 / peaks["hbm_gbs"], "traffic": None, "algorithmic_bytes": alg,
                         "note": "whole step per GPU incl. the two small all-gathers and two barriers"},
            "e2e": {"value": world * n_reads * len(e2e_ms) / (e2e_total / 1e3), "unit": "reads/s", "h2d_bytes_per_step": int(2 * bases + 8 * (n_reads + 1)),
                    "d2h_bytes_per_step": int(16 * len(res["kmers"])), "ms_per_step": e2e_total / len(e2e_ms)},
            "gpu_launches": int(sum(s_["kernel_launches"] for s_ in stats)), "clocks": clocks, "kernel_ms_rank0": kern_ms,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def run_kmercount(args, rank, world, local_rank):
    """--path kmercount: Stage 1 split-k-mer counting (K5, seq_parse::read_to_split_kmers) of the same reads.  Not the
    headline path; one JSON line with the same keys."""
    import torch
    from savont_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    if world > 1:
        return run_kmercount_sharded(args, rank, world, local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    w = _workload(args)
    n_reads = args.reads
    ctx = _lib.Context(local_rank, _lib.default_params(band=w.band))
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    r = w.reads(n_reads, alloc=_lib.pinned_array)
    ctx.upload_reads(r["bases"], r["quals"], r["offsets"])
    sampler = ClockSampler(local_rank); sampler.start()
    cap = max(1 << 24, n_reads * 16)
    for _ in range(args.warmup):
        out = ctx.count_split_kmers_resident(sort=False, cap=cap)
    torch.cuda.synchronize(dev)
    ms, stats = [], []
    for _ in range(args.steps):
        out = ctx.count_split_kmers_resident(sort=False, cap=cap)
        # kernel-only: table clear + counting + compaction (CUDA events inside the library, on this stream); the copy of the
        # surviving entries to the host belongs to the e2e number below
        ms.append(out["stats"]["count_ms"] + out["stats"]["compact_ms"]); stats.append(out["stats"])
    e2e_ms = []
    ctx.count_split_kmers(r["bases"], r["quals"], r["offsets"], sort=False, cap=cap)
    out_bufs = (_lib.pinned_array(8 * cap).view(np.uint64), _lib.pinned_array(8 * cap).view(np.uint32))   # page-locked outputs
    for _ in range(max(1, min(args.steps, args.e2e_steps))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        res = ctx.count_split_kmers(r["bases"], r["quals"], r["offsets"], sort=False, cap=cap, out=out_bufs)
        e1.record(stream); torch.cuda.synchronize(dev)
        e2e_ms.append(e0.elapsed_time(e1))
    clocks = sampler.stop()
    peaks = _peaks()
    st = stats[-1]
    kern_ms = float(np.mean([s["count_ms"] + s["compact_ms"] for s in stats]))
    bases = st["n_bases"]
    n_kmers = bases - 16 * n_reads
    # algorithmic bytes of the partitioned algorithm: bases + qualities read by the histogram and by the scatter pass,
    # 4-byte entries written by both scatter levels and read by the second level and by the counting kernel, the survivors
    alg = 2 * 2 * bases + 4 * 4 * n_kmers + 16 * len(out["kmers"])
    line = {
        "metric": "reads counted/sec (Stage 1 split k-mer counting)", "value": n_reads * args.steps / (sum(ms) / 1e3), "unit": "reads/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(ms) / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 k-mers / u32 counts", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w.cfg['desc']} (k-mer counting of the reads)", "reads_per_gpu_per_step": n_reads,
                   "l2": "inputs (%.1f GB) and partition entries (%.1f GB) larger than L2" % (2 * bases / 1e9, 4 * n_kmers / 1e9)},
        "kernel_ms": {"hist": float(np.mean([s["hist_ms"] for s in stats])), "scatter": float(np.mean([s["scatter_ms"] for s in stats])),
                      "count": float(np.mean([s["count_ms"] - s["hist_ms"] - s["scatter_ms"] for s in stats])),
                      "compact": float(np.mean([s["compact_ms"] for s in stats]))},
        "kmers_per_step": int(n_kmers), "entries_out": int(len(out["kmers"])), "n_distinct": int(st["n_distinct"]),
        "roofline": {"bound": "hbm", "kernel": "kmer_hist + kmer_scatter1/2 + kmer_count_part", "achieved": alg / (kern_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "peak_source": peaks["source"], "frac": alg / (kern_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                     "traffic": None, "algorithmic_bytes": alg,
                     "note": "partitioned counting: sequential traffic only; the single-table algorithm it replaced is profiled in "
                             "profiles/r01c_kmer_count_kernel_full.txt"},
        "e2e": {"value": n_reads * len(e2e_ms) / (sum(e2e_ms) / 1e3), "unit": "reads/s", "h2d_bytes_per_step": int(res["stats"]["h2d_bytes"]),
                "d2h_bytes_per_step": int(res["stats"]["d2h_bytes"]), "ms_per_step": sum(e2e_ms) / len(e2e_ms)},
        "gpu_launches": int(sum(s["kernel_launches"] for s in stats)), "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        ns = min(args.cpu_reads // 4, n_reads)
        t0 = time.perf_counter()
        want = O.count_split_kmers(r["bases"], r["quals"], r["offsets"][: ns + 1], n_threads=cores)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ns / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                                "sample": f"first {ns} reads: extraction on all cores, sort + run-length count on one (oracle/savont_oracle.c)"}
    print(json.dumps(line), flush=True)


def _pileup_clusters(w, n_cons, per, gen_reads):
    """Stage 4 input shaped like the reference's: per consensus (the homopolymer-compressed ASV) the first `per` reads of its
    cluster (reads drawn from that ASV), in cluster order"""
    r = w.reads(gen_reads)
    order = np.argsort(r["truth_asv"], kind="stable")
    ta = r["truth_asv"][order]
    keep = np.zeros(len(order), bool)
    start = np.searchsorted(ta, np.arange(n_cons)); stop = np.searchsorted(ta, np.arange(n_cons) + 1)
    for a in range(n_cons):
        keep[start[a]:min(stop[a], start[a] + per)] = True
    sel = order[keep]
    lens = (r["offsets"][1:] - r["offsets"][:-1]).astype(np.int64)
    ro = np.zeros(len(sel) + 1, np.uint64); ro[1:] = np.cumsum(lens[sel])
    idx = np.concatenate([np.arange(int(r["offsets"][i]), int(r["offsets"][i + 1])) for i in sel]) if len(sel) else np.zeros(0, np.int64)
    rb, rq = r["bases"][idx], r["quals"][idx]
    cons = []
    for a in range(n_cons):
        sq = np.asarray(w.asv_bases[int(w.asv_offsets[a]):int(w.asv_offsets[a + 1])])
        cons.append(sq[np.concatenate([[True], sq[1:] != sq[:-1]])].tobytes())
    cb = np.frombuffer(b"".join(cons), np.uint8); co = np.zeros(n_cons + 1, np.uint64); co[1:] = np.cumsum([len(c_) for c_ in cons])
    return cb, co, rb, rq, ro, ta[keep].astype(np.uint32)


def run_pileup(args, rank, world, local_rank):
    """--path pileup: Stage 4 pileups (alignment::generate_consensus_pileups) through savont_b200_consensus_pileups.  Not the
    headline path; one JSON line with the same keys.  The call takes host buffers, so every number here is end to end."""
    import torch
    from savont_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    if world > 1:
        if rank == 0:
            print(json.dumps({"metric": "reads piled up/sec (Stage 4 pileups)", "unavailable": "clusters are independent: run one process per GPU on disjoint clusters"}))
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    w = _workload(args)
    ctx = _lib.Context(local_rank, _lib.default_params(band=w.band))
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    cb, co, rb, rq, ro, rc = _pileup_clusters(w, w.n_asv, 250, args.reads)
    n_reads = len(rc)

    def pinned(x):
        y = _lib.pinned_array(max(x.nbytes, 1))[:x.nbytes].view(x.dtype); y[:] = x
        return y
    rb, rq, cb = pinned(rb), pinned(rq), pinned(cb)                  # page-locked inputs and entry buffer, as in the headline path
    cap = int(ro[-1]) + int((co[1:] - co[:-1])[rc].sum()) + 64
    ent_buf = _lib.pinned_array(12 * cap).view(_lib.PILEUP_ENTRY)
    sampler = ClockSampler(local_rank); sampler.start()
    for _ in range(args.warmup):
        out = ctx.consensus_pileups(cb, co, rb, rq, ro, rc, entries_out=ent_buf)
    torch.cuda.synchronize(dev)
    ms, stats = [], []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = ctx.consensus_pileups(cb, co, rb, rq, ro, rc, entries_out=ent_buf)
        e1.record(stream); torch.cuda.synchronize(dev)
        ms.append(e0.elapsed_time(e1)); stats.append(out["stats"])
    clocks = sampler.stop()
    peaks = _peaks()
    st = stats[-1]
    km = {k: float(np.mean([s_[k] for s_ in stats])) for k in ("sketch_ms", "align_ms", "filter_ms", "assign_ms")}
    # dominant stage: traceback.  Algorithmic bytes of align_tb_kernel: the per-column history (2 * NW words per column and
    # pair, written once, read back along the traceback path only -> counted once) + the CIGAR slots' ops
    nw = {16: 2, 32: 3, 64: 5, 192: 13}.get(w.band, 5)
    aligned = int((out["read_strand"] != 0xff).sum())
    lut = np.zeros(256, np.uint8); lut[[ord(x) for x in "CcGgTtUu"]] = [1, 1, 2, 2, 3, 3, 3, 3]
    codes = lut[rb]
    run_start = np.ones(len(codes), bool); run_start[1:] = codes[1:] != codes[:-1]; run_start[ro[:-1][ro[:-1] < len(codes)].astype(np.int64)] = True
    hcols = int(run_start.sum())                       # columns of the DP = compressed read bases (all reads are aligned here)
    alg = hcols * 2 * nw * 4
    val = n_reads * args.steps / (sum(ms) / 1e3)
    line = {
        "metric": "reads piled up/sec (Stage 4 pileups)", "value": val, "unit": "reads/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(ms) / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 (bit-parallel INT32 DP) / u8 columns", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w.n_asv} homopolymer-compressed consensus sequences x <= 250 reads of their clusters "
                               f"({n_reads} reads, {int(ro[-1])} bases), band {w.band}",
                   "l2": "every step re-uploads the reads from host memory (H2D inside the timed region)"},
        "stage_ms": {"upload+compress+pack": km["sketch_ms"], "nm_both_strands": km["align_ms"], "traceback": km["filter_ms"],
                     "columns": km["assign_ms"], "host": sum(ms) / args.steps - sum(km.values())},
        "entries_per_step": int(len(out["entries"])), "reads_aligned": aligned,
        "gcups": st["dp_cells"] / (km["align_ms"] / 1e3) / 1e9 if km["align_ms"] > 0 else None,
        "roofline": {"bound": "hbm", "kernel": "align_tb_kernel (traceback history)", "achieved": alg / (km["filter_ms"] / 1e3) / 1e9,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "peak_source": peaks["source"],
                     "frac": alg / (km["filter_ms"] / 1e3) / 1e9 / peaks["hbm_gbs"], "traffic": None, "algorithmic_bytes": alg,
                     "note": "stage time includes CIGAR compaction and the D2H copies of the sub-batches"},
        "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": int(st["h2d_bytes"]), "d2h_bytes_per_step": int(st["d2h_bytes"]),
                "ms_per_step": sum(ms) / args.steps},
        "gpu_launches": int(sum(s_["kernel_launches"] for s_ in stats)), "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        from concurrent.futures import ThreadPoolExecutor
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        nc = min(w.n_asv, max(cores, 2 * cores))
        r_end = np.searchsorted(rc, np.arange(nc + 1))

        def one(a):
            i0, i1 = int(r_end[a]), int(r_end[a + 1])
            b0, b1 = int(ro[i0]), int(ro[i1])
            return O.consensus_pileups(cb[int(co[a]):int(co[a + 1])], [0, int(co[a + 1] - co[a])], rb[b0:b1], rq[b0:b1], ro[i0:i1 + 1] - ro[i0],
                                       np.zeros(i1 - i0, np.uint32), band=w.band)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            parts = list(ex.map(one, range(nc)))
        dt = time.perf_counter() - t0
        ns = int(r_end[nc])
        same = all(np.array_equal(parts[a]["read_nm"], out["read_nm"][int(r_end[a]):int(r_end[a + 1])]) for a in range(nc))
        line["cpu_baseline"] = {"value": ns / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                                "sample": f"first {nc} clusters ({ns} reads), one cluster per thread like the reference's par_iter over consensuses "
                                          "(oracle/savont_oracle.c)", "gpu_matches_oracle_on_sample": bool(same)}
    print(json.dumps(line), flush=True)


def run_cluster(args, rank, world, local_rank):
    """--path cluster: Stage 2 greedy k-mer clustering (asv_cluster::cluster_reads_by_kmers) through
    savont_b200_cluster_reads_by_kmers.  Not the headline path; one JSON line with the same keys.  The call takes host buffers (the
    reads' minimizer k-mers), so every number here is end to end."""
    import torch
    from savont_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    if world > 1:
        if rank == 0:
            print(json.dumps({"metric": "reads clustered/sec (Stage 2 greedy k-mer clustering)", "unavailable": "sequential greedy order: single GPU"}))
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    w = _workload(args)
    ctx = _lib.Context(local_rank, _lib.default_params(band=w.band))
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    n_reads = args.reads
    r = w.reads(n_reads)
    ctx.load_snpmers(w.snpmers)
    kms, offs = [], [np.zeros(1, np.uint64)]
    for b0 in range(0, n_reads, 200_000):                          # TwinRead::minimizer_kmers() of the reads, from the sketch kernel
        b1 = min(n_reads, b0 + 200_000)
        sk = ctx.sketch(r["bases"], r["quals"], r["offsets"][b0:b1 + 1], want_codes=False)
        kms.append(np.asarray(sk["mini_kmer"], np.uint64)); offs.append(np.asarray(sk["mini_offsets"][1:], np.uint64) + offs[-1][-1])
    km = _lib.pinned_array(8 * sum(len(x) for x in kms)).view(np.uint64); km[:] = np.concatenate(kms)
    off = np.concatenate(offs)
    del kms, offs, r
    sampler = ClockSampler(local_rank); sampler.start()
    for _ in range(args.warmup):
        out = ctx.cluster_reads_by_kmers(km, off)
    torch.cuda.synchronize(dev)
    ms, stats = [], []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = ctx.cluster_reads_by_kmers(km, off)
        e1.record(stream); torch.cuda.synchronize(dev)
        ms.append(e0.elapsed_time(e1)); stats.append(out["stats"])
    clocks = sampler.stop()
    peaks = _peaks()
    st = stats[-1]
    km_ms = {k: float(np.mean([s_[k] for s_ in stats])) for k in ("sketch_ms", "filter_ms", "align_ms")}
    val = n_reads * args.steps / (sum(ms) / 1e3)
    # dominant kernel: rep_query_kernel (20-way merge of the read's buckets by representative id, two passes).  Algorithmic bytes:
    # every bucket entry (4 B) is read twice + the per-read result record
    alg = 2 * 4 * int(st["dp_cells"]) + 248 * n_reads
    line = {
        "metric": "reads clustered/sec (Stage 2 greedy k-mer clustering)", "value": val, "unit": "reads/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(ms) / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 k-mers / u32 counts, f64 threshold on the host", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w.cfg['desc']} (minimizer k-mers of the reads: {len(km)} k-mers)",
                   "threshold": 0.95, "l2": "every step re-uploads the k-mers from host memory (H2D inside the timed region)"},
        "stage_ms": {"upload+signatures+sorted_sets": km_ms["sketch_ms"], "bucket_query": km_ms["filter_ms"], "in_block_pairs": km_ms["align_ms"],
                     "host": sum(ms) / args.steps - sum(km_ms.values())},
        "n_reps": int(out["n_reps"]), "blocks": int(st["n_batches"]), "in_block_pairs": int(st["n_pairs"]), "bucket_entries": int(st["dp_cells"]),
        "roofline": {"bound": "hbm", "kernel": "rep_query_kernel", "achieved": alg / (km_ms["filter_ms"] / 1e3) / 1e9, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "peak_source": peaks["source"], "frac": alg / (km_ms["filter_ms"] / 1e3) / 1e9 / peaks["hbm_gbs"],
                     "traffic": None, "algorithmic_bytes": alg,
                     "note": "latency bound, not bandwidth bound: a warp walks ~hundreds of bucket entries serially (REDUX min over the 20 list "
                             "heads per step) and a block holds only as many reads as the sequential greedy order allows (see DESIGN.md K7)"},
        "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": int(st["h2d_bytes"]), "d2h_bytes_per_step": int(st["d2h_bytes"]),
                "ms_per_step": sum(ms) / args.steps},
        "gpu_launches": int(sum(s_["kernel_launches"] for s_ in stats)), "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        ns = min(args.cpu_reads // 2, n_reads)
        t0 = time.perf_counter()
        want = O.cluster_reads_by_kmers(km[:int(off[ns])], off[:ns + 1])
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ns / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                                "sample": f"first {ns} reads (signatures and sorted sets on all cores, the greedy loop on one, like the reference; the cost "
                                          "per read grows with the number of representatives, so a prefix flatters the CPU)",
                                "gpu_matches_oracle_on_sample": bool(np.array_equal(want, out["rep_of"][:ns]))}
    print(json.dumps(line), flush=True)


def run_snpcluster(args, rank, world, local_rank):
    """--path snpcluster: the greedy SNPmer clustering of Stage 3 inside the k-mer clusters (asv_cluster::cluster_reads_by_snpmers)
    through savont_b200_cluster_reads_by_snpmers.  The k-mer clusters of the synthetic workload are its genera (what Stage 2 separates:
    the roots are unrelated sequences), the reads of a cluster in read order.  Host buffers in, so end to end."""
    import torch
    from savont_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    if world > 1:
        if rank == 0:
            print(json.dumps({"metric": "reads clustered/sec (Stage 3 greedy SNPmer clustering)", "unavailable": "k-mer clusters are independent: shard them, one process per GPU"}))
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    w = _workload(args)
    ctx = _lib.Context(local_rank, _lib.default_params(band=w.band))
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    ctx.load_snpmers(w.snpmers)
    n_reads = args.reads
    r = w.reads(n_reads)
    kms, offs = [], [np.zeros(1, np.uint64)]
    for b0 in range(0, n_reads, 200_000):
        b1 = min(n_reads, b0 + 200_000)
        sk = ctx.sketch(r["bases"], r["quals"], r["offsets"][b0:b1 + 1], want_codes=False)
        kms.append(np.asarray(sk["snp_kmer"], np.uint64)); offs.append(np.asarray(sk["snp_offsets"][1:], np.uint64) + offs[-1][-1])
    rk = np.concatenate(kms); ro = np.concatenate(offs)
    n_var = w.cfg["n_var"]
    genus = (r["truth_asv"] // n_var).astype(np.int64)
    order = np.argsort(genus, kind="stable").astype(np.uint32)                    # cluster by cluster, read order inside
    sizes = np.bincount(genus, minlength=w.n_asv // n_var)
    co = np.zeros(len(sizes) + 1, np.uint64); co[1:] = np.cumsum(sizes)
    del kms, offs
    sampler = ClockSampler(local_rank); sampler.start()
    for _ in range(max(1, args.warmup)):
        out = ctx.cluster_reads_by_snpmers(rk, ro, order, co)
    torch.cuda.synchronize(dev)
    ms, stats = [], []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = ctx.cluster_reads_by_snpmers(rk, ro, order, co)
        e1.record(stream); torch.cuda.synchronize(dev)
        ms.append(e0.elapsed_time(e1)); stats.append(out["stats"])
    clocks = sampler.stop()
    kern = float(np.mean([s_["filter_ms"] for s_ in stats]))
    val = n_reads * args.steps / (sum(ms) / 1e3)
    peaks = _peaks()
    alg = 8 * len(rk) + 8 * n_reads
    line = {
        "metric": "reads clustered/sec (Stage 3 greedy SNPmer clustering inside the k-mer clusters)", "value": val, "unit": "reads/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": sum(ms) / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 k-mers / u32 counts", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w.cfg['desc']}: {n_reads} reads ({len(rk)} SNPmers) in {len(sizes)} k-mer clusters (largest "
                               f"{int(sizes.max())} reads), {out['n_representatives']} representatives", "l2": "every step re-uploads the SNPmer lists from host memory"},
        "stage_ms": {"upload": float(np.mean([s_["h2d_ms"] for s_ in stats])), "cluster_kernel": kern},
        "n_representatives": int(out["n_representatives"]),
        "roofline": {"bound": "hbm", "kernel": "snp_round_kernel + snp_resolve_kernel", "achieved": alg / (kern / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "peak_source": peaks["source"], "frac": alg / (kern / 1e3) / 1e9 / peaks["hbm_gbs"], "traffic": None, "algorithmic_bytes": alg,
                     "note": "rounds (one per representative of the cluster with the most representatives; every round streams the lists of the "
                             "reads behind the new representatives) + one in-order pass per cluster for the tie-breaking sizes; the algorithmic "
                             "bytes count every list once"},
        "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": int(stats[-1]["h2d_bytes"]), "d2h_bytes_per_step": int(stats[-1]["d2h_bytes"]),
                "ms_per_step": sum(ms) / args.steps},
        "gpu_launches": int(sum(s_["kernel_launches"] for s_ in stats)), "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        from concurrent.futures import ThreadPoolExecutor
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        def one(c):                                                               # the reference walks the k-mer clusters in a par_iter
            mem = order[int(co[c]):int(co[c + 1])]
            return mem, O.cluster_reads_by_snpmers(rk, ro, mem)
        with ThreadPoolExecutor(max_workers=cores) as ex:
            res = list(ex.map(one, range(len(sizes))))
        dt = time.perf_counter() - t0
        same = all(np.array_equal(out["rep_of"][mem], w_[mem]) for mem, w_ in res)
        line["cpu_baseline"] = {"value": n_reads / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                                "sample": f"all {n_reads} reads, one thread per k-mer cluster like the reference's par_iter (oracle/savont_oracle.c)",
                                "gpu_matches_oracle_on_sample": bool(same)}
    print(json.dumps(line), flush=True)


def run_reassign(args, rank, world, local_rank):
    """--path reassign: the read-vs-consensus SNPmer comparison of Stage 3 reclustering (asv_cluster::reassign_reads_to_best_cluster)
    through savont_b200_snpmer_compare.  Not the headline path; one JSON line with the same keys.  Host buffers in, so end to end."""
    import torch
    from savont_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    if world > 1:
        if rank == 0:
            print(json.dumps({"metric": "reads reassigned/sec (Stage 3 reclustering)", "unavailable": "reads are independent: shard them, one process per GPU"}))
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    w = _workload(args)
    ctx = _lib.Context(local_rank, _lib.default_params(band=w.band))
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    ctx.load_snpmers(w.snpmers)
    n_reads = args.reads
    r = w.reads(n_reads)
    sa = ctx.sketch(w.asv_bases, None, w.asv_offsets, want_codes=False)          # the clusters' consensus SNPmers: those of the ASVs
    ck, co = np.asarray(sa["snp_kmer"], np.uint64), np.asarray(sa["snp_offsets"], np.uint64)
    kms, offs = [], [np.zeros(1, np.uint64)]
    for b0 in range(0, n_reads, 200_000):
        b1 = min(n_reads, b0 + 200_000)
        sk = ctx.sketch(r["bases"], r["quals"], r["offsets"][b0:b1 + 1], want_codes=False)
        kms.append(np.asarray(sk["snp_kmer"], np.uint64)); offs.append(np.asarray(sk["snp_offsets"][1:], np.uint64) + offs[-1][-1])
    rk = _lib.pinned_array(8 * max(1, sum(len(x) for x in kms))).view(np.uint64)[:sum(len(x) for x in kms)]; rk[:] = np.concatenate(kms)
    ro = np.concatenate(offs)
    del kms, offs, r
    sampler = ClockSampler(local_rank); sampler.start()
    for _ in range(args.warmup):
        out = ctx.snpmer_compare(ck, co, rk, ro)
    torch.cuda.synchronize(dev)
    ms, stats = [], []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = ctx.snpmer_compare(ck, co, rk, ro)
        e1.record(stream); torch.cuda.synchronize(dev)
        ms.append(e0.elapsed_time(e1)); stats.append(out["stats"])
    clocks = sampler.stop()
    peaks = _peaks()
    st = stats[-1]
    kern = float(np.mean([s_["filter_ms"] for s_ in stats]))
    val = n_reads * args.steps / (sum(ms) / 1e3)
    # algorithmic bytes of the kernel: the markers (8 B) + per read the three result words; the index is L2 resident
    alg = 8 * len(rk) + 12 * n_reads
    line = {
        "metric": "reads reassigned/sec (Stage 3 reclustering, read-vs-consensus SNPmer comparison)", "value": val, "unit": "reads/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(ms) / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 k-mers / u16 counts", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w.cfg['desc']}: SNPmers of the reads ({len(rk)}) against {len(co) - 1} cluster consensus lists "
                               f"({len(ck)} consensus SNPmers)", "l2": "every step re-uploads the markers from host memory"},
        "stage_ms": {"index+upload": float(np.mean([s_["h2d_ms"] for s_ in stats])), "compare_kernel": kern, "host": sum(ms) / args.steps - float(np.mean([s_["total_ms"] for s_ in stats]))},
        "pair_comparisons_per_step": int(st["n_pairs"]),
        "roofline": {"bound": "hbm", "kernel": "snpmer_compare_kernel", "achieved": alg / (kern / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "peak_source": peaks["source"], "frac": alg / (kern / 1e3) / 1e9 / peaks["hbm_gbs"], "traffic": None, "algorithmic_bytes": alg,
                     "note": "bound by shared-memory work, not HBM: per read a dense row of one word per cluster is cleared, updated by atomics "
                             "and reduced"},
        "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": int(st["h2d_bytes"]), "d2h_bytes_per_step": int(st["d2h_bytes"]),
                "ms_per_step": sum(ms) / args.steps},
        "gpu_launches": int(sum(s_["kernel_launches"] for s_ in stats)), "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        ns = min(args.cpu_reads // 5, n_reads)
        t0 = time.perf_counter()
        want = O.snpmer_compare(ck, co, rk[:int(ro[ns])], ro[:ns + 1])
        dt = time.perf_counter() - t0
        same = all(np.array_equal(want[k_], out[k_][:ns]) for k_ in ("best_cons", "best_mismatches", "best_matches"))
        line["cpu_baseline"] = {"value": ns / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                                "sample": f"first {ns} reads, OpenMP over reads, one hash map per cluster probed with every marker like the reference "
                                          "(oracle/savont_oracle.c)", "gpu_matches_oracle_on_sample": bool(same)}
    print(json.dumps(line), flush=True)


def _bind_to_gpu_numa_node(gpu_index):
    """pin this rank (and therefore the pinned host buffers it first-touches) to the CPU cores next to its GPU, so that
    eight ranks do not pull their H2D traffic across the socket interconnect"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_streamed(args, rank, world, local_rank):
    """--total-reads T > --reads R: a named configuration whose read set does not fit one device batch (a batch is limited to 4 Gi
    bases: config 4 = 2M reads x 4.5 kb, config 5 = 10M reads) is streamed as ceil(T / R) chunks of the deterministic read stream.
    Every chunk is uploaded, then mapped resident (device-timed, one timed pass per chunk; the counters accumulate over the chunks,
    so the final counter vector / per-sample matrix is that of all T reads); `value` = T / sum of the chunks' device times.  e2e legs
    run on the first chunk; the oracle checks the first `--cpu-reads` reads of the LAST chunk (every output) and the per-sample
    matrix of that chunk.  N = 1 only."""
    import math
    import torch
    from savont_b200 import _lib
    if world > 1:
        raise SystemExit("bench.py --total-reads: single GPU only")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    w = _workload(args)
    R, T = args.reads, args.total_reads
    chunks = math.ceil(T / R)
    ctx = _lib.Context(local_rank, _lib.default_params(band=w.band))
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets); ctx.set_batch_reads(args.batch_reads)

    class Reuse:                                     # pinned buffers handed out again for every chunk
        def __init__(self): self.bufs, self.i = [], 0
        def reset(self): self.i = 0
        def __call__(self, nbytes):
            if self.i == len(self.bufs) or len(self.bufs[self.i]) < nbytes:
                buf = _lib.pinned_array(int(nbytes * 1.02) + 4096)
                if self.i == len(self.bufs): self.bufs.append(buf)
                else: self.bufs[self.i] = buf
            self.i += 1
            return self.bufs[self.i - 1][:nbytes]
    alloc = Reuse()
    sampler = ClockSampler(local_rank); sampler.start()
    step_ms, stats, t_gen = [], [], 0.0
    e2e = None; last = None
    for ci in range(chunks):
        n = min(R, T - ci * R)
        alloc.reset()
        t0 = time.perf_counter()
        r = w.reads(n, first=ci * R, alloc=alloc)
        t_gen += time.perf_counter() - t0
        fidx = r["file_idx"] if w.n_samples > 1 else None
        ctx.upload_reads(r["bases"], r["quals"], r["offsets"], file_idx=fidx, n_samples=w.n_samples)
        if ci == 0:
            for _ in range(max(args.warmup, 3)):
                ctx.map_resident()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = ctx.map_resident(zero_counters=(ci == 0))
        e1.record(stream)
        torch.cuda.synchronize(dev)
        step_ms.append(e0.elapsed_time(e1)); stats.append(st)
        if ci == 0 and not args.no_twin:
            from savont_b200 import twin
            tw = twin.twin_reads_from_snpmers(ctx, r["bases"], r["quals"], r["offsets"], file_idx=fidx, kmer48=True)
            heap = twin.scatter_to_heap(tw)
            ctx.map_twin_read_ptrs(heap, n_samples=w.n_samples)
            ms = []
            for _ in range(2):
                torch.cuda.synchronize(dev)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(stream); res = ctx.map_twin_read_ptrs(heap, n_samples=w.n_samples); a1.record(stream)
                torch.cuda.synchronize(dev); ms.append(a0.elapsed_time(a1))
            e2e = {"value": n * len(ms) / (sum(ms) / 1e3), "unit": "reads/s", "h2d_bytes_per_step": int(res["stats"]["h2d_bytes"]),
                   "d2h_bytes_per_step": int(res["stats"]["d2h_bytes"]), "steps": len(ms), "ms_per_step": sum(ms) / len(ms),
                   "gather_ms": res["stats"]["gather_ms"], "chunk_reads": n,
                   "input": "first chunk as &[TwinRead] heap blocks through savont_b200_map_twin_read_ptrs (gather inside the timed region)"}
            del heap, tw
            # the e2e call reset the device counters: redo the chunk's resident pass so that the accumulated counters cover all chunks
            ctx.upload_reads(r["bases"], r["quals"], r["offsets"], file_idx=fidx, n_samples=w.n_samples)
            ctx.map_resident(zero_counters=True)
        if ci == chunks - 1:
            last = (r, fidx, ctx.fetch_resident())
    clocks = sampler.stop()
    total_ms = sum(step_ms)
    r, fidx, got = last
    line = {
        "metric": "reads aligned/sec (read-vs-ASV assignment path)", "value": T / (total_ms / 1e3), "unit": "reads/s", "n_gpus": 1,
        "steps": chunks, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / chunks, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 (bit-parallel INT32 DP) / f64 thresholds", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w.cfg['desc']}", "n_asv": w.n_asv, "band": w.band, "reads_total": T, "reads_per_step": R,
                   "chunks": chunks, "n_samples": w.n_samples, "l2": "every chunk's inputs are GBs, larger than L2"},
        "kernel_ms_total": {k: float(sum(s_[k + "_ms"] for s_ in stats)) for k in ("sketch", "filter", "align", "assign")},
        "pairs_total": int(sum(s_["n_pairs"] for s_ in stats)), "dp_cells_total": int(sum(s_["dp_cells"] for s_ in stats)),
        "gcups": sum(s_["dp_cells"] for s_ in stats) / (sum(s_["align_ms"] for s_ in stats) / 1e3) / 1e9,
        "assigned_total": int(got["totals"][0]), "filtered_total": int(got["totals"][1]),
        "per_sample_singleton_sums": [int(x) for x in got["sample_singleton"].sum(axis=1)],
        "e2e": e2e, "gpu_launches": int(sum(s_["kernel_launches"] for s_ in stats)), "clocks": clocks, "gen_s": t_gen,
    }
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        ns = min(args.cpu_reads, len(r["offsets"]) - 1)
        orc = O.Oracle(O.default_params(band=w.band), w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
        t0 = time.perf_counter()
        want = orc.map_reads(r["bases"], r["quals"], r["offsets"][: ns + 1], file_idx=None if fidx is None else fidx[:ns], n_samples=w.n_samples,
                             n_threads=cores)
        dt = time.perf_counter() - t0
        sub = ctx.map_reads(r["bases"], r["quals"], r["offsets"][: ns + 1], file_idx=None if fidx is None else fidx[:ns], n_samples=w.n_samples)
        keys = ("status", "best_nm", "class_offsets", "class_asv", "unambig", "ambig", "leq10", "totals", "sample_singleton")
        same = all(np.array_equal(np.asarray(want[k]), np.asarray(sub[k])) for k in keys)
        same_res = (np.array_equal(want["status"], got["status"][:ns]) and np.array_equal(want["best_nm"], got["best_nm"][:ns])
                    and np.array_equal(want["class_asv"], got["class_asv"][:int(got["class_offsets"][ns])]))
        line["cpu_baseline"] = {"value": ns / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                                "sample": f"first {ns} reads of the last chunk, OpenMP over reads (oracle/savont_oracle.c)",
                                "gpu_matches_oracle_on_sample": bool(same and same_res)}
    print(json.dumps(line), flush=True)


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from savont_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    # a CPU-side group: the ranks that sit out the single-process leg wait on it without keeping a kernel spinning on their GPU
    gloo = dist.new_group(backend="gloo") if world > 1 else None
    numa = _bind_to_gpu_numa_node(local_rank) if world > 1 else None

    w = _workload(args)
    n_reads = args.reads
    ctx = _lib.Context(local_rank, _lib.default_params(band=w.band))
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    ctx.load_snpmers(w.snpmers)
    ctx.load_asvs(w.asv_bases, w.asv_offsets)
    ctx.set_batch_reads(args.batch_reads)

    # this rank's shard: reads [rank*R, (rank+1)*R) of the deterministic stream, generated into pinned memory
    t_gen = time.perf_counter()
    r = w.reads(n_reads, first=rank * n_reads, alloc=_lib.pinned_array)
    t_gen = time.perf_counter() - t_gen
    fidx = r["file_idx"] if w.n_samples > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    class CountsView:                      # the library's device counter vector as a torch tensor (no copy)
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}

    def allreduce_counts():
        if world > 1:
            ptr, n = ctx.counts_dev()
            t = torch.as_tensor(CountsView(ptr, n), device=dev)
            dist.all_reduce(t)             # one NCCL all-reduce of [unambig|ambig|leq10|totals|singleton]

    # ---------------- kernel-only arm: reads resident in HBM ----------------
    ctx.upload_reads(r["bases"], r["quals"], r["offsets"], file_idx=fidx, n_samples=w.n_samples)
    sampler = ClockSampler(local_rank)
    sampler.start()                       # runs through warm-up, the timed steps and the e2e steps (all under load)
    for _ in range(args.warmup):
        ctx.map_resident(); allreduce_counts()
    barrier()
    step_ms, stats = [], []
    for _ in range(args.steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = ctx.map_resident()
        allreduce_counts()
        e1.record(stream)
        barrier()
        step_ms.append(e0.elapsed_time(e1)); stats.append(st)
    dev_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    total_ms = float(dev_ms.item())
    got = ctx.fetch_resident()

    # ---------------- the same pass WITH the alignment record of every read (strand, coordinates, identity, CIGAR) ----------------
    # north_star names identity / CIGAR per read; Stage 7 itself consumes only NM (src/alignment.rs:1862), so the headline pass leaves the
    # traceback out.  This arm adds savont_b200_best_hits_resident: a second, history-keeping DP pass over the best pair of every read and
    # the walk, the records fetched to host memory.  One pass, rank 0, host wall clock.
    cigar_line = None
    if rank == 0 and not args.no_cigar:
        cig_buf = _lib.pinned_array(4 * (48 * n_reads + 1024)).view(np.uint32)          # the caller's CIGAR buffer, page-locked
        ctx.map_resident()
        ctx.best_hits_resident(cigar_out=cig_buf)                                        # warm-up (workspaces)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        ctx.map_resident()
        bh = ctx.best_hits_resident(cigar_out=cig_buf)
        dt = time.perf_counter() - t0
        hit = bh["best_asv"] != 0xffffffff
        cigar_line = {"value": n_reads / dt, "unit": "reads/s", "ms_per_step": 1e3 * dt, "reads_with_record": int(hit.sum()),
                      "cigar_ops": int(len(bh["cigar"])), "mean_identity": float(np.nanmean(bh["identity"])) if hit.any() else None,
                      "what": "map_resident + best_hits_resident on the resident reads of one GPU: NM pass, then a checkpointing DP pass "
                              "+ block-wise recomputation and walk for the best pair of every read; records copied to host arrays, CIGARs to a page-locked buffer"}
        del bh

    # ---------------- e2e arm: host buffers through the C ABI ----------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    ctx.map_reads(r["bases"], r["quals"], r["offsets"], file_idx=fidx, n_samples=w.n_samples)   # warm-up (allocations)
    e2e_ms, e2e_stats = [], None
    for _ in range(e2e_steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        res = ctx.map_reads(r["bases"], r["quals"], r["offsets"], file_idx=fidx, n_samples=w.n_samples)
        allreduce_counts()
        e1.record(stream)
        barrier()
        e2e_ms.append(e0.elapsed_time(e1))
        e2e_stats = res["stats"]
    e2e_t = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_total_ms = float(e2e_t.item())

    # ---------------- e2e arm 2: the reads as &[TwinRead] (what Stage 7 is handed by the reference) ----------------
    # (a) `e2e` (headline): a Vec<TwinRead> as it lies in the reference's memory -- every read's dna_seq words and snpmer_kmers in
    #     their own block of ordinary pageable host memory (twin.scatter_to_heap), one pointer per read -- through
    #     savont_b200_map_twin_read_ptrs: the library's host threads gather the blocks into page-locked staging INSIDE the timed
    #     region, batch i+1 under the kernels of batch i.
    # (b) `e2e_twin_contiguous`: the same fields already concatenated in pinned memory (savont_b200_map_twin_reads), i.e. without
    #     the gather -- the figure round 1 reported as e2e.
    twin_line = ptr_line = twin_resident = None
    sample_check = None
    if not args.no_twin:
        from savont_b200 import twin
        tw = twin.twin_reads_from_snpmers(ctx, r["bases"], r["quals"], r["offsets"], file_idx=fidx, kmer48=True, alloc=_lib.pinned_array)
        heap = twin.scatter_to_heap(tw, seed=1 + rank)

        ctx.reuse_outputs = True            # the caller's output arrays are allocated once and reused, like its input buffers
        twin_cap = 2 * n_reads + 1024       # class_asv capacity: the library reports ERR_CAPACITY if a read set ever needs more

        def timed(call):
            res = call()                                                                        # warm-up (allocations)
            ms, stats_ = [], None
            for _ in range(e2e_steps):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                res = call()
                allreduce_counts()
                e1.record(stream)
                barrier()
                ms.append(e0.elapsed_time(e1)); stats_ = res["stats"]
            t = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return res, float(t.item()), stats_

        def line_of(res, tot_ms, st_, what):
            same = all(np.array_equal(res[k], got[k]) for k in ("status", "best_nm", "class_offsets", "class_asv"))
            return {"value": world * n_reads * e2e_steps / (tot_ms / 1e3), "unit": "reads/s",
                    "h2d_bytes_per_step": int(st_["h2d_bytes"]), "d2h_bytes_per_step": int(st_["d2h_bytes"]),
                    "steps": e2e_steps, "ms_per_step": tot_ms / e2e_steps, "h2d_ms": st_["h2d_ms"], "gather_ms": st_["gather_ms"],
                    "kernel_ms": {"sketch": st_["sketch_ms"], "filter": st_["filter_ms"], "align": st_["align_ms"]},
                    "input": what, "same_result_as_ascii_entry": bool(same)}

        res, tot, st_ = timed(lambda: ctx.map_twin_reads(tw, n_samples=w.n_samples, class_cap=twin_cap))
        twin_line = line_of(res, tot, st_, "TwinRead.dna_seq words + TwinRead.snpmer_kmers (Kmer48) already concatenated in pinned host memory, "
                                            "savont_b200_map_twin_reads")
        gather_threads = max(1, (os.cpu_count() or 1) // world)                                 # the ranks share the host's cores
        ctx.set_host_threads(gather_threads)
        res, tot, st_ = timed(lambda: ctx.map_twin_read_ptrs(heap, n_samples=w.n_samples, class_cap=twin_cap))
        ptr_line = line_of(res, tot, st_, "&[TwinRead] as it lies in the reference's memory: one pageable heap block per read for dna_seq and for "
                                          "snpmer_kmers, one pointer per read, savont_b200_map_twin_read_ptrs (host-thread gather into pinned "
                                          "staging inside the timed region)")
        ptr_line["host_threads"] = gather_threads
        res = {k: (np.array(v) if isinstance(v, np.ndarray) else v) for k, v in res.items()}   # detach from the reused output arrays
        ctx.reuse_outputs = False
        # the kernel pipeline once more with the TwinReads RESIDENT (what Stage 7 is handed; the sketch then has no SNPmer detection, no
        # quality gate and 2-bit input): informational, `value` stays the FASTQ-resident form that round 1 reported
        ctx.upload_twin_reads(tw, n_samples=w.n_samples)
        for _ in range(2):
            ctx.map_resident()
        tr_ms, tr_st = [], None
        for _ in range(args.steps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); tr_st = ctx.map_resident(); allreduce_counts(); e1.record(stream)
            barrier()
            tr_ms.append(e0.elapsed_time(e1))
        t = torch.tensor([sum(tr_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gtw = ctx.fetch_resident()
        twin_resident = {"value": world * n_reads * args.steps / (float(t.item()) / 1e3), "unit": "reads/s", "ms_per_step": float(t.item()) / args.steps,
                         "kernel_ms": {"sketch": tr_st["sketch_ms"], "filter": tr_st["filter_ms"], "align": tr_st["align_ms"], "assign": tr_st["assign_ms"]},
                         "same_result_as_ascii_entry": bool(all(np.array_equal(gtw[k], got[k]) for k in ("status", "best_nm", "class_offsets", "class_asv"))),
                         "input": "TwinRead fields resident in HBM (savont_b200_upload_twin_reads + savont_b200_map_resident)"}

    clocks = sampler.stop()

    # ---------------- correctness legs (untimed) ----------------
    # N = 1: the first `cpu_reads` reads once more through the C ABI, every output and counter against the oracle (below).
    # N > 1: every rank maps the first cpu_reads / N reads of ITS shard and the counter vectors are all-reduced over NCCL; rank 0
    # then maps the concatenation of those samples on its one GPU (and through the oracle): the reduced vector must be identical.
    multi_check = None
    ns_s = min(args.cpu_reads // max(world, 1), n_reads)
    sub_got = ctx.map_reads(r["bases"], r["quals"], r["offsets"][: ns_s + 1], file_idx=None if fidx is None else fidx[:ns_s], n_samples=w.n_samples)
    if world > 1:
        ptr_c, n_c = ctx.counts_dev()
        vec = torch.as_tensor(CountsView(ptr_c, n_c), device=dev)
        local = vec.clone()
        dist.all_reduce(vec)
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        reduced = vec.cpu().numpy().astype(np.uint64)
        sum_of_ranks = sum(g.cpu().numpy().astype(np.uint64) for g in gathered)
        if rank == 0:
            parts = [w.reads(ns_s, first=q * n_reads) for q in range(world)]
            cat = {k: np.concatenate([p_[k] for p_ in parts]) for k in ("bases", "quals")}
            offs = np.zeros(world * ns_s + 1, np.uint64); base = 0; pos = 1
            for p_ in parts:
                o = p_["offsets"].astype(np.uint64)
                offs[pos:pos + ns_s] = o[1:] + np.uint64(base); base += int(o[-1]); pos += ns_s
            fcat = np.concatenate([p_["file_idx"] for p_ in parts]) if fidx is not None else None
            one = ctx.map_reads(cat["bases"], cat["quals"], offs, file_idx=fcat, n_samples=w.n_samples)
            A = w.n_asv
            single_vec = np.concatenate([one["unambig"], one["ambig"], one["leq10"], one["totals"], one["sample_singleton"].reshape(-1)]).astype(np.uint64)
            multi_check = {"reads_per_rank": ns_s, "allreduce_equals_sum_of_ranks": bool(np.array_equal(reduced, sum_of_ranks)),
                           "multi_gpu_matches_single": bool(np.array_equal(reduced, single_vec)), "counter_vector_len": int(len(reduced)),
                           "assigned": int(reduced[3 * A])}
            if not args.no_cpu_baseline:
                from oracle import oracle as O
                orc = O.Oracle(O.default_params(band=w.band), w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
                want = orc.map_reads(cat["bases"], cat["quals"], offs, file_idx=fcat, n_samples=w.n_samples, n_threads=os.cpu_count() or 1)
                ovec = np.concatenate([want["unambig"], want["ambig"], want["leq10"], want["totals"], want["sample_singleton"].reshape(-1)]).astype(np.uint64)
                multi_check["multi_gpu_matches_oracle"] = bool(np.array_equal(reduced, ovec))
    # ---------------- one process, all N GPUs (savont_b200_create_multi): `savont asv` is a single process ----------------
    # rank 0 alone drives every GPU of the box through the multi-GPU handle while the other ranks wait on the CPU-side group.
    # Every device's shard is the rank-0 shard again (N pointer lists into one 1M-read heap arena): N x the work, known counters.
    sp_line = None
    if world > 1 and not args.no_twin and not args.no_single_process:
        ctx.close()
        torch.cuda.synchronize(dev)
        dist.barrier(group=gloo)
        if rank == 0:
            m = _lib.MultiContext(list(range(world)), _lib.default_params(band=w.band))
            m.load_snpmers(w.snpmers); m.load_asvs(w.asv_bases, w.asv_offsets); m.set_batch_reads(args.batch_reads)
            # every device gets its own copy of the heap arenas (two devices gathering the SAME blocks at the same time measured 35 % slower
            # than two gathering different ones, gpurun r02x: an artefact no real read set has)
            heaps = [heap]
            for _ in range(1, world):
                wa, sa = heap["_arenas"]
                w2, s2 = wa.copy(), sa.copy()
                heaps.append(dict(dna_ptrs=(heap["dna_ptrs"].astype(np.int64) - wa.ctypes.data + w2.ctypes.data).astype(np.uint64),
                                  snpmer_ptrs=(heap["snpmer_ptrs"].astype(np.int64) - sa.ctypes.data + s2.ctypes.data).astype(np.uint64), _arenas=(w2, s2)))
            big = dict(dna_ptrs=np.concatenate([h_["dna_ptrs"] for h_ in heaps]), snpmer_ptrs=np.concatenate([h_["snpmer_ptrs"] for h_ in heaps]),
                       lengths=np.tile(heap["lengths"], world), snpmer_counts=np.tile(heap["snpmer_counts"], world), kmer_bytes=heap["kmer_bytes"],
                       file_idx=None if heap["file_idx"] is None else np.tile(heap["file_idx"], world))
            m.reuse_outputs = True                                                             # a caller mapping into the same Vecs call after call
            m.map_twin_read_ptrs(big, n_samples=w.n_samples, class_cap=2 * world * n_reads + 1024)   # warm-up (allocations)
            walls, st_ = [], None
            for _ in range(e2e_steps):
                t0 = time.perf_counter()
                resm = m.map_twin_read_ptrs(big, n_samples=w.n_samples, class_cap=2 * world * n_reads + 1024)    # returns after every device finished
                walls.append(time.perf_counter() - t0); st_ = resm["stats"]
            ok = bool(np.array_equal(resm["totals"], world * res["totals"]) and np.array_equal(resm["unambig"], world * res["unambig"])
                      and np.array_equal(resm["status"][:n_reads], res["status"]) and np.array_equal(resm["status"][-n_reads:], res["status"]))
            sp_line = {"value": world * n_reads * e2e_steps / sum(walls), "unit": "reads/s", "ms_per_step": 1e3 * sum(walls) / e2e_steps,
                       "timing": "host wall clock around the synchronous call (it contains host work: pointer gather, result concatenation)",
                       "device_ms_max": st_["total_ms"], "gather_ms_max": st_["gather_ms"], "h2d_ms_max": st_["h2d_ms"],
                       "h2d_bytes_per_step": int(st_["h2d_bytes"]), "d2h_bytes_per_step": int(st_["d2h_bytes"]), "steps": e2e_steps,
                       "n_gpus": world, "host_threads": os.cpu_count(), "allreduce": "peer-memory kernel on device 0",
                       "input": f"one process, savont_b200_multi_map_twin_read_ptrs over {world} GPUs; every device's shard = a copy of the rank-0 shard "
                                "in its own heap arena", "counters_equal_n_times_single": ok}
            m.close()
        dist.barrier(group=gloo)

    if rank != 0:
        return
    fastq_line = {"value": world * n_reads * e2e_steps / (e2e_total_ms / 1e3), "unit": "reads/s", "h2d_bytes_per_step": int(e2e_stats["h2d_bytes"]),
                  "d2h_bytes_per_step": int(e2e_stats["d2h_bytes"]), "steps": e2e_steps, "ms_per_step": e2e_total_ms / e2e_steps,
                  "h2d_ms": e2e_stats["h2d_ms"], "d2h_ms": e2e_stats["d2h_ms"],
                  "input": "FASTQ bytes (ASCII bases + Phred) through savont_b200_map_reads"}
    peaks = _peaks()
    K = args.steps
    value = world * n_reads * K / (total_ms / 1e3)
    e2e_value = world * n_reads * e2e_steps / (e2e_total_ms / 1e3)
    align_ms = float(np.mean([s["align_ms"] for s in stats]))
    sketch_ms = float(np.mean([s["sketch_ms"] for s in stats]))
    filter_ms = float(np.mean([s["filter_ms"] for s in stats]))
    assign_ms = float(np.mean([s["assign_ms"] for s in stats]))
    cells = stats[-1]["dp_cells"]; pairs = stats[-1]["n_pairs"]; bases = stats[-1]["n_bases"]
    gcups = cells / (align_ms / 1e3) / 1e9 if align_ms > 0 else 0.0
    # INT32 roofline of the DP kernel: ALU pipe = 16 lanes/clk/SMSP x 4 SMSP x SMs; OPS_PER_WORD INT32 ALU
    # instructions per 32-cell word-column counted from the kernel's SASS (DESIGN.md "K3")
    sm_mhz = clocks.get("sm_mhz") or peaks["sm_max_mhz"]
    alu_ops = 148 * 4 * 16 * sm_mhz * 1e6
    peak_gcups = alu_ops * 32.0 / OPS_PER_WORD / 1e9          # no discount for vector bits the kernel computes beyond the band
    # HBM view of the sketch kernel (algorithmic bytes: bases + quals in, planes + sketch out)
    sk_bytes = bases * 2 + bases / 4 + 12.0 * (stats[-1]["n_reads"] * (bases / max(stats[-1]["n_reads"], 1) / 11.0 + 25))
    line = {
        "metric": "reads aligned/sec (read-vs-ASV assignment path)", "value": value, "unit": "reads/s",
        "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": total_ms / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 (bit-parallel INT32 DP) / f64 thresholds", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w.cfg['desc']}", "n_asv": w.n_asv, "band": w.band,
                   "reads_per_gpu_per_step": n_reads, "parallelism": f"reads sharded x{world}, ASV index replicated",
                   "l2": "inputs (~%.1f GB per step) larger than L2" % (2 * bases / 1e9)},
        "gcups": gcups,
        "kernel_ms": {"sketch": sketch_ms, "filter": filter_ms, "align": align_ms, "assign": assign_ms},
        "pairs_per_step": pairs, "dp_cells_per_step": cells,
        "roofline": {"bound": "int32", "kernel": "align_nm_kernel", "achieved": gcups, "peak": peak_gcups, "unit": "GCUPS",
                     "frac": gcups / peak_gcups if peak_gcups else None,
                     "traffic": _ncu_traffic("align_nm_kernel", align_ms) if (n_reads == 1_000_000 and args.workload == "ont16s") else None,
                     "traffic_source": _ncu_traffic_note("align_nm_kernel", align_ms),
                     "algorithmic_bytes": pairs * (376 + 21),
                     "peak_how": f"148 SM x 64 INT32 ALU lanes x {sm_mhz:.0f} MHz (sampled) x 32 cells / {OPS_PER_WORD} ALU-pipe instructions "
                                 f"per 32 band cells (the recurrence's minimum, DESIGN.md K3; the kernel issues {KERNEL_OPS_PER_COLUMN} per "
                                 f"{2 * 64 + 1}-cell column at band 64); lane rate measured by tools/pipe_bench.cu (profiles/r02_pipe_bench.txt)"},
        "roofline_hbm": {"bound": "hbm", "kernel": "sketch_fast_kernel", "achieved": sk_bytes / (sketch_ms / 1e3) / 1e9 if sketch_ms else None,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "peak_source": peaks["source"],
                         "frac": (sk_bytes / (sketch_ms / 1e3) / 1e9) / peaks["hbm_gbs"] if sketch_ms else None,
                         "traffic": _ncu_traffic("sketch_fast_kernel", sketch_ms) if (n_reads == 1_000_000 and args.workload == "ont16s") else None,
                         "algorithmic_bytes": sk_bytes},
        "e2e": ptr_line if ptr_line is not None else fastq_line,
        "e2e_twin_contiguous": twin_line,
        "value_twin_resident": twin_resident,
        "value_with_cigar": cigar_line,
        "e2e_single_process": sp_line,
        "e2e_fastq": fastq_line,
        "gpu_launches": int(sum(s["kernel_launches"] for s in stats)),
        "clocks": clocks,
        "assigned_frac": float(got["totals"][0]) / max(n_reads * (world if world > 1 else 1), 1),
        "host_cores_per_rank": numa,
        "gen_s": t_gen,
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        ns = min(args.cpu_reads, n_reads)
        sub = {k: r[k] for k in ("bases", "quals")}
        offs = r["offsets"][: ns + 1]
        orc = O.Oracle(O.default_params(band=w.band), w.snpmers)
        orc.load_asvs(w.asv_bases, w.asv_offsets)
        t0 = time.perf_counter()
        want = orc.map_reads(sub["bases"], sub["quals"], offs, file_idx=None if fidx is None else fidx[:ns], n_samples=w.n_samples,
                             n_threads=cores)
        dt = time.perf_counter() - t0
        keys = ("status", "best_nm", "class_offsets", "class_asv", "unambig", "ambig", "leq10", "totals", "sample_singleton")
        same = bool(ns == ns_s and all(np.array_equal(np.asarray(want[k]), np.asarray(sub_got[k])) for k in keys))
        same_resident = bool(np.array_equal(want["status"], got["status"][:ns]) and np.array_equal(want["best_nm"], got["best_nm"][:ns])
                             and np.array_equal(want["class_offsets"], got["class_offsets"][:ns + 1])
                             and np.array_equal(want["class_asv"], got["class_asv"][:int(got["class_offsets"][ns])]))
        line["cpu_baseline"] = {"value": ns / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                                "sample": f"first {ns} reads of the same workload, OpenMP over reads (oracle/savont_oracle.c)",
                                "gpu_matches_oracle_on_sample": bool(same and same_resident),
                                "compared": "status, best_nm, class_offsets, class_asv of the timed resident pass; the same plus unambig / ambig / "
                                            "leq10 / totals / per-sample singleton counters of a separate C-ABI call on the sample"}
    if multi_check is not None:
        line["multi_gpu_check"] = multi_check
    print(json.dumps(line), flush=True)


# INT32 ALU-pipe instructions per 32-cell word-column that the banded Myers/Hyyro recurrence as formulated in align.cu
# cannot avoid (NW = 5 words): 2 LOP3 per word-aligned ASV word for the mismatch bits (2*(NW+1)/NW = 2.4), 1 SHF to align
# them to the band, X, t, IADD3.X, D0, Ph, Mh, the re-alignment SHF, Pv', Mv' = 12.4.  The SASS inner loop of
# align_nm_kernel<5> is 71 instructions per column (66 ALU-pipe, 4 IMAD, 1 BRA: `python tools/sass_loops.py`) = 13.2 ALU
# per word including the per-column overhead (read-base broadcast, band-top score, loop counter).
OPS_PER_WORD = 12.4


KERNEL_OPS_PER_COLUMN = 55.9     # ALU-pipe instructions per column in the SASS interior loop of align_nm_kernel<4, true> (tools/sass_loops.py)
PROFILE_TAG = "r02t"             # profiles/<tag>_<kernel>_full.txt: ncu --set full summaries of THIS round's kernels (profiles/refresh.py)


def _ncu_profile(kernel):
    """(dram bytes read + written, kernel duration in ms) of the committed ncu --set full summary of a 1M-read launch, or None"""
    path = os.path.join(ROOT, "profiles", f"{PROFILE_TAG}_{kernel}_full.txt")
    try:
        tot, dur = 0.0, None
        for line in open(path):
            f = line.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(f[2], 1.0)
                tot += float(f[1]) * mult
            if len(f) >= 3 and f[0] == "gpu__time_duration.sum":
                dur = float(f[1]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(f[2], 1.0)
        return (tot or None), dur
    except Exception:
        return None, None


def _ncu_traffic(kernel, measured_ms):
    """DRAM traffic per launch from the committed capture -- only if that capture is of the kernel as it runs now: its recorded
    duration must be within 10 % of this run's CUDA-event time, otherwise the profile is stale and the traffic is not reported"""
    tot, dur = _ncu_profile(kernel)
    if tot is None or dur is None or not measured_ms or abs(dur - measured_ms) > 0.10 * measured_ms:
        return None
    return tot


def _ncu_traffic_note(kernel, measured_ms):
    tot, dur = _ncu_profile(kernel)
    if tot is None:
        return f"no profile profiles/{PROFILE_TAG}_{kernel}_full.txt"
    ok = dur is not None and measured_ms and abs(dur - measured_ms) <= 0.10 * measured_ms
    return (f"profiles/{PROFILE_TAG}_{kernel}_full.txt (ncu --set full, one launch of the same 1M-read step: {dur:.3f} ms there, {measured_ms:.3f} ms here"
            + ("" if ok else "; MORE THAN 10 % APART: profile is stale, traffic withheld") + ")")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ont16s")
    ap.add_argument("--reads", type=int, default=1_000_000, help="reads per GPU per step")
    ap.add_argument("--total-reads", type=int, default=0, help="stream a read set larger than one step as ceil(total / reads) chunks (N = 1)")
    ap.add_argument("--batch-reads", type=int, default=131072)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-reads", type=int, default=100_000, help="sample size of the CPU baseline / reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-twin", action="store_true", help="skip the TwinRead-input e2e arms")
    ap.add_argument("--no-cigar", action="store_true", help="skip the arm that also fetches strand / coordinates / identity / CIGAR of every read")
    ap.add_argument("--no-single-process", action="store_true", help="N > 1: skip the leg in which rank 0 drives all GPUs through the multi-GPU handle")
    ap.add_argument("--path", default="assign", choices=["assign", "kmercount", "pileup", "cluster", "reassign", "snpcluster"],
                    help="assign: the headline read-vs-ASV path; kmercount: Stage 1 split-k-mer counting (K5)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        print(f"bench.py: WORLD_SIZE={world} != --gpus {args.gpus}", file=sys.stderr)
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.path == "kmercount":
        run_kmercount(args, rank, world, local_rank)
    elif args.path == "pileup":
        run_pileup(args, rank, world, local_rank)
    elif args.path == "cluster":
        run_cluster(args, rank, world, local_rank)
    elif args.path == "reassign":
        run_reassign(args, rank, world, local_rank)
    elif args.path == "snpcluster":
        run_snpcluster(args, rank, world, local_rank)
    elif args.total_reads > args.reads:
        run_streamed(args, rank, world, local_rank)
    else:
        run_b200(args, rank, world, local_rank)
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
