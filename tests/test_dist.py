"""N>1 path on CPU: world_size-2 gloo run of the sharding / all-reduce / class-gather logic of
savont_b200/dist.py, with the CPU oracle standing in for the per-rank mapper (the CUDA context needs a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class OracleMapper:
    def __init__(self, w):
        from oracle import oracle as O
        self.orc = O.Oracle(O.default_params(band=w.band), w.snpmers)
        self.orc.load_asvs(w.asv_bases, w.asv_offsets)

    def map_shard(self, bases, quals, offsets, file_idx, n_samples):
        return self.orc.map_reads(bases, quals, offsets, file_idx=file_idx, n_samples=n_samples, n_threads=1)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from savont_b200 import dist as sd, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.Workload("pooled", n_roots=3, n_samples=2, sites_per_genus=4)
    reads = w.reads(301)
    out = sd.map_reads_sharded(OracleMapper(w), reads, w.n_asv, n_samples=2)
    payload = dict(rank=rank, shard=out["shard"], counters={k: v.tolist() for k, v in out["counters"].items()},
                   multi=None if out["multi_classes"] is None else {repr(k): v for k, v in out["multi_classes"].items()})
    if rank == 0:
        payload["depths"] = sd.depths_from_counts(out["counters"], out["multi_classes"], w.n_asv).tolist()
        payload["depths_s1"] = sd.depths_from_counts(out["counters"], out["multi_classes"], w.n_asv, sample=1).tolist()
    q.put(payload)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_all_reads():
    from savont_b200 import dist as sd
    for n in (0, 1, 7, 100, 301):
        for world in (1, 2, 3, 8):
            b = [sd.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def test_two_rank_gloo_matches_single_process():
    pytest.importorskip("torch")
    from oracle import oracle as O
    from savont_b200 import api, synth
    if not os.path.exists(api._HOST_PATH):
        pytest.skip("libsavont_host.so not built")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    outs = [q.get(timeout=240) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    outs.sort(key=lambda d: d["rank"])
    assert outs[0]["shard"] == (0, 151) and outs[1]["shard"] == (151, 301)
    assert outs[0]["counters"] == outs[1]["counters"]                # every rank holds the reduced counters
    # single-process truth
    w = synth.Workload("pooled", n_roots=3, n_samples=2, sites_per_genus=4)
    reads = w.reads(301)
    orc = O.Oracle(O.default_params(band=w.band), w.snpmers); orc.load_asvs(w.asv_bases, w.asv_offsets)
    full = orc.map_reads(reads["bases"], reads["quals"], reads["offsets"], file_idx=reads["file_idx"], n_samples=2)
    c = outs[0]["counters"]
    assert c["unambig"] == full["unambig"].tolist() and c["ambig"] == full["ambig"].tolist() and c["leq10"] == full["leq10"].tolist()
    assert c["totals"] == full["totals"].tolist() and c["sample_singleton"] == full["sample_singleton"].tolist()
    offs, flat, counts, total = api.eq_classes(full["status"], full["class_offsets"], full["class_asv"])
    _, depths, _ = O.em(w.n_asv, offs, flat, counts, total)
    assert outs[0]["depths"] == depths.tolist()
    offs, flat, counts, total = api.eq_classes(full["status"], full["class_offsets"], full["class_asv"], reads["file_idx"], 1)
    _, d1, _ = O.em(w.n_asv, offs, flat, counts, total)
    assert outs[0]["depths_s1"] == d1.tolist()
    assert outs[1]["multi"] is None and outs[0]["multi"] is not None


def test_kmer_shard_layout_is_a_partition_of_the_owner_buffers():
    """index arithmetic of the sharded k-mer counting: every (rank, partition) run lands in its owner's buffer, runs are
    disjoint, ordered by rank inside a partition, and together cover the buffer exactly"""
    from savont_b200 import dist as sd
    rng = np.random.default_rng(5)
    for world in (1, 2, 3, 8):
        P = 8192
        H = rng.integers(0, 50, size=(world, P)).astype(np.uint64)
        H[:, rng.integers(0, P, 300)] = 0
        owner, part_off, dst_off = sd.kmer_shard_layout(H)
        assert owner.min() == 0 and owner.max() == world - 1 and np.all(np.diff(owner) >= 0)
        assert np.all(np.bincount(owner, minlength=world) >= P // world)
        for o in range(world):
            own = owner == o
            assert part_off[o, -1] == H[:, own].sum()
            assert np.all((part_off[o, 1:] - part_off[o, :-1])[~own] == 0)
            filled = np.zeros(int(part_off[o, -1]), np.int32)
            for r in range(world):
                for p in np.flatnonzero(own & (H[r] > 0)):
                    a = int(dst_off[r, p]); b = a + int(H[r, p])
                    assert int(part_off[o, p]) <= a and b <= int(part_off[o, p + 1])
                    filled[a:b] += 1
            assert np.all(filled == 1)


class StubShardCtx:
    """stands in for the CUDA context in count_split_kmers_sharded: 'device buffers' are ids in a table shared through the
    handles, the scatter / count phases only record what they were given"""
    P = 64

    def __init__(self, rank, hists):
        self.rank, self.hists, self.step, self.cap, self.gen, self.device = rank, hists, 0, 0, 0, 0
        self.log = []

    def kmer_shard_hist(self):
        return self.hists[self.step].astype(np.uint32)

    def kmer_shard_alloc(self, n, with_capacity=False):
        if n > self.cap:
            self.cap = int(n); self.gen += 1
            self.log.append(("alloc", self.step, int(n)))
        ptr = 1000 * (self.rank + 1) + self.gen
        return (ptr, self.cap) if with_capacity else ptr

    def ipc_export(self, ptr):
        return int(ptr).to_bytes(8, "little") * 8

    def ipc_open(self, handle):
        self.log.append(("open", self.step, int.from_bytes(handle[:8], "little")))
        return int.from_bytes(handle[:8], "little")

    def ipc_close(self, ptr):
        self.log.append(("close", self.step, int(ptr)))

    def kmer_shard_scatter(self, ptrs, dst_off):
        self.log.append(("scatter", self.step, [int(x) for x in ptrs], np.asarray(dst_off).tolist()))

    def kmer_shard_count(self, part_off, single_strand=False, cap=0, sort=False, fetch=True):
        self.log.append(("count", self.step, np.asarray(part_off).tolist()))
        self.step += 1
        return dict(n=0)


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    from savont_b200 import dist as sd
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    hists = [rng.integers(0, 20, StubShardCtx.P), rng.integers(0, 20, StubShardCtx.P), rng.integers(50, 90, StubShardCtx.P)]   # step 2 grows
    ctx = StubShardCtx(rank, hists)
    peers = sd.KmerShardPeers(ctx)
    for _ in range(3):
        sd.count_split_kmers_sharded(ctx, peers, fetch=False)
    peers.close()
    q.put(dict(rank=rank, hists=[h.tolist() for h in hists], log=ctx.log))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_kmer_counting_driver_two_ranks_gloo():
    """count_split_kmers_sharded over gloo with stub contexts: every rank derives the same layout from the all-gathered histograms,
    peers are re-mapped only when a buffer grew, and a mapping is closed before its owner re-allocates"""
    from savont_b200 import dist as sd
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    outs = sorted([q.get(timeout=240) for _ in range(world)], key=lambda d: d["rank"])
    for p in procs: p.join(timeout=60)
    for step in range(3):
        H = np.array([outs[r]["hists"][step] for r in range(world)], np.uint64)
        owner, part_off, dst_off = sd.kmer_shard_layout(H)
        for r in range(world):
            sc = [e for e in outs[r]["log"] if e[0] == "scatter" and e[1] == step][0]
            ct = [e for e in outs[r]["log"] if e[0] == "count" and e[1] == step][0]
            assert sc[3] == dst_off[r].tolist() and ct[2] == part_off[r].tolist()
            peer = 1 - r
            assert sc[2][r] // 1000 == r + 1 and sc[2][peer] // 1000 == peer + 1      # own buffer, the peer's mapped buffer
    for r in range(world):
        log = outs[r]["log"]
        opens = [e for e in log if e[0] == "open"]; closes = [e for e in log if e[0] == "close"]; allocs = [e for e in log if e[0] == "alloc"]
        assert [e[1] for e in opens] == [0, 2] and [e[1] for e in allocs] == [0, 2]     # mapped at the first call and after the growth only
        assert len(closes) == 2 and closes[0][1] == 2                                    # old mapping closed in step 2 (before the re-allocation) ...
        assert log.index(closes[0]) < log.index(allocs[1]) and closes[0][2] == opens[0][2]
        assert closes[1][2] == opens[1][2]                                               # ... and the new one by peers.close()
