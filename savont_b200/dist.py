"""Multi-GPU plumbing for the read-vs-ASV path: one process per GPU under torch.distributed.

Reads shard into contiguous ranges (order preserving, reference src/main.rs:538); the SNPmer set, the ASVs
and their index are replicated; there is no data-path exchange.  The only collective is ONE all-reduce (sum)
of the u64 counter vector [unambig | ambig | leq10 | assigned | filtered | singleton S*A]; the sparse
multi-ASV equivalence classes are gathered to rank 0 for the EM.

`mapper` is anything with `map_shard(bases, quals, offsets, file_idx, n_samples) -> dict` returning the
per-read results and the counter arrays: the CUDA context in production (GpuMapper), a stub in the
CPU-only gloo tests."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_reads: int, world: int, rank: int):
    """contiguous, order-preserving ranges; the first n % world ranks get one extra read"""
    base, rem = divmod(n_reads, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def slice_reads(reads: dict, lo: int, hi: int) -> dict:
    offs = np.asarray(reads["offsets"], np.uint64)
    b0, b1 = int(offs[lo]), int(offs[hi])
    out = dict(bases=np.asarray(reads["bases"])[b0:b1], offsets=offs[lo:hi + 1] - offs[lo],
               quals=None if reads.get("quals") is None else np.asarray(reads["quals"])[b0:b1],
               file_idx=None if reads.get("file_idx") is None else np.asarray(reads["file_idx"])[lo:hi])
    return out


class GpuMapper:
    """libsavont_b200 context on this rank's device; counters are all-reduced in place on the device"""

    def __init__(self, ctx):
        self.ctx = ctx

    def map_shard(self, bases, quals, offsets, file_idx, n_samples):
        self.ctx.upload_reads(bases, quals, offsets, file_idx=file_idx, n_samples=n_samples)
        self.ctx.map_resident()
        return None   # results fetched after the all-reduce

    def allreduce_counts(self, group=None):
        ptr, n = self.ctx.counts_dev()

        class _View:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}
        t = torch.as_tensor(_View(), device=torch.device("cuda", self.ctx.device))
        dist.all_reduce(t, group=group)          # NCCL, one launch, in place

    def fetch(self):
        return self.ctx.fetch_resident()


def _counter_vector(res, n_asv, n_samples):
    return np.concatenate([res["unambig"], res["ambig"], res["leq10"], res["totals"], res["sample_singleton"].reshape(-1)]).astype(np.int64)


def map_reads_sharded(mapper, reads: dict, n_asv: int, n_samples: int = 1, group=None):
    """Every rank maps its shard; returns on every rank the globally reduced counters and, on rank 0, the
    merged multi-ASV class histogram {(sample, asv tuple): count}; per-read outputs stay local."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(reads["offsets"]) - 1
    lo, hi = shard_bounds(n, world, rank)
    sh = slice_reads(reads, lo, hi)
    res = mapper.map_shard(sh["bases"], sh["quals"], sh["offsets"], sh["file_idx"], n_samples)
    if res is None:                       # GPU path: counters live on the device
        if world > 1:
            mapper.allreduce_counts(group)
        res = mapper.fetch()
        vec = _counter_vector(res, n_asv, n_samples)
    else:                                 # host path (tests): reduce a CPU tensor (gloo)
        vec = _counter_vector(res, n_asv, n_samples)
        if world > 1:
            t = torch.from_numpy(vec.copy())
            dist.all_reduce(t, group=group)
            vec = t.numpy()
    A = n_asv
    counters = dict(unambig=vec[:A], ambig=vec[A:2 * A], leq10=vec[2 * A:3 * A], totals=vec[3 * A:3 * A + 2],
                    sample_singleton=vec[3 * A + 2:].reshape(n_samples, A))
    # sparse part: classes with more than one ASV, keyed by (sample, ids)
    co = res["class_offsets"].astype(np.int64)
    sizes = co[1:] - co[:-1]
    fi = sh["file_idx"] if sh["file_idx"] is not None else np.zeros(hi - lo, np.uint32)
    local = {}
    for i in np.flatnonzero(sizes > 1):
        key = (int(fi[i]), tuple(int(x) for x in res["class_asv"][co[i]:co[i + 1]]))
        local[key] = local.get(key, 0) + 1
    merged = None
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(local, gathered, dst=0, group=group)
        if rank == 0:
            merged = {}
            for d in gathered:
                for k, v in d.items():
                    merged[k] = merged.get(k, 0) + v
    else:
        merged = local
    return dict(counters=counters, multi_classes=merged, local=res, shard=(lo, hi))


def depths_from_counts(counters, multi_classes, n_asv, sample=None):
    """EM depths from the reduced singleton counts + merged multi-ASV classes (rank 0)"""
    from . import api
    if sample is None:
        single = counters["unambig"]
        multi = {}
        for (s, ids), c in multi_classes.items():
            multi[ids] = multi.get(ids, 0) + c
    else:
        single = counters["sample_singleton"][sample]
        multi = {ids: c for (s, ids), c in multi_classes.items() if s == sample}
    classes = {(a,): int(single[a]) for a in range(n_asv) if single[a] > 0}
    classes.update({k: int(v) for k, v in multi.items()})
    keys = sorted(classes)
    if not keys:
        return np.zeros(n_asv, np.uint64)
    offs = np.zeros(len(keys) + 1, np.uint64)
    offs[1:] = np.cumsum([len(k) for k in keys])
    flat = np.array([a for k in keys for a in k], np.uint32)
    counts = np.array([classes[k] for k in keys], np.uint64)
    _, depths, _ = api.em(n_asv, offs, flat, counts, int(counts.sum()))
    return depths


# ---------------------------------------------------------------------------------------------------------
# Sharded Stage 1 k-mer counting (K5 over several GPUs).  Every rank counts the occurrences of ITS reads per hash partition;
# partition p belongs to rank p * world // P; the scatter kernel of every rank writes its run of partition p directly into the
# owner's entry buffer through NVLink peer memory (savont_b200_kmer_shard_scatter), so the only collectives are one all-gather
# of the P-entry histograms and one of the 64-byte IPC handles, plus a barrier between scatter and count.
# ---------------------------------------------------------------------------------------------------------
def kmer_shard_layout(hists: np.ndarray):
    """hists [world, P] occurrences per rank and partition -> (owner [P], part_off [world, P + 1], dst_off [world, P]):
    part_off[o] = offsets of the partitions inside owner o's entry buffer (empty ranges for partitions it does not own),
    dst_off[r][p] = where rank r's run of partition p starts inside the owner's buffer (ranks in order inside a partition)"""
    H = np.asarray(hists, np.uint64)
    world, P = H.shape
    owner = (np.arange(P, dtype=np.uint64) * np.uint64(world) // np.uint64(P)).astype(np.int64)
    tot = H.sum(axis=0)
    part_off = np.zeros((world, P + 1), np.uint64)
    for o in range(world):
        part_off[o, 1:] = np.cumsum(np.where(owner == o, tot, 0).astype(np.uint64))
    before = np.cumsum(H, axis=0) - H                                   # occurrences of the lower ranks
    start = part_off[owner, np.arange(P)]                               # start of p inside its owner's buffer
    dst_off = (start[None, :] + before).astype(np.uint64)
    return owner, part_off, dst_off


class KmerShardPeers:
    """the peers' entry buffers, cudaIpc-mapped; re-opened only when a peer re-allocated its buffer"""

    def __init__(self, ctx):
        self.ctx, self.handles, self.ptrs = ctx, {}, {}

    def update(self, rank, handles, own_ptr):
        out = []
        for r, hd in enumerate(handles):
            if r == rank:
                out.append(own_ptr); continue
            if self.handles.get(r) != hd:
                if r in self.ptrs:
                    self.ctx.ipc_close(self.ptrs[r])
                self.ptrs[r] = self.ctx.ipc_open(hd); self.handles[r] = hd
            out.append(self.ptrs[r])
        return out

    def drop(self, r):
        if r in self.ptrs:
            self.ctx.ipc_close(self.ptrs.pop(r)); self.handles.pop(r, None)

    def close(self):
        for p in self.ptrs.values():
            self.ctx.ipc_close(p)
        self.handles, self.ptrs = {}, {}


def count_split_kmers_sharded(ctx, peers: KmerShardPeers | None = None, single_strand=False, cap=1 << 24, sort=False, group=None,
                              fetch=True):
    """K5 over the ranks of `group`; the reads of this rank are resident (ctx.upload_reads).  Returns this rank's shard of the
    result (the k-mers of the partitions it owns) as ctx.count_split_kmers does (fetch=False: only their number and the times).
    Collectives per call: ONE all-gather of [histogram | buffer capacity | IPC handle] (a second one only when some rank had to
    grow its buffer, i.e. on the first call), a barrier after the scatter and one after the count."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    # NCCL moves device tensors; the CPU-only tests drive the same code over gloo with a stub context
    dev = torch.device("cuda", ctx.device) if dist.is_initialized() and dist.get_backend(group) == "nccl" else torch.device("cpu")
    hist = ctx.kmer_shard_hist()
    P = len(hist)
    if world == 1:
        owner, part_off, dst_off = kmer_shard_layout(hist[None, :])
        own_ptr = ctx.kmer_shard_alloc(int(part_off[0, -1]))
        ctx.kmer_shard_scatter([own_ptr], dst_off[0])
        return ctx.kmer_shard_count(part_off[0], single_strand=single_strand, cap=cap, sort=sort, fetch=fetch)
    own_peers = peers or KmerShardPeers(ctx)

    def exchange(own_ptr, own_cap):
        payload = np.zeros(P + 9, np.int64)
        payload[:P] = hist; payload[P] = own_cap
        payload[P + 1:] = np.frombuffer(ctx.ipc_export(own_ptr), np.int64)
        t = torch.from_numpy(payload).to(dev)
        g = torch.empty(world * (P + 9), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(g, t, group=group)
        return g.cpu().numpy().reshape(world, P + 9)
    own_ptr, own_cap = ctx.kmer_shard_alloc(0, with_capacity=True)          # the buffer of the previous call, if any
    G = exchange(own_ptr, own_cap)
    owner, part_off, dst_off = kmer_shard_layout(G[:, :P].astype(np.uint64))
    need = part_off[:, -1].astype(np.int64)
    grow = need > G[:, P]
    if np.any(grow):                                                        # somebody has to grow: everybody sees it
        for r in np.flatnonzero(grow):                                      # unmap a buffer before its owner frees it
            if r != rank:
                own_peers.drop(int(r))
        dist.barrier(group=group)
        if grow[rank]:
            own_ptr, own_cap = ctx.kmer_shard_alloc(int(need[rank] + need[rank] // 8), with_capacity=True)
        G = exchange(own_ptr, own_cap)
    handles = [G[r, P + 1:].tobytes() for r in range(world)]
    ptrs = own_peers.update(rank, handles, own_ptr)
    ctx.kmer_shard_scatter(ptrs, dst_off[rank])
    dist.barrier(group=group)                                # every rank's runs have landed in the owners' buffers
    out = ctx.kmer_shard_count(part_off[rank], single_strand=single_strand, cap=cap, sort=sort, fetch=fetch)
    dist.barrier(group=group)                                # nobody overwrites a buffer a peer still counts from
    if peers is None:
        own_peers.close()
    return out
