"""torchrun --nproc-per-node N tools/shard_check.py [n_reads]: the sharded Stage 1 k-mer counting on N GPUs (cudaIpc peer buffers,
one process per GPU) against the single-GPU count of the same reads (rank 0).  SHARD_CHECK_BACKEND=gloo runs the N processes on
GPU 0 with gloo collectives (a one-GPU box is enough)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from savont_b200 import _lib, dist as sd, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if os.environ.get("SHARD_CHECK_BACKEND") == "gloo":     # all ranks on GPU 0 (cudaIpc between processes on one device), gloo collectives
        lr = 0
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    w = synth.Workload("ont16s", n_roots=5)
    r = w.reads(n)
    lo, hi = sd.shard_bounds(n, world, rank)
    sh = sd.slice_reads(r, lo, hi)
    ctx = _lib.Context(lr, _lib.default_params(band=w.band))
    ctx.upload_reads(sh["bases"], sh["quals"], sh["offsets"])
    peers = sd.KmerShardPeers(ctx)
    # a small shard first, then the full one twice: the entry buffers grow while the peers have them mapped, then are reused
    sm = sd.slice_reads(r, lo, lo + max(1, (hi - lo) // 5))
    ctx.upload_reads(sm["bases"], sm["quals"], sm["offsets"])
    sd.count_split_kmers_sharded(ctx, peers, cap=1 << 22, fetch=False)
    ctx.upload_reads(sh["bases"], sh["quals"], sh["offsets"])
    for it in range(2):
        out = sd.count_split_kmers_sharded(ctx, peers, cap=1 << 22, sort=True)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((out["kmers"], out["counts"]), gathered, dst=0)
    ok = True
    if rank == 0:
        km = np.concatenate([g[0] for g in gathered]); ct = np.concatenate([g[1] for g in gathered])
        o = np.argsort(km, kind="stable"); km, ct = km[o], ct[o]
        ref = _lib.Context(lr, _lib.default_params(band=w.band))
        want = ref.count_split_kmers(r["bases"], r["quals"], r["offsets"])
        ok = np.array_equal(km, want["kmers"]) and np.array_equal(ct, want["counts"])
        print(f"shard_check world={world} reads={n} survivors={len(km)} per-rank={[len(g[0]) for g in gathered]} "
              f"{'OK' if ok else 'MISMATCH'}", flush=True)
        ref.close()
    peers.close()
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
