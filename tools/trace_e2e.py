import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from savont_b200 import _lib, synth, twin
w = synth.Workload("ont16s")
r = w.reads(1_000_000, alloc=_lib.pinned_array)
ctx = _lib.Context(0, _lib.default_params(band=w.band))
ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
tw = twin.twin_reads_from_snpmers(ctx, r["bases"], r["quals"], r["offsets"], kmer48=True, alloc=_lib.pinned_array)
heap = twin.scatter_to_heap(tw)
for thr in (0, 8, 4):
    ctx.set_host_threads(thr)
    ctx.map_twin_read_ptrs(heap); 
    t0 = time.perf_counter(); res = ctx.map_twin_read_ptrs(heap); dt = time.perf_counter() - t0
    print("threads", thr, "wall ms", round(dt * 1e3, 2), {k: round(v, 2) if isinstance(v, float) else v for k, v in res["stats"].items() if k in ("gather_ms", "h2d_ms", "total_ms", "n_batches")}, flush=True)
os.environ["X"] = "1"
t0 = time.perf_counter(); res = ctx.map_twin_reads(tw); dt = time.perf_counter() - t0
print("contiguous wall ms", round(dt * 1e3, 2), flush=True)
