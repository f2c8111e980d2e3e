import sys, time, numpy as np
sys.path.insert(0, '.')
from savont_b200 import _lib, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
w = synth.Workload("ont16s")
r = w.reads(n, alloc=_lib.pinned_array)
ctx = _lib.Context(0, _lib.default_params(band=w.band))
ctx.upload_reads(r["bases"], r["quals"], r["offsets"])
for it in range(3):
    t = time.time(); out = ctx.count_split_kmers_resident(sort=False); dt = time.time() - t
    print("resident", it, round(dt * 1e3, 1), "ms wall", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in out["stats"].items()}, len(out["kmers"]), flush=True)
for it in range(2):
    t = time.time(); out = ctx.count_split_kmers(r["bases"], r["quals"], r["offsets"], sort=False); dt = time.time() - t
    print("e2e", it, round(dt * 1e3, 1), "ms wall", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in out["stats"].items()}, len(out["kmers"]), flush=True)
