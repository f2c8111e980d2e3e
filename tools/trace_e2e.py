"""e2e pipeline probe (GPU box): the TwinRead pointer entry point under different gather settings, optionally with the per-batch
host timeline (SAVONT_B200_TRACE=1).  usage: python tools/trace_e2e.py [sweep]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from savont_b200 import _lib, synth, twin
w = synth.Workload("ont16s")
r = w.reads(1_000_000, alloc=_lib.pinned_array)
ctx = _lib.Context(0, _lib.default_params(band=w.band))
ctx.load_snpmers(w.snpmers); ctx.load_asvs(w.asv_bases, w.asv_offsets)
tw = twin.twin_reads_from_snpmers(ctx, r["bases"], r["quals"], r["offsets"], kmer48=True, alloc=_lib.pinned_array)
heap = twin.scatter_to_heap(tw)
def run(label):
    ctx.map_twin_read_ptrs(heap)
    best = None
    for _ in range(3):
        t0 = time.perf_counter(); res = ctx.map_twin_read_ptrs(heap); dt = time.perf_counter() - t0
        st = res["stats"]
        if best is None or st["total_ms"] < best[1]["total_ms"]: best = (dt, st)
    print(label, "wall ms", round(best[0] * 1e3, 2), {k: round(v, 2) if isinstance(v, float) else v for k, v in best[1].items() if k in ("gather_ms", "h2d_ms", "total_ms", "n_batches")}, flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "threads":
    for thr in (16, 32, 64, 128):
        ctx.set_host_threads(thr); run(f"threads {thr}")
    # the same heap in 2 MB pages (transparent huge pages requested for the arenas)
    import ctypes, mmap
    libc = ctypes.CDLL(None)
    def huge_copy(a):
        n = (a.nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20)
        m = mmap.mmap(-1, n + (2 << 20))
        buf = np.frombuffer(m, dtype=np.uint8)
        off = (-buf.ctypes.data) % (2 << 20)
        libc.madvise(ctypes.c_void_p(buf.ctypes.data + off), ctypes.c_size_t(n), 14)      # MADV_HUGEPAGE
        out = buf[off:off + a.nbytes]; out[:] = a
        return out, m
    wa, sa = heap["_arenas"]
    w2, m1 = huge_copy(wa); s2, m2 = huge_copy(sa)
    heap2 = dict(heap)
    heap2["dna_ptrs"] = (heap["dna_ptrs"].astype(np.int64) - wa.ctypes.data + w2.ctypes.data).astype(np.uint64)
    heap2["snpmer_ptrs"] = (heap["snpmer_ptrs"].astype(np.int64) - sa.ctypes.data + s2.ctypes.data).astype(np.uint64)
    heap_small = heap; heap = heap2
    for thr in (16, 32):
        ctx.set_host_threads(thr); run(f"hugepage arenas, threads {thr}")
    print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), flush=True)
    heap = heap_small
elif len(sys.argv) > 1 and sys.argv[1] == "multi":
    # one process driving 2 GPUs: where does the gather time go compared with one context per process?
    import threading
    cores = os.cpu_count() or 1
    half = max(1, cores // 2)
    ctx.set_host_threads(half); run(f"one context, {half} threads")
    def copy_heap(h):
        wa, sa = h["_arenas"]
        w2 = wa.copy(); s2 = sa.copy()
        o = dict(h)
        o["dna_ptrs"] = (h["dna_ptrs"].astype(np.int64) - wa.ctypes.data + w2.ctypes.data).astype(np.uint64)
        o["snpmer_ptrs"] = (h["snpmer_ptrs"].astype(np.int64) - sa.ctypes.data + s2.ctypes.data).astype(np.uint64)
        o["_arenas"] = (w2, s2)
        return o
    heapB = copy_heap(heap)
    ctx1 = _lib.Context(1, _lib.default_params(band=w.band))
    ctx1.load_snpmers(w.snpmers); ctx1.load_asvs(w.asv_bases, w.asv_offsets); ctx1.set_host_threads(half)
    for hb, label in ((heap, "same heap"), (heapB, "own heap")):
        ctx1.map_twin_read_ptrs(hb)
        out = [None, None]
        def go(i, cx, hh):
            t0 = time.perf_counter(); res = cx.map_twin_read_ptrs(hh); out[i] = (time.perf_counter() - t0, res["stats"])
        for _ in range(2):
            th = [threading.Thread(target=go, args=(0, ctx, heap)), threading.Thread(target=go, args=(1, ctx1, hb))]
            t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; dt = time.perf_counter() - t0
        print(f"two contexts from two threads ({label}): wall ms", round(dt * 1e3, 2), [(round(o[0] * 1e3, 2), round(o[1]["gather_ms"], 2), round(o[1]["total_ms"], 2)) for o in out], flush=True)
    ctx1.close()
    m = _lib.MultiContext([0, 1], _lib.default_params(band=w.band))
    m.load_snpmers(w.snpmers); m.load_asvs(w.asv_bases, w.asv_offsets)
    for hb, label in ((heap, "same heap"), (heapB, "own heap")):
        big = {k: np.concatenate([heap[k], hb[k]]) for k in ("dna_ptrs", "snpmer_ptrs", "lengths", "snpmer_counts")}
        big["kmer_bytes"] = heap["kmer_bytes"]; big["file_idx"] = None
        for thr in (0, cores, 2 * cores):
            m.set_host_threads(thr)
            m.map_twin_read_ptrs(big)
            t0 = time.perf_counter(); res = m.map_twin_read_ptrs(big); dt = time.perf_counter() - t0
            print(f"MultiContext 2 devices ({label}), host threads {thr}: wall ms", round(dt * 1e3, 2), "gather_ms", round(res["stats"]["gather_ms"], 2), "total_ms", round(res["stats"]["total_ms"], 2), flush=True)
    m.close()
elif len(sys.argv) > 1 and sys.argv[1] == "sweep":
    for pf, pf_all, nt in [(8, 1, 1), (8, 1, 0), (0, 1, 1), (4, 1, 1), (16, 1, 1), (32, 1, 1), (16, 0, 1), (32, 0, 1), (16, 0, 0)]:
        os.environ["SAVONT_B200_GATHER_PF"] = str(pf); os.environ["SAVONT_B200_GATHER_PF_ALL"] = str(pf_all); os.environ["SAVONT_B200_GATHER_NT"] = str(nt)
        run(f"pf={pf} all={pf_all} nt={nt}")
    for b in (65536, 262144):
        ctx.set_batch_reads(b); run(f"batch_reads={b} (pf=16 all=0 nt=0)")
    ctx.set_batch_reads(0)
else:
    for thr in (0, 8, 4):
        ctx.set_host_threads(thr); run(f"threads {thr}")
ctx.set_host_threads(0)
ctx.map_twin_reads(tw)
t0 = time.perf_counter(); res = ctx.map_twin_reads(tw); dt = time.perf_counter() - t0
print("contiguous wall ms", round(dt * 1e3, 2), round(res["stats"]["total_ms"], 2), flush=True)
