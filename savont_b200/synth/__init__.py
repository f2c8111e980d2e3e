"""Synthetic 16S-like workloads for tests and bench.py.  Harness only."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libsavont_synth.so")
_SRC = os.path.join(_HERE, "synth.cpp")

# name -> generator settings + path params; seeds follow SURVEY.md 8d (20260817 + config id)
CONFIGS = {
    "ont16s": dict(id=2, n_roots=50, n_var=10, len_mean=1500, len_jitter=0, err_rate=0.02, q_lo=25, q_hi=40, band=64,
                   n_reads=1_000_000, n_samples=1, desc="1M ONT full-length 16S reads (~1500 bp, 98% identity) vs 500 ASVs"),
    "hifi16s": dict(id=3, n_roots=50, n_var=10, len_mean=1500, len_jitter=0, err_rate=0.001, q_lo=40, q_hi=93, band=32,
                    n_reads=1_000_000, n_samples=1, desc="1M PacBio HiFi 16S reads (99.9% identity) vs 500 ASVs"),
    "operon": dict(id=4, n_roots=200, n_var=10, len_mean=4500, len_jitter=100, err_rate=0.02, q_lo=25, q_hi=40, band=192,
                   n_reads=2_000_000, n_samples=1, desc="2M rRNA-operon ONT reads (~4.5 kb) vs 2000 ASVs"),
    "pooled": dict(id=5, n_roots=500, n_var=10, len_mean=1500, len_jitter=0, err_rate=0.02, q_lo=25, q_hi=40, band=64,
                   n_reads=10_000_000, n_samples=20, desc="10M pooled reads, 20 samples vs 5000 ASVs"),
}


def build(force=False):
    if os.path.exists(_SRC) and (force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC)):
        cxx = shutil.which("g++") or "/usr/bin/g++"
        subprocess.check_call([cxx, "-O2", "-fPIC", "-fopenmp", "-std=c++17", "-shared", "-o", _LIB, _SRC])
    return _LIB


_lib = None


def _L():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.synth_create.restype = C.c_void_p
        L.synth_create.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                   C.c_double, C.c_int, C.c_int, C.c_int]
        L.synth_destroy.argtypes = [C.c_void_p]
        L.synth_num_asvs.restype = C.c_size_t; L.synth_num_asvs.argtypes = [C.c_void_p]
        L.synth_asv_bytes.restype = C.c_size_t; L.synth_asv_bytes.argtypes = [C.c_void_p]
        L.synth_get_asvs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.synth_snpmers.restype = C.c_size_t; L.synth_snpmers.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.synth_snpmers_from.restype = C.c_size_t
        L.synth_snpmers_from.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
        L.synth_read_lengths.argtypes = [C.c_void_p, C.c_uint64, C.c_size_t, C.c_void_p]
        L.synth_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class Workload:
    """ASVs + SNPmer set + a deterministic stream of reads."""

    def __init__(self, name="ont16s", seed=None, n_roots=None, n_var=None, k=17, **over):
        cfg = dict(CONFIGS[name]); cfg.update(over)
        if n_roots is not None: cfg["n_roots"] = n_roots
        if n_var is not None: cfg["n_var"] = n_var
        self.name, self.cfg, self.k = name, cfg, k
        self.seed = int(seed if seed is not None else 20260817 + cfg["id"])
        L = _L()
        self.h = L.synth_create(self.seed, cfg["n_roots"], cfg["n_var"], cfg["len_mean"], cfg["len_jitter"], 0.525,
                                cfg["err_rate"], cfg["q_lo"], cfg["q_hi"], cfg.get("abund_sigma", 1.5), cfg["n_samples"],
                                cfg.get("trim_max", 20), cfg.get("sites_per_genus", 40))
        self.n_asv = int(L.synth_num_asvs(self.h))
        self.asv_bases = np.zeros(int(L.synth_asv_bytes(self.h)), np.uint8)
        self.asv_offsets = np.zeros(self.n_asv + 1, np.uint64)
        L.synth_get_asvs(self.h, self.asv_bases.ctypes.data, self.asv_offsets.ctypes.data)
        n = int(L.synth_snpmers(self.h, k, None, 0))
        self.snpmers = np.zeros(max(n, 1), np.uint64)
        L.synth_snpmers(self.h, k, self.snpmers.ctypes.data, n)
        self.snpmers = self.snpmers[:n]
        self.band = cfg["band"]
        self.n_samples = cfg["n_samples"]

    def __del__(self):
        try:
            _L().synth_destroy(self.h)
        except Exception:
            pass

    def reads(self, n, first=0, alloc=None):
        """-> dict(bases, quals, offsets, truth_asv, truth_strand, file_idx).  `alloc(nbytes)` may return
        a uint8 array (e.g. pinned memory) for bases / quals."""
        L = _L()
        lens = np.zeros(n, np.uint32)
        L.synth_read_lengths(self.h, first, n, lens.ctypes.data)
        offs = np.zeros(n + 1, np.uint64)
        np.cumsum(lens, out=offs[1:])
        total = int(offs[-1])
        mk = alloc or (lambda nb: np.zeros(max(nb, 1), np.uint8))
        bases, quals = mk(total), mk(total)
        ta = np.zeros(n, np.uint32); ts = np.zeros(n, np.uint8); fi = np.zeros(n, np.uint32)
        L.synth_reads(self.h, first, n, offs.ctypes.data, bases.ctypes.data, quals.ctypes.data, ta.ctypes.data, ts.ctypes.data,
                      fi.ctypes.data)
        return dict(bases=bases[:total], quals=quals[:total], offsets=offs, truth_asv=ta, truth_strand=ts, file_idx=fi)


def snpmers_from_sequences(bases, offsets, k=17):
    """SNPmer set the way Stage 1 would call it from a set of ASV sequences (see synth.cpp header)"""
    L = _L()
    b = np.ascontiguousarray(bases, np.uint8); o = np.ascontiguousarray(offsets, np.uint64)
    n = int(L.synth_snpmers_from(b.ctypes.data, o.ctypes.data, len(o) - 1, k, None, 0))
    out = np.zeros(max(n, 1), np.uint64)
    L.synth_snpmers_from(b.ctypes.data, o.ctypes.data, len(o) - 1, k, out.ctypes.data, n)
    return out[:n]
