"""ctypes binding of libsavont_b200.so (include/savont_b200.h).

The library is the product; this module is only the Python doorway used by the tests, the bench
and the multi-GPU host layer.  It never falls back to a CPU implementation: if the shared library
or a CUDA device is missing the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsavont_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "savont_b200.h")

READ_ASSIGNED, READ_NO_CANDIDATE, READ_RATIO_FILTERED, READ_NO_ALIGNMENT, READ_TOO_SHORT = 0, 1, 2, 3, 4
READ_LENGTH_FILTERED, READ_REPETITIVE, READ_LOW_QUALITY = 5, 6, 7
OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_OVERFLOW, ERR_NOMEM = 0, -1, -2, -3, -4, -5


class Params(C.Structure):
    _fields_ = [
        ("k", C.c_int32), ("c", C.c_int32), ("l", C.c_int32), ("min_bq", C.c_int32),
        ("band", C.c_int32), ("strand_mode", C.c_int32), ("candidate_mode", C.c_int32), ("reserved", C.c_int32),
        ("ratio_thresh", C.c_double), ("min_mini_frac_base", C.c_double),
    ]


class ReadFilter(C.Structure):
    """savont_b200_read_filter: the options of kmer_comp::twin_reads_from_snpmers (src/cli.rs:87-96, src/constants.rs)"""
    _fields_ = [("min_read_length", C.c_uint32), ("max_read_length", C.c_uint32), ("max_kmer_count_in_read", C.c_uint32),
                ("reserved", C.c_uint32), ("quality_value_cutoff", C.c_double)]


def read_filter(min_read_length=1100, max_read_length=2000, quality_value_cutoff=98.0, max_kmer_count_in_read=500) -> ReadFilter:
    return ReadFilter(min_read_length, max_read_length, max_kmer_count_in_read, 0, quality_value_cutoff)


class Stats(C.Structure):
    _fields_ = [
        ("sketch_ms", C.c_double), ("filter_ms", C.c_double), ("align_ms", C.c_double), ("assign_ms", C.c_double),
        ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("total_ms", C.c_double),
        ("n_reads", C.c_uint64), ("n_bases", C.c_uint64), ("n_pairs", C.c_uint64), ("dp_cells", C.c_uint64),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("kernel_launches", C.c_uint64), ("n_batches", C.c_uint64),
        ("gather_ms", C.c_double),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class KmerStats(C.Structure):
    _fields_ = [("count_ms", C.c_double), ("compact_ms", C.c_double), ("h2d_ms", C.c_double), ("total_ms", C.c_double),
                ("hist_ms", C.c_double), ("scatter_ms", C.c_double),
                ("n_reads", C.c_uint64), ("n_bases", C.c_uint64), ("table_slots", C.c_uint64), ("n_distinct", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("kernel_launches", C.c_uint64), ("table_regrows", C.c_uint64),
                ("partitions", C.c_uint64), ("spilled_distinct", C.c_uint64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


# savont_b200_pileup_entry
PILEUP_ENTRY = np.dtype([("read", np.uint32), ("kind", np.uint8), ("n", np.uint8), ("base", np.uint8, 2), ("qual", np.uint8, 2),
                         ("hp", np.uint8, 2)])
assert PILEUP_ENTRY.itemsize == 12


class SavontB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsavont_b200 error {code}: {msg}")
        self.code = code


def declared_symbols():
    """every function declared in include/savont_b200.h"""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(savont_b200_[a-z_0-9]+)\s*\(", txt)))


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SavontB200Error(ERR_CUDA, f"{LIB_PATH} is missing: run `python -m savont_b200.build` (nvcc, sm_100a); "
                                            "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.savont_b200_last_error.restype = C.c_char_p
        L.savont_b200_abi_version.restype = C.c_int
        for name in declared_symbols():
            fn = getattr(L, name)
            if name not in ("savont_b200_last_error", "savont_b200_default_params"):
                fn.restype = C.c_int
        L.savont_b200_default_params.restype = None
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise SavontB200Error(rc, lib().savont_b200_last_error().decode())


def _p(arr, typ):
    if arr is None:
        return None
    return arr.ctypes.data_as(C.POINTER(typ))


def _u8(x):
    if x is None:
        return None
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    return np.ascontiguousarray(x, dtype=np.uint8)


def default_params(**kw) -> Params:
    p = Params()
    lib().savont_b200_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def pinned_array(nbytes: int) -> np.ndarray:
    """uint8 numpy array backed by cudaMallocHost memory (never freed explicitly: process lifetime)"""
    ptr = C.c_void_p()
    _check(lib().savont_b200_host_alloc(C.c_size_t(max(nbytes, 1)), C.byref(ptr)))
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(ptr.value)
    return np.frombuffer(buf, dtype=np.uint8, count=nbytes)


def _heap_args(heap):
    dp = np.ascontiguousarray(heap["dna_ptrs"], dtype=np.uint64); sp = np.ascontiguousarray(heap["snpmer_ptrs"], dtype=np.uint64)
    ln = np.ascontiguousarray(heap["lengths"], dtype=np.uint32); sc = np.ascontiguousarray(heap["snpmer_counts"], dtype=np.uint32)
    fi = heap.get("file_idx")
    fi = None if fi is None else np.ascontiguousarray(fi, dtype=np.uint32)
    return dp, ln, sp, sc, int(heap["kmer_bytes"]), fi


class Context:
    """One GPU.  Mirrors the order of calls in refine_asv_depths_with_em (reference
    src/alignment.rs:1723-1765): params -> SNPmer set -> ASVs -> reads."""

    def __init__(self, device: int = 0, params: Params | None = None):
        self.L = lib()
        h = C.c_void_p()
        _check(self.L.savont_b200_create(C.c_int(device), C.byref(h)))
        self.h = h
        self.device = device
        self.n_asv = 0
        self.params = default_params()
        if params is not None:
            self.set_params(params)

    def close(self):
        if getattr(self, "h", None):
            self.L.savont_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, p: Params):
        _check(self.L.savont_b200_set_params(self.h, C.byref(p)))
        self.params = p

    def set_stream(self, cuda_stream_ptr: int | None):
        _check(self.L.savont_b200_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)))

    def set_batch_reads(self, n: int):
        _check(self.L.savont_b200_set_batch_reads(self.h, C.c_size_t(n)))

    def load_snpmers(self, kmers):
        k = np.ascontiguousarray(kmers, dtype=np.uint64)
        _check(self.L.savont_b200_load_snpmers(self.h, _p(k, C.c_uint64), C.c_size_t(len(k))))

    # ---- a4 ----
    def load_high_freq_kmers(self, kmers):
        k = np.ascontiguousarray(kmers if kmers is not None else [], dtype=np.uint64)
        _check(self.L.savont_b200_load_high_freq_kmers(self.h, _p(k, C.c_uint64), C.c_size_t(len(k))))

    def set_read_filter(self, rf: ReadFilter | None):
        _check(self.L.savont_b200_set_read_filter(self.h, None if rf is None else C.byref(rf)))

    def twin_reads_from_snpmers(self, bases, quals, offsets):
        """per-read part of kmer_comp::twin_reads_from_snpmers: status (0 kept / 4 / 5 / 6 / 7), est_id, solid counts"""
        b = _u8(bases); q = _u8(quals)
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(o) - 1
        st = np.zeros(max(n, 1), np.uint8); est = np.zeros(max(n, 1), np.float64)
        sm = np.zeros(max(n, 1), np.uint32); ss = np.zeros(max(n, 1), np.uint32)
        _check(self.L.savont_b200_twin_reads_from_snpmers(self.h, _p(b, C.c_uint8), _p(q, C.c_uint8), _p(o, C.c_uint64), C.c_size_t(n),
                                                          _p(st, C.c_uint8), _p(est, C.c_double), _p(sm, C.c_uint32), _p(ss, C.c_uint32)))
        return dict(status=st[:n], est_id=est[:n], n_solid_mini=sm[:n], n_solid_snp=ss[:n])

    def qual_seq(self, quals, offsets):
        """TwinRead.qual_seq (Seq<QualCompact3>) of a batch of reads: (words u64, word_offsets u64[n+1])"""
        q = _u8(quals); o = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(o) - 1
        lens = (o[1:] - o[:-1]).astype(np.int64)
        cap = int((((lens + 3) // 4 + 15) // 16).sum())
        wo = np.zeros(n + 1, np.uint64); words = np.zeros(max(cap, 1), np.uint64)
        _check(self.L.savont_b200_qual_seq(self.h, _p(q, C.c_uint8), _p(o, C.c_uint64), C.c_size_t(n), _p(wo, C.c_uint64),
                                           _p(words, C.c_uint64), C.c_size_t(cap)))
        return words[:cap], wo

    # ---- Stage 1 counting (K5) ----
    def _kmer_out(self, call, cap, sort, out=None):
        for _ in range(3):
            if out is not None and len(out[0]) >= cap and len(out[1]) >= 2 * cap:
                km, ct = out        # caller's buffers (e.g. page-locked: the D2H of the survivors then runs at PCIe speed)
            else:
                km = np.zeros(max(cap, 1), np.uint64); ct = np.zeros(2 * max(cap, 1), np.uint32)
            n = C.c_size_t(0); st = KmerStats()
            rc = call(_p(km, C.c_uint64), _p(ct, C.c_uint32), C.c_size_t(cap), C.byref(n), C.byref(st))
            if rc == ERR_CAPACITY and n.value > cap:
                cap = int(n.value) + 16
                continue
            _check(rc)
            break
        m = int(n.value)
        km = km[:m]; ct = ct[:2 * m].reshape(m, 2)
        if sort:                                      # the library's order is unspecified (like the reference's hash maps)
            o = np.argsort(km, kind="stable"); km = km[o]; ct = ct[o]
        return dict(kmers=km, counts=ct, stats=st.as_dict())

    def count_split_kmers(self, bases, quals, offsets, rc=None, single_strand=False, cap=None, sort=True, out=None):
        """seq_parse::read_to_split_kmers (counting part): -> dict(kmers, counts [n, 2], stats)"""
        b = _u8(bases); q = _u8(quals)
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        r = None if rc is None else np.ascontiguousarray(rc, dtype=np.uint8)
        n = len(o) - 1
        cap = cap or max(int(o[-1] - o[0]) // 64 + 4096, 4096)
        call = lambda km, ct, cp, nn, st: self.L.savont_b200_count_split_kmers(
            self.h, _p(b, C.c_uint8), _p(q, C.c_uint8), _p(o, C.c_uint64), _p(r, C.c_uint8), C.c_size_t(n),
            C.c_int(1 if single_strand else 0), km, ct, cp, nn, st)
        return self._kmer_out(call, cap, sort, out)

    def count_split_kmers_resident(self, single_strand=False, cap=1 << 22, sort=True):
        call = lambda km, ct, cp, nn, st: self.L.savont_b200_count_split_kmers_resident(
            self.h, C.c_int(1 if single_strand else 0), km, ct, cp, nn, st)
        return self._kmer_out(call, cap, sort)

    def load_asvs(self, bases, offsets):
        b = _u8(bases)
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        _check(self.L.savont_b200_load_asvs(self.h, _p(b, C.c_uint8), _p(o, C.c_uint64), C.c_size_t(len(o) - 1)))
        n = C.c_size_t()
        _check(self.L.savont_b200_num_asvs(self.h, C.byref(n)))
        self.n_asv = int(n.value)

    # ---- a3 ----
    def sketch(self, bases, quals, offsets, want_codes=True):
        b = _u8(bases); q = _u8(quals)
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(o) - 1
        total = int(o[-1] - o[0]) if n > 0 else 0
        spacing = (self.params.c - 1) // 2 + 1
        mcap = max(total // spacing + 2 * n + 16, 16)
        scap = max(total, 16)
        st = np.zeros(max(n, 1), np.uint8); est = np.zeros(max(n, 1), np.float64)
        mo = np.zeros(n + 1, np.uint64); so = np.zeros(n + 1, np.uint64)
        mpos = np.zeros(mcap, np.uint32); mk = np.zeros(mcap, np.uint64); mf = np.zeros(mcap, np.uint8)
        spos = np.zeros(scap, np.uint32); sk = np.zeros(scap, np.uint64)
        codes = np.zeros(max(int(o[-1]) if n > 0 else 0, 1), np.uint8) if want_codes else None
        _check(self.L.savont_b200_sketch(self.h, _p(b, C.c_uint8), _p(q, C.c_uint8), _p(o, C.c_uint64), C.c_size_t(n),
                                         _p(st, C.c_uint8), _p(est, C.c_double),
                                         _p(mo, C.c_uint64), _p(mpos, C.c_uint32), _p(mk, C.c_uint64), _p(mf, C.c_uint8), C.c_size_t(mcap),
                                         _p(so, C.c_uint64), _p(spos, C.c_uint32), _p(sk, C.c_uint64), C.c_size_t(scap),
                                         _p(codes, C.c_uint8)))
        nm, ns = int(mo[-1]), int(so[-1])
        out = dict(status=st[:n], est_id=est[:n], mini_offsets=mo, mini_pos=mpos[:nm], mini_kmer=mk[:nm], mini_fwd=mf[:nm],
                   snp_offsets=so, snp_pos=spos[:ns], snp_kmer=sk[:ns])
        if want_codes:
            out["codes"] = codes[:int(o[-1]) if n > 0 else 0]
        return out

    # ---- a8-a10 ----
    def find_candidates(self, bases, quals, offsets, cap=None):
        b = _u8(bases); q = _u8(quals)
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(o) - 1
        cap = cap or max(n * max(min(self.n_asv, 64), 1), 16)
        st = np.zeros(max(n, 1), np.uint8); co = np.zeros(n + 1, np.uint64)
        ca = np.zeros(cap, np.uint32); cm = np.zeros(cap, np.uint32); ci = np.zeros(cap, np.uint32); cs = np.zeros(cap, np.uint8)
        _check(self.L.savont_b200_find_candidates(self.h, _p(b, C.c_uint8), _p(q, C.c_uint8), _p(o, C.c_uint64), C.c_size_t(n),
                                                  _p(st, C.c_uint8), _p(co, C.c_uint64), _p(ca, C.c_uint32), _p(cm, C.c_uint32),
                                                  _p(ci, C.c_uint32), _p(cs, C.c_uint8), C.c_size_t(cap)))
        t = int(co[-1])
        return dict(status=st[:n], offsets=co, asv=ca[:t], mm=cm[:t], mini=ci[:t], strand=cs[:t])

    # ---- a11 ----
    def align_pairs(self, bases, offsets, pair_read, pair_asv, pair_strand, want_cigar=False, cigar_cap=None):
        b = _u8(bases)
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        pr = np.ascontiguousarray(pair_read, dtype=np.uint32)
        pa = np.ascontiguousarray(pair_asv, dtype=np.uint32)
        ps = np.ascontiguousarray(pair_strand, dtype=np.uint8)
        n = len(pr)
        nm = np.zeros(max(n, 1), np.int32); ae = np.zeros(max(n, 1), np.int32); re_ = np.zeros(max(n, 1), np.int32)
        as_ = np.zeros(max(n, 1), np.int32); rs = np.zeros(max(n, 1), np.int32)
        co = np.zeros(n + 1, np.uint64)
        if want_cigar and cigar_cap is None:
            lens = (o[1:] - o[:-1]).astype(np.int64)
            cigar_cap = int(lens[pr].sum() * 2 + 16 * n + 16) if n else 16
        cg = np.zeros(cigar_cap or 1, np.uint32)
        _check(self.L.savont_b200_align_pairs(self.h, _p(b, C.c_uint8), _p(o, C.c_uint64), C.c_size_t(len(o) - 1),
                                              _p(pr, C.c_uint32), _p(pa, C.c_uint32), _p(ps, C.c_uint8), C.c_size_t(n),
                                              _p(nm, C.c_int32), _p(ae, C.c_int32), _p(re_, C.c_int32),
                                              C.c_int(1 if want_cigar else 0), _p(as_, C.c_int32), _p(rs, C.c_int32),
                                              _p(co, C.c_uint64), _p(cg, C.c_uint32), C.c_size_t(len(cg))))
        out = dict(nm=nm[:n], asv_end=ae[:n], read_end=re_[:n])
        if want_cigar:
            out.update(asv_start=as_[:n], read_start=rs[:n], cigar_offsets=co, cigar=cg[:int(co[-1])])
        return out

    # ---- whole path ----
    # reuse_outputs = True: the output arrays of the map_* calls are kept and handed out again for calls of the same shape (a caller that
    # maps batch after batch into the same Vecs); the arrays of a result are then only valid until the next call
    reuse_outputs = False

    def _out_buffers(self, n, n_samples, class_cap):
        A = self.n_asv
        if self.reuse_outputs:
            key = (n, n_samples, class_cap, A)
            cache = self.__dict__.setdefault("_out_cache", {})
            if key not in cache:
                cache.clear()
                self.reuse_outputs = False
                try:
                    cache[key] = self._out_buffers(n, n_samples, class_cap)
                finally:
                    self.reuse_outputs = True
            B = cache[key]
            for k in ("unambig", "ambig", "leq10", "totals", "sample_singleton"):      # the library ADDS into the counter arrays
                B[k][:] = 0
            return B
        return dict(status=np.zeros(max(n, 1), np.uint8), best_nm=np.zeros(max(n, 1), np.int32),
                    class_offsets=np.zeros(n + 1, np.uint64), class_asv=np.zeros(max(class_cap, 1), np.uint32),
                    unambig=np.zeros(max(A, 1), np.uint64), ambig=np.zeros(max(A, 1), np.uint64), leq10=np.zeros(max(A, 1), np.uint64),
                    totals=np.zeros(2, np.uint64), sample_singleton=np.zeros(max(A * n_samples, 1), np.uint64))

    def map_reads(self, bases, quals, offsets, file_idx=None, n_samples=1, class_cap=None):
        b = _u8(bases); q = _u8(quals)
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(o) - 1
        fi = None if file_idx is None else np.ascontiguousarray(file_idx, dtype=np.uint32)
        class_cap = class_cap or max(8 * n + 1024, 16)
        B = self._out_buffers(n, n_samples, class_cap)
        st = Stats()
        _check(self.L.savont_b200_map_reads(self.h, _p(b, C.c_uint8), _p(q, C.c_uint8), _p(o, C.c_uint64), _p(fi, C.c_uint32),
                                            C.c_size_t(n), C.c_int(n_samples), _p(B["status"], C.c_uint8), _p(B["best_nm"], C.c_int32),
                                            _p(B["class_offsets"], C.c_uint64), _p(B["class_asv"], C.c_uint32), C.c_size_t(class_cap),
                                            _p(B["unambig"], C.c_uint64), _p(B["ambig"], C.c_uint64), _p(B["leq10"], C.c_uint64),
                                            _p(B["totals"], C.c_uint64), _p(B["sample_singleton"], C.c_uint64), C.byref(st)))
        return self._finish(B, n, n_samples, st)

    # ---- whole path, reads handed over as TwinRead fields (src/types.rs:386-412) ----
    @staticmethod
    def _twin_args(twin):
        w = np.ascontiguousarray(twin["dna_words"], dtype=np.uint64)
        wo = np.ascontiguousarray(twin["dna_word_offsets"], dtype=np.uint64)
        ln = np.ascontiguousarray(twin["lengths"], dtype=np.uint32)
        so = np.ascontiguousarray(twin["snpmer_offsets"], dtype=np.uint64)
        sk = twin["snpmer_kmers"]
        if sk.dtype == np.uint8:            # Vec<Kmer48> bytes
            sk = np.ascontiguousarray(sk); kb = 6
            assert len(sk) == 6 * int(so[-1])
        else:
            sk = np.ascontiguousarray(sk, dtype=np.uint64); kb = 8
        fi = twin.get("file_idx")
        fi = None if fi is None else np.ascontiguousarray(fi, dtype=np.uint32)
        return w, wo, ln, sk, kb, so, fi

    def map_twin_reads(self, twin, n_samples=1, class_cap=None):
        """twin: dict(dna_words u64, dna_word_offsets u64[n+1], lengths u32[n], snpmer_kmers u64 | u8 (Kmer48 bytes),
        snpmer_offsets u64[n+1], file_idx u32[n] | None) -- see savont_b200/twin.py"""
        w, wo, ln, sk, kb, so, fi = self._twin_args(twin)
        n = len(ln)
        class_cap = class_cap or max(8 * n + 1024, 16)
        B = self._out_buffers(n, n_samples, class_cap)
        st = Stats()
        _check(self.L.savont_b200_map_twin_reads(self.h, _p(w, C.c_uint64), _p(wo, C.c_uint64), _p(ln, C.c_uint32),
                                                 C.c_void_p(sk.ctypes.data), C.c_int(kb), _p(so, C.c_uint64), _p(fi, C.c_uint32),
                                                 C.c_size_t(n), C.c_int(n_samples), _p(B["status"], C.c_uint8), _p(B["best_nm"], C.c_int32),
                                                 _p(B["class_offsets"], C.c_uint64), _p(B["class_asv"], C.c_uint32), C.c_size_t(class_cap),
                                                 _p(B["unambig"], C.c_uint64), _p(B["ambig"], C.c_uint64), _p(B["leq10"], C.c_uint64),
                                                 _p(B["totals"], C.c_uint64), _p(B["sample_singleton"], C.c_uint64), C.byref(st)))
        return self._finish(B, n, n_samples, st)

    def set_host_threads(self, n):
        _check(self.L.savont_b200_set_host_threads(self.h, C.c_int(int(n))))

    def map_twin_read_ptrs(self, heap, n_samples=1, class_cap=None):
        """heap: dict(dna_ptrs u64[n], lengths u32[n], snpmer_ptrs u64[n], snpmer_counts u32[n], kmer_bytes, file_idx | None) --
        one pointer per read into caller-owned host memory (savont_b200/twin.py: scatter_to_heap)"""
        dp, ln, sp, sc, kb, fi = _heap_args(heap)
        n = len(ln)
        class_cap = class_cap or max(8 * n + 1024, 16)
        B = self._out_buffers(n, n_samples, class_cap)
        st = Stats()
        _check(self.L.savont_b200_map_twin_read_ptrs(self.h, C.c_void_p(dp.ctypes.data), _p(ln, C.c_uint32), C.c_void_p(sp.ctypes.data),
                                                     _p(sc, C.c_uint32), C.c_int(kb), _p(fi, C.c_uint32), C.c_size_t(n), C.c_int(n_samples),
                                                     _p(B["status"], C.c_uint8), _p(B["best_nm"], C.c_int32), _p(B["class_offsets"], C.c_uint64),
                                                     _p(B["class_asv"], C.c_uint32), C.c_size_t(class_cap), _p(B["unambig"], C.c_uint64),
                                                     _p(B["ambig"], C.c_uint64), _p(B["leq10"], C.c_uint64), _p(B["totals"], C.c_uint64),
                                                     _p(B["sample_singleton"], C.c_uint64), C.byref(st)))
        return self._finish(B, n, n_samples, st)

    def upload_twin_reads(self, twin, n_samples=1):
        w, wo, ln, sk, kb, so, fi = self._twin_args(twin)
        _check(self.L.savont_b200_upload_twin_reads(self.h, _p(w, C.c_uint64), _p(wo, C.c_uint64), _p(ln, C.c_uint32),
                                                    C.c_void_p(sk.ctypes.data), C.c_int(kb), _p(so, C.c_uint64), _p(fi, C.c_uint32),
                                                    C.c_size_t(len(ln)), C.c_int(n_samples)))
        self._resident = (len(ln), n_samples)

    def _finish(self, B, n, n_samples, st):
        A = self.n_asv
        t = int(B["class_offsets"][-1])
        return dict(status=B["status"][:n], best_nm=B["best_nm"][:n], class_offsets=B["class_offsets"], class_asv=B["class_asv"][:t],
                    unambig=B["unambig"][:A], ambig=B["ambig"][:A], leq10=B["leq10"][:A], totals=B["totals"],
                    sample_singleton=B["sample_singleton"][:A * n_samples].reshape(n_samples, A), stats=st.as_dict())

    # ---- device-resident variant ----
    def upload_reads(self, bases, quals, offsets, file_idx=None, n_samples=1):
        b = _u8(bases); q = _u8(quals)
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        fi = None if file_idx is None else np.ascontiguousarray(file_idx, dtype=np.uint32)
        _check(self.L.savont_b200_upload_reads(self.h, _p(b, C.c_uint8), _p(q, C.c_uint8), _p(o, C.c_uint64), _p(fi, C.c_uint32),
                                               C.c_size_t(len(o) - 1), C.c_int(n_samples)))
        self._resident = (len(o) - 1, n_samples)

    def set_resident_parts(self, n_parts):
        _check(self.L.savont_b200_set_resident_parts(self.h, C.c_int(int(n_parts))))

    def map_resident(self, zero_counters=True):
        st = Stats()
        _check(self.L.savont_b200_map_resident(self.h, C.c_int(1 if zero_counters else 0), C.byref(st)))
        return st.as_dict()

    def fetch_resident(self, class_cap=None):
        n, n_samples = self._resident
        class_cap = class_cap or max(8 * n + 1024, 16)
        B = self._out_buffers(n, n_samples, class_cap)
        _check(self.L.savont_b200_fetch_resident(self.h, _p(B["status"], C.c_uint8), _p(B["best_nm"], C.c_int32),
                                                 _p(B["class_offsets"], C.c_uint64), _p(B["class_asv"], C.c_uint32), C.c_size_t(class_cap)))
        _check(self.L.savont_b200_read_counts(self.h, _p(B["unambig"], C.c_uint64), _p(B["ambig"], C.c_uint64), _p(B["leq10"], C.c_uint64),
                                              _p(B["totals"], C.c_uint64), _p(B["sample_singleton"], C.c_uint64)))
        return self._finish(B, n, n_samples, Stats())

    def best_hits_resident(self, cigar_cap=None, cigar_out=None):
        """best-hit alignment record (ASV, strand, coordinates, identity, CIGAR) per read of the resident batch.  cigar_out: a caller's
        uint32 array for the CIGAR ops (e.g. a view of pinned_array memory; its length is the capacity)"""
        n, _ = self._resident
        m = max(n, 1)
        out = dict(best_asv=np.zeros(m, np.uint32), strand=np.zeros(m, np.uint8), asv_start=np.zeros(m, np.int32),
                   asv_end=np.zeros(m, np.int32), read_start=np.zeros(m, np.int32), read_end=np.zeros(m, np.int32),
                   identity=np.zeros(m, np.float64))
        co = np.zeros(n + 1, np.uint64)
        cg = cigar_out if cigar_out is not None else np.empty(cigar_cap or max(64 * n, 1024), np.uint32)
        _check(self.L.savont_b200_best_hits_resident(self.h, _p(out["best_asv"], C.c_uint32), _p(out["strand"], C.c_uint8),
                                                     _p(out["asv_start"], C.c_int32), _p(out["asv_end"], C.c_int32),
                                                     _p(out["read_start"], C.c_int32), _p(out["read_end"], C.c_int32),
                                                     _p(out["identity"], C.c_double), _p(co, C.c_uint64), _p(cg, C.c_uint32),
                                                     C.c_size_t(len(cg))))
        out = {k: v[:n] for k, v in out.items()}
        out.update(cigar_offsets=co, cigar=cg[:int(co[-1])])
        return out

    def mapping_rows_resident(self, max_rows=5):
        """rows of temp/read_to_asv_mappings.tsv per read of the resident batch: dict(row_offsets, asv, nm, snp_mismatches)"""
        n, _ = self._resident
        cap = max(n * max_rows, 1)
        ro = np.zeros(n + 1, np.uint64); ra = np.zeros(cap, np.uint32); rn = np.zeros(cap, np.int32); rm = np.zeros(cap, np.uint32)
        _check(self.L.savont_b200_mapping_rows_resident(self.h, C.c_uint32(max_rows), _p(ro, C.c_uint64), _p(ra, C.c_uint32), _p(rn, C.c_int32),
                                                        _p(rm, C.c_uint32), C.c_size_t(cap)))
        t = int(ro[-1])
        return dict(row_offsets=ro, asv=ra[:t], nm=rn[:t], snp_mismatches=rm[:t])

    def consensus_pileups(self, cons_bases, cons_offsets, read_bases, read_quals, read_offsets, read_cons, use_hpc=True, cap=None,
                          entries_out=None):
        """Stage 4 pileups (generate_consensus_pileups, src/alignment.rs:415-641) -> dict(pos_offsets, entries
        (PILEUP_ENTRY records), ref_hp_length, read_strand, read_nm).  entries_out: caller-provided PILEUP_ENTRY array
        (e.g. a view of pinned_array memory), its length is the capacity"""
        cb = _u8(cons_bases); co = np.ascontiguousarray(cons_offsets, dtype=np.uint64)
        rb = _u8(read_bases); rq = None if read_quals is None else _u8(read_quals)
        ro = np.ascontiguousarray(read_offsets, dtype=np.uint64)
        rc = np.ascontiguousarray(read_cons, dtype=np.uint32)
        n = len(ro) - 1
        npos = int(co[-1] - co[0])
        cap = cap if cap is not None else int((ro[-1] - ro[0]) + (co[1:] - co[:-1])[rc].sum() + 64)
        if entries_out is not None:
            ent, cap = entries_out, len(entries_out)
        else:
            ent = np.zeros(max(cap, 1), PILEUP_ENTRY)
        po = np.zeros(npos + 1, np.uint64)
        hp = np.zeros(max(npos, 1), np.uint8); st = np.zeros(max(n, 1), np.uint8); nm = np.zeros(max(n, 1), np.int32)
        ne = C.c_uint64(0); stt = Stats()
        _check(self.L.savont_b200_consensus_pileups(self.h, C.c_int(1 if use_hpc else 0), _p(cb, C.c_uint8), _p(co, C.c_uint64),
                                                    C.c_size_t(len(co) - 1), _p(rb, C.c_uint8), _p(rq, C.c_uint8), _p(ro, C.c_uint64),
                                                    _p(rc, C.c_uint32), C.c_size_t(n), _p(po, C.c_uint64), C.c_void_p(ent.ctypes.data),
                                                    C.c_size_t(cap), C.byref(ne), _p(hp, C.c_uint8), _p(st, C.c_uint8), _p(nm, C.c_int32),
                                                    C.byref(stt)))
        return dict(pos_offsets=po, entries=ent[:int(ne.value)], ref_hp_length=hp[:npos], read_strand=st[:n], read_nm=nm[:n],
                    stats=stt.as_dict())

    def lsh_signatures(self, kmers, offsets):
        """TwinRead::compute_lsh_signatures (src/types.rs:719-747) -> (sig [n, 20] u64, valid [n] u8)"""
        km = np.ascontiguousarray(kmers, dtype=np.uint64); off = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(off) - 1
        sig = np.zeros((max(n, 1), 20), np.uint64); valid = np.zeros(max(n, 1), np.uint8)
        _check(self.L.savont_b200_lsh_signatures(self.h, _p(km, C.c_uint64), _p(off, C.c_uint64), C.c_size_t(n), _p(sig, C.c_uint64),
                                                 _p(valid, C.c_uint8)))
        return sig[:n], valid[:n]

    def cluster_reads_by_kmers(self, kmers, offsets, threshold=0.95):
        """greedy loop of asv_cluster::cluster_reads_by_kmers (src/asv_cluster.rs:72-190) -> dict(rep_of, n_reps, stats)"""
        km = np.ascontiguousarray(kmers, dtype=np.uint64); off = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(off) - 1
        rep = np.zeros(max(n, 1), np.uint32); nr = C.c_uint64(0); stt = Stats()
        _check(self.L.savont_b200_cluster_reads_by_kmers(self.h, _p(km, C.c_uint64), _p(off, C.c_uint64), C.c_size_t(n),
                                                         C.c_double(threshold), _p(rep, C.c_uint32), C.byref(nr), C.byref(stt)))
        return dict(rep_of=rep[:n], n_reps=int(nr.value), stats=stt.as_dict())

    # ---- sharded Stage 1 counting (one context per GPU; savont_b200/dist.py drives the phases) ----
    def kmer_shard_hist(self):
        P = self.L.savont_b200_kmer_shard_partitions()
        hist = np.zeros(P, np.uint32)
        _check(self.L.savont_b200_kmer_shard_hist(self.h, _p(hist, C.c_uint32)))
        return hist

    def kmer_shard_alloc(self, n_entries: int, with_capacity=False):
        ptr = C.c_void_p(); cap = C.c_uint64(0)
        _check(self.L.savont_b200_kmer_shard_alloc(self.h, C.c_uint64(int(n_entries)), C.byref(ptr), C.byref(cap)))
        return (int(ptr.value), int(cap.value)) if with_capacity else int(ptr.value)

    def ipc_export(self, dev_ptr: int) -> bytes:
        buf = (C.c_uint8 * 64)()
        _check(self.L.savont_b200_ipc_export(self.h, C.c_void_p(dev_ptr), buf))
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        ptr = C.c_void_p(); buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        _check(self.L.savont_b200_ipc_open(self.h, buf, C.byref(ptr)))
        return int(ptr.value)

    def ipc_close(self, dev_ptr: int):
        _check(self.L.savont_b200_ipc_close(self.h, C.c_void_p(dev_ptr)))

    def kmer_shard_scatter(self, peer_entries, dst_off):
        n = len(peer_entries)
        arr = (C.c_void_p * n)(*[C.c_void_p(int(x)) for x in peer_entries])
        do = np.ascontiguousarray(dst_off, dtype=np.uint64)
        _check(self.L.savont_b200_kmer_shard_scatter(self.h, C.c_int(n), arr, _p(do, C.c_uint64)))

    def kmer_shard_count(self, part_off, single_strand=False, cap=1 << 22, sort=True, fetch=True):
        po = np.ascontiguousarray(part_off, dtype=np.uint64)
        if not fetch:                                   # count only: the number of survivors and the stage times
            n = C.c_size_t(0); st = KmerStats()
            _check(self.L.savont_b200_kmer_shard_count(self.h, _p(po, C.c_uint64), C.c_int(1 if single_strand else 0), None, None,
                                                       C.c_size_t(0), C.byref(n), C.byref(st)))
            return dict(n=int(n.value), stats=st.as_dict())
        call = lambda km, ct, cp, nn, st: self.L.savont_b200_kmer_shard_count(
            self.h, _p(po, C.c_uint64), C.c_int(1 if single_strand else 0), km, ct, cp, nn, st)
        return self._kmer_out(call, cap, sort)

    def snpmer_compare(self, cons_kmers, cons_offsets, read_kmers, read_offsets, want_matrix=False):
        """reassign_reads_to_best_cluster / compare_consensus_snpmers (src/asv_cluster.rs:968-1098) ->
        dict(best_cons, best_mismatches, best_matches[, matrix [n_reads, n_cons, 2]], stats)"""
        ck = np.ascontiguousarray(cons_kmers, dtype=np.uint64); co = np.ascontiguousarray(cons_offsets, dtype=np.uint64)
        rk = np.ascontiguousarray(read_kmers, dtype=np.uint64); ro = np.ascontiguousarray(read_offsets, dtype=np.uint64)
        nc, n = len(co) - 1, len(ro) - 1
        bc = np.zeros(max(n, 1), np.uint32); bmm = np.zeros(max(n, 1), np.uint32); bm = np.zeros(max(n, 1), np.uint32)
        mat = np.zeros((max(n, 1), max(nc, 1), 2), np.uint16) if want_matrix else None
        stt = Stats()
        _check(self.L.savont_b200_snpmer_compare(self.h, _p(ck, C.c_uint64), _p(co, C.c_uint64), C.c_size_t(nc), _p(rk, C.c_uint64),
                                                 _p(ro, C.c_uint64), C.c_size_t(n), _p(bc, C.c_uint32), _p(bmm, C.c_uint32),
                                                 _p(bm, C.c_uint32), _p(mat, C.c_uint16), C.byref(stt)))
        out = dict(best_cons=bc[:n], best_mismatches=bmm[:n], best_matches=bm[:n], stats=stt.as_dict())
        if want_matrix:
            out["matrix"] = mat[:n, :nc]
        return out

    def cluster_reads_by_snpmers(self, snp_kmers, snp_offsets, cluster_reads, cluster_offsets, iterative_above=1000):
        """per-cluster loop of asv_cluster::cluster_reads_by_snpmers (src/asv_cluster.rs:561-700) -> dict(rep_of, n_representatives, stats)"""
        km = np.ascontiguousarray(snp_kmers, dtype=np.uint64); so = np.ascontiguousarray(snp_offsets, dtype=np.uint64)
        cr = np.ascontiguousarray(cluster_reads, dtype=np.uint32); co = np.ascontiguousarray(cluster_offsets, dtype=np.uint64)
        n, nc = len(so) - 1, len(co) - 1
        rep = np.zeros(max(n, 1), np.uint32); nr = C.c_uint64(0); st = Stats()
        _check(self.L.savont_b200_cluster_reads_by_snpmers(self.h, _p(km, C.c_uint64), _p(so, C.c_uint64), C.c_size_t(n), _p(cr, C.c_uint32),
                                                           _p(co, C.c_uint64), C.c_size_t(nc), C.c_uint32(int(iterative_above)),
                                                           _p(rep, C.c_uint32), C.byref(nr), C.byref(st)))
        return dict(rep_of=rep[:n], n_representatives=int(nr.value), stats=st.as_dict())

    def counts_dev(self):
        """(device pointer, length in u64) of [unambig|ambig|leq10|assigned|filtered|singleton S*A]"""
        ptr = C.c_void_p(); n = C.c_size_t()
        _check(self.L.savont_b200_counts_dev(self.h, C.byref(ptr), C.byref(n)))
        return int(ptr.value), int(n.value)


class MultiContext:
    """One process, several GPUs (savont_b200_create_multi): a context per device, reads sharded in read order, counters summed
    across the devices, results concatenated.  Same call order and result dictionaries as Context."""

    def __init__(self, devices, params: Params | None = None, allreduce: str = "peer"):
        self.L = lib()
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        _check(self.L.savont_b200_create_multi(arr, C.c_int(len(self.devices)), C.byref(h)))
        self.h = h
        self.n_asv = 0
        if params is not None:
            self.set_params(params)
        if allreduce != "peer":
            self.set_allreduce(allreduce)

    def close(self):
        if getattr(self, "h", None):
            self.L.savont_b200_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, p: Params):
        _check(self.L.savont_b200_multi_set_params(self.h, C.byref(p)))

    def set_allreduce(self, mode):
        _check(self.L.savont_b200_multi_set_allreduce(self.h, C.c_int({"peer": 0, "nccl": 1}[mode])))

    def set_host_threads(self, n):
        _check(self.L.savont_b200_multi_set_host_threads(self.h, C.c_int(int(n))))

    def set_batch_reads(self, n):
        _check(self.L.savont_b200_multi_set_batch_reads(self.h, C.c_size_t(n)))

    def load_snpmers(self, kmers):
        k = np.ascontiguousarray(kmers, dtype=np.uint64)
        _check(self.L.savont_b200_multi_load_snpmers(self.h, _p(k, C.c_uint64), C.c_size_t(len(k))))

    def load_asvs(self, bases, offsets):
        b = _u8(bases); o = np.ascontiguousarray(offsets, dtype=np.uint64)
        _check(self.L.savont_b200_multi_load_asvs(self.h, _p(b, C.c_uint8), _p(o, C.c_uint64), C.c_size_t(len(o) - 1)))
        self.n_asv = len(o) - 1

    reuse_outputs = False
    _out_buffers = Context._out_buffers
    _finish = Context._finish

    def _outs(self, B, class_cap, st):
        return (_p(B["status"], C.c_uint8), _p(B["best_nm"], C.c_int32), _p(B["class_offsets"], C.c_uint64), _p(B["class_asv"], C.c_uint32),
                C.c_size_t(class_cap), _p(B["unambig"], C.c_uint64), _p(B["ambig"], C.c_uint64), _p(B["leq10"], C.c_uint64),
                _p(B["totals"], C.c_uint64), _p(B["sample_singleton"], C.c_uint64), C.byref(st))

    def map_reads(self, bases, quals, offsets, file_idx=None, n_samples=1, class_cap=None):
        b = _u8(bases); q = _u8(quals); o = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(o) - 1
        fi = None if file_idx is None else np.ascontiguousarray(file_idx, dtype=np.uint32)
        class_cap = class_cap or max(8 * n + 1024, 16)
        B = self._out_buffers(n, n_samples, class_cap); st = Stats()
        _check(self.L.savont_b200_multi_map_reads(self.h, _p(b, C.c_uint8), _p(q, C.c_uint8), _p(o, C.c_uint64), _p(fi, C.c_uint32),
                                                  C.c_size_t(n), C.c_int(n_samples), *self._outs(B, class_cap, st)))
        return self._finish(B, n, n_samples, st)

    def map_twin_reads(self, twin, n_samples=1, class_cap=None):
        w, wo, ln, sk, kb, so, fi = Context._twin_args(twin)
        n = len(ln)
        class_cap = class_cap or max(8 * n + 1024, 16)
        B = self._out_buffers(n, n_samples, class_cap); st = Stats()
        _check(self.L.savont_b200_multi_map_twin_reads(self.h, _p(w, C.c_uint64), _p(wo, C.c_uint64), _p(ln, C.c_uint32),
                                                       C.c_void_p(sk.ctypes.data), C.c_int(kb), _p(so, C.c_uint64), _p(fi, C.c_uint32),
                                                       C.c_size_t(n), C.c_int(n_samples), *self._outs(B, class_cap, st)))
        return self._finish(B, n, n_samples, st)

    def map_twin_read_ptrs(self, heap, n_samples=1, class_cap=None):
        dp, ln, sp, sc, kb, fi = _heap_args(heap)
        n = len(ln)
        class_cap = class_cap or max(8 * n + 1024, 16)
        B = self._out_buffers(n, n_samples, class_cap); st = Stats()
        _check(self.L.savont_b200_multi_map_twin_read_ptrs(self.h, C.c_void_p(dp.ctypes.data), _p(ln, C.c_uint32), C.c_void_p(sp.ctypes.data),
                                                           _p(sc, C.c_uint32), C.c_int(kb), _p(fi, C.c_uint32), C.c_size_t(n),
                                                           C.c_int(n_samples), *self._outs(B, class_cap, st)))
        return self._finish(B, n, n_samples, st)
