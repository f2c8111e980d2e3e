"""ctypes binding of the CPU oracle (oracle/libsavont_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under savont_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsavont_oracle.so")

STATUS_ASSIGNED, STATUS_NO_CANDIDATE, STATUS_RATIO_FILTERED, STATUS_NO_ALIGNMENT, STATUS_TOO_SHORT = 0, 1, 2, 3, 4
STATUS_LENGTH_FILTERED, STATUS_REPETITIVE, STATUS_LOW_QUALITY = 5, 6, 7


class Params(C.Structure):
    _fields_ = [
        ("k", C.c_int32), ("c", C.c_int32), ("l", C.c_int32), ("min_bq", C.c_int32),
        ("band", C.c_int32), ("strand_mode", C.c_int32), ("candidate_mode", C.c_int32), ("nm_mode", C.c_int32),
        ("ratio_thresh", C.c_double), ("min_mini_frac_base", C.c_double),
    ]


class ReadFilter(C.Structure):
    """a4 options (so_read_filter): the Cli fields / constants kmer_comp::twin_reads_from_snpmers reads"""
    _fields_ = [("min_read_length", C.c_uint32), ("max_read_length", C.c_uint32), ("max_kmer_count_in_read", C.c_uint32),
                ("reserved", C.c_uint32), ("quality_value_cutoff", C.c_double)]


def read_filter(min_read_length=1100, max_read_length=2000, quality_value_cutoff=98.0, max_kmer_count_in_read=500) -> ReadFilter:
    """defaults of src/cli.rs:87-96 and MAX_KMER_COUNT_IN_READ (src/constants.rs)"""
    return ReadFilter(min_read_length, max_read_length, max_kmer_count_in_read, 0, quality_value_cutoff)


def default_params(**kw) -> Params:
    p = Params(17, 11, 3, 25, 64, 0, 0, 0, 0.005, 0.95)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "savont_oracle.c")
    hdr = os.path.join(_HERE, "savont_oracle.h")
    if os.path.exists(src) and (force or not os.path.exists(_LIB_PATH)
                                or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def _p(arr, typ):
    if arr is None:
        return None
    return arr.ctypes.data_as(C.POINTER(typ))


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.so_encode_base.restype = C.c_uint8
        L.so_encode_base.argtypes = [C.c_uint8]
        L.so_mm_hash64.restype = C.c_uint64
        L.so_mm_hash64.argtypes = [C.c_uint64]
        L.so_kmer_from_ascii.restype = C.c_uint64
        L.so_kmer_from_ascii.argtypes = [C.c_char_p, C.c_int]
        L.so_kmer_revcomp.restype = C.c_uint64
        L.so_kmer_revcomp.argtypes = [C.c_uint64, C.c_int]
        L.so_kmer48_from_u64.restype = None
        L.so_kmer48_from_u64.argtypes = [C.c_uint64, C.POINTER(C.c_uint8)]
        L.so_kmer48_to_u64.restype = C.c_uint64
        L.so_kmer48_to_u64.argtypes = [C.POINTER(C.c_uint8)]
        L.so_powi.restype = C.c_double
        L.so_powi.argtypes = [C.c_double, C.c_int]
        L.so_split_mask.restype = C.c_uint64
        L.so_split_mask.argtypes = [C.c_int]
        L.so_create.restype = C.c_void_p
        L.so_create.argtypes = [C.POINTER(Params), C.POINTER(C.c_uint64), C.c_size_t]
        L.so_destroy.restype = None
        L.so_destroy.argtypes = [C.c_void_p]
        L.so_load_asvs.restype = C.c_int
        L.so_load_asvs.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.c_size_t]
        L.so_num_asvs.restype = C.c_size_t
        L.so_num_asvs.argtypes = [C.c_void_p]
        L.so_sketch_batch.restype = C.c_int
        L.so_find_candidates.restype = C.c_int
        L.so_align_pairs.restype = C.c_int
        L.so_map_reads.restype = C.c_int
        L.so_em.restype = C.c_int
        L.so_twin_reads_filter.restype = C.c_int
        L.so_count_split_kmers.restype = C.c_longlong
        L.so_consensus_pileups.restype = C.c_longlong
        L.so_high_freq_kmers.restype = C.c_uint32
        L.so_set_read_filter.restype = None
        L.so_nm_pair.restype = C.c_int
        L.so_nm_pair.argtypes = [C.POINTER(Params), C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_int32)]
        _lib = L
    return _lib


def as_u8(x) -> np.ndarray:
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8).copy()
    return np.ascontiguousarray(x, dtype=np.uint8)


def concat(seqs):
    """list of bytes -> (uint8 concatenation, uint64 offsets[n+1])"""
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if len(seqs):
        offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(bytes(s) for s in seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
    return buf, offs


def nm_pair(params: Params, asv: bytes, read: bytes):
    """NM of one pair (both given in alignment orientation, ASCII) under params.nm_mode -> (has_alignment, nm); a11 sensitivity study"""
    L = lib()
    a = np.array([L.so_encode_base(b) for b in bytes(asv)], np.uint8)
    r = np.array([L.so_encode_base(b) for b in bytes(read)], np.uint8)
    nm = C.c_int32(0)
    ok = L.so_nm_pair(C.byref(params), _p(a, C.c_uint8), C.c_int(len(a)), _p(r, C.c_uint8), C.c_int(len(r)), C.byref(nm))
    return bool(ok), int(nm.value)


def mm_hash64(x: int) -> int:
    return int(lib().so_mm_hash64(C.c_uint64(x)))


def sketch_batch(params: Params, snpmers, bases, quals, offsets):
    L = lib()
    snp = np.ascontiguousarray(snpmers, dtype=np.uint64)
    bases = as_u8(bases)
    quals = None if quals is None else as_u8(quals)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    total = int(offsets[-1]) if n >= 0 else 0
    cap = max(total, 1)
    status = np.zeros(max(n, 1), np.uint8)
    est = np.zeros(max(n, 1), np.float64)
    mo = np.zeros(n + 1, np.uint64)
    so = np.zeros(n + 1, np.uint64)
    mpos = np.zeros(cap, np.uint32); mk = np.zeros(cap, np.uint64); mf = np.zeros(cap, np.uint8)
    spos = np.zeros(cap, np.uint32); sk = np.zeros(cap, np.uint64)
    codes = np.zeros(cap, np.uint8)
    rc = L.so_sketch_batch(C.byref(params), _p(snp, C.c_uint64), C.c_size_t(len(snp)),
                           _p(bases, C.c_uint8), _p(quals, C.c_uint8), _p(offsets, C.c_uint64), C.c_size_t(n),
                           _p(status, C.c_uint8), _p(est, C.c_double),
                           _p(mo, C.c_uint64), _p(mpos, C.c_uint32), _p(mk, C.c_uint64), _p(mf, C.c_uint8), C.c_size_t(cap),
                           _p(so, C.c_uint64), _p(spos, C.c_uint32), _p(sk, C.c_uint64), C.c_size_t(cap),
                           _p(codes, C.c_uint8))
    assert rc == 0, rc
    nm, ns = int(mo[-1]), int(so[-1])
    return dict(status=status[:n], est_id=est[:n], mini_offsets=mo, mini_pos=mpos[:nm], mini_kmer=mk[:nm],
                mini_fwd=mf[:nm], snp_offsets=so, snp_pos=spos[:ns], snp_kmer=sk[:ns], codes=codes[:total])


def concat_ids(ids):
    """list of str/bytes read ids -> (uint8 concatenation, uint64 offsets[n+1])"""
    return concat([i.encode() if isinstance(i, str) else bytes(i) for i in ids])


def twin_reads_filter(params: Params, rf: ReadFilter, snpmers, high_freq, bases, quals, offsets, ids=None):
    """a4: kmer_comp::twin_reads_from_snpmers + the read order of src/main.rs:538 (so_twin_reads_filter)"""
    L = lib()
    snp = np.ascontiguousarray(snpmers, dtype=np.uint64)
    hf = np.ascontiguousarray(high_freq if high_freq is not None else [], dtype=np.uint64)
    bases = as_u8(bases)
    quals = None if quals is None else as_u8(quals)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    idb, ido = (None, None) if ids is None else concat_ids(ids)
    st = np.zeros(max(n, 1), np.uint8); est = np.zeros(max(n, 1), np.float64)
    sm = np.zeros(max(n, 1), np.uint32); ss = np.zeros(max(n, 1), np.uint32); order = np.zeros(max(n, 1), np.uint32)
    nk = C.c_size_t(0)
    rc = L.so_twin_reads_filter(C.byref(params), C.byref(rf), _p(snp, C.c_uint64), C.c_size_t(len(snp)), _p(hf, C.c_uint64),
                                C.c_size_t(len(hf)), _p(bases, C.c_uint8), _p(quals, C.c_uint8), _p(offsets, C.c_uint64), C.c_size_t(n),
                                _p(idb, C.c_uint8), _p(ido, C.c_uint64), _p(st, C.c_uint8), _p(est, C.c_double),
                                _p(sm, C.c_uint32), _p(ss, C.c_uint32), _p(order, C.c_uint32), C.byref(nk))
    assert rc == 0, rc
    return dict(status=st[:n], est_id=est[:n], n_solid_mini=sm[:n], n_solid_snp=ss[:n], order=order[:int(nk.value)])


def count_split_kmers(bases, quals, offsets, rc=None, k=17, min_bq=25, single_strand=False, n_threads=0, cap=None):
    """Stage 1 counting (so_count_split_kmers): -> dict(kmers u64 sorted, counts u32 [n,2], n_distinct)"""
    L = lib()
    bases = as_u8(bases); quals = None if quals is None else as_u8(quals)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    rcf = None if rc is None else np.ascontiguousarray(rc, dtype=np.uint8)
    cap = cap or max(int(offsets[-1] - offsets[0]) // 2 + 16, 16)
    km = np.zeros(cap, np.uint64); ct = np.zeros(2 * cap, np.uint32)
    nd = C.c_uint64(0)
    m = L.so_count_split_kmers(C.c_int(k), C.c_int(min_bq), C.c_int(1 if single_strand else 0), _p(bases, C.c_uint8), _p(quals, C.c_uint8),
                               _p(offsets, C.c_uint64), _p(rcf, C.c_uint8), C.c_size_t(n), _p(km, C.c_uint64), _p(ct, C.c_uint32),
                               C.c_size_t(cap), C.byref(nd), C.c_int(n_threads))
    assert m >= 0, m
    return dict(kmers=km[:m], counts=ct[:2 * m].reshape(m, 2), n_distinct=int(nd.value))


def high_freq_kmers(kmers, counts):
    """src/kmer_comp.rs:459-500 -> (high_freq_thresh, high-frequency k-mers in input order)"""
    L = lib()
    km = np.ascontiguousarray(kmers, np.uint64); ct = np.ascontiguousarray(counts, np.uint32).reshape(-1)
    n = len(km)
    assert n > 0
    out = np.zeros(n, np.uint64); nh = C.c_size_t(0)
    th = L.so_high_freq_kmers(_p(km, C.c_uint64), _p(ct, C.c_uint32), C.c_size_t(n), _p(out, C.c_uint64), C.byref(nh))
    return int(th), out[:int(nh.value)]


PILEUP_ENTRY = np.dtype([("read", np.uint32), ("kind", np.uint8), ("n", np.uint8), ("base", np.uint8, 2), ("qual", np.uint8, 2),
                         ("hp", np.uint8, 2)])      # so_pileup_entry / savont_b200_pileup_entry, 12 bytes
assert PILEUP_ENTRY.itemsize == 12


def consensus_pileups(cons_bases, cons_offsets, read_bases, read_quals, read_offsets, read_cons, band=64, use_hpc=True, cap=None):
    """Stage 4 pileups (so_consensus_pileups) -> dict(pos_offsets, entries (PILEUP_ENTRY), ref_hp_length, read_strand, read_nm)"""
    L = lib()
    cb = as_u8(cons_bases); co = np.ascontiguousarray(cons_offsets, dtype=np.uint64)
    rb = as_u8(read_bases); rq = None if read_quals is None else as_u8(read_quals)
    ro = np.ascontiguousarray(read_offsets, dtype=np.uint64)
    rc_ = np.ascontiguousarray(read_cons, dtype=np.uint32)
    n = len(ro) - 1
    npos = int(co[-1] - co[0])
    cap = cap or int((ro[-1] - ro[0]) + sum(int(co[ci + 1] - co[ci]) for ci in rc_) + npos + 64)
    po = np.zeros(npos + 1, np.uint64); ent = np.zeros(cap, PILEUP_ENTRY)
    hp = np.zeros(max(npos, 1), np.uint8); st = np.zeros(max(n, 1), np.uint8); nm = np.zeros(max(n, 1), np.int32)
    m = L.so_consensus_pileups(C.c_int(band), C.c_int(1 if use_hpc else 0), _p(cb, C.c_uint8), _p(co, C.c_uint64), C.c_size_t(len(co) - 1),
                               _p(rb, C.c_uint8), _p(rq, C.c_uint8), _p(ro, C.c_uint64), _p(rc_, C.c_uint32), C.c_size_t(n),
                               _p(po, C.c_uint64), C.c_void_p(ent.ctypes.data), C.c_size_t(cap), _p(hp, C.c_uint8), _p(st, C.c_uint8),
                               _p(nm, C.c_int32))
    assert m >= 0, m
    return dict(pos_offsets=po, entries=ent[:m], ref_hp_length=hp[:npos], read_strand=st[:n], read_nm=nm[:n])


def lsh_signatures(kmers, offsets):
    """so_lsh_signatures -> (sig [n, 20] u64, valid [n] u8)"""
    L = lib()
    km = np.ascontiguousarray(kmers, dtype=np.uint64); off = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(off) - 1
    sig = np.zeros((max(n, 1), 20), np.uint64); valid = np.zeros(max(n, 1), np.uint8)
    rc = L.so_lsh_signatures(_p(km, C.c_uint64), _p(off, C.c_uint64), C.c_size_t(n), _p(sig, C.c_uint64), _p(valid, C.c_uint8))
    assert rc == 0
    return sig[:n], valid[:n]


def cluster_reads_by_kmers(kmers, offsets, k=17, threshold=0.95):
    """so_cluster_reads_by_kmers -> rep_of [n] u32"""
    L = lib()
    km = np.ascontiguousarray(kmers, dtype=np.uint64); off = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(off) - 1
    rep = np.zeros(max(n, 1), np.uint32); nr = C.c_uint64(0)
    rc = L.so_cluster_reads_by_kmers(C.c_int(k), C.c_double(threshold), _p(km, C.c_uint64), _p(off, C.c_uint64), C.c_size_t(n),
                                     _p(rep, C.c_uint32), C.byref(nr))
    assert rc == 0
    return rep[:n]


def snpmer_compare(cons_kmers, cons_offsets, read_kmers, read_offsets, k=17, want_matrix=False):
    """so_snpmer_compare -> dict(best_cons, best_mismatches, best_matches[, matrix [n_reads, n_cons, 2]])"""
    L = lib()
    ck = np.ascontiguousarray(cons_kmers, dtype=np.uint64); co = np.ascontiguousarray(cons_offsets, dtype=np.uint64)
    rk = np.ascontiguousarray(read_kmers, dtype=np.uint64); ro = np.ascontiguousarray(read_offsets, dtype=np.uint64)
    nc, n = len(co) - 1, len(ro) - 1
    bc = np.zeros(max(n, 1), np.uint32); bmm = np.zeros(max(n, 1), np.uint32); bm = np.zeros(max(n, 1), np.uint32)
    mat = np.zeros((max(n, 1), max(nc, 1), 2), np.uint16) if want_matrix else None
    rc = L.so_snpmer_compare(C.c_int(k), _p(ck, C.c_uint64), _p(co, C.c_uint64), C.c_size_t(nc), _p(rk, C.c_uint64), _p(ro, C.c_uint64),
                             C.c_size_t(n), _p(bc, C.c_uint32), _p(bmm, C.c_uint32), _p(bm, C.c_uint32), _p(mat, C.c_uint16))
    assert rc == 0
    out = dict(best_cons=bc[:n], best_mismatches=bmm[:n], best_matches=bm[:n])
    if want_matrix:
        out["matrix"] = mat[:n, :nc]
    return out


def cluster_reads_by_snpmers(snp_kmers, snp_offsets, cluster, k=17, iterative_above=1000):
    """so_cluster_reads_by_snpmers for one k-mer cluster -> rep_of (u32 per read id; UINT32_MAX outside the cluster)"""
    L = lib()
    sk = np.ascontiguousarray(snp_kmers, dtype=np.uint64); so = np.ascontiguousarray(snp_offsets, dtype=np.uint64)
    cl = np.ascontiguousarray(cluster, dtype=np.uint32)
    rep = np.full(max(len(so) - 1, 1), 0xffffffff, np.uint32)
    rc = L.so_cluster_reads_by_snpmers(C.c_int(k), _p(sk, C.c_uint64), _p(so, C.c_uint64), _p(cl, C.c_uint32), C.c_size_t(len(cl)),
                                       C.c_uint32(iterative_above), _p(rep, C.c_uint32))
    assert rc == 0
    return rep[:len(so) - 1]


class Oracle:
    def set_read_filter(self, rf: ReadFilter | None, high_freq=None):
        hf = np.ascontiguousarray(high_freq if high_freq is not None else [], dtype=np.uint64)
        self.L.so_set_read_filter(C.c_void_p(self.h), None if rf is None else C.byref(rf), _p(hf, C.c_uint64), C.c_size_t(len(hf)))

    def __init__(self, params: Params, snpmers):
        self.L = lib()
        self.params = params
        self.snpmers = np.ascontiguousarray(snpmers, dtype=np.uint64)
        self.h = self.L.so_create(C.byref(params), _p(self.snpmers, C.c_uint64), C.c_size_t(len(self.snpmers)))
        self.n_asv = 0

    def close(self):
        if self.h:
            self.L.so_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_asvs(self, bases, offsets):
        bases = as_u8(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        rc = self.L.so_load_asvs(C.c_void_p(self.h), _p(bases, C.c_uint8), _p(offsets, C.c_uint64), C.c_size_t(len(offsets) - 1))
        if rc != 0:
            raise RuntimeError(f"so_load_asvs failed: {rc}")
        self.n_asv = int(self.L.so_num_asvs(C.c_void_p(self.h)))

    def find_candidates(self, bases, quals, offsets):
        bases = as_u8(bases)
        quals = None if quals is None else as_u8(quals)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        cap = max(n * max(self.n_asv, 1), 1)
        st = np.zeros(max(n, 1), np.uint8)
        co = np.zeros(n + 1, np.uint64)
        ca = np.zeros(cap, np.uint32); cm = np.zeros(cap, np.uint32); ci = np.zeros(cap, np.uint32); cs = np.zeros(cap, np.uint8)
        rc = self.L.so_find_candidates(C.c_void_p(self.h), _p(bases, C.c_uint8), _p(quals, C.c_uint8), _p(offsets, C.c_uint64),
                                       C.c_size_t(n), _p(st, C.c_uint8), _p(co, C.c_uint64), _p(ca, C.c_uint32),
                                       _p(cm, C.c_uint32), _p(ci, C.c_uint32), _p(cs, C.c_uint8), C.c_size_t(cap))
        assert rc == 0, rc
        t = int(co[-1])
        return dict(status=st[:n], offsets=co, asv=ca[:t], mm=cm[:t], mini=ci[:t], strand=cs[:t])

    def align_pairs(self, bases, offsets, pair_read, pair_asv, pair_strand, want_cigar=False):
        bases = as_u8(bases)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        pr = np.ascontiguousarray(pair_read, dtype=np.uint32)
        pa = np.ascontiguousarray(pair_asv, dtype=np.uint32)
        ps = np.ascontiguousarray(pair_strand, dtype=np.uint8)
        n = len(pr)
        nm = np.zeros(max(n, 1), np.int32); ae = np.zeros(max(n, 1), np.int32); re = np.zeros(max(n, 1), np.int32)
        as_ = np.zeros(max(n, 1), np.int32); rs = np.zeros(max(n, 1), np.int32)
        cap = 1
        if want_cigar:
            lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
            cap = int(sum(int(lens[r]) for r in pr) * 2 + 16 * n + 16)
        co = np.zeros(n + 1, np.uint64)
        cg = np.zeros(cap, np.uint32)
        rc = self.L.so_align_pairs(C.c_void_p(self.h), _p(bases, C.c_uint8), _p(offsets, C.c_uint64), C.c_size_t(len(offsets) - 1),
                                   _p(pr, C.c_uint32), _p(pa, C.c_uint32), _p(ps, C.c_uint8), C.c_size_t(n),
                                   _p(nm, C.c_int32), _p(ae, C.c_int32), _p(re, C.c_int32),
                                   C.c_int(1 if want_cigar else 0), _p(as_, C.c_int32), _p(rs, C.c_int32),
                                   _p(co, C.c_uint64), _p(cg, C.c_uint32), C.c_size_t(cap))
        assert rc == 0, rc
        out = dict(nm=nm[:n], asv_end=ae[:n], read_end=re[:n])
        if want_cigar:
            out.update(asv_start=as_[:n], read_start=rs[:n], cigar_offsets=co, cigar=cg[:int(co[-1])])
        return out

    def map_reads(self, bases, quals, offsets, file_idx=None, n_samples=1, n_threads=0):
        bases = as_u8(bases)
        quals = None if quals is None else as_u8(quals)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        A = self.n_asv
        fi = None if file_idx is None else np.ascontiguousarray(file_idx, dtype=np.uint32)
        cap = max(n * max(min(A, 64), 1), 1)
        st = np.zeros(max(n, 1), np.uint8); bn = np.zeros(max(n, 1), np.int32)
        co = np.zeros(n + 1, np.uint64); ca = np.zeros(cap, np.uint32)
        un = np.zeros(max(A, 1), np.uint64); am = np.zeros(max(A, 1), np.uint64); le = np.zeros(max(A, 1), np.uint64)
        tot = np.zeros(2, np.uint64)
        ss = np.zeros(max(A * n_samples, 1), np.uint64)
        cells = C.c_uint64(0)
        rc = self.L.so_map_reads(C.c_void_p(self.h), _p(bases, C.c_uint8), _p(quals, C.c_uint8), _p(offsets, C.c_uint64),
                                 _p(fi, C.c_uint32), C.c_size_t(n), C.c_int(n_samples),
                                 _p(st, C.c_uint8), _p(bn, C.c_int32), _p(co, C.c_uint64), _p(ca, C.c_uint32), C.c_size_t(cap),
                                 _p(un, C.c_uint64), _p(am, C.c_uint64), _p(le, C.c_uint64), _p(tot, C.c_uint64),
                                 _p(ss, C.c_uint64), C.byref(cells), C.c_int(n_threads))
        assert rc == 0, rc
        t = int(co[-1])
        return dict(status=st[:n], best_nm=bn[:n], class_offsets=co, class_asv=ca[:t], unambig=un[:A], ambig=am[:A],
                    leq10=le[:A], totals=tot, sample_singleton=ss[:A * n_samples].reshape(n_samples, A), dp_cells=int(cells.value))


def qual_seq(quals, offsets):
    """TwinRead.qual_seq of a batch of reads (seeding::get_twin_read_syncmer, reference src/seeding.rs:578-602): minimum quality byte of
    every QUALITY_SEQ_BIN = 4 bases (src/constants.rs:48), rounded by QualCompact3::try_from_bits (src/types.rs:437-459), packed as a
    bio-seq Seq of a 4-bit codec (bin i of a read in bits 4(i % 16).. of the read's word i / 16).  -> (words u64, word_offsets u64[n+1])"""
    q = as_u8(quals); o = np.asarray(offsets, np.int64)
    n = len(o) - 1
    lens = o[1:] - o[:-1]
    bins = (lens + 3) // 4
    nwords = (bins + 15) // 16
    wo = np.zeros(n + 1, np.uint64); wo[1:] = np.cumsum(nwords)
    words = np.zeros(int(wo[-1]), np.uint64)
    # src/types.rs:441-458: 0..=34 -> 0, 35..=37 -> 1, ... 74..=76 -> 14, everything above -> 15
    table = np.array([0 if b <= 34 else min(15, (b - 35) // 3 + 1) for b in range(256)], np.uint64)
    for r in range(n):
        L = int(lens[r])
        if L == 0:
            continue
        x = q[int(o[r]):int(o[r + 1])]
        pad = np.full(int(bins[r]) * 4, 255, np.uint8); pad[:L] = x
        mn = pad.reshape(-1, 4).min(axis=1)                                    # :585-599, the last bin may hold fewer than 4
        codes = table[mn]
        cp = np.zeros(int(nwords[r]) * 16, np.uint64); cp[:len(codes)] = codes
        sh = (np.arange(16, dtype=np.uint64) * np.uint64(4))
        words[int(wo[r]):int(wo[r + 1])] = np.bitwise_or.reduce(cp.reshape(-1, 16) << sh, axis=1)
    return words, wo


def em(n_asv, cls_offsets, cls_asv, counts, total_assigned):
    L = lib()
    co = np.ascontiguousarray(cls_offsets, dtype=np.uint64)
    ca = np.ascontiguousarray(cls_asv, dtype=np.uint32)
    cn = np.ascontiguousarray(counts, dtype=np.uint64)
    ab = np.zeros(max(n_asv, 1), np.float64); dp = np.zeros(max(n_asv, 1), np.uint64)
    it = L.so_em(C.c_size_t(n_asv), _p(co, C.c_uint64), _p(ca, C.c_uint32), _p(cn, C.c_uint64), C.c_size_t(len(cn)),
                 C.c_uint64(int(total_assigned)), _p(ab, C.c_double), _p(dp, C.c_uint64))
    return ab[:n_asv], dp[:n_asv], it
