"""TwinRead field encodings (reference src/types.rs:386-412) for the C ABI's TwinRead entry points
(savont_b200_map_twin_reads / savont_b200_upload_twin_reads).

The Rust shim passes `TwinRead.dna_seq` (bio-seq `Seq<Dna>` raw words) and `TwinRead.snpmer_kmers`
(`Vec<Kmer48>`) as they lie in memory; this module builds the same byte layouts from sketch output so that
the Python host layer, the tests and bench.py can exercise that entry point.  Host-side packing only (numpy);
nothing here computes any part of the path.
"""
from __future__ import annotations

import numpy as np


def pack_dna_words(codes: np.ndarray, offsets: np.ndarray, chunk_reads: int = 32768):
    """2-bit codes (A0 C1 G2 T3, one per byte, concatenated) -> (words u64, word_offsets u64[n+1], lengths u32[n]).
    Base i of a read sits in bits 2(i%32).. of the read's word i/32 (first base in the LOW bits, src/types.rs:575-579);
    every read starts at a word boundary.  Works through the reads in chunks to bound the scatter's index memory."""
    offsets = np.asarray(offsets, np.int64)
    n = len(offsets) - 1
    lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
    nwords = (lens + 31) // 32
    woffs = np.zeros(n + 1, np.int64)
    np.cumsum(nwords, out=woffs[1:])
    words = np.zeros(int(woffs[-1]), np.uint64)
    for r0 in range(0, n, chunk_reads):
        r1 = min(n, r0 + chunk_reads)
        nb = int(offsets[r1] - offsets[r0]); nw = int(woffs[r1] - woffs[r0])
        if nb == 0:
            continue
        padded = np.zeros(nw * 32, np.uint8)
        shift = np.repeat((woffs[r0:r1] - woffs[r0]) * 32 - (offsets[r0:r1] - offsets[r0]), lens[r0:r1])
        idx = np.arange(nb, dtype=np.int64) + shift
        padded[idx] = np.asarray(codes[int(offsets[r0]):int(offsets[r1])], np.uint8) & 3
        # 4 bases -> one byte (little endian bit order), 8 bytes -> one little-endian u64
        q = padded.reshape(-1, 4)
        by = (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8)
        words[int(woffs[r0]):int(woffs[r1])] = by.view("<u8")
    return words, woffs.astype(np.uint64), lens.astype(np.uint32)


def kmer48_bytes(kmers: np.ndarray) -> np.ndarray:
    """u64 k-mers -> the bytes of a Vec<Kmer48> (low 6 bytes, little endian; src/types.rs:104-134)"""
    k = np.ascontiguousarray(kmers, dtype="<u8")
    return np.ascontiguousarray(k.view(np.uint8).reshape(-1, 8)[:, :6]).reshape(-1)


def from_sketch(sk: dict, offsets: np.ndarray, file_idx=None, kmer48: bool = False) -> dict:
    """sketch output of the same reads (Context.sketch(..., want_codes=True), i.e. what Stage 1.5 stored in the
    TwinReads) -> the argument dict of Context.map_twin_reads"""
    words, woffs, lens = pack_dna_words(sk["codes"], offsets)
    sk_k = np.ascontiguousarray(sk["snp_kmer"], np.uint64)
    return dict(dna_words=words, dna_word_offsets=woffs, lengths=lens,
                snpmer_kmers=kmer48_bytes(sk_k) if kmer48 else sk_k,
                snpmer_offsets=np.ascontiguousarray(sk["snp_offsets"], np.uint64),
                file_idx=None if file_idx is None else np.ascontiguousarray(file_idx, np.uint32))


def twin_reads_from_snpmers(ctx, bases, quals, offsets, file_idx=None, kmer48: bool = False, chunk_reads: int = 65536, alloc=None) -> dict:
    """Read side of Stage 1.5 (kmer_comp::twin_reads_from_snpmers, src/kmer_comp.rs:68-258, the part Stage 7 consumes):
    sketch the reads on the GPU (Context.sketch) and lay dna_seq / snpmer_kmers out as the TwinRead fields.
    `alloc(nbytes)` may return a uint8 array (e.g. pinned memory) for the two big buffers."""
    offsets = np.asarray(offsets, np.uint64)
    n = len(offsets) - 1
    words, woffs, lens, kmers = [], [np.zeros(1, np.uint64)], [], []
    soffs = [np.zeros(1, np.uint64)]
    wbase = sbase = 0
    for r0 in range(0, n, chunk_reads):
        r1 = min(n, r0 + chunk_reads)
        o = offsets[r0:r1 + 1]
        b0, b1 = int(o[0]), int(o[-1])
        sk = ctx.sketch(bases[b0:b1], None if quals is None else quals[b0:b1], o - o[0])
        w, wo, ln = pack_dna_words(sk["codes"], o - o[0])
        words.append(w); woffs.append(wo[1:] + np.uint64(wbase)); lens.append(ln)
        kmers.append(np.ascontiguousarray(sk["snp_kmer"], np.uint64)); soffs.append(sk["snp_offsets"][1:] + np.uint64(sbase))
        wbase += int(wo[-1]); sbase += int(sk["snp_offsets"][-1])
    cat = lambda xs, dt: np.concatenate(xs).astype(dt, copy=False) if xs else np.zeros(0, dt)
    words = cat(words, np.uint64); kmers = cat(kmers, np.uint64)
    kk = kmer48_bytes(kmers) if kmer48 else kmers
    if alloc is not None:
        pw = alloc(words.nbytes)[:words.nbytes].view(np.uint64); pw[:] = words; words = pw
        pk = alloc(kk.nbytes)[:kk.nbytes].view(kk.dtype); pk[:] = kk; kk = pk
    return dict(dna_words=words, dna_word_offsets=cat(woffs, np.uint64), lengths=cat(lens, np.uint32), snpmer_kmers=kk,
                snpmer_offsets=cat(soffs, np.uint64), file_idx=None if file_idx is None else np.ascontiguousarray(file_idx, np.uint32))


def scatter_to_heap(twin: dict, seed: int = 1) -> dict:
    """Emulates a Vec<TwinRead> in the reference's memory: every read's dna_seq words and snpmer_kmers in its OWN block of ordinary
    (pageable) host memory, blocks in shuffled order with allocator-like gaps, and one pointer per read -- the argument of
    Context.map_twin_read_ptrs / savont_b200_map_twin_read_ptrs.  The returned dict keeps the arenas alive."""
    rng = np.random.default_rng(seed)
    wo = np.asarray(twin["dna_word_offsets"], np.int64); so = np.asarray(twin["snpmer_offsets"], np.int64)
    n = len(wo) - 1
    nw = wo[1:] - wo[:-1]; ns = so[1:] - so[:-1]
    sk = twin["snpmer_kmers"]
    kb = 6 if sk.dtype == np.uint8 else 8
    order = rng.permutation(n)

    def place(sizes_bytes, align):
        # block i of the shuffled order starts behind block i-1 plus a malloc-like header / padding of 16..64 bytes
        gaps = rng.integers(16, 65, size=n).astype(np.int64)
        sz = sizes_bytes[order]
        start_sh = np.zeros(n, np.int64)
        step = (sz + gaps + align - 1) // align * align
        np.cumsum(step[:-1], out=start_sh[1:])
        start = np.zeros(n, np.int64); start[order] = start_sh
        return start, int(start_sh[-1] + step[-1]) if n else 0

    wstart, wtot = place(nw * 8, 16)
    sstart, stot = place(ns * kb, 2)
    warena = np.zeros(max(wtot, 16), np.uint8); sarena = np.zeros(max(stot, 16), np.uint8)
    words = np.ascontiguousarray(twin["dna_words"], np.uint64).view(np.uint8)
    skb = np.ascontiguousarray(sk).view(np.uint8)
    # scatter (vectorised through flat byte indices, in chunks to bound the index memory)
    for r0 in range(0, n, 65536):
        r1 = min(n, r0 + 65536)
        for arena, src, off, cnt, unit, dst in ((warena, words, wo, nw, 8, wstart), (sarena, skb, so, ns, kb, sstart)):
            c = cnt[r0:r1] * unit
            tot = int(c.sum())
            if tot == 0:
                continue
            rel = np.arange(tot, dtype=np.int64) - np.repeat(np.cumsum(c) - c, c)
            arena[np.repeat(dst[r0:r1], c) + rel] = src[int(off[r0]) * unit:int(off[r1]) * unit]
    fi = twin.get("file_idx")
    return dict(dna_ptrs=(warena.ctypes.data + wstart).astype(np.uint64), snpmer_ptrs=(sarena.ctypes.data + sstart).astype(np.uint64),
                lengths=np.ascontiguousarray(twin["lengths"], np.uint32), snpmer_counts=ns.astype(np.uint32), kmer_bytes=kb,
                file_idx=None if fi is None else np.ascontiguousarray(fi, np.uint32), _arenas=(warena, sarena))
