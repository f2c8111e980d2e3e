"""TwinRead field encodings (reference src/types.rs:386-412) for the C ABI's TwinRead entry points
(savont_b200_map_twin_reads / savont_b200_upload_twin_reads).

The Rust shim passes `TwinRead.dna_seq` (bio-seq `Seq<Dna>` raw words) and `TwinRead.snpmer_kmers`
(`Vec<Kmer48>`) as they lie in memory; this module builds the same byte layouts from sketch output so that
the Python host layer, the tests and bench.py can exercise that entry point.  Host-side packing only (numpy);
nothing here computes any part of the path.
"""
from __future__ import annotations

import numpy as np


def pack_dna_words(codes: np.ndarray, offsets: np.ndarray):
    """2-bit codes (A0 C1 G2 T3, one per byte, concatenated) -> (words u64, word_offsets u64[n+1], lengths u32[n]).
    Base i of a read sits in bits 2(i%32).. of the read's word i/32 (first base in the LOW bits, src/types.rs:575-579);
    every read starts at a word boundary."""
    offsets = np.asarray(offsets, np.int64)
    n = len(offsets) - 1
    lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
    nwords = (lens + 31) // 32
    woffs = np.zeros(n + 1, np.int64)
    np.cumsum(nwords, out=woffs[1:])
    total_words = int(woffs[-1])
    padded = np.zeros(total_words * 32, np.uint8)
    if n and int(lens.sum()):
        shift = np.repeat(woffs[:-1] * 32 - (offsets[:-1] - offsets[0]), lens)
        idx = np.arange(int(lens.sum()), dtype=np.int64) + shift
        padded[idx] = np.asarray(codes[int(offsets[0]):int(offsets[-1])], np.uint8) & 3
    # 4 bases -> one byte (little endian bit order), 8 bytes -> one little-endian u64
    q = padded.reshape(-1, 4)
    by = (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8)
    words = by.view("<u8").copy() if total_words else np.zeros(0, np.uint64)
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
