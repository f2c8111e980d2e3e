"""helpers shared by the parity tests"""
import random

import numpy as np

from tests import pyref


def rand_seq(rng, n, alphabet=b"ACGT"):
    return bytes(rng.choice(alphabet) for _ in range(n))


def snpmer_set_from(seqs, k, every=7):
    """some canonical k-mers of the sequences, both alleles of the middle base"""
    s = set()
    mask = ((1 << 64) - 1) ^ (3 << (k - 1))
    for sq in seqs:
        for p in range(0, len(sq) - k + 1, every):
            f = 0
            for b in sq[p:p + k]:
                f = (f << 2) | pyref.enc(b)
            r = pyref.revcomp_kmer(f, k)
            canon = f if (f & mask) < (r & mask) else r
            s.add(canon)
            s.add((canon & mask) | (((canon >> (k - 1)) + 1) & 3) << (k - 1))
    return s


def mutate(rng, s, nsub, nins, ndel):
    s = bytearray(s)
    for _ in range(nsub):
        i = rng.randrange(len(s)); s[i] = rng.choice(b"ACGT")
    for _ in range(nins):
        i = rng.randrange(len(s)); s.insert(i, rng.choice(b"ACGT"))
    for _ in range(ndel):
        if len(s) > 1:
            i = rng.randrange(len(s)); del s[i]
    return bytes(s)


def weird_sequences(seed, k, n=40, lengths=(0, 1, 16, 17, 18, 26, 27, 31, 32, 33, 63, 64, 65, 100, 257, 800, 1500, 2100)):
    """ragged / degenerate inputs: too short, exactly k, non-ACGT, low complexity, repeats"""
    rng = random.Random(seed)
    seqs, quals = [], []
    for t in range(n):
        L = lengths[t % len(lengths)] if t < len(lengths) else rng.choice([k - 1, k, k + 1, 30, 64, 200, 333, 800])
        alpha = b"ACGT" if t % 3 else b"ACGTNacgtRYu"
        sq = rand_seq(rng, L, alpha)
        if t % 5 == 0 and L > 60:
            sq = sq[:20] + b"ACACACACACACACACACACACACACACAC" + sq[50:]
        if t % 7 == 0 and L > 100:
            sq = sq[:50] + sq[:50] + sq[100:]
        if t % 11 == 0 and L > 40:
            sq = b"A" * 25 + sq[25:]
        seqs.append(sq)
        if t % 4 == 3:
            quals.append(bytes([40 + 33] * L))
        else:
            quals.append(bytes(33 + rng.choice([5, 12, 25, 26, 30, 40, 50]) for _ in range(L)))
    return seqs, quals


def assert_same(a, b, keys, ctx=""):
    for key in keys:
        x, y = np.asarray(a[key]), np.asarray(b[key])
        if x.dtype.kind == "f":
            ok = x.shape == y.shape and np.array_equal(x.view(np.uint64), y.view(np.uint64)) or \
                (x.shape == y.shape and np.all((x == y) | (np.isnan(x) & np.isnan(y))))
        else:
            ok = x.shape == y.shape and np.array_equal(x, y)
        if not ok:
            bad = None
            if x.shape == y.shape and x.size:
                idx = np.flatnonzero(x != y)
                bad = (idx[:8], x[idx[:8]], y[idx[:8]])
            raise AssertionError(f"{ctx}: mismatch in '{key}': shapes {x.shape} vs {y.shape}; first diffs {bad}")


def pileup_case(seed, n_cons=5, reads_per=12, cons_len=(60, 200), long_runs=False):
    """clusters for the Stage 4 pileups: homopolymer-compressed consensus sequences and raw reads (with homopolymer runs put
    back, errors, both strands, N bases, clipped / extended ends, empty reads) in cluster order"""
    rng = random.Random(seed)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    cons, reads, quals, rcons = [], [], [], []
    for ci in range(n_cons):
        L = rng.randrange(*cons_len) if ci != 2 else 0          # one empty consensus
        s = bytearray()
        while len(s) < L:
            b = rng.choice(b"ACGT")
            if not s or s[-1] != b:
                s.append(b)
        cons.append(bytes(s))
        for t in range(reads_per if ci != 3 else 0):            # one consensus without reads
            raw = bytearray()
            for b in s:
                run = rng.choice([1, 1, 1, 2, 3, 5]) if not (long_runs and rng.random() < 0.01) else rng.choice([254, 255, 256, 300, 520])
                raw += bytes([b]) * run
            rd = mutate(rng, bytes(raw), rng.randrange(4), rng.randrange(3), rng.randrange(3)) if raw else b""
            if t % 5 == 1 and len(rd) > 20:
                rd = rd[rng.randrange(1, 9):]
            if t % 5 == 2 and len(rd) > 20:
                rd = rd[:-rng.randrange(1, 9)] + rand_seq(rng, 4)
            if t % 7 == 3 and len(rd) > 10:
                i = rng.randrange(len(rd)); rd = rd[:i] + b"N" + rd[i + 1:]
            if t % 2:
                rd = rd.translate(comp)[::-1]
            if t == 6:
                rd = b""
            reads.append(rd); rcons.append(ci)
            quals.append(bytes(33 + rng.choice([3, 10, 20, 30, 40]) for _ in range(len(rd))))
    return cons, reads, quals, rcons


def clustering_case(seed, n_species=6, n_reads=300, length=700, err=0.02, k=17, c=11):
    """minimizer lists (CSR) of reads drawn from a few related 'species' + degenerate reads (no / few / duplicated minimizers)"""
    from oracle import oracle as O
    rng = random.Random(seed)
    roots = [rand_seq(rng, length) for _ in range(max(1, n_species // 2))]
    species = [roots[s % len(roots)] if s < len(roots) else mutate(rng, roots[s % len(roots)], 12, 2, 2) for s in range(n_species)]
    reads = []
    for i in range(n_reads):
        s = species[rng.randrange(n_species)]
        ne = int(len(s) * err * rng.choice([0.2, 1, 1, 2]))
        rd = mutate(rng, s, ne // 2, ne // 4, ne // 4)
        if i % 37 == 5:
            rd = rd[:rng.choice([5, 17, 27, 30, 40])]            # too short for any / for three minimizers
        if i % 41 == 7:
            rd = reads[-1]                                        # identical to the previous read
        if i % 43 == 9:
            rd = rd[:200] + rd[:200] + rd[:200]                   # repeated k-mers: duplicates in the minimizer list
        reads.append(rd)
    b, o = O.concat(reads)
    sk = O.sketch_batch(O.default_params(k=k, c=c), [], b, None, o)
    return np.asarray(sk["mini_kmer"], np.uint64), np.asarray(sk["mini_offsets"], np.uint64), reads


def reassign_case(seed, n_cons=12, n_sites=60, n_reads=150, k=17):
    """cluster consensus SNPmer lists over a shared set of SNP sites (splitmers) and reads carrying subsets of the sites with a
    few allele flips; some sites unknown to every consensus, empty reads / consensus lists, duplicated splitmers in a list"""
    rng = random.Random(seed)
    mask = ((1 << 64) - 1) ^ (3 << (k - 1))
    sites = []
    while len(sites) < n_sites:
        s = rng.getrandbits(2 * k) & mask
        if s not in sites:
            sites.append(s)
    def km(site, allele):
        return site | (allele << (k - 1))
    cons_alleles = [[rng.randrange(4) if rng.random() < 0.8 else None for _ in sites] for _ in range(n_cons)]
    cons = [[km(s, a) for s, a in zip(sites, al) if a is not None] for al in cons_alleles]
    cons[3] = []                                                   # a cluster without consensus SNPmers
    cons[5] = cons[5] + [km(sites[0], 1), km(sites[0], 2)]         # duplicated splitmer: the last entry wins (HashMap::insert)
    cons[7] = list(cons[6])                                        # identical lists: the lower index wins ties
    reads = []
    for r in range(n_reads):
        src = cons_alleles[rng.randrange(n_cons)]
        rd = []
        for s, a in zip(sites, src):
            if rng.random() < 0.5:
                al = a if a is not None else rng.randrange(4)
                if rng.random() < 0.05:
                    al = (al + 1) & 3
                rd.append(km(s, al))
        if r % 17 == 3:
            rd = []
        if r % 19 == 4:
            rd = [rng.getrandbits(2 * k) for _ in range(8)]        # sites nobody knows
        rng.shuffle(rd)
        reads.append(rd)
    return cons, reads


def csr_u64(lists):
    off = np.zeros(len(lists) + 1, np.uint64)
    off[1:] = np.cumsum([len(x) for x in lists])
    flat = np.array([v for x in lists for v in x], np.uint64)
    return flat, off
