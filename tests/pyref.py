"""Pure-Python second restatement of the sketch and of the banded
both-ends-free edit distance, written independently of oracle/savont_oracle.c and used only
to cross-check the C oracle on small inputs (tests/test_oracle.py)."""
import math

M64 = (1 << 64) - 1
ENC = {ord('A'): 0, ord('a'): 0, ord('C'): 1, ord('c'): 1, ord('G'): 2, ord('g'): 2,
       ord('T'): 3, ord('t'): 3, ord('U'): 3, ord('u'): 3, 1: 1, 2: 2, 3: 3}


def enc(b):
    return ENC.get(b, 0)


def mm_hash64(key):
    key = (~key + (key << 21)) & M64
    key ^= key >> 24
    key = (key + (key << 3) + (key << 8)) & M64
    key ^= key >> 14
    key = (key + (key << 2) + (key << 4)) & M64
    key ^= key >> 28
    key = (key + (key << 31)) & M64
    return key


def revcomp_kmer(x, k):
    out = 0
    for _ in range(k):
        out = (out << 2) | (3 - (x & 3))
        x >>= 2
    return out


def sketch(seq: bytes, qual, k, c, min_bq, snpmer_set):
    L = len(seq)
    if L < k:
        return None
    s = k - c + 1
    W = k - s + 1
    mid = (k - s) // 2
    split_mask = M64 ^ (3 << (k - 1))
    mask2k = (1 << (2 * k)) - 1
    mask2s = (1 << (2 * s)) - 1
    flat = qual is not None and all(q == qual[0] for q in qual)
    F = R = f7 = r7 = 0
    for i in range(k - 1):
        e = enc(seq[i])
        F = (F << 2) | e
        R = (R >> 2) | ((3 - e) << (2 * (k - 1)))
        if i < s - 1:
            f7 = (f7 << 2) | e
            r7 = (r7 >> 2) | ((3 - e) << (2 * (s - 1)))
    window = []
    mini, snp, dup = [], [], {}
    for i in range(k - 1, L):
        e = enc(seq[i])
        F = ((F << 2) | e) & mask2k
        R = ((R >> 2) & ~(3 << (2 * k - 2))) | ((3 - e) << (2 * (k - 1)))
        fwd = (F & split_mask) < (R & split_mask)
        canon = F if fwd else R
        f7 = ((f7 << 2) | e) & mask2s
        r7 = ((r7 >> 2) & ~(3 << (2 * s - 2))) | ((3 - e) << (2 * (s - 1)))
        window.append(mm_hash64(min(f7, r7)))
        if len(window) > W:
            window.pop(0)
        if canon in snpmer_set:
            q = (qual[i + 1 + k // 2 - k] - 33) if qual is not None else 60
            if q > min_bq or flat:
                snp.append((i + 1 - k, canon))
            dup[canon & split_mask] = dup.get(canon & split_mask, 0) + 1
        if len(window) == W and all(window[j] > window[mid] for j in range(W) if j != mid):
            mini.append((i + 1 - k, canon, int(fwd)))
    snp = [x for x in snp if dup[x[1] & split_mask] == 1]
    est = None
    if qual is not None and not flat:
        sm = 0.0
        for q in qual:
            sm += math.pow(10.0, -float(q - 33) / 10.0)
        est = 100. - (sm / len(qual) * 100.)
    return dict(mini=mini, snp=snp, est_id=est, codes=[enc(b) for b in seq])


def banded_nm(a, r, B):
    """Full (n+1)x(m+1) matrix with out-of-band cells = inf; returns (nm, i_end, j_end)."""
    n, m = len(a), len(r)
    INF = 10 ** 9
    D = [[INF] * (m + 1) for _ in range(n + 1)]
    for i in range(n + 1):
        for j in range(m + 1):
            if abs(i - j) > B:
                continue
            if i == 0 or j == 0:
                D[i][j] = 0
                continue
            D[i][j] = min(D[i - 1][j - 1] + (a[i - 1] != r[j - 1]), D[i - 1][j] + 1, D[i][j - 1] + 1)
    best = (INF, -1, -1)
    cells = [(n, j) for j in range(m + 1) if abs(n - j) <= B] + [(i, m) for i in range(n + 1) if abs(i - m) <= B]
    for (i, j) in cells:
        d = D[i][j]
        if d < best[0] or (d == best[0] and i + j > best[1] + best[2]):
            best = (d, i, j)
    return best


def twin_reads_filter(reads, quals, ids, k, c, min_bq, snpmer_set, high_freq, min_len, max_len, cutoff, max_count=500):
    """Independent restatement of kmer_comp::twin_reads_from_snpmers (reference src/kmer_comp.rs:68-258) + the read order
    of src/main.rs:538.  Returns (status list, kept indices in Stage-7 order)."""
    from collections import Counter
    status, kept = [], []
    for idx, (sq, ql) in enumerate(zip(reads, quals)):
        if len(sq) < min_len or (max_len and len(sq) > max_len):
            status.append(5); continue
        sk = sketch(sq, ql, k, c, min_bq, snpmer_set)
        if sk is None:
            status.append(4); continue
        kmers = [m[1] for m in sk["mini"]]
        cnt = Counter(kmers)
        solid = sum(1 for x in kmers if cnt[x] <= max_count and x not in high_freq)
        if solid < len(sq) // c // 20:
            status.append(6); continue
        status.append(0)
        kept.append((idx, sk["est_id"]))
    kept.sort(key=lambda t: (ids[t[0]].encode() if isinstance(ids[t[0]], str) else ids[t[0]]))       # stable
    kept = [t for t in kept if t[1] is None or t[1] >= cutoff or cutoff <= 0]
    dropped = set(i for i in range(len(reads)) if status[i] == 0) - set(t[0] for t in kept)
    for i in dropped:
        status[i] = 7
    kept.sort(key=lambda t: -(t[1] if t[1] is not None else 100.0))                                    # stable, descending
    return status, [t[0] for t in kept]


def split_kmer_counts(reads, quals, rc, k, min_bq, single_strand=False):
    """Independent restatement of seeding::split_kmer_mid (reference src/seeding.rs:975-1068) +
    seq_parse::read_to_split_kmers with bloom_filter_size == 0 (src/seq_parse.rs:12-83, :315-497).
    Returns (sorted list of (kmer, c0, c1) passing the filter, number of distinct k-mers)."""
    comp = {ord('A'): ord('T'), ord('a'): ord('T'), ord('T'): ord('A'), ord('t'): ord('A'), ord('C'): ord('G'), ord('c'): ord('G'),
            ord('G'): ord('C'), ord('g'): ord('C')}
    split_mask = M64 ^ (3 << (k - 1))
    counts = {}
    for sq, ql, r in zip(reads, quals, rc):
        if r:                                          # id ends in "rc": utils::reverse_complement + reversed qualities
            sq = bytes(comp.get(b, ord('N')) for b in reversed(sq))
            ql = None if ql is None else bytes(reversed(ql))
        L = len(sq)
        if L < k:
            continue
        use_q = ql is not None and any(q != ql[0] for q in ql)
        for p in range(L - k + 1):
            f = 0
            for b in sq[p:p + k]:
                f = (f << 2) | enc(b)
            rv = revcomp_kmer(f, k)
            sf, sr = f & split_mask, rv & split_mask
            if sf == sr:
                continue
            if use_q and ql[p + k // 2] - 33 < min_bq:
                continue
            fwd = sf < sr
            c = counts.setdefault(f if fwd else rv, [0, 0])
            c[1 if fwd else 0] += 1
    out = []
    for kmer in sorted(counts):
        c0, c1 = counts[kmer]
        if (c0 > 2) if single_strand else (c0 > 0 and c1 > 0 and c0 + c1 > 2):
            out.append((kmer, c0, c1))
    return out, len(counts)


def hpc_with_quality(seq: bytes, qual: bytes, do_hpc=True):
    """utils::homopolymer_compress_with_quality (src/utils.rs:135-184) on the upper-case ACGT decoded from 2-bit codes"""
    codes = [enc(b) for b in seq]
    if not codes:
        return [], [], []
    if not do_hpc:
        return codes, list(qual), [1] * len(codes)
    hc, hq, hl = [], [], []
    cur, ln, mq = codes[0], 1, qual[0]
    for c_, q in zip(codes[1:], qual[1:]):
        if c_ == cur and ln < 255:
            ln += 1; mq = min(mq, q)
        else:
            hc.append(cur); hq.append(mq); hl.append(ln)
            cur, ln, mq = c_, 1, q
    hc.append(cur); hq.append(mq); hl.append(ln)
    return hc, hq, hl


def pileup_columns(cons_lens, reads, aln):
    """CIGAR walk + median run length of generate_consensus_pileups (src/alignment.rs:524-622).
    reads[r] = (cons index, hc, hq, hl) compressed read; aln[r] = None or (strand, cons_start, read_start, read_end, cigar ops
    [(len, op)]).  Returns (columns[cons][pos] = list of (read, kind, bases, quals, hps), ref_hp[cons][pos])."""
    cols = [[[] for _ in range(n)] for n in cons_lens]
    for r, (ci, hc, hq, hl) in enumerate(reads):
        if aln[r] is None:
            continue
        sd, cs, qs, qe, cig = aln[r]
        if sd:
            fs = [3 - x for x in reversed(hc)]; fq = list(reversed(hq)); fl = list(reversed(hl))
        else:
            fs, fq, fl = hc, hq, hl
        ms, mq, ml = fs[qs:qe], fq[qs:qe], fl[qs:qe]
        ref, qp = cs, 0
        col = cols[ci]
        for ln, op in cig:
            if op == 0:
                for j in range(ln):
                    if ref + j < len(col) and qp + j < len(ms):
                        col[ref + j].append((r, 0, [b"ACGT"[ms[qp + j]]], [mq[qp + j]], [ml[qp + j]]))
                ref += ln; qp += ln
            elif op == 1:
                if ref > 0 and ref - 1 < len(col) and qp + ln <= len(ms):
                    m = min(ln, 2)
                    col[ref - 1].append((r, 2, [b"ACGT"[x] for x in ms[qp:qp + m]], mq[qp:qp + m], ml[qp:qp + m]))
                qp += ln
            else:
                for j in range(ln):
                    if ref + j < len(col):
                        col[ref + j].append((r, 1, [], [], []))
                ref += ln
    hp = []
    for col in cols:
        out = []
        for entries in col:
            v = sorted(e[4][0] for e in entries if e[1] == 0)
            if not v:
                out.append(1)
            elif len(v) % 2 == 0:
                out.append((v[len(v) // 2 - 1] + v[len(v) // 2]) // 2)
            else:
                out.append(v[len(v) // 2])
        hp.append(out)
    return cols, hp


_M64 = (1 << 64) - 1


def fx_hash2(seed, kmer):
    """FxHasher64 (fxhash 0.2.1): write_u64(seed); write_u64(kmer); finish()"""
    h = 0
    for w in (seed, kmer):
        h = ((((h << 5) | (h >> 59)) & _M64) ^ w) * 0x517cc1b727220a95 & _M64
    return h


def lsh_signatures(minis, n_tables=20, bucket=3):
    """TwinRead::compute_lsh_signatures (src/types.rs:719-747) -> list of 20 Option<u64>"""
    out = []
    for t in range(n_tables):
        if len(minis) < bucket:
            out.append(None); continue
        hashed = sorted(((fx_hash2(t, km), km) for km in minis), key=lambda x: x[0])      # stable, by hash only
        sg = 0
        for i in range(bucket):
            sg ^= (hashed[i][1] * (i + 1)) & _M64
        out.append(sg)
    return out


def cluster_reads_by_kmers(reads_minis, k=17, threshold=0.95, top_n=10):
    """asv_cluster::cluster_reads_by_kmers (src/asv_cluster.rs:72-190, bucketed branch) -> rep_of"""
    sigs = [lsh_signatures(m) for m in reads_minis]
    tables = [dict() for _ in range(20)]
    rep_of = []
    for i, minis in enumerate(reads_minis):
        hits = {}
        for t in range(20):
            sg = sigs[i][t]
            if sg is not None:
                for c in tables[t].get(sg, ()):
                    hits[c] = hits.get(c, 0) + 1
        best = None
        if hits:
            cands = sorted(hits.items(), key=lambda x: (x[1], x[0]), reverse=True)
            max_hits = cands[0][1]
            check = []
            for c, h in cands:
                if h == max_hits or len(check) < top_n:
                    check.append(c)
                else:
                    break
            best_sim = 0.0
            rset = set(minis)
            for c in check:
                rep = reads_minis[c]
                count = sum(1 for km in rset if km in rep)
                ratio = count / max(len(rset), len(rep))
                sim = ratio ** (1.0 / k)
                if sim > best_sim:
                    best_sim, best = sim, c
            if not best_sim > threshold:
                best = None
        if best is None:
            for t in range(20):
                if sigs[i][t] is not None:
                    tables[t].setdefault(sigs[i][t], []).append(i)
            rep_of.append(i)
        else:
            rep_of.append(best)
    return rep_of


def reassign_reads(cons, reads, k=17):
    """reassign_reads_to_best_cluster (src/asv_cluster.rs:1052-1098, SNPmer markers) + the (matches, mismatches) matrix of
    compare_consensus_snpmers (:968-994): cons / reads = lists of k-mer lists -> (best cluster, mismatches, matches, matrix)"""
    mask = _M64 ^ (3 << (k - 1))
    maps = []
    for c in cons:
        m = {}
        for km in c:
            m[km & mask] = km
        maps.append(m)
    out, mat = [], []
    for rd in reads:
        best, score, row = None, (1 << 64, 0), []
        for ci, m in enumerate(maps):
            ma = mm = 0
            for km in rd:
                ck = m.get(km & mask)
                if ck is not None:
                    if ck == km:
                        ma += 1
                    else:
                        mm += 1
            row.append((ma, mm))
            if mm < score[0] or (mm == score[0] and ma > score[1]):
                score, best = (mm, ma), ci
        out.append((best, score[0], score[1])); mat.append(row)
    return out, mat


def cluster_reads_by_snpmers(read_snpmers, cluster, k=17, iterative_above=1000):
    """the greedy loop of asv_cluster::cluster_reads_by_snpmers for one k-mer cluster (src/asv_cluster.rs:561-700, blockmers off);
    above `iterative_above` representatives the first compatible one in creation order (the reference's find_any is not
    deterministic) -> {read: representative}"""
    mask = _M64 ^ (3 << (k - 1))
    index, reps, size, out = {}, [], {}, {}
    for read in cluster:
        q = read_snpmers[read]
        best = None
        if len(reps) > iterative_above:
            qmap = {}
            for km in q:
                qmap[km & mask] = km
            for rep in reps:
                ma = mi = 0
                for km in read_snpmers[rep]:
                    hit = qmap.get(km & mask)
                    if hit is not None:
                        if hit == km:
                            ma += 1
                        else:
                            mi += 1
                if mi == 0 and ma > 0:
                    best = rep; break
        else:
            stats = {}
            for km in q:
                for rep, rk in index.get(km & mask, ()):
                    s = stats.setdefault(rep, [0, 0])
                    s[0 if rk == km else 1] += 1
            cands = sorted((-s[0], size[rep], rep) for rep, s in stats.items() if s[1] == 0 and s[0] > 0)
            if cands:
                best = cands[0][2]
        if best is not None:
            out[read] = best; size[best] += 1
        else:
            reps.append(read); size[read] = 1; out[read] = read
            for km in q:
                index.setdefault(km & mask, []).append((read, km))
    return out


def qual_seq(qual: bytes):
    """TwinRead.qual_seq of one read, element by element (reference src/seeding.rs:578-602 + QualCompact3::try_from_bits src/types.rs:437-459)
    -> list of 4-bit codes, one per bin of 4 bases"""
    binned, counter, mn = [], 0, 255
    for b in qual:
        if counter == 4:
            binned.append(mn); counter = 0; mn = 255
        counter += 1
        if b < mn:
            mn = b
    if counter != 0:
        binned.append(mn)
    ranges = [(0, 34)] + [(35 + 3 * i, 37 + 3 * i) for i in range(14)]
    out = []
    for v in binned:
        code = 15
        for c, (lo, hi) in enumerate(ranges):
            if lo <= v <= hi:
                code = c; break
        out.append(code)
    return out
