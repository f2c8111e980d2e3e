/*
 * savont_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see savont_oracle.h).
 *
 * Scalar restatement of savont v0.6.4's read-vs-ASV assignment path.  Written to be
 * obviously equivalent to the Rust source (rolling registers, a deque of s-mer hashes,
 * per-pair set intersections, a cell-by-cell banded DP), NOT to be fast, and NOT sharing
 * any code or data structure with the CUDA product path.
 *
 * NM parity with the reference (minimap2 lr:hq) is UNPINNED -- see the header.
 */
#include "savont_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------ */
/* primitives                                                                           */
/* ------------------------------------------------------------------------------------ */

/* src/types.rs:92-101 BYTE_TO_SEQ: A/a=0 C/c=1 G/g=2 T/t/U/u=3, everything else 0 */
uint8_t so_encode_base(uint8_t b) {
    switch (b) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        /* the table also maps raw bytes 0,1,2,3 to themselves */
        case 1: return 1;
        case 2: return 2;
        case 3: return 3;
        default: return 0;
    }
}

/* src/seeding.rs:18-28 */
uint64_t so_mm_hash64(uint64_t key) {
    key = (~key) + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}

/* first base in the highest used bit pair (src/types.rs:1118-1128: ACGTG -> 0b00_01_10_11_10) */
uint64_t so_kmer_from_ascii(const char *s, int k) {
    uint64_t v = 0;
    for (int i = 0; i < k; i++) v = (v << 2) | so_encode_base((uint8_t)s[i]);
    return v;
}

/* src/types.rs:559-568 reverse_bit_pairs */
static uint64_t reverse_bits64(uint64_t n) {
    uint64_t r = 0;
    for (int i = 0; i < 64; i++) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}
static uint64_t reverse_bit_pairs(uint64_t n, int k) {
    const uint64_t even_mask = 0xAAAAAAAAAAAAAAAAull, odd_mask = 0x5555555555555555ull;
    uint64_t odd_bits_rev = reverse_bits64(n & odd_mask) >> (64 - 2 * k);
    uint64_t even_bits_rev = reverse_bits64(n & even_mask) >> (64 - 2 * k);
    return (odd_bits_rev >> 1) | (even_bits_rev << 1);
}
/* src/types.rs:654 reverse_kmer = reverse_bit_pairs(kmer ^ u64::MAX, k) */
uint64_t so_kmer_revcomp(uint64_t kmer, int k) { return reverse_bit_pairs(kmer ^ UINT64_MAX, k); }

/* src/types.rs:112-127 */
void so_kmer48_from_u64(uint64_t v, uint8_t out[6]) {
    for (int i = 0; i < 6; i++) out[i] = (uint8_t)(v >> (8 * i));
}
uint64_t so_kmer48_to_u64(const uint8_t in[6]) {
    uint64_t v = 0;
    for (int i = 0; i < 6; i++) v |= (uint64_t)in[i] << (8 * i);
    return v;
}

/* f64::powi lowers to llvm.powi -> compiler-rt __powidf2 (repeated squaring) */
double so_powi(double a, int b) {
    const int recip = b < 0;
    double r = 1;
    while (1) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1 / r : r;
}

/* src/seeding.rs:353  split_mask = !(3 << (k-1)) */
uint64_t so_split_mask(int k) { return ~((uint64_t)3 << (k - 1)); }

/* ------------------------------------------------------------------------------------ */
/* small containers                                                                     */
/* ------------------------------------------------------------------------------------ */
typedef struct { uint64_t *keys; uint32_t *vals; size_t cap; size_t n; } u64map; /* key -> u32, EMPTY = UINT64_MAX */

static void u64map_init(u64map *m, size_t expect) {
    size_t cap = 16;
    while (cap < expect * 2 + 2) cap <<= 1;
    m->cap = cap; m->n = 0;
    m->keys = (uint64_t *)malloc(cap * sizeof(uint64_t));
    m->vals = (uint32_t *)calloc(cap, sizeof(uint32_t));
    for (size_t i = 0; i < cap; i++) m->keys[i] = UINT64_MAX;
}
static void u64map_free(u64map *m) { free(m->keys); free(m->vals); m->keys = NULL; m->vals = NULL; }
static uint32_t *u64map_slot(u64map *m, uint64_t key, int insert) {
    size_t i = (size_t)so_mm_hash64(key) & (m->cap - 1);
    while (1) {
        if (m->keys[i] == key) return &m->vals[i];
        if (m->keys[i] == UINT64_MAX) {
            if (!insert) return NULL;
            m->keys[i] = key; m->vals[i] = 0; m->n++;
            return &m->vals[i];
        }
        i = (i + 1) & (m->cap - 1);
    }
}

typedef struct {
    uint32_t *mini_pos; uint64_t *mini_kmer; uint8_t *mini_fwd; size_t n_mini, cap_mini;
    uint32_t *snp_pos; uint64_t *snp_kmer; size_t n_snp, cap_snp;
    uint8_t *codes; size_t len;
    double est_id; /* NaN = None */
    int ok;
} sketch_t;

static void sketch_free(sketch_t *s) {
    free(s->mini_pos); free(s->mini_kmer); free(s->mini_fwd); free(s->snp_pos); free(s->snp_kmer); free(s->codes);
    memset(s, 0, sizeof(*s));
}
static void push_mini(sketch_t *s, uint32_t pos, uint64_t kmer, uint8_t fwd) {
    if (s->n_mini == s->cap_mini) {
        s->cap_mini = s->cap_mini ? s->cap_mini * 2 : 64;
        s->mini_pos = (uint32_t *)realloc(s->mini_pos, s->cap_mini * sizeof(uint32_t));
        s->mini_kmer = (uint64_t *)realloc(s->mini_kmer, s->cap_mini * sizeof(uint64_t));
        s->mini_fwd = (uint8_t *)realloc(s->mini_fwd, s->cap_mini);
    }
    s->mini_pos[s->n_mini] = pos; s->mini_kmer[s->n_mini] = kmer; s->mini_fwd[s->n_mini] = fwd; s->n_mini++;
}
static void push_snp(sketch_t *s, uint32_t pos, uint64_t kmer) {
    if (s->n_snp == s->cap_snp) {
        s->cap_snp = s->cap_snp ? s->cap_snp * 2 : 64;
        s->snp_pos = (uint32_t *)realloc(s->snp_pos, s->cap_snp * sizeof(uint32_t));
        s->snp_kmer = (uint64_t *)realloc(s->snp_kmer, s->cap_snp * sizeof(uint64_t));
    }
    s->snp_pos[s->n_snp] = pos; s->snp_kmer[s->n_snp] = kmer; s->n_snp++;
}

/* ------------------------------------------------------------------------------------ */
/* a3: get_twin_read_syncmer  (src/seeding.rs:317-658)                                  */
/* ------------------------------------------------------------------------------------ */
static void sketch_one(const so_params *p, const u64map *snpmer_set, const uint8_t *string, const uint8_t *qualities,
                       size_t len, sketch_t *out) {
    memset(out, 0, sizeof(*out));
    const int k = p->k;
    out->est_id = NAN;
    if (len < (size_t)k) { out->ok = 0; return; }           /* :339-341 */
    out->ok = 1;
    out->len = len;

    uint64_t rolling_kmer_f = 0, rolling_kmer_r = 0;
    const int marker_reverse_shift_dist = 2 * (k - 1);
    const uint64_t split_mask = so_split_mask(k);                          /* :353 */
    const uint64_t marker_mask = UINT64_MAX >> (64 - 2 * k);               /* :354 */
    const uint64_t marker_rev_mask = ~((uint64_t)3 << (2 * k - 2));        /* :355 */
    const int mid_k = k / 2;

    const int s = k - p->c + 1;                                            /* :363 */
    const uint64_t s_mask = UINT64_MAX >> (64 - 2 * s);
    const uint64_t s_rev_mask = ~((uint64_t)3 << (2 * s - 2));
    const int s_reverse_shift_dist = 2 * (s - 1);
    const int win = k - s + 1;                                             /* :368 */

    uint64_t *window = (uint64_t *)malloc((size_t)(win + 1) * sizeof(uint64_t)); /* the VecDeque */
    int wlen = 0;
    uint64_t rolling_s_f = 0, rolling_s_r = 0;

    int all_equal_q = 0;                                                   /* :372-380 */
    if (qualities) {
        all_equal_q = 1;
        for (size_t i = 1; i < len; i++) if (qualities[i] != qualities[0]) { all_equal_q = 0; break; }
    }

    for (int i = 0; i < k - 1; i++) {                                      /* :383-398 */
        uint64_t nuc_f = so_encode_base(string[i]);
        uint64_t nuc_r = 3 - nuc_f;
        rolling_kmer_f <<= 2; rolling_kmer_f |= nuc_f;
        rolling_kmer_r >>= 2; rolling_kmer_r |= nuc_r << marker_reverse_shift_dist;
        if (i < s - 1) {
            rolling_s_f <<= 2; rolling_s_f |= nuc_f;
            rolling_s_r >>= 2; rolling_s_r |= nuc_r << s_reverse_shift_dist;
        }
    }

    /* all SNPmer hits (pre-dedup) + splitmer multiplicities */
    u64map dedup; u64map_init(&dedup, 64);
    sketch_t hits; memset(&hits, 0, sizeof(hits));

    for (size_t i = (size_t)k - 1; i < len; i++) {                         /* :413 */
        uint64_t nuc_f = so_encode_base(string[i]);
        uint64_t nuc_r = 3 - nuc_f;
        rolling_kmer_f <<= 2; rolling_kmer_f |= nuc_f; rolling_kmer_f &= marker_mask;
        rolling_kmer_r >>= 2; rolling_kmer_r &= marker_rev_mask; rolling_kmer_r |= nuc_r << marker_reverse_shift_dist;

        uint64_t split_f = rolling_kmer_f & split_mask, split_r = rolling_kmer_r & split_mask;
        int canonical_marker = split_f < split_r;                          /* :429 tie -> reverse */
        uint64_t canonical_kmer = canonical_marker ? rolling_kmer_f : rolling_kmer_r;

        rolling_s_f <<= 2; rolling_s_f |= nuc_f; rolling_s_f &= s_mask;   /* :437-443 */
        rolling_s_r >>= 2; rolling_s_r &= s_rev_mask; rolling_s_r |= nuc_r << s_reverse_shift_dist;
        uint64_t canonical_s = rolling_s_f < rolling_s_r ? rolling_s_f : rolling_s_r;
        uint64_t hash = so_mm_hash64(canonical_s);

        window[wlen++] = hash;                                             /* :455-458 */
        if (wlen > win) { memmove(window, window + 1, (size_t)(wlen - 1) * sizeof(uint64_t)); wlen--; }

        /* blockmers (:461-505) are skipped: the blockmer set is empty on this path
         * (src/kmer_comp.rs:59; --use-blockmers is a hidden flag, src/cli.rs:160-162). */

        if (snpmer_set && u64map_slot((u64map *)snpmer_set, canonical_kmer, 0)) {   /* :509 */
            int mid_base_qval = 60;
            if (qualities) mid_base_qval = (int)qualities[i + 1 + mid_k - k] - 33;  /* :510-515 */
            if (mid_base_qval > p->min_bq || all_equal_q)                            /* :517 strict > */
                push_snp(&hits, (uint32_t)(i + 1 - k), canonical_kmer);
            if (dedup.n * 2 + 2 >= dedup.cap) {                                     /* grow */
                u64map bigger; u64map_init(&bigger, dedup.cap);
                for (size_t t = 0; t < dedup.cap; t++)
                    if (dedup.keys[t] != UINT64_MAX) *u64map_slot(&bigger, dedup.keys[t], 1) = dedup.vals[t];
                u64map_free(&dedup); dedup = bigger;
            }
            *u64map_slot(&dedup, canonical_kmer & split_mask, 1) += 1;             /* :522-524 */
        }

        if (wlen == win) {                                                 /* :527-543 */
            int middle_idx = (k - s) / 2;
            uint64_t middle_hash = window[middle_idx];
            int syncmer = 1;
            for (int j = 0; j < wlen; j++)
                if (j != middle_idx && window[j] <= middle_hash) { syncmer = 0; break; }
            if (syncmer) push_mini(out, (uint32_t)(i + 1 - k), canonical_kmer, (uint8_t)canonical_marker);
        }
    }

    for (size_t t = 0; t < hits.n_snp; t++) {                              /* :550-568 DEDUP_SNPMERS */
        uint32_t *cnt = u64map_slot(&dedup, hits.snp_kmer[t] & split_mask, 0);
        if (cnt && *cnt == 1) push_snp(out, hits.snp_pos[t], hits.snp_kmer[t]);
    }
    sketch_free(&hits);
    u64map_free(&dedup);
    free(window);

    if (qualities && !all_equal_q) {                                       /* :570-576, :801-817 */
        double sum = 0.0; size_t count = 0;
        for (size_t i = 0; i < len; i++) {
            double q = (double)((int)qualities[i] - 33);
            double pr = pow(10.0, -q / 10.0);
            sum += pr; count++;
        }
        out->est_id = 100. - (sum / (double)count * 100.);
    }

    out->codes = (uint8_t *)malloc(len);                                   /* :604-626 non-ACGT -> A */
    for (size_t i = 0; i < len; i++) out->codes[i] = so_encode_base(string[i]);
}

int so_sketch_batch(const so_params *p, const uint64_t *snpmers, size_t n_snpmers,
                    const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets, size_t n_seqs,
                    uint8_t *status, double *est_id,
                    uint64_t *mini_offsets, uint32_t *mini_pos, uint64_t *mini_kmer, uint8_t *mini_fwd, size_t mini_cap,
                    uint64_t *snp_offsets, uint32_t *snp_pos, uint64_t *snp_kmer, size_t snp_cap,
                    uint8_t *codes) {
    u64map set; u64map_init(&set, n_snpmers);
    for (size_t i = 0; i < n_snpmers; i++) u64map_slot(&set, snpmers[i], 1);
    size_t nm = 0, ns = 0; int rc = 0;
    mini_offsets[0] = 0; snp_offsets[0] = 0;
    for (size_t r = 0; r < n_seqs; r++) {
        size_t off = offsets[r], len = offsets[r + 1] - off;
        sketch_t sk; sketch_one(p, &set, bases + off, quals ? quals + off : NULL, len, &sk);
        status[r] = sk.ok ? 0 : SO_TOO_SHORT;
        if (est_id) est_id[r] = sk.est_id;
        if (sk.ok) {
            if (nm + sk.n_mini > mini_cap || ns + sk.n_snp > snp_cap) { rc = -1; sketch_free(&sk); break; }
            for (size_t t = 0; t < sk.n_mini; t++) {
                mini_pos[nm + t] = sk.mini_pos[t]; mini_kmer[nm + t] = sk.mini_kmer[t];
                if (mini_fwd) mini_fwd[nm + t] = sk.mini_fwd[t];
            }
            for (size_t t = 0; t < sk.n_snp; t++) { snp_pos[ns + t] = sk.snp_pos[t]; snp_kmer[ns + t] = sk.snp_kmer[t]; }
            nm += sk.n_mini; ns += sk.n_snp;
            if (codes) memcpy(codes + off, sk.codes, len);
        } else if (codes) {
            for (size_t t = 0; t < len; t++) codes[off + t] = so_encode_base(bases[off + t]);
        }
        mini_offsets[r + 1] = nm; snp_offsets[r + 1] = ns;
        sketch_free(&sk);
    }
    u64map_free(&set);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* context: ASV sketches + SNPmer inverted index                                        */
/* ------------------------------------------------------------------------------------ */
typedef struct { uint64_t splitmer; uint32_t asv; uint64_t kmer; } idx_entry;

typedef struct {
    uint8_t *codes; size_t len;
    uint64_t *mset; uint8_t *mset_fwd; size_t n_mset;   /* distinct minimizer k-mers, sorted; fwd-any flag */
} asv_t;

struct so_ctx {
    so_params p;
    u64map snpmer_set;
    asv_t *asvs; size_t n_asv;
    idx_entry *index; size_t n_index;                   /* sorted by splitmer */
    int has_filter; so_read_filter rf; u64map high_freq; /* a4 tests inside so_map_reads */
};

so_ctx *so_create(const so_params *p, const uint64_t *snpmers, size_t n_snpmers) {
    so_ctx *c = (so_ctx *)calloc(1, sizeof(so_ctx));
    c->p = *p;
    u64map_init(&c->snpmer_set, n_snpmers);
    for (size_t i = 0; i < n_snpmers; i++) u64map_slot(&c->snpmer_set, snpmers[i], 1);
    return c;
}

static void free_asvs(so_ctx *c) {
    for (size_t i = 0; i < c->n_asv; i++) { free(c->asvs[i].codes); free(c->asvs[i].mset); free(c->asvs[i].mset_fwd); }
    free(c->asvs); free(c->index);
    c->asvs = NULL; c->index = NULL; c->n_asv = 0; c->n_index = 0;
}
void so_destroy(so_ctx *c) {
    if (c && c->high_freq.cap) u64map_free(&c->high_freq);
    if (!c) return;
    free_asvs(c); u64map_free(&c->snpmer_set); free(c);
}
size_t so_num_asvs(const so_ctx *c) { return c->n_asv; }

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}
static int cmp_idx(const void *a, const void *b) {
    const idx_entry *x = (const idx_entry *)a, *y = (const idx_entry *)b;
    if (x->splitmer != y->splitmer) return x->splitmer < y->splitmer ? -1 : 1;
    return x->asv < y->asv ? -1 : x->asv > y->asv;
}

/* distinct k-mers of a sketch, sorted; flag = OR of the forward flags of all occurrences */
static void build_mset(const sketch_t *sk, uint64_t **mset, uint8_t **fwd, size_t *n) {
    size_t cnt = sk->n_mini;
    uint64_t *tmp = (uint64_t *)malloc((cnt + 1) * sizeof(uint64_t));
    for (size_t i = 0; i < cnt; i++) tmp[i] = (sk->mini_kmer[i] << 1) | (uint64_t)(sk->mini_fwd[i] & 1);
    qsort(tmp, cnt, sizeof(uint64_t), cmp_u64);
    uint64_t *keys = (uint64_t *)malloc((cnt + 1) * sizeof(uint64_t));
    uint8_t *flags = (uint8_t *)malloc(cnt + 1);
    size_t m = 0;
    for (size_t i = 0; i < cnt; i++) {
        uint64_t key = tmp[i] >> 1; uint8_t f = (uint8_t)(tmp[i] & 1);
        if (m > 0 && keys[m - 1] == key) flags[m - 1] |= f;
        else { keys[m] = key; flags[m] = f; m++; }
    }
    free(tmp);
    *mset = keys; *fwd = flags; *n = m;
}

/* src/types.rs:622-663 kmer_from_position: canonical by splitmer, tie -> forward */
static uint64_t kmer_from_position(const uint8_t *codes, uint32_t pos, int k) {
    uint64_t kmer = 0;
    for (int t = 0; t < k; t++) kmer = (kmer << 2) | codes[pos + t];
    uint64_t reverse_kmer = so_kmer_revcomp(kmer, k);
    uint64_t mid_mask = so_split_mask(k);
    if ((reverse_kmer & mid_mask) < (kmer & mid_mask)) return reverse_kmer;
    return kmer;
}

/* src/kmer_comp.rs:39-66 (twin_reads_from_fasta) + src/alignment.rs:1754-1765 (index) */
int so_load_asvs(so_ctx *c, const uint8_t *bases, const uint64_t *offsets, size_t n_asv) {
    free_asvs(c);
    c->asvs = (asv_t *)calloc(n_asv ? n_asv : 1, sizeof(asv_t));
    size_t cap = 1024; c->index = (idx_entry *)malloc(cap * sizeof(idx_entry)); c->n_index = 0;
    const uint64_t mask = so_split_mask(c->p.k);
    size_t kept = 0;
    for (size_t a = 0; a < n_asv; a++) {
        size_t off = offsets[a], len = offsets[a + 1] - off;
        sketch_t sk; sketch_one(&c->p, &c->snpmer_set, bases + off, NULL, len, &sk);
        if (!sk.ok) { sketch_free(&sk); return -2; }   /* reference would silently drop it; the ABI rejects */
        asv_t *A = &c->asvs[kept];
        A->codes = sk.codes; sk.codes = NULL; A->len = len;
        build_mset(&sk, &A->mset, &A->mset_fwd, &A->n_mset);
        for (size_t t = 0; t < sk.n_snp; t++) {        /* asv_read.snpmers_vec(): re-extracted from positions */
            uint64_t kmer = kmer_from_position(A->codes, sk.snp_pos[t], c->p.k);
            if (c->n_index == cap) { cap *= 2; c->index = (idx_entry *)realloc(c->index, cap * sizeof(idx_entry)); }
            c->index[c->n_index].splitmer = kmer & mask;
            c->index[c->n_index].asv = (uint32_t)kept;
            c->index[c->n_index].kmer = kmer;
            c->n_index++;
        }
        sketch_free(&sk);
        kept++;
    }
    c->n_asv = kept;
    qsort(c->index, c->n_index, sizeof(idx_entry), cmp_idx);
    return 0;
}

static void index_range(const so_ctx *c, uint64_t splitmer, size_t *lo_out, size_t *hi_out) {
    size_t lo = 0, hi = c->n_index;
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (c->index[mid].splitmer < splitmer) lo = mid + 1; else hi = mid; }
    size_t start = lo; hi = c->n_index;
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (c->index[mid].splitmer <= splitmer) lo = mid + 1; else hi = mid; }
    *lo_out = start; *hi_out = lo;
}

/* ------------------------------------------------------------------------------------ */
/* a8-a10: candidates for one sketched read                                             */
/* ------------------------------------------------------------------------------------ */
typedef struct { uint32_t asv, mm, mini; uint8_t strand; } cand_t;
typedef struct { cand_t *v; size_t n, cap; } cand_vec;
static void cand_push(cand_vec *cv, cand_t x) {
    if (cv->n == cv->cap) { cv->cap = cv->cap ? cv->cap * 2 : 8; cv->v = (cand_t *)realloc(cv->v, cv->cap * sizeof(cand_t)); }
    cv->v[cv->n++] = x;
}
static int cmp_cand_asv(const void *a, const void *b) {
    const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
    return x->asv < y->asv ? -1 : x->asv > y->asv;
}

/* scratch: hits[A], mism[A] zeroed on entry and exit */
static int candidates_one(const so_ctx *c, const sketch_t *rd, uint32_t *hits, uint32_t *mism, cand_vec *kept) {
    const int k = c->p.k;
    const uint64_t mask = so_split_mask(k);
    kept->n = 0;

    /* find_compatible_candidates, src/asv_cluster.rs:356-383 (matches are unused downstream) */
    uint32_t *touched = (uint32_t *)malloc((c->n_asv + 1) * sizeof(uint32_t)); size_t n_touched = 0;
    for (size_t t = 0; t < rd->n_snp; t++) {
        uint64_t q = rd->snp_kmer[t];
        size_t lo, hi; index_range(c, q & mask, &lo, &hi);
        for (size_t e = lo; e < hi; e++) {
            uint32_t a = c->index[e].asv;
            if (hits[a] == 0) touched[n_touched++] = a;
            hits[a]++;
            if (q != c->index[e].kmer) mism[a]++;
        }
    }

    /* read_minimizers: FxHashSet<Kmer48> of the UNPRUNED minimizer_kmers, src/alignment.rs:1788 */
    uint64_t *rset; uint8_t *rfwd; size_t n_rset;
    build_mset(rd, &rset, &rfwd, &n_rset);

    const double min_frac = so_powi(c->p.min_mini_frac_base, k);          /* (0.950f64).powi(k) :1806 */
    cand_vec scores = {0, 0, 0};
    if (c->p.candidate_mode == 1) {
        /* low-polymorphism mode (refine_asv_depths_with_minimap2, src/alignment.rs:1527-1719): no SNPmer index;
         * the reference maps the read against ALL ASVs with minimap2 and keeps hits with mapq > 0.  North-star
         * restatement: every ASV sharing minimizers with the read and passing the same containment test is a
         * candidate with 0 SNPmer mismatches (the mapq filter has no counterpart here). */
        for (size_t ti = 0; ti < n_touched; ti++) { hits[touched[ti]] = 0; mism[touched[ti]] = 0; }
        n_touched = 0;
        for (uint32_t a = 0; a < (uint32_t)c->n_asv; a++) touched[n_touched++] = a;
    }
    for (size_t ti = 0; ti < n_touched; ti++) {                            /* :1796-1814 */
        uint32_t a = touched[ti];
        const asv_t *A = &c->asvs[a];
        size_t m = 0, same = 0, opp = 0;
        for (size_t x = 0; x < n_rset; x++) {
            size_t lo = 0, hi = A->n_mset;
            while (lo < hi) { size_t mid = (lo + hi) / 2; if (A->mset[mid] < rset[x]) lo = mid + 1; else hi = mid; }
            if (lo < A->n_mset && A->mset[lo] == rset[x]) {
                m++;
                if (rfwd[x] == A->mset_fwd[lo]) same++; else opp++;
            }
        }
        if (m == 0) continue;
        size_t mn = n_rset < A->n_mset ? n_rset : A->n_mset;
        if ((double)m / (double)mn < min_frac) continue;
        cand_t cd; cd.asv = a; cd.mm = mism[a]; cd.mini = (uint32_t)m; cd.strand = (uint8_t)(opp > same);
        cand_push(&scores, cd);
    }
    for (size_t ti = 0; ti < n_touched; ti++) { hits[touched[ti]] = 0; mism[touched[ti]] = 0; }
    free(touched); free(rset); free(rfwd);

    if (scores.n == 0) { free(scores.v); return SO_NO_CANDIDATE; }        /* :1817-1820 */

    uint32_t lowest = UINT32_MAX; size_t n_ok = 0;                         /* :1829-1843 */
    for (size_t i = 0; i < scores.n; i++) {
        double ratio = (double)scores.v[i].mm / (double)scores.v[i].mini / (double)c->p.c;   /* :1811 */
        if (ratio <= c->p.ratio_thresh) { n_ok++; if (scores.v[i].mm < lowest) lowest = scores.v[i].mm; }
        else scores.v[i].asv = UINT32_MAX;
    }
    if (n_ok == 0) { free(scores.v); return SO_RATIO_FILTERED; }
    for (size_t i = 0; i < scores.n; i++)
        if (scores.v[i].asv != UINT32_MAX && scores.v[i].mm == lowest) cand_push(kept, scores.v[i]);
    free(scores.v);
    qsort(kept->v, kept->n, sizeof(cand_t), cmp_cand_asv);
    return 0;
}

int so_find_candidates(so_ctx *c, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets, size_t n_reads,
                       uint8_t *cand_status, uint64_t *cand_offsets,
                       uint32_t *cand_asv, uint32_t *cand_mm, uint32_t *cand_mini, uint8_t *cand_strand, size_t cand_cap) {
    uint32_t *hits = (uint32_t *)calloc(c->n_asv + 1, sizeof(uint32_t));
    uint32_t *mism = (uint32_t *)calloc(c->n_asv + 1, sizeof(uint32_t));
    cand_vec kept = {0, 0, 0};
    size_t n = 0; int rc = 0;
    cand_offsets[0] = 0;
    for (size_t r = 0; r < n_reads; r++) {
        size_t off = offsets[r], len = offsets[r + 1] - off;
        sketch_t sk; sketch_one(&c->p, &c->snpmer_set, bases + off, quals ? quals + off : NULL, len, &sk);
        if (!sk.ok) { cand_status[r] = SO_TOO_SHORT; cand_offsets[r + 1] = n; sketch_free(&sk); continue; }
        int st = candidates_one(c, &sk, hits, mism, &kept);
        cand_status[r] = (uint8_t)st;
        if (st == 0) {
            if (n + kept.n > cand_cap) { rc = -1; sketch_free(&sk); break; }
            for (size_t i = 0; i < kept.n; i++) {
                cand_asv[n + i] = kept.v[i].asv; cand_mm[n + i] = kept.v[i].mm;
                cand_mini[n + i] = kept.v[i].mini; cand_strand[n + i] = kept.v[i].strand;
            }
            n += kept.n;
        }
        cand_offsets[r + 1] = n;
        sketch_free(&sk);
    }
    free(hits); free(mism); free(kept.v);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* a11: banded, both-ends-free unit-cost edit distance (SURVEY.md 8c "S2")              */
/*                                                                                      */
/*   rows i = 0..n over the ASV a[1..n], columns j = 0..m over the read r[1..m]         */
/*   (reverse-complemented when strand==1); cell (i,j) exists iff |i-j| <= B.           */
/*   D[0][j] = 0, D[i][0] = 0 (in band); D[i][j] = min(D[i-1][j-1] + (a_i != r_j),      */
/*   D[i-1][j] + 1, D[i][j-1] + 1) over existing cells.                                 */
/*   NM = min over end cells {(n,j)} U {(i,m)}; tie -> largest i+j, then the last-row   */
/*   cell.  Traceback preference: diagonal, then D (consume ASV), then I (consume read);*/
/*   stops at the first cell with i==0 or j==0.                                         */
/* ------------------------------------------------------------------------------------ */
#define SO_INF (1 << 29)

typedef struct {
    int32_t nm, asv_end, read_end, asv_start, read_start;
    uint32_t *cigar; size_t n_cigar, cap_cigar;
} aln_t;

static void cigar_push(aln_t *al, int op) {
    if (al->n_cigar > 0 && (int)(al->cigar[al->n_cigar - 1] & 15u) == op) { al->cigar[al->n_cigar - 1] += 16u; return; }
    if (al->n_cigar == al->cap_cigar) {
        al->cap_cigar = al->cap_cigar ? al->cap_cigar * 2 : 16;
        al->cigar = (uint32_t *)realloc(al->cigar, al->cap_cigar * sizeof(uint32_t));
    }
    al->cigar[al->n_cigar++] = (1u << 4) | (uint32_t)op;
}

static void banded_align(const uint8_t *a, int n, const uint8_t *r, int m, int B, int want_cigar, aln_t *out) {
    const int W = 2 * B + 1;
    int32_t *D = (int32_t *)malloc((size_t)(n + 1) * (size_t)W * sizeof(int32_t));
#define CELL(i, j) D[(size_t)(i) * (size_t)W + (size_t)((j) - (i) + B)]
#define INBAND(i, j) ((i) >= 0 && (j) >= 0 && (i) <= n && (j) <= m && (j) - (i) <= B && (i) - (j) <= B)
    for (int i = 0; i <= n; i++) {
        int jlo = i - B < 0 ? 0 : i - B, jhi = i + B > m ? m : i + B;
        for (int j = jlo; j <= jhi; j++) {
            if (i == 0 || j == 0) { CELL(i, j) = 0; continue; }
            int32_t best = CELL(i - 1, j - 1) + (a[i - 1] != r[j - 1]);
            if (INBAND(i - 1, j)) { int32_t v = CELL(i - 1, j) + 1; if (v < best) best = v; }
            if (INBAND(i, j - 1)) { int32_t v = CELL(i, j - 1) + 1; if (v < best) best = v; }
            CELL(i, j) = best;
        }
    }
    int32_t bd = SO_INF; int bi = -1, bj = -1;
    {
        int jlo = n - B < 0 ? 0 : n - B, jhi = n + B > m ? m : n + B;
        for (int j = jlo; j <= jhi; j++) {           /* last-row cells */
            int32_t d = CELL(n, j);
            if (d < bd || (d == bd && n + j > bi + bj)) { bd = d; bi = n; bj = j; }
        }
        int ilo = m - B < 0 ? 0 : m - B, ihi = m + B > n ? n : m + B;
        for (int i = ilo; i <= ihi; i++) {           /* last-column cells; a last-row cell wins an i+j tie */
            int32_t d = CELL(i, m);
            if (d < bd || (d == bd && i + m > bi + bj)) { bd = d; bi = i; bj = m; }
        }
    }
    out->nm = bd; out->asv_end = bi; out->read_end = bj;
    out->asv_start = -1; out->read_start = -1; out->n_cigar = 0;
    if (want_cigar) {
        int i = bi, j = bj;
        aln_t rev; memset(&rev, 0, sizeof(rev));
        while (i > 0 && j > 0) {
            int32_t d = CELL(i, j);
            if (d == CELL(i - 1, j - 1) + (a[i - 1] != r[j - 1])) { cigar_push(&rev, 0); i--; j--; }
            else if (INBAND(i - 1, j) && d == CELL(i - 1, j) + 1) { cigar_push(&rev, 2); i--; }
            else { cigar_push(&rev, 1); j--; }
        }
        out->asv_start = i; out->read_start = j;
        for (size_t t = rev.n_cigar; t > 0; t--) {
            uint32_t e = rev.cigar[t - 1];
            if (out->n_cigar == out->cap_cigar) {
                out->cap_cigar = out->cap_cigar ? out->cap_cigar * 2 : 16;
                out->cigar = (uint32_t *)realloc(out->cigar, out->cap_cigar * sizeof(uint32_t));
            }
            out->cigar[out->n_cigar++] = e;
        }
        free(rev.cigar);
    }
    free(D);
#undef CELL
#undef INBAND
}

/* ------------------------------------------------------------------------------------ */
/* NM under OTHER definitions -- a sensitivity study of a11, not the product's definition */
/* (so_params.nm_mode; tests/test_nm_semantics.py, tools/nm_semantics.py).   */
/*   mode 1  the S2 recurrence above with a four times wider band (4B).  Without any band */
/*           S2 degenerates: a corner alignment of one base pair scores 0, so the band --  */
/*           every alignment covers >= min(n, m) - B columns -- is part of the definition. */
/*   mode 2  "minimap2-like": the best LOCAL alignment under the scoring of the presets   */
/*           the reference uses (lr:hq / map-ont: match +2, mismatch -4, dual affine gaps */
/*           min(4 + 2l, 24 + l); SURVEY.md 8c / Appendix D -- values from upstream       */
/*           documentation, not checkable here) by Gotoh DP over the full matrix;         */
/*           NM = mismatches + gap bases on the traceback of the best cell (the first one */
/*           in row-major order; traceback preference diagonal, then the short-gap        */
/*           states, then the long-gap states).  minimap2 itself chains exact 19-mer      */
/*           anchors and fills between / beyond them with the same scoring (z-drop 400),  */
/*           which for >= 97 % identical 16S pairs is this optimum up to tie-breaking;    */
/*           a candidate "has no alignment" when the score is below min_dp_max = 80.      */
/* ------------------------------------------------------------------------------------ */
static int affine_local_nm(const uint8_t *a, int n, const uint8_t *r, int m, int32_t *nm_out, int32_t *score_out) {
    const int MA = 2, MI = 4, Q1 = 4, E1 = 2, Q2 = 24, E2 = 1, NEG = -(1 << 28);
    uint8_t *tb = (uint8_t *)malloc((size_t)(n + 1) * (size_t)(m + 1));
    int32_t *H = (int32_t *)malloc((size_t)(m + 1) * sizeof(int32_t));
    int32_t *F1 = (int32_t *)malloc((size_t)(m + 1) * sizeof(int32_t)), *F2 = (int32_t *)malloc((size_t)(m + 1) * sizeof(int32_t));
    for (int j = 0; j <= m; j++) { H[j] = 0; F1[j] = NEG; F2[j] = NEG; tb[j] = 0; }
    int32_t best = 0; int bi = 0, bj = 0;
    for (int i = 1; i <= n; i++) {
        uint8_t *row = tb + (size_t)i * (size_t)(m + 1);
        int32_t diag = H[0], e1 = NEG, e2 = NEG;
        H[0] = 0; row[0] = 0;
        for (int j = 1; j <= m; j++) {
            uint8_t fl = 0;
            /* E: gap consuming read bases (insertion w.r.t. the ASV), along the row */
            int32_t o = H[j - 1] - Q1 - E1, x = e1 - E1;
            if (x > o) { e1 = x; fl |= 8; } else e1 = o;
            o = H[j - 1] - Q2 - E2; x = e2 - E2;
            if (x > o) { e2 = x; fl |= 32; } else e2 = o;
            /* F: gap consuming ASV bases (deletion), along the column; H[j] still holds row i-1 */
            o = H[j] - Q1 - E1; x = F1[j] - E1;
            int32_t f1, f2;
            if (x > o) { f1 = x; fl |= 16; } else f1 = o;
            o = H[j] - Q2 - E2; x = F2[j] - E2;
            if (x > o) { f2 = x; fl |= 64; } else f2 = o;
            F1[j] = f1; F2[j] = f2;
            int32_t h = diag + (a[i - 1] == r[j - 1] ? MA : -MI);
            uint8_t src = 1;
            if (e1 > h) { h = e1; src = 2; }
            if (f1 > h) { h = f1; src = 3; }
            if (e2 > h) { h = e2; src = 4; }
            if (f2 > h) { h = f2; src = 5; }
            if (h <= 0) { h = 0; src = 0; }
            diag = H[j]; H[j] = h;
            row[j] = (uint8_t)(fl | src);
            if (h > best) { best = h; bi = i; bj = j; }
        }
    }
    int32_t nm = 0;
    {
        int i = bi, j = bj, state = 0;                       /* 0 = H, 2..5 = inside a gap of that kind */
        while (i > 0 && j > 0) {
            const uint8_t t = tb[(size_t)i * (size_t)(m + 1) + (size_t)j];
            if (state == 0) {
                const int src = t & 7;
                if (src == 0) break;
                if (src == 1) { nm += a[i - 1] != r[j - 1]; i--; j--; }
                else state = src;
            } else if (state == 2) { nm++; const int ext = t & 8;  j--; if (!ext) state = 0; }
            else if (state == 3)   { nm++; const int ext = t & 16; i--; if (!ext) state = 0; }
            else if (state == 4)   { nm++; const int ext = t & 32; j--; if (!ext) state = 0; }
            else                   { nm++; const int ext = t & 64; i--; if (!ext) state = 0; }
        }
    }
    free(tb); free(H); free(F1); free(F2);
    *nm_out = nm; *score_out = best;
    return best >= 80;
}

/* NM of one oriented pair under so_params.nm_mode; returns 0 when the candidate has no alignment (mode 2 only) */
static int nm_by_mode(const so_params *p, const uint8_t *a, int n, const uint8_t *r, int m, int32_t *nm) {
    if (p->nm_mode == 2) { int32_t sc; return affine_local_nm(a, n, r, m, nm, &sc); }
    aln_t al; memset(&al, 0, sizeof(al));
    banded_align(a, n, r, m, p->nm_mode == 1 ? 4 * p->band : p->band, 0, &al);
    *nm = al.nm;
    return 1;
}

int so_nm_pair(const so_params *p, const uint8_t *asv_codes, int n, const uint8_t *read_codes, int m, int32_t *nm) {
    return nm_by_mode(p, asv_codes, n, read_codes, m, nm);
}

static uint8_t *oriented_codes(const uint8_t *codes, size_t len, int strand) {
    uint8_t *o = (uint8_t *)malloc(len ? len : 1);
    if (!strand) memcpy(o, codes, len);
    else for (size_t i = 0; i < len; i++) o[i] = (uint8_t)(3 - codes[len - 1 - i]);
    return o;
}

int so_align_pairs(so_ctx *c, const uint8_t *bases, const uint64_t *offsets, size_t n_reads,
                   const uint32_t *pair_read, const uint32_t *pair_asv, const uint8_t *pair_strand, size_t n_pairs,
                   int32_t *nm, int32_t *asv_end, int32_t *read_end,
                   int want_cigar, int32_t *asv_start, int32_t *read_start,
                   uint64_t *cigar_offsets, uint32_t *cigar, size_t cigar_cap) {
    size_t nc = 0; int rc = 0;
    if (want_cigar) cigar_offsets[0] = 0;
    for (size_t pi = 0; pi < n_pairs; pi++) {
        uint32_t r = pair_read[pi], a = pair_asv[pi];
        if (r >= n_reads || a >= c->n_asv) return -3;
        size_t off = offsets[r], len = offsets[r + 1] - off;
        uint8_t *codes = (uint8_t *)malloc(len ? len : 1);
        for (size_t t = 0; t < len; t++) codes[t] = so_encode_base(bases[off + t]);
        uint8_t *rd = oriented_codes(codes, len, pair_strand[pi]);
        aln_t al; memset(&al, 0, sizeof(al));
        banded_align(c->asvs[a].codes, (int)c->asvs[a].len, rd, (int)len, c->p.band, want_cigar, &al);
        nm[pi] = al.nm; asv_end[pi] = al.asv_end; read_end[pi] = al.read_end;
        if (want_cigar) {
            asv_start[pi] = al.asv_start; read_start[pi] = al.read_start;
            if (nc + al.n_cigar > cigar_cap) rc = -1;
            else { memcpy(cigar + nc, al.cigar, al.n_cigar * sizeof(uint32_t)); nc += al.n_cigar; }
            cigar_offsets[pi + 1] = nc;
        }
        free(al.cigar); free(rd); free(codes);
        if (rc) break;
    }
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* Stage 4 pileups, alignment::generate_consensus_pileups          */
/* (src/alignment.rs:415-641) over utils::homopolymer_compress_with_quality              */
/* (src/utils.rs:135-184).  The minimap2 `map_ont` call of :439-483 is replaced by the   */
/* north-star DP above on both strands (smaller NM wins, tie -> forward): parity of the  */
/* alignments with minimap2 is UNPINNED, everything around them is restated line by line.*/
/* ------------------------------------------------------------------------------------ */
/* src/utils.rs:135-184 on 2-bit codes (the reference compares the upper-case ACGT characters decoded from the TwinRead's
 * 2-bit dna_seq, src/alignment.rs:452-454); returns the compressed length */
static size_t hpc_with_quality(const uint8_t *codes, const uint8_t *quals, size_t len, int do_hpc, uint8_t *hc, uint8_t *hq,
                               uint8_t *hl) {
    if (len == 0) return 0;
    if (!do_hpc) {
        for (size_t i = 0; i < len; i++) { hc[i] = codes[i]; hq[i] = quals[i]; hl[i] = 1; }
        return len;
    }
    size_t n = 0;
    uint8_t cur = codes[0], cur_len = 1, cur_min = quals[0];
    for (size_t i = 1; i < len; i++) {
        if (codes[i] == cur && cur_len < 255) { cur_len++; if (quals[i] < cur_min) cur_min = quals[i]; }
        else { hc[n] = cur; hq[n] = cur_min; hl[n] = cur_len; n++; cur = codes[i]; cur_len = 1; cur_min = quals[i]; }
    }
    hc[n] = cur; hq[n] = cur_min; hl[n] = cur_len; n++;
    return n;
}

typedef struct { so_pileup_entry *v; size_t n, cap; } pile_col;
static void col_push(pile_col *c, so_pileup_entry e) {
    if (c->n == c->cap) { c->cap = c->cap ? c->cap * 2 : 8; c->v = (so_pileup_entry *)realloc(c->v, c->cap * sizeof(so_pileup_entry)); }
    c->v[c->n++] = e;
}

long long so_consensus_pileups(int band, int use_hpc, const uint8_t *cons_bases, const uint64_t *cons_offsets, size_t n_cons,
                               const uint8_t *read_bases, const uint8_t *read_quals, const uint64_t *read_offsets,
                               const uint32_t *read_cons, size_t n_reads, uint64_t *pos_offsets, so_pileup_entry *entries,
                               size_t entry_cap, uint8_t *ref_hp_length, uint8_t *read_strand, int32_t *read_nm) {
    static const char ACGT[4] = {'A', 'C', 'G', 'T'};
    const size_t total_pos = (size_t)(cons_offsets[n_cons] - cons_offsets[0]);
    pile_col *cols = (pile_col *)calloc(total_pos ? total_pos : 1, sizeof(pile_col));
    for (size_t r = 0; r < n_reads; r++) {                                    /* :447-450, reads in cluster order */
        read_strand[r] = 0xff; read_nm[r] = -1;
        const uint32_t ci = read_cons[r];
        if (ci >= n_cons) { free(cols); return -3; }
        const size_t coff = (size_t)(cons_offsets[ci] - cons_offsets[0]), clen = (size_t)(cons_offsets[ci + 1] - cons_offsets[ci]);
        const size_t off = read_offsets[r], len = read_offsets[r + 1] - off;
        uint8_t *codes = (uint8_t *)malloc(len + 1), *q = (uint8_t *)malloc(len + 1);
        for (size_t t = 0; t < len; t++) { codes[t] = so_encode_base(read_bases[off + t]); q[t] = read_quals ? read_quals[off + t] : 33; }
        uint8_t *hc = (uint8_t *)malloc(len + 1), *hq = (uint8_t *)malloc(len + 1), *hl = (uint8_t *)malloc(len + 1);
        const size_t hn = hpc_with_quality(codes, q, len, use_hpc, hc, hq, hl);             /* :481 */
        if (hn > 0 && clen > 0) {              /* minimap2 returns Err for an empty sequence: the read is skipped (:486) */
            uint8_t *ccodes = (uint8_t *)malloc(clen);
            for (size_t t = 0; t < clen; t++) ccodes[t] = so_encode_base(cons_bases[cons_offsets[ci] + t]);
            aln_t al[2]; memset(al, 0, sizeof(al));
            uint8_t *ori[2];
            for (int sd = 0; sd < 2; sd++) {
                ori[sd] = oriented_codes(hc, hn, sd);
                banded_align(ccodes, (int)clen, ori[sd], (int)hn, band, 1, &al[sd]);
            }
            const int sd = al[1].nm < al[0].nm ? 1 : 0;
            read_strand[r] = (uint8_t)sd; read_nm[r] = al[sd].nm;
            /* :497-521: the read in alignment orientation; qualities and run lengths reversed with it */
            const uint8_t *fseq = ori[sd];
            uint8_t *fq = (uint8_t *)malloc(hn), *fl = (uint8_t *)malloc(hn);
            for (size_t t = 0; t < hn; t++) { fq[t] = sd ? hq[hn - 1 - t] : hq[t]; fl[t] = sd ? hl[hn - 1 - t] : hl[t]; }
            const size_t qs = (size_t)al[sd].read_start, qe = (size_t)al[sd].read_end;
            const size_t mlen = qe - qs;                                               /* mapped_seq = final_seq[qs..qe] */
            size_t ref_pos = (size_t)al[sd].asv_start, query_pos = 0;                  /* :524-525 */
            for (size_t t = 0; t < al[sd].n_cigar; t++) {
                const size_t ln = al[sd].cigar[t] >> 4; const int op = (int)(al[sd].cigar[t] & 15u);
                if (op == 0) {                                                          /* :531-543 */
                    for (size_t j = 0; j < ln; j++)
                        if (ref_pos + j < clen && query_pos + j < mlen) {
                            so_pileup_entry e; memset(&e, 0, sizeof(e));
                            e.read = (uint32_t)r; e.kind = 0; e.n = 1;
                            e.base[0] = (uint8_t)ACGT[fseq[qs + query_pos + j]]; e.qual[0] = fq[qs + query_pos + j]; e.hp[0] = fl[qs + query_pos + j];
                            col_push(&cols[coff + ref_pos + j], e);
                        }
                    ref_pos += ln; query_pos += ln;
                } else if (op == 1) {                                                   /* :544-557, MAX_INSERTION_LENGTH = 2 */
                    if (ref_pos > 0 && ref_pos - 1 < clen && query_pos + ln <= mlen) {
                        so_pileup_entry e; memset(&e, 0, sizeof(e));
                        e.read = (uint32_t)r; e.kind = 2; e.n = (uint8_t)(ln < 2 ? ln : 2);
                        for (size_t j = 0; j < e.n; j++) {
                            e.base[j] = (uint8_t)ACGT[fseq[qs + query_pos + j]]; e.qual[j] = fq[qs + query_pos + j]; e.hp[j] = fl[qs + query_pos + j];
                        }
                        col_push(&cols[coff + ref_pos - 1], e);
                    }
                    query_pos += ln;
                } else {                                                                /* :558-566 */
                    for (size_t j = 0; j < ln; j++)
                        if (ref_pos + j < clen) {
                            so_pileup_entry e; memset(&e, 0, sizeof(e));
                            e.read = (uint32_t)r; e.kind = 1;
                            col_push(&cols[coff + ref_pos + j], e);
                        }
                    ref_pos += ln;
                }
            }
            for (int t = 0; t < 2; t++) { free(al[t].cigar); free(ori[t]); }
            free(fq); free(fl); free(ccodes);
        }
        free(codes); free(q); free(hc); free(hq); free(hl);
    }
    /* flatten + the median run length per position (:589-622: sorted hp lengths of the Base entries, even count ->
     * (s[mid-1] + s[mid]) / 2 in u16, no bases -> 1) */
    size_t ne = 0; long long rc = 0;
    pos_offsets[0] = 0;
    for (size_t p = 0; p < total_pos; p++) {
        uint32_t hist[256]; memset(hist, 0, sizeof(hist));
        size_t nb = 0;
        for (size_t t = 0; t < cols[p].n; t++) {
            if (rc == 0) { if (ne < entry_cap) entries[ne] = cols[p].v[t]; else rc = -1; }
            ne++;
            if (cols[p].v[t].kind == 0) { hist[cols[p].v[t].hp[0]]++; nb++; }
        }
        pos_offsets[p + 1] = ne;
        if (nb == 0) ref_hp_length[p] = 1;
        else {
            const size_t mid = nb / 2;
            int lo = -1, hi = -1; size_t acc = 0;                    /* sorted[mid-1], sorted[mid] */
            for (int v = 0; v < 256; v++) {
                acc += hist[v];
                if (lo < 0 && mid >= 1 && acc >= mid) lo = v;       /* index mid-1 is covered once acc >= mid */
                if (hi < 0 && acc >= mid + 1) { hi = v; break; }
            }
            ref_hp_length[p] = (nb % 2 == 0) ? (uint8_t)(((uint16_t)lo + (uint16_t)hi) / 2) : (uint8_t)hi;
        }
        free(cols[p].v);
    }
    free(cols);
    return rc ? rc : (long long)ne;
}

/* ------------------------------------------------------------------------------------ */
/* SURVEY.md 8(f) row 3, Stage 2: TwinRead::compute_lsh_signatures (src/types.rs:719-747) */
/* and asv_cluster::cluster_reads_by_kmers (src/asv_cluster.rs:72-249, bucketed branch).  */
/* FxHasher64 is the fxhash crate 0.2.1 (Cargo.lock:503-505; not under /root/reference):  */
/* write_u64(w): hash = (rotl(hash, 5) ^ w) * 0x517cc1b727220a95, finish() = hash.        */
/* ------------------------------------------------------------------------------------ */
#define SO_LSH_NUM_TABLES 20            /* src/constants.rs:67 */
#define SO_LSH_BUCKET_SIZE 3            /* src/constants.rs:68 */
static uint64_t fx_write_u64(uint64_t h, uint64_t w) { return (((h << 5) | (h >> 59)) ^ w) * 0x517cc1b727220a95ull; }

typedef struct { uint64_t hash, kmer; size_t idx; } lsh_item;
static int cmp_lsh_item(const void *a, const void *b) {          /* sort_by_key(hash) is stable: ties keep input order */
    const lsh_item *x = (const lsh_item *)a, *y = (const lsh_item *)b;
    if (x->hash != y->hash) return x->hash < y->hash ? -1 : 1;
    return x->idx < y->idx ? -1 : x->idx > y->idx;
}

int so_lsh_signatures(const uint64_t *kmers, const uint64_t *offsets, size_t n_reads, uint64_t *sig, uint8_t *valid) {
#pragma omp parallel for schedule(dynamic, 64)
    for (size_t r = 0; r < n_reads; r++) {
        const uint64_t *km = kmers + offsets[r];
        const size_t n = (size_t)(offsets[r + 1] - offsets[r]);
        valid[r] = n >= SO_LSH_BUCKET_SIZE;                                              /* :728-730 */
        lsh_item *it = (lsh_item *)malloc((n ? n : 1) * sizeof(lsh_item));
        for (int t = 0; t < SO_LSH_NUM_TABLES; t++) {
            sig[r * SO_LSH_NUM_TABLES + t] = 0;
            if (!valid[r]) continue;
            for (size_t i = 0; i < n; i++) { it[i].hash = fx_write_u64(fx_write_u64(0, (uint64_t)t), km[i]); it[i].kmer = km[i]; it[i].idx = i; }
            qsort(it, n, sizeof(lsh_item), cmp_lsh_item);                                /* :741 */
            uint64_t sg = 0;
            for (int i = 0; i < SO_LSH_BUCKET_SIZE; i++) sg ^= it[i].kmer * (uint64_t)(i + 1);   /* :743-745 */
            sig[r * SO_LSH_NUM_TABLES + t] = sg;
        }
        free(it);
    }
    return 0;
}

typedef struct { uint32_t *v; size_t n, cap; } u32vec;
static void u32vec_push(u32vec *a, uint32_t x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 4; a->v = (uint32_t *)realloc(a->v, a->cap * sizeof(uint32_t)); }
    a->v[a->n++] = x;
}
typedef struct { uint32_t id, hits; } lsh_cand;
static int cmp_lsh_cand(const void *a, const void *b) {          /* (b.1, b.0).cmp(&(a.1, a.0)): hits desc, then id desc */
    const lsh_cand *x = (const lsh_cand *)a, *y = (const lsh_cand *)b;
    if (x->hits != y->hits) return x->hits > y->hits ? -1 : 1;
    return x->id > y->id ? -1 : x->id < y->id;
}

int so_cluster_reads_by_kmers(int k, double threshold, const uint64_t *kmers, const uint64_t *offsets, size_t n_reads, uint32_t *rep_of,
                              uint64_t *n_reps_out) {
    const size_t top_n_candidates = 10;                                                  /* :85 */
    uint64_t *sig = (uint64_t *)malloc((n_reads ? n_reads : 1) * SO_LSH_NUM_TABLES * sizeof(uint64_t));
    uint8_t *valid = (uint8_t *)malloc(n_reads ? n_reads : 1);
    so_lsh_signatures(kmers, offsets, n_reads, sig, valid);
    /* sorted distinct k-mers per read (the reference's HashSet of the query / Vec::contains on the representative) */
    uint64_t *dk = (uint64_t *)malloc((size_t)(offsets[n_reads] ? offsets[n_reads] : 1) * sizeof(uint64_t));
    uint32_t *nd = (uint32_t *)malloc((n_reads ? n_reads : 1) * sizeof(uint32_t));
#pragma omp parallel for schedule(dynamic, 64)
    for (size_t r = 0; r < n_reads; r++) {
        const size_t n = (size_t)(offsets[r + 1] - offsets[r]);
        uint64_t *d = dk + offsets[r];
        memcpy(d, kmers + offsets[r], n * sizeof(uint64_t));
        qsort(d, n, sizeof(uint64_t), cmp_u64);
        size_t m = 0;
        for (size_t i = 0; i < n; i++) if (i == 0 || d[i] != d[i - 1]) d[m++] = d[i];
        nd[r] = (uint32_t)m;
    }
    u64map tab[SO_LSH_NUM_TABLES];                         /* signature -> index into lists (+1) */
    u32vec *lists = NULL; size_t n_lists = 0, cap_lists = 0;
    for (int t = 0; t < SO_LSH_NUM_TABLES; t++) u64map_init(&tab[t], 1024);
    uint8_t *hits = (uint8_t *)calloc(n_reads ? n_reads : 1, 1);
    lsh_cand *cand = (lsh_cand *)malloc((n_reads ? n_reads : 1) * sizeof(lsh_cand));
    uint64_t n_reps = 0;
    for (size_t i = 0; i < n_reads; i++) {                                               /* :100 */
        size_t nc = 0;
        if (valid[i])
            for (int t = 0; t < SO_LSH_NUM_TABLES; t++) {                                /* query_read_against_bucket_index :299-337 */
                const uint32_t *slot = u64map_slot(&tab[t], sig[i * SO_LSH_NUM_TABLES + t], 0);
                if (!slot) continue;
                const u32vec *l = &lists[*slot - 1];
                for (size_t q = 0; q < l->n; q++) { if (hits[l->v[q]]++ == 0) cand[nc++].id = l->v[q]; }
            }
        long best = -1;
        if (nc) {
            for (size_t q = 0; q < nc; q++) { cand[q].hits = hits[cand[q].id]; hits[cand[q].id] = 0; }
            qsort(cand, nc, sizeof(lsh_cand), cmp_lsh_cand);                             /* :111 */
            const uint32_t max_hits = cand[0].hits;
            double best_similarity = 0.0;
            size_t checked = 0;
            for (size_t q = 0; q < nc; q++) {                                            /* :118-125 */
                if (!(cand[q].hits == max_hits || checked < top_n_candidates)) break;
                checked++;
                const uint32_t c = cand[q].id;
                size_t count = 0;                                                        /* :133-141 */
                for (uint32_t a = 0; a < nd[i]; a++)
                    if (bsearch(&dk[offsets[i] + a], dk + offsets[c], nd[c], sizeof(uint64_t), cmp_u64)) count++;
                const size_t rep_len = (size_t)(offsets[c + 1] - offsets[c]);
                const size_t den = nd[i] > rep_len ? nd[i] : rep_len;
                const double ratio = (double)count / (double)den;                        /* :143 */
                const double similarity = pow(ratio, 1.0 / (double)k);                   /* :144 */
                if (similarity > best_similarity) { best_similarity = similarity; best = (long)c; }
            }
            if (!(best_similarity > threshold)) best = -1;                               /* :152-156 */
        }
        if (best >= 0) rep_of[i] = (uint32_t)best;
        else {                                                                           /* :178-187 */
            if (valid[i])
                for (int t = 0; t < SO_LSH_NUM_TABLES; t++) {
                    if (tab[t].n * 2 + 2 >= tab[t].cap) {                                /* grow */
                        u64map g; u64map_init(&g, tab[t].cap);
                        for (size_t q = 0; q < tab[t].cap; q++)
                            if (tab[t].keys[q] != UINT64_MAX) *u64map_slot(&g, tab[t].keys[q], 1) = tab[t].vals[q];
                        u64map_free(&tab[t]); tab[t] = g;
                    }
                    uint32_t *slot = u64map_slot(&tab[t], sig[i * SO_LSH_NUM_TABLES + t], 1);
                    if (*slot == 0) {
                        if (n_lists == cap_lists) { cap_lists = cap_lists ? cap_lists * 2 : 1024; lists = (u32vec *)realloc(lists, cap_lists * sizeof(u32vec)); }
                        memset(&lists[n_lists], 0, sizeof(u32vec));
                        *slot = (uint32_t)(++n_lists);
                    }
                    u32vec_push(&lists[*slot - 1], (uint32_t)i);
                }
            rep_of[i] = (uint32_t)i; n_reps++;
        }
    }
    for (size_t q = 0; q < n_lists; q++) free(lists[q].v);
    for (int t = 0; t < SO_LSH_NUM_TABLES; t++) u64map_free(&tab[t]);
    free(lists); free(hits); free(cand); free(sig); free(valid); free(dk); free(nd);
    if (n_reps_out) *n_reps_out = n_reps;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* SURVEY.md 8(f) row 3, Stage 3 reclustering: the read-vs-consensus SNPmer comparison of */
/* asv_cluster::reassign_reads_to_best_cluster (src/asv_cluster.rs:1007-1131, SNPmer       */
/* markers) and compare_consensus_snpmers (:968-994).  The consensus lists (built by        */
/* build_consensus_snpmers, :840-893, whose ties follow hash-map iteration order) are input.*/
/* ------------------------------------------------------------------------------------ */
int so_snpmer_compare(int k, const uint64_t *cons_kmers, const uint64_t *cons_offsets, size_t n_cons, const uint64_t *read_kmers,
                      const uint64_t *read_offsets, size_t n_reads, uint32_t *best_cons, uint32_t *best_mismatches, uint32_t *best_matches,
                      uint16_t *matrix) {
    const uint64_t mask = so_split_mask(k);
    /* splitmer_to_consensus_kmers[c]: insert(cs.splitmer, cs.kmer), a later entry of the same splitmer replaces an earlier (:1027-1037) */
    u64map *maps = (u64map *)malloc((n_cons ? n_cons : 1) * sizeof(u64map));
    for (size_t c = 0; c < n_cons; c++) {
        const size_t n = (size_t)(cons_offsets[c + 1] - cons_offsets[c]);
        u64map_init(&maps[c], n);
        for (size_t i = 0; i < n; i++) *u64map_slot(&maps[c], cons_kmers[cons_offsets[c] + i] & mask, 1) = (uint32_t)(i + 1);
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t r = 0; r < n_reads; r++) {
        size_t best = 0; uint64_t b_mm = UINT64_MAX, b_m = 0;                       /* best_score = (usize::MAX, 0), :1055 */
        for (size_t c = 0; c < n_cons; c++) {
            uint32_t m = 0, mm = 0;
            for (uint64_t i = read_offsets[r]; i < read_offsets[r + 1]; i++) {         /* :1064-1084 */
                const uint32_t *slot = u64map_slot(&maps[c], read_kmers[i] & mask, 0);
                if (slot) { if (cons_kmers[cons_offsets[c] + *slot - 1] == read_kmers[i]) m++; else mm++; }
            }
            if (matrix) { matrix[(r * n_cons + c) * 2] = (uint16_t)(m > 65535 ? 65535 : m); matrix[(r * n_cons + c) * 2 + 1] = (uint16_t)(mm > 65535 ? 65535 : mm); }
            if (mm < b_mm || (mm == b_mm && m > b_m)) { b_mm = mm; b_m = m; best = c; }   /* :1087-1091 */
        }
        best_cons[r] = n_cons ? (uint32_t)best : UINT32_MAX;
        best_mismatches[r] = n_cons ? (uint32_t)b_mm : 0; best_matches[r] = (uint32_t)b_m;
    }
    for (size_t c = 0; c < n_cons; c++) u64map_free(&maps[c]);
    free(maps);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Stage 3 greedy loop, asv_cluster::cluster_reads_by_snpmers (src/asv_cluster.rs:561-700) */
/* for one k-mer cluster -- restated ahead of a device implementation (no product code    */
/* calls this yet).  Indexed branch (:467-510, blockmers off): candidates with 0          */
/* mismatches and > 0 matches, ordered by (-matches, cluster size, read id).  Above 1000   */
/* representatives the reference switches to rayon find_any (:413-464), whose choice is    */
/* not deterministic; this restatement takes the FIRST compatible representative in        */
/* creation order there (what find_first would return).                                    */
/* ------------------------------------------------------------------------------------ */
int so_cluster_reads_by_snpmers(int k, const uint64_t *snp_kmers, const uint64_t *snp_offsets, const uint32_t *cluster, size_t n,
                                uint32_t iterative_above, uint32_t *rep_of) {
    const uint64_t mask = so_split_mask(k);
    /* index: splitmer -> list of (rep slot, kmer), src/asv_cluster.rs:340-353 */
    u64map idx; u64map_init(&idx, 1024);
    typedef struct { uint32_t *rep; uint64_t *kmer; size_t n, cap; } bucket_t;
    bucket_t *buckets = NULL; size_t n_buckets = 0, cap_buckets = 0;
    uint32_t *reps = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));          /* representative read ids in creation order */
    uint32_t *rep_size = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
    uint32_t *m = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t)), *mm = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
    uint32_t *touched = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    size_t n_reps = 0;
    for (size_t ci = 0; ci < n; ci++) {
        const uint32_t read = cluster[ci];
        const uint64_t *q = snp_kmers + snp_offsets[read];
        const size_t nq = (size_t)(snp_offsets[read + 1] - snp_offsets[read]);
        long best = -1;
        if (n_reps > iterative_above) {                                           /* :616-632 */
            for (size_t r = 0; r < n_reps && best < 0; r++) {
                const uint32_t rid = reps[r];
                uint32_t ma = 0, mi = 0;
                for (uint64_t i = snp_offsets[rid]; i < snp_offsets[rid + 1]; i++)   /* rep SNPmers against the read's map (later entries win) */
                    for (size_t j = nq; j-- > 0;)
                        if ((q[j] & mask) == (snp_kmers[i] & mask)) { if (q[j] == snp_kmers[i]) ma++; else mi++; break; }
                if (mi == 0 && ma > 0) best = (long)r;
            }
        } else {                                                                  /* find_compatible_candidates + :483-509 */
            size_t nt = 0;
            for (size_t j = 0; j < nq; j++) {
                const uint32_t *slot = u64map_slot(&idx, q[j] & mask, 0);
                if (!slot) continue;
                const bucket_t *b = &buckets[*slot - 1];
                for (size_t e = 0; e < b->n; e++) {
                    const uint32_t r = b->rep[e];
                    if (m[r] == 0 && mm[r] == 0) touched[nt++] = r;
                    if (b->kmer[e] == q[j]) m[r]++; else mm[r]++;
                }
            }
            for (size_t t = 0; t < nt; t++) {
                const uint32_t r = touched[t];
                if (mm[r] == 0 && m[r] > 0) {
                    if (best < 0) best = (long)r;
                    else {
                        const uint32_t b = (uint32_t)best;     /* sort key (-matches, size, read id) */
                        if (m[r] > m[b] || (m[r] == m[b] && (rep_size[r] < rep_size[b] || (rep_size[r] == rep_size[b] && reps[r] < reps[b])))) best = (long)r;
                    }
                }
            }
            for (size_t t = 0; t < nt; t++) { m[touched[t]] = 0; mm[touched[t]] = 0; }
        }
        if (best >= 0) { rep_of[read] = reps[best]; rep_size[best]++; }           /* :386-394 */
        else {                                                                    /* :397-410 */
            const uint32_t r = (uint32_t)n_reps;
            reps[n_reps++] = read; rep_size[r] = 1; rep_of[read] = read;
            for (size_t j = 0; j < nq; j++) {
                if (idx.n * 2 + 2 >= idx.cap) {
                    u64map g; u64map_init(&g, idx.cap);
                    for (size_t x = 0; x < idx.cap; x++) if (idx.keys[x] != UINT64_MAX) *u64map_slot(&g, idx.keys[x], 1) = idx.vals[x];
                    u64map_free(&idx); idx = g;
                }
                uint32_t *slot = u64map_slot(&idx, q[j] & mask, 1);
                if (*slot == 0) {
                    if (n_buckets == cap_buckets) { cap_buckets = cap_buckets ? cap_buckets * 2 : 256; buckets = (bucket_t *)realloc(buckets, cap_buckets * sizeof(bucket_t)); }
                    memset(&buckets[n_buckets], 0, sizeof(bucket_t));
                    *slot = (uint32_t)(++n_buckets);
                }
                bucket_t *b = &buckets[*slot - 1];
                if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 4; b->rep = (uint32_t *)realloc(b->rep, b->cap * 4); b->kmer = (uint64_t *)realloc(b->kmer, b->cap * 8); }
                b->rep[b->n] = r; b->kmer[b->n] = q[j]; b->n++;
            }
        }
    }
    for (size_t x = 0; x < n_buckets; x++) { free(buckets[x].rep); free(buckets[x].kmer); }
    free(buckets); u64map_free(&idx); free(reps); free(rep_size); free(m); free(mm); free(touched);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* a4: kmer_comp::twin_reads_from_snpmers, src/kmer_comp.rs:68-258                      */
/* ------------------------------------------------------------------------------------ */
/* the tests a read goes through before it becomes a TwinRead; the quality cutoff (:248) is applied after the id sort in
 * the reference but is a per-read predicate, so it is evaluated here as the last test */
static int read_filter_one(const so_params *p, const so_read_filter *f, const u64map *high_freq, size_t len, const sketch_t *sk,
                           uint32_t *n_solid_mini, uint32_t *n_solid_snp) {
    *n_solid_mini = 0; *n_solid_snp = 0;
    if (len < f->min_read_length || (f->max_read_length && len > f->max_read_length)) return SO_LENGTH_FILTERED;   /* :117-120 */
    if (!sk->ok) return SO_TOO_SHORT;                                            /* :160-161 twin_read.is_some() */
    u64map counter; u64map_init(&counter, sk->n_mini + 1);                       /* :163-166 kmer_counter_map */
    for (size_t i = 0; i < sk->n_mini; i++) *u64map_slot(&counter, sk->mini_kmer[i], 1) += 1;
    uint32_t solid = 0;
    for (size_t i = 0; i < sk->n_mini; i++) {                                    /* :168-183 */
        if (*u64map_slot(&counter, sk->mini_kmer[i], 0) > f->max_kmer_count_in_read) continue;
        if (!(high_freq->cap && u64map_slot((u64map *)high_freq, sk->mini_kmer[i], 0))) solid++;   /* USE_SOLID_KMERS = false */
    }
    u64map_free(&counter);
    if (solid < len / (size_t)p->c / 20) return SO_REPETITIVE;                   /* :185-188 */
    uint32_t ssnp = 0;
    for (size_t i = 0; i < sk->n_snp; i++)                                       /* :190-203 */
        if (!(high_freq->cap && u64map_slot((u64map *)high_freq, sk->snp_kmer[i], 0))) ssnp++;
    *n_solid_mini = solid; *n_solid_snp = ssnp;
    if (f->quality_value_cutoff > 0.0 && !isnan(sk->est_id) && sk->est_id < f->quality_value_cutoff) return SO_LOW_QUALITY; /* :248 */
    return 0;
}

typedef struct { const uint8_t *id; size_t id_len; double est; uint32_t idx; } order_t;
static int cmp_order_id(const void *a, const void *b) {      /* String::cmp = byte-wise, shorter first on a common prefix */
    const order_t *x = (const order_t *)a, *y = (const order_t *)b;
    size_t m = x->id_len < y->id_len ? x->id_len : y->id_len;
    int c = m ? memcmp(x->id, y->id, m) : 0;
    if (c) return c;
    if (x->id_len != y->id_len) return x->id_len < y->id_len ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);        /* stable */
}
static void merge_sort_est_desc(order_t *v, order_t *tmp, size_t n) {   /* stable, src/main.rs:538 */
    if (n < 2) return;
    size_t h = n / 2;
    merge_sort_est_desc(v, tmp, h); merge_sort_est_desc(v + h, tmp, n - h);
    size_t i = 0, j = h, o = 0;
    while (i < h && j < n) tmp[o++] = (v[j].est > v[i].est) ? v[j++] : v[i++];   /* take right only if strictly larger */
    while (i < h) tmp[o++] = v[i++];
    while (j < n) tmp[o++] = v[j++];
    memcpy(v, tmp, n * sizeof(order_t));
}

int so_twin_reads_filter(const so_params *p, const so_read_filter *f, const uint64_t *snpmers, size_t n_snpmers,
                         const uint64_t *high_freq, size_t n_high_freq,
                         const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets, size_t n_reads,
                         const uint8_t *ids, const uint64_t *id_offsets,
                         uint8_t *status, double *est_id, uint32_t *n_solid_mini, uint32_t *n_solid_snp,
                         uint32_t *order, size_t *n_kept) {
    u64map set; u64map_init(&set, n_snpmers);
    for (size_t i = 0; i < n_snpmers; i++) u64map_slot(&set, snpmers[i], 1);
    u64map hf; memset(&hf, 0, sizeof(hf));
    if (n_high_freq) { u64map_init(&hf, n_high_freq); for (size_t i = 0; i < n_high_freq; i++) u64map_slot(&hf, high_freq[i], 1); }
    order_t *ord = (order_t *)malloc((n_reads + 1) * sizeof(order_t));
    size_t nk = 0;
    for (size_t r = 0; r < n_reads; r++) {
        size_t off = offsets[r], len = offsets[r + 1] - off;
        sketch_t sk; sketch_one(p, &set, bases + off, quals ? quals + off : NULL, len, &sk);
        uint32_t sm = 0, ss = 0;
        int st = read_filter_one(p, f, &hf, len, &sk, &sm, &ss);
        status[r] = (uint8_t)st;
        if (est_id) est_id[r] = sk.ok ? sk.est_id : NAN;
        if (n_solid_mini) n_solid_mini[r] = sm;
        if (n_solid_snp) n_solid_snp[r] = ss;
        if (st == 0) {
            ord[nk].idx = (uint32_t)r; ord[nk].est = isnan(sk.est_id) ? 100.0 : sk.est_id;   /* unwrap_or(100.0) */
            ord[nk].id = ids ? ids + id_offsets[r] : NULL; ord[nk].id_len = ids ? (size_t)(id_offsets[r + 1] - id_offsets[r]) : 0;
            nk++;
        }
        sketch_free(&sk);
    }
    if (ids) qsort(ord, nk, sizeof(order_t), cmp_order_id);                      /* :233 */
    order_t *tmp = (order_t *)malloc((nk + 1) * sizeof(order_t));
    merge_sort_est_desc(ord, tmp, nk);                                           /* src/main.rs:538 */
    if (order) for (size_t i = 0; i < nk; i++) order[i] = ord[i].idx;
    if (n_kept) *n_kept = nk;
    free(tmp); free(ord); u64map_free(&set); if (hf.cap) u64map_free(&hf);
    return 0;
}

void so_set_read_filter(so_ctx *c, const so_read_filter *f, const uint64_t *high_freq, size_t n_high_freq) {
    if (c->high_freq.cap) { u64map_free(&c->high_freq); memset(&c->high_freq, 0, sizeof(c->high_freq)); }
    c->has_filter = f != NULL;
    if (!f) return;
    c->rf = *f;
    if (n_high_freq) { u64map_init(&c->high_freq, n_high_freq); for (size_t i = 0; i < n_high_freq; i++) u64map_slot(&c->high_freq, high_freq[i], 1); }
}

/* ------------------------------------------------------------------------------------ */
/* whole path, src/alignment.rs:1786-1925 (Stage 7b duplicate :2095-2171)               */
/* ------------------------------------------------------------------------------------ */
int so_map_reads(so_ctx *c, const uint8_t *bases, const uint8_t *quals, const uint64_t *offsets,
                 const uint32_t *file_idx, size_t n_reads, int n_samples,
                 uint8_t *status, int32_t *best_nm_out, uint64_t *class_offsets, uint32_t *class_asv, size_t class_cap,
                 uint64_t *unambig, uint64_t *ambig, uint64_t *leq10, uint64_t *totals, uint64_t *sample_singleton,
                 uint64_t *dp_cells, int n_threads) {
    const size_t A = c->n_asv;
    uint32_t **cls = (uint32_t **)calloc(n_reads ? n_reads : 1, sizeof(uint32_t *));
    uint32_t *cls_n = (uint32_t *)calloc(n_reads ? n_reads : 1, sizeof(uint32_t));
    uint64_t cells_total = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel reduction(+ : cells_total)
    {
        uint32_t *hits = (uint32_t *)calloc(A + 1, sizeof(uint32_t));
        uint32_t *mism = (uint32_t *)calloc(A + 1, sizeof(uint32_t));
        cand_vec kept = {0, 0, 0};
#pragma omp for schedule(dynamic, 16)
        for (long long rr = 0; rr < (long long)n_reads; rr++) {
            size_t r = (size_t)rr;
            size_t off = offsets[r], len = offsets[r + 1] - off;
            best_nm_out[r] = INT32_MAX;
            sketch_t sk; sketch_one(&c->p, &c->snpmer_set, bases + off, quals ? quals + off : NULL, len, &sk);
            if (c->has_filter) {                       /* a4: the read must survive twin_reads_from_snpmers to reach Stage 7 */
                uint32_t sm, ss;
                int fs = read_filter_one(&c->p, &c->rf, &c->high_freq, len, &sk, &sm, &ss);
                if (fs != 0) { status[r] = (uint8_t)fs; sketch_free(&sk); continue; }
            }
            if (!sk.ok) { status[r] = SO_TOO_SHORT; sketch_free(&sk); continue; }
            int st = candidates_one(c, &sk, hits, mism, &kept);
            if (st != 0) { status[r] = (uint8_t)st; sketch_free(&sk); continue; }

            /* :1846-1863 -- NM per kept candidate (north-star DP instead of minimap2) */
            int32_t *nms = (int32_t *)malloc(kept.n * sizeof(int32_t));
            uint8_t *fw = oriented_codes(sk.codes, len, 0), *rv = oriented_codes(sk.codes, len, 1);
            for (size_t i = 0; i < kept.n; i++) {
                const asv_t *Av = &c->asvs[kept.v[i].asv];
                if (c->p.strand_mode == 1) {
                    int32_t n1 = INT32_MAX, n2 = INT32_MAX;
                    const int ok1 = nm_by_mode(&c->p, Av->codes, (int)Av->len, fw, (int)len, &n1);
                    const int ok2 = nm_by_mode(&c->p, Av->codes, (int)Av->len, rv, (int)len, &n2);
                    if (!ok1) n1 = INT32_MAX;
                    if (!ok2) n2 = INT32_MAX;
                    nms[i] = n2 < n1 ? n2 : n1;
                    cells_total += 2ull * Av->len * (uint64_t)(2 * c->p.band + 1);
                } else {
                    int32_t n1 = INT32_MAX;
                    if (!nm_by_mode(&c->p, Av->codes, (int)Av->len, kept.v[i].strand ? rv : fw, (int)len, &n1)) n1 = INT32_MAX;
                    nms[i] = n1;
                    cells_total += (uint64_t)Av->len * (uint64_t)(2 * c->p.band + 1);
                }
            }
            free(fw); free(rv);
            /* :1865-1873 best_nm and its ties; kept is sorted by asv id so the class is sorted (:1893) */
            int32_t best = INT32_MAX;
            for (size_t i = 0; i < kept.n; i++) if (nms[i] < best) best = nms[i];
            size_t cn = 0;
            cls[r] = (uint32_t *)malloc((kept.n + 1) * sizeof(uint32_t));
            for (size_t i = 0; i < kept.n; i++) if (nms[i] == best && best != INT32_MAX) cls[r][cn++] = kept.v[i].asv;
            cls_n[r] = (uint32_t)cn;
            free(nms);
            if (cn == 0) { status[r] = SO_NO_ALIGNMENT; free(cls[r]); cls[r] = NULL; }
            else { status[r] = SO_ASSIGNED; best_nm_out[r] = best; }
            sketch_free(&sk);
        }
        free(hits); free(mism); free(kept.v);
    }
    /* :1890-1920 counters, serial so the output is deterministic */
    size_t n = 0; int rc = 0;
    class_offsets[0] = 0;
    for (size_t r = 0; r < n_reads; r++) {
        if (status[r] == SO_ASSIGNED) {
            uint32_t cn = cls_n[r];
            if (n + cn > class_cap) { rc = -1; }
            else {
                memcpy(class_asv + n, cls[r], cn * sizeof(uint32_t)); n += cn;
                if (cn == 1) {
                    unambig[cls[r][0]]++;
                    if (sample_singleton) {
                        uint32_t s = file_idx ? file_idx[r] : 0;
                        if ((int)s < n_samples) sample_singleton[(size_t)s * A + cls[r][0]]++;
                    }
                } else for (uint32_t t = 0; t < cn; t++) ambig[cls[r][t]]++;
                if (best_nm_out[r] <= 10) for (uint32_t t = 0; t < cn; t++) leq10[cls[r][t]]++;
                totals[0]++;
            }
        } else if (status[r] < SO_TOO_SHORT) {
            totals[1]++;
        }
        class_offsets[r + 1] = n;
        free(cls[r]);
    }
    free(cls); free(cls_n);
    if (dp_cells) *dp_cells += cells_total;
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* Stage 1 counting: seeding::split_kmer_mid (src/seeding.rs:975-1068) +                 */
/* seq_parse::read_to_split_kmers (src/seq_parse.rs:12-83, :315-497), bloom_filter_size 0 */
/* ------------------------------------------------------------------------------------ */
/* marked k-mers of one read: canonical k-mer | (forward is canonical) << 63; returns how many were written */
static size_t split_kmer_mid(const uint8_t *string, const uint8_t *qualities, size_t len, int k, int minimum_bq, uint64_t *out) {
    if (len < (size_t)k) return 0;                                           /* :982-984 */
    size_t n = 0;
    uint64_t rolling_kmer_f = 0, rolling_kmer_r = 0;
    const int marker_reverse_shift_dist = 2 * (k - 1);
    const uint64_t marker_mask = UINT64_MAX >> (64 - 2 * k);
    const uint64_t marker_rev_mask = ~((uint64_t)3 << (2 * k - 2));
    const uint64_t split_mask = ~((uint64_t)3 << (k - 1));
    const int mid_k = k / 2;
    int use_q = 0;                                                           /* :1004-1018 positions_to_skip */
    if (qualities) for (size_t i = 1; i < len; i++) if (qualities[i] != qualities[0]) { use_q = 1; break; }
    for (int i = 0; i < k - 1; i++) {                                        /* :1021-1028 */
        uint64_t nuc_f = so_encode_base(string[i]), nuc_r = 3 - nuc_f;
        rolling_kmer_f <<= 2; rolling_kmer_f |= nuc_f;
        rolling_kmer_r >>= 2; rolling_kmer_r |= nuc_r << marker_reverse_shift_dist;
    }
    for (size_t i = (size_t)k - 1; i < len; i++) {                           /* :1030-1065 */
        uint64_t nuc_f = so_encode_base(string[i]), nuc_r = 3 - nuc_f;
        rolling_kmer_f <<= 2; rolling_kmer_f |= nuc_f; rolling_kmer_f &= marker_mask;
        rolling_kmer_r >>= 2; rolling_kmer_r &= marker_rev_mask; rolling_kmer_r |= nuc_r << marker_reverse_shift_dist;
        uint64_t split_f = rolling_kmer_f & split_mask, split_r = rolling_kmer_r & split_mask;
        if (split_f == split_r) continue;                                    /* palindromic splitmer, :1044-1046 */
        if (use_q) {
            size_t mid_pos = i + 1 + (size_t)mid_k - (size_t)k;
            if ((int)qualities[mid_pos] - 33 < minimum_bq) continue;         /* :1012, :1049-1051 */
        }
        int canonical_marker = split_f < split_r;
        uint64_t canonical = canonical_marker ? rolling_kmer_f : rolling_kmer_r;
        out[n++] = canonical | ((uint64_t)canonical_marker << 63);           /* :1063 */
    }
    return n;
}

static int cmp_marked(const void *a, const void *b) {
    const uint64_t m = ~((uint64_t)1 << 63);
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    if ((x & m) != (y & m)) return (x & m) < (y & m) ? -1 : 1;
    return (x >> 63) < (y >> 63) ? -1 : ((x >> 63) > (y >> 63));
}

long long so_count_split_kmers(int k, int min_bq, int single_strand, const uint8_t *bases, const uint8_t *quals,
                               const uint64_t *offsets, const uint8_t *rc, size_t n_reads, uint64_t *out_kmers,
                               uint32_t *out_counts, size_t cap, uint64_t *n_distinct, int n_threads) {
    /* per-read output segments: at most len - k + 1 marked k-mers each */
    uint64_t *seg = (uint64_t *)malloc((n_reads + 1) * sizeof(uint64_t));
    seg[0] = 0;
    for (size_t r = 0; r < n_reads; r++) {
        size_t len = offsets[r + 1] - offsets[r];
        seg[r + 1] = seg[r] + (len >= (size_t)k ? len - (size_t)k + 1 : 0);
    }
    uint64_t *all = (uint64_t *)malloc((seg[n_reads] + 1) * sizeof(uint64_t));
    uint64_t *cnt = (uint64_t *)calloc(n_reads + 1, sizeof(uint64_t));
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel for schedule(dynamic, 64)
    for (long long rr = 0; rr < (long long)n_reads; rr++) {
        size_t r = (size_t)rr, off = offsets[r], len = offsets[r + 1] - off;
        const uint8_t *sq = bases + off, *ql = quals ? quals + off : NULL;
        uint8_t *tmp = NULL;
        if (rc && rc[r]) {                                                   /* src/seq_parse.rs:361-369 + utils::reverse_complement */
            tmp = (uint8_t *)malloc(2 * len + 2);
            for (size_t i = 0; i < len; i++) {
                uint8_t b = sq[len - 1 - i], c;
                switch (b) {                                                 /* src/utils.rs:51-65 reverse_complement */
                    case 'A': case 'a': c = 'T'; break;
                    case 'T': case 't': c = 'A'; break;
                    case 'C': case 'c': c = 'G'; break;
                    case 'G': case 'g': c = 'C'; break;
                    default: c = 'N';                                        /* N, n and anything else (also U) */
                }
                tmp[i] = c;
                if (ql) tmp[len + i] = ql[len - 1 - i];
            }
            sq = tmp; if (ql) ql = tmp + len;
        }
        cnt[r] = split_kmer_mid(sq, ql, len, k, min_bq, all + seg[r]);
        free(tmp);
    }
    /* compact, sort by (k-mer, flag), run-length count: the FxHashMap<u64,[u32;2]> of :455-460 */
    size_t n = 0;
    for (size_t r = 0; r < n_reads; r++) { memmove(all + n, all + seg[r], cnt[r] * sizeof(uint64_t)); n += cnt[r]; }
    qsort(all, n, sizeof(uint64_t), cmp_marked);
    const uint64_t mask = ~((uint64_t)1 << 63);
    long long n_out = 0; uint64_t distinct = 0; int rcode = 0;
    for (size_t i = 0; i < n;) {
        uint64_t kmer = all[i] & mask; uint32_t c[2] = {0, 0};
        size_t j = i;
        while (j < n && (all[j] & mask) == kmer) { c[all[j] >> 63]++; j++; }
        distinct++;
        int keep = single_strand ? (c[0] > 2) : (c[0] > 0 && c[1] > 0 && c[0] + c[1] > 2);   /* :30-44 */
        if (keep) {
            if ((size_t)n_out >= cap) { rcode = -1; break; }
            out_kmers[n_out] = kmer; out_counts[2 * n_out] = c[0]; out_counts[2 * n_out + 1] = c[1]; n_out++;
        }
        i = j;
    }
    if (n_distinct) *n_distinct = distinct;
    free(all); free(seg); free(cnt);
    return rcode ? -1 : n_out;
}

static int cmp_u32(const void *a, const void *b) { uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b; return x < y ? -1 : x > y; }
uint32_t so_high_freq_kmers(const uint64_t *kmers, const uint32_t *counts, size_t n, uint64_t *hf_out, size_t *n_hf) {
    uint32_t *kc = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));                  /* src/kmer_comp.rs:463-466 */
    for (size_t i = 0; i < n; i++) kc[i] = counts[2 * i] + counts[2 * i + 1];
    qsort(kc, n, sizeof(uint32_t), cmp_u32);
    uint32_t thresh = kc[n - (n / 100000) - 1];                                     /* :474 */
    if (thresh < 100) thresh = 100;
    free(kc);
    size_t m = 0;
    for (size_t i = 0; i < n; i++)
        if (counts[2 * i] + counts[2 * i + 1] > thresh) { if (hf_out) hf_out[m] = kmers[i]; m++; }   /* :497-499 */
    if (n_hf) *n_hf = m;
    return thresh;
}

/* ------------------------------------------------------------------------------------ */
/* a13-a14: EM, src/alignment.rs:1957-2038                                              */
/* ------------------------------------------------------------------------------------ */
int so_em(size_t n_asv, const uint64_t *cls_offsets, const uint32_t *cls_asv, const uint64_t *counts, size_t n_cls,
          uint64_t total_assigned, double *abundances, uint64_t *depths) {
    double *ab = (double *)malloc((n_asv + 1) * sizeof(double));
    double *nw = (double *)malloc((n_asv + 1) * sizeof(double));
    for (size_t a = 0; a < n_asv; a++) ab[a] = 1.0 / (double)n_asv;
    const double convergence_threshold = 0.01 / (double)total_assigned;
    int iteration = 0;
    const int MAX_ITERATIONS = 10000;
    while (1) {
        iteration++;
        for (size_t a = 0; a < n_asv; a++) nw[a] = 0.0;
        for (size_t q = 0; q < n_cls; q++) {
            double denominator = 0.0;
            for (uint64_t e = cls_offsets[q]; e < cls_offsets[q + 1]; e++) denominator += ab[cls_asv[e]];
            if (denominator > 0.0)
                for (uint64_t e = cls_offsets[q]; e < cls_offsets[q + 1]; e++)
                    nw[cls_asv[e]] += (double)counts[q] * ab[cls_asv[e]] / denominator;
        }
        double total_counts = 0.0;
        for (size_t a = 0; a < n_asv; a++) total_counts += nw[a];
        if (total_counts > 0.0) for (size_t a = 0; a < n_asv; a++) nw[a] /= (double)total_assigned;
        double max_change = 0.0;
        for (size_t a = 0; a < n_asv; a++) { double d = fabs(ab[a] - nw[a]); if (d > max_change) max_change = d; }
        memcpy(ab, nw, n_asv * sizeof(double));
        if (max_change < convergence_threshold || iteration >= MAX_ITERATIONS) break;
    }
    for (size_t a = 0; a < n_asv; a++) {
        if (abundances) abundances[a] = ab[a];
        if (depths) depths[a] = (uint64_t)round(ab[a] * (double)total_assigned);   /* f64::round: half away from zero */
    }
    free(ab); free(nw);
    return iteration;
}
