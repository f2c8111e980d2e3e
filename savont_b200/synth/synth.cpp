// synth.cpp -- synthetic 16S-like workload generator for the parity tests and bench.py
// (SURVEY.md 8d / BASELINE.md configs 2-5).  Harness code: not part of the product path and not
// part of the oracle.  Deterministic for a given seed regardless of the thread count (every ASV
// and every read has its own RNG stream).
//
//   ASVs   n_roots random "genus" roots (uniform ACGT at the requested GC) x n_var variants; every
//          genus owns a pool of biallelic SNP sites, each variant flips 1..15 of them (plus 0..2
//          single-base indels), so SNPmers are shared inside a genus the way real 16S copies are.
//   reads  ASV ~ log-normal(sigma) abundance (per sample), random strand, 0..20 bases trimmed at
//          either end, errors at `err_rate` split 40/30/30 sub/ins/del (half of the indels extend or
//          contract a homopolymer), Phred 10..15 at error sites and q_lo..q_hi elsewhere.
//   SNPmer set derived the way Stage 1 would call it from the ASVs: canonical-by-splitmer k-mers
//          grouped by splitmer (palindromic splitmers skipped, reference src/seeding.rs:1044-1046),
//          every splitmer with >= 2 mid bases contributes its two most frequent alleles
//          (reference src/kmer_comp.rs:552-605).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cctype>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed ? seed : 0x9E3779B97F4A7C15ull) { next(); next(); }
    uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    double normal() { double u1 = uniform(), u2 = uniform(); if (u1 < 1e-300) u1 = 1e-300; return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2); }
};

uint64_t mix(uint64_t a, uint64_t b) { Rng r(a * 0xD1342543DE82EF95ull + b + 0x632BE59BD9B4E019ull); return r.next(); }

struct Config {
    uint64_t seed; int n_roots, n_var, len_mean, len_jitter; double gc;
    double err_rate; int q_lo, q_hi; double abund_sigma; int n_samples; int trim_max; int sites_per_genus;
};

struct Synth {
    Config cfg;
    std::vector<std::string> asvs;
    std::vector<std::vector<double>> cum_abund;   // per sample cumulative abundance over ASVs
    std::vector<double> sample_cum;               // cumulative sample sizes
};

const char ACGT[4] = {'A', 'C', 'G', 'T'};
inline int code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }

void build_asvs(Synth &S) {
    const Config &c = S.cfg;
    S.asvs.assign((size_t)c.n_roots * c.n_var, std::string());
    for (int g = 0; g < c.n_roots; ++g) {
        Rng r(mix(c.seed, 1000003ull * (uint64_t)g + 17));
        const int len = c.len_mean + (c.len_jitter ? (int)r.below(2 * c.len_jitter + 1) - c.len_jitter : 0);
        std::string root((size_t)len, 'A');
        for (int i = 0; i < len; ++i) {
            const bool gcb = r.uniform() < c.gc;
            const bool hi = r.uniform() < 0.5;
            root[i] = gcb ? (hi ? 'G' : 'C') : (hi ? 'T' : 'A');
        }
        // pool of biallelic sites, at least 40 bases from the ends and 20 apart
        std::vector<int> sites; std::vector<char> alt;
        int guard = 0;
        while ((int)sites.size() < c.sites_per_genus && guard++ < 100000) {
            const int p = 40 + (int)r.below((uint32_t)std::max(1, len - 80));
            bool ok = true;
            for (int q : sites) if (std::abs(q - p) < 20) { ok = false; break; }
            if (!ok) continue;
            sites.push_back(p);
            char a;
            do { a = ACGT[r.below(4)]; } while (a == root[p]);
            alt.push_back(a);
        }
        for (int v = 0; v < c.n_var; ++v) {
            Rng rv(mix(c.seed, 7919ull * (uint64_t)(g * c.n_var + v) + 3));
            std::string s = root;
            const int nsub = 1 + (int)rv.below(15);
            std::vector<int> idx(sites.size());
            for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
            for (int t = 0; t < nsub && t < (int)idx.size(); ++t) {
                const int j = t + (int)rv.below((uint32_t)(idx.size() - t));
                std::swap(idx[t], idx[j]);
                s[sites[idx[t]]] = alt[idx[t]];
            }
            const int nindel = (int)rv.below(3);
            for (int t = 0; t < nindel; ++t) {
                const int p = 30 + (int)rv.below((uint32_t)std::max(1, (int)s.size() - 60));
                if (rv.uniform() < 0.5) s.insert(s.begin() + p, ACGT[rv.below(4)]);
                else s.erase(s.begin() + p);
            }
            S.asvs[(size_t)g * c.n_var + v] = s;
        }
    }
    // abundances: log-normal per (sample, ASV); sample sizes ~ Dirichlet(1) (normalised exponentials)
    const size_t A = S.asvs.size();
    S.cum_abund.assign((size_t)c.n_samples, std::vector<double>(A));
    S.sample_cum.assign((size_t)c.n_samples, 0.0);
    Rng rs(mix(c.seed, 424242));
    double tot = 0;
    for (int s = 0; s < c.n_samples; ++s) { tot += -std::log(std::max(rs.uniform(), 1e-12)); S.sample_cum[s] = tot; }
    for (int s = 0; s < c.n_samples; ++s) S.sample_cum[s] /= tot;
    for (int s = 0; s < c.n_samples; ++s) {
        Rng ra(mix(c.seed, 99991ull * (uint64_t)s + 5));
        double acc = 0;
        for (size_t a = 0; a < A; ++a) { acc += std::exp(c.abund_sigma * ra.normal()); S.cum_abund[s][a] = acc; }
        for (size_t a = 0; a < A; ++a) S.cum_abund[s][a] /= acc;
    }
}

// simulate one read; if out == nullptr only the length is computed
size_t make_read(const Synth &S, uint64_t idx, uint8_t *bases, uint8_t *quals, uint32_t *truth_asv, uint8_t *truth_strand,
                 uint32_t *sample_out) {
    const Config &c = S.cfg;
    Rng r(mix(c.seed ^ 0xABCDEF12345ull, idx));
    int sample = 0;
    if (c.n_samples > 1) {
        const double u = r.uniform();
        sample = (int)(std::lower_bound(S.sample_cum.begin(), S.sample_cum.end(), u) - S.sample_cum.begin());
        if (sample >= c.n_samples) sample = c.n_samples - 1;
    }
    const std::vector<double> &cum = S.cum_abund[(size_t)sample];
    size_t a = (size_t)(std::lower_bound(cum.begin(), cum.end(), r.uniform()) - cum.begin());
    if (a >= cum.size()) a = cum.size() - 1;
    const std::string &src = S.asvs[a];
    const int t0 = (int)r.below((uint32_t)c.trim_max + 1), t1 = (int)r.below((uint32_t)c.trim_max + 1);
    const int lo = t0, hi = (int)src.size() - t1;
    const bool rev = r.uniform() < 0.5;
    size_t n = 0;
    auto emit = [&](char b, bool err) {
        if (bases) {
            bases[n] = (uint8_t)b;
            const int q = err ? 10 + (int)r.below(6) : c.q_lo + (int)r.below((uint32_t)(c.q_hi - c.q_lo + 1));
            quals[n] = (uint8_t)(33 + std::min(q, 93));
        } else {
            if (err) r.below(6); else r.below((uint32_t)(c.q_hi - c.q_lo + 1));
        }
        ++n;
    };
    char prev = 'A';
    for (int i = lo; i < hi; ++i) {
        const char b = src[(size_t)i];
        const double u = r.uniform();
        if (u < c.err_rate) {
            const double t = r.uniform();
            if (t < 0.4) {                       // substitution
                char x; do { x = ACGT[r.below(4)]; } while (x == b);
                emit(x, true); prev = x;
            } else if (t < 0.7) {                // insertion (before this base)
                const char x = (r.uniform() < 0.5) ? prev : ACGT[r.below(4)];
                emit(x, true); emit(b, false); prev = b;
            } else {                             // deletion
                r.uniform();
            }
        } else { emit(b, false); prev = b; }
    }
    if (bases && rev) {
        for (size_t i = 0, j = n ? n - 1 : 0; i < j; ++i, --j) { std::swap(bases[i], bases[j]); std::swap(quals[i], quals[j]); }
        for (size_t i = 0; i < n; ++i) bases[i] = (uint8_t)ACGT[3 - code((char)bases[i])];
    }
    if (truth_asv) *truth_asv = (uint32_t)a;
    if (truth_strand) *truth_strand = rev ? 1 : 0;
    if (sample_out) *sample_out = (uint32_t)sample;
    return n;
}

}  // namespace

extern "C" {

void *synth_create(uint64_t seed, int n_roots, int n_var, int len_mean, int len_jitter, double gc, double err_rate, int q_lo,
                   int q_hi, double abund_sigma, int n_samples, int trim_max, int sites_per_genus) {
    Synth *S = new Synth();
    S->cfg = Config{seed, n_roots, n_var, len_mean, len_jitter, gc, err_rate, q_lo, q_hi, abund_sigma, n_samples < 1 ? 1 : n_samples,
                    trim_max, sites_per_genus};
    build_asvs(*S);
    return S;
}
void synth_destroy(void *h) { delete (Synth *)h; }
size_t synth_num_asvs(void *h) { return ((Synth *)h)->asvs.size(); }
size_t synth_asv_bytes(void *h) { size_t t = 0; for (auto &s : ((Synth *)h)->asvs) t += s.size(); return t; }
void synth_get_asvs(void *h, uint8_t *bases, uint64_t *offsets) {
    Synth *S = (Synth *)h;
    uint64_t off = 0; offsets[0] = 0;
    for (size_t a = 0; a < S->asvs.size(); ++a) { memcpy(bases + off, S->asvs[a].data(), S->asvs[a].size()); off += S->asvs[a].size(); offsets[a + 1] = off; }
}

static size_t snpmers_of(const std::vector<std::string> &asvs, int k, uint64_t *out, size_t cap);

// SNPmer set from the ASVs; returns the number written (or needed if cap too small)
size_t synth_snpmers(void *h, int k, uint64_t *out, size_t cap) { return snpmers_of(((Synth *)h)->asvs, k, out, cap); }

// same derivation for caller-supplied sequences (any non-ACGT base is read as A, like BYTE_TO_SEQ)
size_t synth_snpmers_from(const uint8_t *bases, const uint64_t *offsets, size_t n, int k, uint64_t *out, size_t cap) {
    std::vector<std::string> v(n);
    for (size_t i = 0; i < n; ++i) {
        v[i].assign((const char *)bases + offsets[i], (const char *)bases + offsets[i + 1]);
        for (char &c : v[i]) { c = (char)toupper((unsigned char)c); if (c == 'U') c = 'T'; if (c != 'A' && c != 'C' && c != 'G' && c != 'T') c = 'A'; }
    }
    return snpmers_of(v, k, out, cap);
}

static size_t snpmers_of(const std::vector<std::string> &asvs, int k, uint64_t *out, size_t cap) {
    const uint64_t split_mask = ~(3ull << (k - 1));
    const uint64_t kmask = (1ull << (2 * k)) - 1;
    std::vector<uint64_t> all;   // canonical k-mers (one per ASV occurrence)
    for (auto &s : asvs) {
        uint64_t f = 0, r = 0;
        for (size_t i = 0; i < s.size(); ++i) {
            const uint64_t c = (uint64_t)code(s[i]);
            f = ((f << 2) | c) & kmask;
            r = (r >> 2) | ((3 - c) << (2 * (k - 1)));
            if (i + 1 < (size_t)k) continue;
            const uint64_t sf = f & split_mask, sr = r & split_mask;
            if (sf == sr) continue;                                   // palindromic splitmer
            all.push_back(sf < sr ? f : r);
        }
    }
    std::sort(all.begin(), all.end(), [&](uint64_t x, uint64_t y) {
        const uint64_t sx = x & split_mask, sy = y & split_mask;
        return sx != sy ? sx < sy : x < y;
    });
    size_t n = 0;
    for (size_t i = 0; i < all.size();) {
        size_t j = i;
        uint64_t cnt[4] = {0, 0, 0, 0};
        while (j < all.size() && (all[j] & split_mask) == (all[i] & split_mask)) { cnt[(all[j] >> (k - 1)) & 3]++; ++j; }
        int distinct = 0; for (int b = 0; b < 4; ++b) distinct += cnt[b] > 0;
        if (distinct >= 2) {
            int best = -1, second = -1;
            for (int b = 0; b < 4; ++b) if (cnt[b] && (best < 0 || cnt[b] > cnt[best])) best = b;
            for (int b = 0; b < 4; ++b) if (b != best && cnt[b] && (second < 0 || cnt[b] > cnt[second])) second = b;
            const uint64_t sp = all[i] & split_mask;
            if (n < cap) out[n] = sp | ((uint64_t)best << (k - 1)); ++n;
            if (n < cap) out[n] = sp | ((uint64_t)second << (k - 1)); ++n;
        }
        i = j;
    }
    return n;
}

// lengths[n] of reads first..first+n
void synth_read_lengths(void *h, uint64_t first, size_t n, uint32_t *lengths) {
    Synth *S = (Synth *)h;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) lengths[i] = (uint32_t)make_read(*S, first + (uint64_t)i, nullptr, nullptr, nullptr, nullptr, nullptr);
}
// fills bases/quals at offsets[i] (offsets from the prefix sum of lengths)
void synth_reads(void *h, uint64_t first, size_t n, const uint64_t *offsets, uint8_t *bases, uint8_t *quals, uint32_t *truth_asv,
                 uint8_t *truth_strand, uint32_t *file_idx) {
    Synth *S = (Synth *)h;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i)
        make_read(*S, first + (uint64_t)i, bases + offsets[i], quals + offsets[i], truth_asv ? truth_asv + i : nullptr,
                  truth_strand ? truth_strand + i : nullptr, file_idx ? file_idx + i : nullptr);
}

}  // extern "C"
